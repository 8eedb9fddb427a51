"""GeoPandas / shapely objects in, GeoArrow buffers for the hot path out (SURVEY.md section 8f
row 4; reference: python/cuspatial/cuspatial/io/geopandas_reader.py:27-84).

The reference walks a geopandas.GeoSeries and reads every geometry through
`shapely.geometry.mapping(geom)["coordinates"]`, i.e. through the `__geo_interface__` protocol.
This reader does the same walk over ANY iterable of objects that implement that protocol
(shapely geometries, the elements of a GeoSeries, GeoJSON-like dicts with "type" and
"coordinates"), so it neither imports nor requires geopandas/shapely -- they are simply the usual
producers of such objects -- and it builds the three single-type containers of `geoarrow.py`
that the spatial-join functions accept:

    POINT                         -> from_points_xy
    LINESTRING / MULTILINESTRING  -> from_linestrings_xy
    POLYGON / MULTIPOLYGON        -> from_polygons_xy   (rings as given: exterior, then holes)

A series mixing these families, or holding None / MULTIPOINT / collections, is rejected: the
join path of the reference rejects it as well (`contains_only_points`, `contains_only_polygons`,
`contains_only_linestrings`, core/spatial/join.py:62-73,224-228,305-318).
"""
import numpy as np

from . import geoarrow

_FAMILY = {"Point": "points", "LineString": "linestrings", "MultiLineString": "linestrings",
           "Polygon": "polygons", "MultiPolygon": "polygons"}


def _mapping(geom):
    if isinstance(geom, dict):
        return geom
    gi = getattr(geom, "__geo_interface__", None)
    if gi is None:
        raise TypeError(type(geom))  # geopandas_reader.py:75
    return gi


def parse_geometries(geoseries):
    """One pass over the geometries.  Returns (family, xy float64[2n], ring_offset | None,
    part_offset | None, geometry_offset | None) with int32 offsets in the GeoArrow nesting the
    reference uses (coordinates <- rings <- parts <- geometries)."""
    family = None
    xy, ring_off, part_off, geom_off = [], [0], [0], [0]
    n_coords = 0
    for geom in geoseries:
        if geom is None:
            raise ValueError("null geometries are not supported on the spatial-join path")
        m = _mapping(geom)
        kind, coords = m["type"], m["coordinates"]
        fam = _FAMILY.get(kind)
        if fam is None:
            raise TypeError("unsupported geometry type on the spatial-join path: %s" % kind)
        if family is None:
            family = fam
        elif fam != family:
            raise TypeError("the series mixes %s and %s" % (family, fam))
        if fam == "points":
            xy.append(np.asarray(coords, dtype=np.float64)[:2].reshape(1, 2))
            n_coords += 1
            continue
        if kind == "LineString":
            coords = [coords]
        elif kind == "Polygon":
            coords = [coords]
        if fam == "linestrings":
            for part in coords:  # a part is a list of coordinates
                a = np.asarray(part, dtype=np.float64)[:, :2]
                xy.append(a)
                n_coords += len(a)
                part_off.append(n_coords)
            geom_off.append(len(part_off) - 1)
        else:
            for part in coords:  # a part (polygon) is a list of rings
                for ring in part:
                    a = np.asarray(ring, dtype=np.float64)[:, :2]
                    xy.append(a)
                    n_coords += len(a)
                    ring_off.append(n_coords)
                part_off.append(len(ring_off) - 1)
            geom_off.append(len(part_off) - 1)
    flat = np.concatenate(xy).reshape(-1) if xy else np.empty(0, np.float64)
    i32 = lambda v: np.asarray(v, dtype=np.int32)  # noqa: E731
    if family in (None, "points"):
        return "points", flat, None, None, None
    if family == "linestrings":
        return family, flat, None, i32(part_off), i32(geom_off)
    return family, flat, i32(ring_off), i32(part_off), i32(geom_off)


def from_geopandas(geoseries, dtype=np.float64):
    """cuspatial.from_geopandas for the single-type series the join path takes: a
    geoarrow.GeoArrowSeries (device buffers when a GPU is present)."""
    family, xy, ring_off, part_off, geom_off = parse_geometries(geoseries)
    xy = xy.astype(dtype, copy=False)
    if family == "points":
        return geoarrow.from_points_xy(xy)
    if family == "linestrings":
        return geoarrow.from_linestrings_xy(xy, part_off, geom_off)
    return geoarrow.from_polygons_xy(xy, ring_off, part_off, geom_off)

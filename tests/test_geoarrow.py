"""GeoArrow ingestion: the from_*_xy constructors and the accessor shape the
reference's wrappers read.  Container logic runs on CPU tensors; the API itself still insists on
device memory (no CPU fallback)."""
import numpy as np
import pytest
import torch

import cuspatial_b200 as cs
from cuspatial_b200 import geoarrow as ga

SQUARES_XY = np.array([0, 0, 1, 0, 1, 1, 0, 1, 0, 0, 5, 5, 6, 5, 6, 6, 5, 6, 5, 5], dtype=np.float64)


def test_constructors_keep_the_interleaved_buffer_and_split_lazily():
    xy = torch.arange(10, dtype=torch.float64)
    p = cs.from_points_xy(xy)
    assert len(p) == 5 and p.points.xy.data_ptr() == xy.data_ptr()        # zero-copy
    assert p.points.x.tolist() == [0, 2, 4, 6, 8] and p.points.y.tolist() == [1, 3, 5, 7, 9]
    assert p.points.x is p.points.x                                       # materialised once
    poly = cs.from_polygons_xy(torch.as_tensor(SQUARES_XY), [0, 5, 10], [0, 1, 2], [0, 1, 2])
    a = poly.polygons
    assert len(poly) == 2 and a.ring_offset.tolist() == [0, 5, 10] and a.part_offset.dtype == torch.int32
    assert a.x.tolist() == SQUARES_XY[0::2].tolist() and a.y.tolist() == SQUARES_XY[1::2].tolist()
    assert not ga.is_multi(poly)
    lines = cs.from_linestrings_xy(torch.as_tensor(SQUARES_XY), [0, 5, 10], [0, 2])
    assert len(lines) == 1 and ga.is_multi(lines) and lines.lines.geometry_offset.tolist() == [0, 2]


def test_wrong_kind_odd_length_and_integer_coordinates_are_rejected():
    p = cs.from_points_xy(torch.zeros(4, dtype=torch.float32))
    with pytest.raises(ValueError, match="holds points"):
        p.polygons
    with pytest.raises(ValueError, match="even number"):
        cs.from_points_xy(torch.zeros(3, dtype=torch.float64))
    with pytest.raises(TypeError, match="float32 or float64"):
        cs.from_points_xy(torch.zeros(4, dtype=torch.int32))


def test_multipolygons_and_multilinestrings_are_refused_like_the_reference():
    pts = cs.from_points_xy(torch.zeros(4, dtype=torch.float64))
    mp = cs.from_polygons_xy(torch.as_tensor(SQUARES_XY), [0, 5, 10], [0, 1, 2], [0, 2])
    with pytest.raises(ValueError, match="cannot contain multipolygon"):        # join.py:75-78
        cs.point_in_polygon(pts, mp)
    ml = cs.from_linestrings_xy(torch.as_tensor(SQUARES_XY), [0, 5, 10], [0, 2])
    with pytest.raises(ValueError, match="cannot contain multilinestrings"):    # join.py:323-326
        cs.quadtree_point_to_nearest_linestring((None, None), None, None, pts, ml)


def test_cpu_buffers_never_reach_a_cpu_fallback():
    pts = cs.from_points_xy(torch.zeros(4, dtype=torch.float64))
    with pytest.raises(ValueError, match="CUDA tensor"):
        cs.quadtree_on_points(pts, 0, 1, 0, 1, 1, 3, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_geoarrow_series_through_the_whole_path(golden, dtype):
    """The reference's 71-point / 4-polygon case fed as GeoArrow buffers gives the golden rows;
    linestring boxes of a multilinestring series are per geometry (bounding.py:123-125)."""
    sj = golden["small_join"]
    b = sj["bbox"]
    p = np.array(sj["points"], dtype=dtype)
    v = np.array(sj["vertices"], dtype=dtype)
    dev = "cuda"
    pts = cs.from_points_xy(torch.as_tensor(p.reshape(-1), device=dev))
    polys = cs.from_polygons_xy(torch.as_tensor(v.reshape(-1), device=dev), sj["ring_offsets"],
                                sj["part_offsets"], sj["geometry_offsets"])
    pidx, tree = cs.quadtree_on_points(pts, b[0], b[1], b[2], b[3], sj["scale"], sj["max_depth"],
                                       sj["max_size"])
    bb = cs.polygon_bounding_boxes(polys)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, b[0], b[1], b[2], b[3], sj["scale"],
                                                sj["max_depth"])
    assert pairs["bbox_offset"].cpu().numpy().tolist() == sj["pair_poly"]
    assert pairs["quad_offset"].cpu().numpy().tolist() == sj["pair_quad"]
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, pts, polys)
    assert hits["polygon_index"].cpu().numpy().tolist() == sj["pip_poly"]
    assert hits["point_index"].cpu().numpy().tolist() == sj["pip_point"]
    rows = cs.contains_properly(polys, pts, mode="quadtree")
    assert len(rows) == len(sj["pip_poly"])
    # two linestrings as ONE multilinestring geometry: a single box over both parts
    n = golden["nearest_linestring"]
    lv = np.array(n["vertices"], dtype=dtype)
    ml = cs.from_linestrings_xy(torch.as_tensor(lv.reshape(-1), device=dev), n["line_offsets"],
                                [0, 2, 3, 4])
    lb = cs.linestring_bounding_boxes(ml, 0.0)
    first_two = lv[: n["line_offsets"][2]]
    assert len(lb) == 3
    assert float(lb["minx"][0]) == first_two[:, 0].min() and float(lb["maxy"][0]) == first_two[:, 1].max()


class _Geo:
    """Stand-in for a shapely geometry: all the reader needs is `__geo_interface__`."""

    def __init__(self, kind, coords):
        self.__geo_interface__ = {"type": kind, "coordinates": coords}


def test_geopandas_reader_walks_the_geo_interface_like_the_reference():
    """geopandas_reader.py:27-84 reads shapely.geometry.mapping(geom)['coordinates']; so does this
    reader, for any object speaking __geo_interface__ (shapely itself is not installed here)."""
    sq = lambda x0, y0: ((x0, y0), (x0 + 1, y0), (x0 + 1, y0 + 1), (x0, y0 + 1), (x0, y0))  # noqa: E731
    hole = ((0.2, 0.2), (0.2, 0.4), (0.4, 0.4), (0.4, 0.2), (0.2, 0.2))
    polys = cs.from_geopandas([_Geo("Polygon", (sq(0, 0), hole)),
                               {"type": "MultiPolygon", "coordinates": [(sq(5, 5),), (sq(8, 8),)]}])
    a = polys.polygons
    assert len(polys) == 2 and ga.is_multi(polys)
    assert a.geometry_offset.tolist() == [0, 1, 3] and a.part_offset.tolist() == [0, 2, 3, 4]
    assert a.ring_offset.tolist() == [0, 5, 10, 15, 20] and a.xy.dtype == torch.float64
    assert a.x[:5].tolist() == [0, 1, 1, 0, 0] and a.y[5:10].tolist() == [0.2, 0.4, 0.4, 0.2, 0.2]
    pts = cs.from_geopandas([_Geo("Point", (1.5, 2.5)), _Geo("Point", (3.0, 4.0, 9.0))], np.float32)
    assert pts.points.xy.tolist() == [1.5, 2.5, 3.0, 4.0] and pts.points.xy.dtype == torch.float32
    ls = cs.from_geopandas([_Geo("LineString", ((0, 0), (1, 1), (2, 0))),
                            _Geo("MultiLineString", (((5, 5), (6, 6)), ((7, 7), (8, 8), (9, 9))))])
    assert ls.lines.part_offset.tolist() == [0, 3, 5, 8] and ls.lines.geometry_offset.tolist() == [0, 1, 3]
    with pytest.raises(TypeError, match="mixes"):
        cs.from_geopandas([_Geo("Point", (0, 0)), _Geo("Polygon", (sq(0, 0),))])
    with pytest.raises(TypeError, match="unsupported geometry type"):
        cs.from_geopandas([_Geo("MultiPoint", ((0, 0), (1, 1)))])
    with pytest.raises(TypeError):
        cs.from_geopandas([object()])


def test_geopandas_reader_on_real_shapely_objects_when_installed():
    """Run-time probe: with shapely / geopandas importable the same reader takes their objects
    (and a GeoSeries) unchanged.  Neither can be installed offline in this image -> skipped."""
    shapely_geometry = pytest.importorskip("shapely.geometry")
    polys = [shapely_geometry.Polygon([(0, 0), (1, 0), (1, 1), (0, 1)]),
             shapely_geometry.Polygon([(5, 5), (6, 5), (6, 6), (5, 6)])]
    try:
        import geopandas

        polys = geopandas.GeoSeries(polys)
    except ImportError:
        pass
    s = cs.from_geopandas(polys)
    assert len(s) == 2 and s.polygons.ring_offset.tolist() == [0, 5, 10]

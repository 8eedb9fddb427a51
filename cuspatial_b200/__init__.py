"""cuspatial_b200 -- B200-native quadtree point-in-polygon spatial join.

Drop-in for one hot path of rapidsai/cuspatial (quadtree_on_points ->
join_quadtree_and_bounding_boxes -> quadtree_point_in_polygon, plus the bitmask and pairwise
point_in_polygon, polygon_bounding_boxes and the contains_properly driver).  Hand-written sm_100a CUDA behind a C ABI
(include/cuspatial_b200.h); this package is the thin ctypes layer mirroring the reference's Python
functions.  There is no CPU fallback.
"""
from .api import (contains_properly, join_quadtree_and_bounding_boxes,
                  linestring_bounding_boxes, quadtree_point_to_nearest_linestring,
                  pairwise_point_in_polygon, point_in_polygon, point_in_polygon_bitmask,
                  polygon_bounding_boxes, quadtree_on_points, quadtree_point_in_polygon)
from .frame import Frame
from .geoarrow import from_linestrings_xy, from_points_xy, from_polygons_xy
from .geopandas_reader import from_geopandas

__all__ = [
    "quadtree_on_points", "join_quadtree_and_bounding_boxes", "quadtree_point_in_polygon",
    "point_in_polygon", "point_in_polygon_bitmask", "polygon_bounding_boxes",
    "pairwise_point_in_polygon", "contains_properly", "quadtree_point_to_nearest_linestring",
    "linestring_bounding_boxes", "from_points_xy", "from_polygons_xy", "from_linestrings_xy",
    "from_geopandas", "Frame",
]

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

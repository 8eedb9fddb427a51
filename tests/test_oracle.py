"""CPU tests (no GPU): pin the oracle.

1. oracle/oracle.cpp against every golden vector harvested from the reference's own tests;
2. the reference's own header-only implementation compiled for the host (oracle/_ref) against the
   same vectors -- proves the harness drives the reference correctly;
3. oracle == reference host build, bit for bit, on randomised inputs (fp32 + fp64; uniform and
   clustered; out-of-bbox points; duplicates; shallow/deep trees; max_size 1).
"""
import numpy as np
import pytest

from util import assert_same, make_case, pairwise_case, run_host


def _check_golden(lib, golden):
    for dt in (np.float32, np.float64):
        for c in golden["quadtree_cases"]:
            pts = np.array(c["points"], dtype=dt).reshape(-1, 2)
            t = lib.quadtree_on_points(pts[:, 0].copy(), pts[:, 1].copy(), c["v_min"][0],
                                       c["v_max"][0], c["v_min"][1], c["v_max"][1], c["scale"],
                                       c["max_depth"], c["max_size"])
            for k in ("key", "level", "is_internal_node", "length", "offset"):
                assert list(t[k]) == c[k], (c["name"], k, dt)
            assert len(t["point_indices"]) == len(pts)
        sj = golden["small_join"]
        pts = np.array(sj["points"], dtype=dt)
        x, y = pts[:, 0].copy(), pts[:, 1].copy()
        t = lib.quadtree_on_points(x, y, 0, 8, 0, 8, sj["scale"], sj["max_depth"], sj["max_size"])
        v = np.array(sj["vertices"], dtype=dt)
        vx, vy = v[:, 0].copy(), v[:, 1].copy()
        bb = lib.polygon_bounding_boxes(sj["part_offsets"], sj["ring_offsets"], vx, vy)
        pp, pq = lib.join_quadtree_and_bounding_boxes(t, *bb, 0, 0, sj["scale"], sj["max_depth"])
        assert list(pp) == sj["pair_poly"] and list(pq) == sj["pair_quad"]
        a, b = lib.quadtree_point_in_polygon(pp, pq, t, t["point_indices"], x, y,
                                             sj["part_offsets"], sj["ring_offsets"], vx, vy)
        assert list(a) == sj["pip_poly"] and list(b) == sj["pip_point"]
        # linestring bboxes expanded by 2.0 (test_spatial_join.py:321-432)
        lj = golden["linestring_join"]
        r = dt(lj["expansion_radius"])
        ro = sj["ring_offsets"]
        lb = [np.array([f(vv[ro[i]:ro[i + 1]]) for i in range(len(ro) - 1)], dtype=dt)
              for vv, f in ((vx - r, np.min), (vy - r, np.min), (vx + r, np.max), (vy + r, np.max))]
        pp, pq = lib.join_quadtree_and_bounding_boxes(t, *lb, 0, 0, sj["scale"], sj["max_depth"])
        assert list(pp) == lj["bbox_offset"] and list(pq) == lj["quad_offset"]
        for c in golden["pip_cases"]:
            p = np.array(c["points"], dtype=dt)
            v = np.array(c["vertices"], dtype=dt)
            m = lib.point_in_polygon(p[:, 0].copy(), p[:, 1].copy(), c["part_offsets"],
                                     c["ring_offsets"], v[:, 0].copy(), v[:, 1].copy())
            assert list(m) == c["expected_mask"], (c["name"], dt)
        for c in golden["pairwise_cases"]:
            v = np.array(c["vertices"], dtype=dt)
            for call in c["calls"]:
                p = np.array(call["points"], dtype=dt)
                k = len(p)  # point i vs polygon i of the first k polygons
                got = lib.pairwise_point_in_polygon(
                    p[:, 0].copy(), p[:, 1].copy(), c["part_offsets"][:k + 1], c["ring_offsets"],
                    v[:, 0].copy(), v[:, 1].copy())
                assert list(got) == call["expected"], (c["name"], dt)


def test_oracle_matches_reference_golden_vectors(oracle_lib, golden):
    _check_golden(oracle_lib, golden)


def test_reference_host_build_matches_its_own_golden_vectors(reference_lib, golden):
    _check_golden(reference_lib, golden)


CASES = [
    # n, n_poly, depth, max_size, kind, oob, dups
    (20000, 30, 15, 64, "u", 0, 0),
    (50000, 263, 15, 512, "u", 0, 0),
    (30000, 50, 8, 20, "c", 50, 100),
    (5000, 10, 3, 5, "u", 0, 0),
    (10000, 20, 15, 1, "c", 10, 300),
    (1000, 5, 1, 10, "u", 0, 0),
    (3000, 7, 2, 1, "u", 3, 0),
    (40000, 100, 12, 100, "c", 0, 0),
    (1, 3, 15, 1, "u", 0, 0),
    (2, 3, 4, 1, "u", 1, 0),
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_host_build(oracle_lib, reference_lib, case, dtype):
    n, n_poly, depth, max_size, kind, oob, dups = case
    c = make_case(n, n_poly, depth, kind, dtype, seed=n + depth, oob=oob, dups=dups)
    assert_same(run_host(oracle_lib, c, max_size), run_host(reference_lib, c, max_size),
                "oracle vs reference")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bitmask_oracle_equals_reference(oracle_lib, reference_lib, dtype):
    c = make_case(20000, 31, 8, "u", dtype, seed=3)
    po, ro = c["po"].astype(np.int32), c["ro"].astype(np.int32)
    a = oracle_lib.point_in_polygon(c["x"], c["y"], po, ro, c["vx"], c["vy"])
    b = reference_lib.point_in_polygon(c["x"], c["y"], po, ro, c["vx"], c["vy"])
    np.testing.assert_array_equal(a, b)
    assert np.count_nonzero(a) > 1000


def test_near_edge_points_oracle_equals_reference(oracle_lib, reference_lib):
    """Points a few ULP around edges/vertices, incl. vertical edges: the 4-ULP on-edge rule."""
    rng = np.random.default_rng(7)
    for dt in (np.float32, np.float64):
        c = make_case(10, 12, 6, "u", dt, seed=9, median_vertices=16)
        vx, vy, ro = c["vx"], c["vy"], c["ro"]
        # an axis-aligned square as an extra polygon => vertical/horizontal edges
        sq = np.array([[0.3, 0.3], [0.6, 0.3], [0.6, 0.6], [0.3, 0.6], [0.3, 0.3]], dtype=dt)
        vx = np.concatenate([vx, sq[:, 0]]); vy = np.concatenate([vy, sq[:, 1]])
        ro = np.concatenate([ro, [ro[-1] + 5]]).astype(np.uint32)
        po = np.concatenate([c["po"], [c["po"][-1] + 1]]).astype(np.uint32)
        xs, ys = [], []
        for i in range(len(vx)):
            j = i + 1 if i + 1 < len(vx) else i
            for t in (0.0, 0.25, 0.5, 1.0):
                px, py = vx[i] + t * (vx[j] - vx[i]), vy[i] + t * (vy[j] - vy[i])
                for k in range(-6, 7, 3):
                    xs.append(np.nextafter(px, dt(np.inf)) if k > 0 else px)
                    ys.append(py + dt(k) * np.spacing(py))
                    xs.append(px + dt(k) * np.spacing(px)); ys.append(py)
        x = np.array(xs, dtype=dt); y = np.array(ys, dtype=dt)
        keep = (x > c["ext"][0]) & (x < c["ext"][1]) & (y > c["ext"][2]) & (y < c["ext"][3])
        c2 = dict(c, x=x[keep], y=y[keep], po=po, ro=ro, vx=vx, vy=vy)
        assert_same(run_host(oracle_lib, c2, 8), run_host(reference_lib, c2, 8), "near-edge")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("degenerate", [False, True])
def test_pairwise_oracle_equals_reference_host_build(oracle_lib, reference_lib, dtype, degenerate):
    px, py, po, ro, vx, vy = pairwise_case(2000, dtype, 5, degenerate)
    a = oracle_lib.pairwise_point_in_polygon(px, py, po, ro, vx, vy)
    b = reference_lib.pairwise_point_in_polygon(px, py, po, ro, vx, vy)
    np.testing.assert_array_equal(a, b)
    assert 0 < a.sum() < len(a)


def test_pairwise_size_mismatch_is_an_error(oracle_lib):
    px, py, po, ro, vx, vy = pairwise_case(10, np.float64, 1)
    with pytest.raises(RuntimeError, match="same number of points as polygons"):
        oracle_lib.pairwise_point_in_polygon(px[:5], py[:5], po, ro, vx, vy)

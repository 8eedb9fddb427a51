"""CPU tests (no GPU): pin the oracle.

1. oracle/oracle.cpp against every golden vector harvested from the reference's own tests;
2. the reference's own header-only implementation compiled for the host (oracle/_ref) against the
   same vectors -- proves the harness drives the reference correctly;
3. oracle == reference host build, bit for bit, on randomised inputs (fp32 + fp64; uniform and
   clustered; out-of-bbox points; duplicates; shallow/deep trees; max_size 1).
"""
import numpy as np
import pytest

from util import (assert_same, covered_positions, make_case, make_linestrings, pairwise_case,
                  run_host, run_host_nearest)


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


def _golden_nearest(lib, golden, dt):
    n = golden["nearest_linestring"]
    p = np.array(n["points"], dtype=dt)
    v = np.array(n["vertices"], dtype=dt)
    b = n["bbox"]
    t = lib.quadtree_on_points(p[:, 0].copy(), p[:, 1].copy(), b[0], b[1], b[2], b[3], n["scale"],
                               n["max_depth"], n["max_size"])
    bb = lib.linestring_bounding_boxes(n["line_offsets"], v[:, 0].copy(), v[:, 1].copy(),
                                       n["expansion_radius"])
    pl, pq = lib.join_quadtree_and_bounding_boxes(t, *bb, b[0], b[2], n["scale"], n["max_depth"])
    assert list(pl) == n["pair_line"] and list(pq) == n["pair_quad"]
    return lib.quadtree_point_to_nearest_linestring(pl, pq, t, t["point_indices"], p[:, 0].copy(),
                                                    p[:, 1].copy(), n["line_offsets"],
                                                    v[:, 0].copy(), v[:, 1].copy())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_oracle_nearest_linestring_matches_golden_bit_for_bit(oracle_lib, golden, dt):
    """quadtree_point_to_nearest_linestring_test_small.cu: the expected distances were produced by
    the reference's CUDA build; the restatement (with nvcc's FMA contraction of dot()) reproduces
    every one of them exactly."""
    n = golden["nearest_linestring"]
    pi, li, d = _golden_nearest(oracle_lib, golden, dt)
    assert list(pi) == n["point_index"] and list(li) == n["linestring_index"]
    want = np.array(n["distance_f32" if dt == np.float32 else "distance_f64"], dtype=dt)
    np.testing.assert_array_equal(d, want)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_reference_host_build_nearest_linestring_matches_golden(reference_lib, golden, dt):
    """The host build contracts dot() the way gcc does, not the way nvcc does: indices must be
    exact, distances agree to rounding (d0 - r cancels, hence the loose relative bound)."""
    n = golden["nearest_linestring"]
    pi, li, d = _golden_nearest(reference_lib, golden, dt)
    assert list(pi) == n["point_index"] and list(li) == n["linestring_index"]
    want = np.array(n["distance_f32" if dt == np.float32 else "distance_f64"], dtype=dt)
    np.testing.assert_allclose(d, want, rtol=2e-5 if dt == np.float32 else 1e-12)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_oracle_nearest_linestring_equals_reference_host_build(oracle_lib, reference_lib, dt):
    """Every point covered (radius = the whole extent).  With partial coverage the reference
    scatters its reduced rows IN PLACE (quadtree_point_to_nearest_linestring.cuh:309-316: source
    and destination are the same arrays), which leaves stale rows behind and races; the port
    defines those rows as zeros, so parity is claimed for covered tables only."""
    c = make_case(20000, 5, 10, "u", dt, seed=31)
    lines = make_linestrings(40, c["ext"], 7, dt)
    radius = c["ext"][1] - c["ext"][0] + c["ext"][3] - c["ext"][2]
    a = run_host_nearest(oracle_lib, c, lines, 50, radius)
    b = run_host_nearest(reference_lib, c, lines, 50, radius)
    for i in range(2):
        np.testing.assert_array_equal(a["pairs"][i], b["pairs"][i])
    assert covered_positions(a["tree"], a["pairs"][1], 20000).all()
    np.testing.assert_array_equal(a["nearest"][0], b["nearest"][0])
    # gcc and nvcc contract dot() differently, and d0 - r cancels for points next to a segment:
    # distances agree to sqrt(eps * d0); a distance that rounds to exactly 0 on one side only is
    # skipped by the selection rule there (:288-291), so a handful of rows may pick another line
    close = np.isclose(a["nearest"][2], b["nearest"][2], rtol=1e-4 if dt == np.float32 else 1e-10,
                       atol=2e-4 if dt == np.float32 else 1e-8)
    assert (~close).mean() < 1e-3
    assert (a["nearest"][1] != b["nearest"][1]).mean() < 1e-3


def test_oracle_nearest_linestring_partial_coverage_is_zero_filled(oracle_lib):
    c = make_case(5000, 5, 8, "u", np.float64, seed=3)
    lines = make_linestrings(10, c["ext"], 9, np.float64)
    a = run_host_nearest(oracle_lib, c, lines, 20, 0.01 * (c["ext"][1] - c["ext"][0]))
    m = covered_positions(a["tree"], a["pairs"][1], 5000)
    assert 0 < m.sum() < 5000
    assert (a["nearest"][2][~m] == 0).all() and (a["nearest"][0][~m] == 0).all()
    np.testing.assert_array_equal(a["nearest"][0][m], np.nonzero(m)[0])
    assert (a["nearest"][2][m] > 0).all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_lattice_polygons_oracle_equals_reference(oracle_lib, reference_lib, dtype):
    """Adversarial exact-arithmetic territory: integer-lattice polygons (collinear runs, repeated
    vertices, self-intersections, unclosed rings, a hole touching the shell) against every lattice
    and half-lattice point -- most points sit exactly on an edge, a vertex or an edge's line."""
    rng = np.random.default_rng(2024)
    rings, po, ro = [], [0], [0]
    for p in range(31):
        k = int(rng.integers(3, 9))
        v = rng.integers(0, 9, size=(k, 2)).astype(np.float64)
        if p % 5 == 0:
            v = np.repeat(v, 2, axis=0)[: k + 2]          # repeated vertices (zero-length segments)
        if p % 3 == 0:
            v = np.vstack([v, v[:1]])                      # closed ring; the others stay unclosed
        rings.append(v)
        ro.append(ro[-1] + len(v))
        if p % 7 == 0:                                     # a hole sharing a vertex with the shell
            h = np.array([v[0], v[0] + [1, 0], v[0] + [0, 1], v[0]])
            rings.append(h)
            ro.append(ro[-1] + len(h))
        po.append(len(ro) - 1)
    vv = np.vstack(rings).astype(dtype)
    gx, gy = np.meshgrid(np.arange(-1, 19) / 2.0, np.arange(-1, 19) / 2.0)
    px, py = gx.reshape(-1).astype(dtype), gy.reshape(-1).astype(dtype)
    po, ro = np.array(po, np.int32), np.array(ro, np.int32)
    a = oracle_lib.point_in_polygon(px, py, po, ro, vv[:, 0].copy(), vv[:, 1].copy())
    b = reference_lib.point_in_polygon(px, py, po, ro, vv[:, 0].copy(), vv[:, 1].copy())
    np.testing.assert_array_equal(a, b)
    assert 0 < np.count_nonzero(a) < len(a)
    # the same polygons pairwise against a point each
    qx, qy = px[:31].copy(), py[200:231].copy()
    np.testing.assert_array_equal(
        oracle_lib.pairwise_point_in_polygon(qx, qy, po, ro, vv[:, 0].copy(), vv[:, 1].copy()),
        reference_lib.pairwise_point_in_polygon(qx, qy, po, ro, vv[:, 0].copy(), vv[:, 1].copy()))

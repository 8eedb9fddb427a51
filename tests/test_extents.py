"""Parity away from the unit square.

Every exactness shortcut of the CUDA path (whole-quadrant classification, cell-centre test, exact
edge skipping, the guarded Morton encode) is a MARGIN argument scaled by eps * |coordinate|.  The
round-1 fixtures all lived in [0,1]^2 or [0,64]^2; these cases put the same workloads where
|coordinate| >> extent (lon/lat, UTM), at negative coordinates, at a tiny extent around the origin
and at a tiny extent far from it, and in float32 where the margin exceeds the finest cell.

CPU part (no GPU): the oracle restatement equals the reference's own host build there, so the
oracle stays pinned in these regimes.  GPU part: the CUDA path equals the oracle (and the
reference host build) bit for bit, with the cell-geometry hint, without it and without the
sorted-key hint.  Reference predicate: cpp/include/cuspatial/detail/algorithm/
is_point_in_polygon.cuh:46-101, detail/utility/floating_point.cuh:118-130.
"""
import numpy as np
import pytest

from util import assert_same, in_contract, make_case, run_gpu, run_host

# (name, extent, dtypes)
EXTENTS = [
    ("lonlat", (-74.3, -73.6, 40.4, 41.0), (np.float32, np.float64)),
    ("utm", (5.8e5, 6.1e5, 4.49e6, 4.52e6), (np.float32, np.float64)),
    ("tiny_origin", (-1e-3, 1e-3, -1e-3, 1e-3), (np.float32, np.float64)),
    ("tiny_far", (1e6, 1e6 + 1e-3, 1e6, 1e6 + 1e-3), (np.float64,)),
    ("f32_offset_1024", (1024.0, 1025.0, 1024.0, 1025.0), (np.float32, np.float64)),
    ("negative_large", (-3.0e6, -2.9e6, -8.0e5, -7.2e5), (np.float32, np.float64)),
]
EXT_PARAMS = [pytest.param(e, dt, id="%s-%s" % (n, dt.__name__)) for n, e, dts in EXTENTS
              for dt in dts]

SHAPES = [
    # n, n_poly, depth, max_size, kind, oob, dups, median_vertices
    (60000, 40, 15, 64, "u", 0, 0, 60),
    (40000, 20, 9, 16, "c", 30, 200, 30),
    (20000, 31, 12, 1, "u", 5, 0, 24),
]


def near_edge_case(ext, dtype, depth=6):
    """Points on polygon edges and vertices and +-1, +-4, +-5 ULP off them, plus points sharing
    the x of vertical edges at arbitrary y, at `ext`; polygons include an axis-aligned square
    whose edges lie on cell boundaries of the depth-`depth` grid."""
    c = make_case(10, 12, depth, "u", dtype, seed=9, median_vertices=16, extent=ext)
    x0, x1, y0, y1 = c["ext"]
    cell = dtype(c["scale"])
    # a square on cell boundaries: corners at min + k * scale computed the reference's way
    k0, k1 = (1 << depth) // 4, (3 << depth) // 8
    ax, bx = dtype(x0) + dtype(k0) * cell, dtype(x0) + dtype(k1) * cell
    ay, by = dtype(y0) + dtype(k0) * cell, dtype(y0) + dtype(k1) * cell
    sq = np.array([[ax, ay], [bx, ay], [bx, by], [ax, by], [ax, ay]], dtype=dtype)
    vx = np.concatenate([c["vx"], sq[:, 0]])
    vy = np.concatenate([c["vy"], sq[:, 1]])
    ro = np.concatenate([c["ro"], [c["ro"][-1] + 5]]).astype(np.uint32)
    po = np.concatenate([c["po"], [c["po"][-1] + 1]]).astype(np.uint32)
    xs, ys = [], []
    for i in range(len(vx)):
        j = i + 1 if i + 1 < len(vx) else i
        for t in (0.0, 0.25, 0.5, 1.0):
            px = dtype(vx[i] + dtype(t) * (vx[j] - vx[i]))
            py = dtype(vy[i] + dtype(t) * (vy[j] - vy[i]))
            for k in (-5, -4, -1, 0, 1, 4, 5):
                xs.append(px); ys.append(py + dtype(k) * np.spacing(py))
                xs.append(px + dtype(k) * np.spacing(px)); ys.append(py)
    for yy in np.linspace(y0 + 0.05 * (y1 - y0), y1 - 0.05 * (y1 - y0), 40):
        xs += [ax, bx]; ys += [dtype(yy), dtype(yy)]
    x, y = np.array(xs, dtype=dtype), np.array(ys, dtype=dtype)
    keep = (x > x0) & (x < x1) & (y > y0) & (y < y1) & in_contract(x, y, c["ext"], c["scale"],
                                                                    depth, dtype)
    return dict(c, x=x[keep], y=y[keep], po=po, ro=ro, vx=vx, vy=vy)


def lattice_case(ext, dtype, depth):
    """Polygon edges exactly on quadtree cell boundaries at an offset, and lattice points that hit
    cell borders, vertices and vertical edges."""
    x0, y0 = dtype(ext[0]), dtype(ext[2])
    w = dtype(max(ext[1] - ext[0], ext[3] - ext[2]))
    u = w / dtype(64)
    scale = float(w) / (1 << depth)
    rings = [
        [(8, 8), (24, 8), (24, 24), (8, 24), (8, 8)],
        [(32, 4), (60, 4), (60, 30), (46, 30), (46, 18), (32, 18), (32, 4)],
        [(4, 40), (28, 40), (16, 60), (4, 40)],
        [(36, 36), (60, 36), (60, 60), (36, 60), (36, 36)],
        [(44, 44), (44, 52), (52, 52), (52, 44), (44, 44)],
    ]
    po = np.array([0, 1, 2, 3, 5], dtype=np.uint32)
    ro = np.cumsum([0] + [len(r) for r in rings]).astype(np.uint32)
    v = np.array([p for r in rings for p in r], dtype=dtype)
    vx, vy = x0 + v[:, 0] * u, y0 + v[:, 1] * u
    g = np.arange(0, 64, 0.5, dtype=dtype)
    gx, gy = np.meshgrid(g, g)
    rng = np.random.default_rng(5)
    x = np.concatenate([x0 + gx.ravel() * u, x0 + rng.uniform(0, 64, 20000).astype(dtype) * u])
    y = np.concatenate([y0 + gy.ravel() * u, y0 + rng.uniform(0, 64, 20000).astype(dtype) * u])
    e = (float(x0), float(x0 + w), float(y0), float(y0 + w))
    keep = in_contract(x, y, e, scale, depth, dtype)
    return dict(x=x[keep], y=y[keep], po=po, ro=ro, vx=vx.astype(dtype), vy=vy.astype(dtype),
                ext=e, scale=scale, depth=depth)


# ------------------------------------------------------------------------------------------------
# CPU: the oracle stays pinned on the reference's own host build in these regimes
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ext,dtype", EXT_PARAMS)
def test_oracle_equals_reference_host_build_at_extent(oracle_lib, reference_lib, ext, dtype):
    for n, n_poly, depth, max_size, kind, oob, dups, mv in SHAPES:
        c = make_case(n // 4, n_poly, depth, kind, dtype, seed=n + depth, oob=oob, dups=dups,
                      median_vertices=mv, extent=ext)
        assert_same(run_host(oracle_lib, c, max_size), run_host(reference_lib, c, max_size),
                    "oracle vs reference host build")
    c = near_edge_case(ext, dtype)
    for max_size in (8, 100000):
        assert_same(run_host(oracle_lib, c, max_size), run_host(reference_lib, c, max_size),
                    "near-edge: oracle vs reference host build")
    c = lattice_case(ext, dtype, 6)
    assert_same(run_host(oracle_lib, c, 4), run_host(reference_lib, c, 4),
                "lattice: oracle vs reference host build")


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
def _gpu_all_modes(c, max_size, want, what):
    assert_same(run_gpu(c, max_size), want, what)
    assert_same(run_gpu(c, max_size, use_grid_hint=False), want, what + " (no grid hint)")
    assert_same(run_gpu(c, max_size, use_grid_hint="no_keys"), want, what + " (no keys)")


@pytest.mark.gpu
@pytest.mark.parametrize("ext,dtype", EXT_PARAMS)
def test_random_inputs_at_extent(oracle_lib, ext, dtype):
    from oracle import hostlib

    for n, n_poly, depth, max_size, kind, oob, dups, mv in SHAPES:
        c = make_case(n, n_poly, depth, kind, dtype, seed=n + depth, oob=oob, dups=dups,
                      median_vertices=mv, extent=ext)
        want = run_host(oracle_lib, c, max_size)
        assert len(want["hits"][0]) > 0
        _gpu_all_modes(c, max_size, want, "gpu vs oracle")
        if hostlib.reference_available() and depth == 15:
            assert_same(run_gpu(c, max_size), run_host(hostlib.reference(), c, max_size),
                        "gpu vs reference host build")


@pytest.mark.gpu
@pytest.mark.parametrize("ext,dtype", EXT_PARAMS)
def test_near_edge_points_at_extent(oracle_lib, ext, dtype):
    c = near_edge_case(ext, dtype)
    for max_size in (8, 100000):
        want = run_host(oracle_lib, c, max_size)
        _gpu_all_modes(c, max_size, want, "near-edge")


@pytest.mark.gpu
@pytest.mark.parametrize("ext,dtype", EXT_PARAMS)
@pytest.mark.parametrize("depth,max_size", [(6, 4), (8, 1)])
def test_cell_aligned_polygons_at_extent(oracle_lib, ext, dtype, depth, max_size):
    c = lattice_case(ext, dtype, depth)
    want = run_host(oracle_lib, c, max_size)
    _gpu_all_modes(c, max_size, want, "lattice")


@pytest.mark.gpu
@pytest.mark.parametrize("grid_log2", ["0", "5", "10"])
@pytest.mark.parametrize("ext,dtype", EXT_PARAMS)
def test_bitmask_at_extent(oracle_lib, ext, dtype, grid_log2, monkeypatch):
    """grid_log2: side of the bitmask kernel's cell-class grid (0 = off; the library only turns
    it on by itself from 2^20 points)."""
    import torch

    monkeypatch.setenv("BSJ_BITMASK_GRID_LOG2", grid_log2)

    import cuspatial_b200 as cs

    c = near_edge_case(ext, dtype)
    c2 = make_case(100000, 31, 8, "u", dtype, seed=3, median_vertices=60, extent=ext)
    for cc in (c, c2):
        po, ro = cc["po"][:32].astype(np.int32), cc["ro"].astype(np.int32)
        want = oracle_lib.point_in_polygon(cc["x"], cc["y"], po, ro, cc["vx"], cc["vy"])
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")  # noqa: E731
        got = cs.point_in_polygon_bitmask((t(cc["x"]), t(cc["y"])),
                                          (t(po), t(ro), t(cc["vx"]), t(cc["vy"])))
        np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_long_edges_reaching_far_outside_the_area_of_interest(oracle_lib, dtype):
    """Polygons whose vertices lie up to 1000 extents outside the area of interest cross it with
    long, nearly straight edges: the products of the reference's edge function are then huge
    compared with the cell size, which is where the float32 allowance of the cell-centre test
    (relative to |u|, not to the cell) has to hold.  Points cluster along those edges."""
    rng = np.random.default_rng(17)
    c = make_case(30000, 6, 10, "u", dtype, seed=23, median_vertices=12)
    x0, x1, y0, y1 = c["ext"]
    w, h = x1 - x0, y1 - y0
    polys_x, polys_y, ro, po = [c["vx"]], [c["vy"]], list(c["ro"]), list(c["po"])
    for k in range(6):  # thin slivers and huge triangles through the area of interest
        far = 10.0 ** rng.uniform(0.5, 3.0)
        ang = rng.uniform(0, np.pi)
        cx, cy = x0 + rng.uniform(0.2, 0.8) * w, y0 + rng.uniform(0.2, 0.8) * h
        dx, dy = np.cos(ang) * far * w, np.sin(ang) * far * h
        nx, ny = -np.sin(ang) * rng.uniform(0.01, 0.3) * w, np.cos(ang) * rng.uniform(0.01, 0.3) * h
        ring = np.array([[cx - dx, cy - dy], [cx + dx, cy + dy], [cx + nx, cy + ny],
                         [cx - dx, cy - dy]], dtype=dtype)
        polys_x.append(ring[:, 0]); polys_y.append(ring[:, 1])
        ro.append(ro[-1] + 4); po.append(po[-1] + 1)
    vx, vy = np.concatenate(polys_x).astype(dtype), np.concatenate(polys_y).astype(dtype)
    # points on and next to the long edges, inside the area of interest
    xs, ys = [c["x"]], [c["y"]]
    for r in range(len(c["ro"]) - 1, len(ro) - 1):
        a = ro[r]
        for e in range(3):
            t = rng.uniform(0.0, 1.0, 4000)
            px = vx[a + e] + t * (vx[a + e + 1] - vx[a + e])
            py = vy[a + e] + t * (vy[a + e + 1] - vy[a + e])
            keep = (px > x0) & (px < x1) & (py > y0) & (py < y1)
            px, py = px[keep].astype(dtype), py[keep].astype(dtype)
            for ulps in (0, 1, -1, 4, -5, 50):
                xs.append(px); ys.append(np.nextafter(py, dtype(np.inf) if ulps >= 0 else dtype(-np.inf))
                                         if abs(ulps) == 1 else py + dtype(ulps) * np.spacing(py))
    x, y = np.concatenate(xs).astype(dtype), np.concatenate(ys).astype(dtype)
    keep = in_contract(x, y, c["ext"], c["scale"], c["depth"], dtype)
    c2 = dict(c, x=x[keep], y=y[keep], vx=vx, vy=vy, ro=np.asarray(ro, np.uint32),
              po=np.asarray(po, np.uint32))
    for max_size in (16, 100000):
        want = run_host(oracle_lib, c2, max_size)
        assert len(want["hits"][0]) > 0
        _gpu_all_modes(c2, max_size, want, "long edges")

"""GPU parity tests (B200): the CUDA path, called through the Python API -> C ABI, must equal the
CPU oracle bit for bit -- quadtree arrays, point_indices, the ordered (bbox, quad) pair table and
the ordered (polygon_index, point_index) rows -- on the reference's golden vectors, on randomised
inputs at sizes the oracle finishes in seconds, and (size-independent properties) at full size.
The reference's own host build (oracle/_ref) is used as a second checker when present.
"""
import numpy as np
import pytest

from util import (TREE_COLS, assert_same, assert_same_nearest, covered_positions, make_case,
                  make_linestrings, run_gpu, run_gpu_nearest, run_host, run_host_nearest)

pytestmark = pytest.mark.gpu


def _t(a, dev="cuda"):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def test_extension_is_loaded_and_launches_kernels():
    from cuspatial_b200 import _lib

    before = _lib.kernel_launch_count()
    c = make_case(1000, 5, 5, "u", np.float64, seed=1)
    run_gpu(c, 16)
    assert _lib.kernel_launch_count() > before


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_quadtrees(golden, dtype):
    import cuspatial_b200 as cs

    for c in golden["quadtree_cases"]:
        pts = np.array(c["points"], dtype=dtype).reshape(-1, 2)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pidx, tree = cs.quadtree_on_points((_t(pts[:, 0]), _t(pts[:, 1])), c["v_min"][0],
                                               c["v_max"][0], c["v_min"][1], c["v_max"][1],
                                               c["scale"], c["max_depth"], c["max_size"])
        for k in TREE_COLS:
            assert tree[k].cpu().numpy().astype(np.int64).tolist() == c[k], (c["name"], k)
        assert pidx.shape[0] == len(pts)
        # reference dtypes (python/cuspatial/cuspatial/tests/spatial/indexing/test_indexing.py:27-31)
        import torch

        assert tree["key"].dtype == torch.uint32 and tree["level"].dtype == torch.uint8
        assert tree["is_internal_node"].dtype == torch.bool
        assert tree["length"].dtype == torch.uint32 and tree["offset"].dtype == torch.uint32
        assert pidx.dtype == torch.uint32


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_small_join_and_pip(golden, dtype):
    import cuspatial_b200 as cs

    sj = golden["small_join"]
    pts = np.array(sj["points"], dtype=dtype)
    x, y = _t(pts[:, 0]), _t(pts[:, 1])
    pidx, tree = cs.quadtree_on_points((x, y), 0, 8, 0, 8, sj["scale"], sj["max_depth"],
                                       sj["max_size"])
    v = np.array(sj["vertices"], dtype=dtype)
    polys = (_t(np.array(sj["part_offsets"], np.uint32)), _t(np.array(sj["ring_offsets"], np.uint32)),
             _t(v[:, 0]), _t(v[:, 1]))
    bb = cs.polygon_bounding_boxes(polys)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, 0, 8, 0, 8, sj["scale"], sj["max_depth"])
    assert pairs["bbox_offset"].cpu().numpy().tolist() == sj["pair_poly"]
    assert pairs["quad_offset"].cpu().numpy().tolist() == sj["pair_quad"]
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polys)
    assert hits["polygon_index"].cpu().numpy().tolist() == sj["pip_poly"]
    assert hits["point_index"].cpu().numpy().tolist() == sj["pip_point"]
    # linestring bounding boxes expanded by 2.0: 21 pairs incl. non-bottom leaves and ties
    lj = golden["linestring_join"]
    lb = cs.polygon_bounding_boxes(polys, lj["expansion_radius"])
    pairs = cs.join_quadtree_and_bounding_boxes(tree, lb, 0, 8, 0, 8, sj["scale"], sj["max_depth"])
    assert pairs["bbox_offset"].cpu().numpy().tolist() == lj["bbox_offset"]
    assert pairs["quad_offset"].cpu().numpy().tolist() == lj["quad_offset"]


@pytest.mark.parametrize("grid_log2", ["0", "3", "8"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_bitmask_predicate_cases(golden, dtype, grid_log2, monkeypatch):
    monkeypatch.setenv("BSJ_BITMASK_GRID_LOG2", grid_log2)
    import cuspatial_b200 as cs

    for c in golden["pip_cases"]:
        p = np.array(c["points"], dtype=dtype)
        v = np.array(c["vertices"], dtype=dtype)
        polys = (_t(np.array(c["part_offsets"], np.int32)), _t(np.array(c["ring_offsets"], np.int32)),
                 _t(v[:, 0]), _t(v[:, 1]))
        m = cs.point_in_polygon_bitmask((_t(p[:, 0]), _t(p[:, 1])), polys)
        assert m.cpu().numpy().tolist() == c["expected_mask"], c["name"]
        frame = cs.point_in_polygon((_t(p[:, 0]), _t(p[:, 1])), polys)
        assert len(frame.columns) == len(c["part_offsets"]) - 1


CASES = [
    # n, n_poly, depth, max_size, kind, oob, dups, median_vertices
    (20000, 30, 15, 64, "u", 0, 0, 40),
    (200000, 263, 15, 512, "u", 0, 0, 200),
    (300000, 50, 8, 20, "c", 50, 1000, 60),
    (5000, 10, 3, 5, "u", 0, 0, 12),
    (100000, 20, 15, 1, "c", 10, 300, 30),
    (1000, 5, 1, 10, "u", 0, 0, 20),
    (3000, 7, 2, 1, "u", 3, 0, 20),
    (400000, 100, 12, 100, "c", 0, 0, 100),
    (1, 3, 15, 1, "u", 0, 0, 10),
    (2, 3, 4, 1, "u", 1, 0, 10),
    (100000, 40, 15, 100000, "u", 0, 0, 50),   # a single huge leaf per top cell: many tiles per run
    (3000, 300, 15, 512, "u", 5, 50, 60),      # large sparse quadrants under many polygons: tiles
    (60000, 2000, 15, 256, "c", 20, 300, 40),  # span whole polygons (point-by-point evaluation)
]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("case", CASES)
def test_random_inputs_equal_oracle(oracle_lib, case, dtype):
    n, n_poly, depth, max_size, kind, oob, dups, mv = case
    c = make_case(n, n_poly, depth, kind, dtype, seed=n + depth, oob=oob, dups=dups,
                  median_vertices=mv)
    want = run_host(oracle_lib, c, max_size)
    assert_same(run_gpu(c, max_size), want, "gpu vs oracle")
    assert_same(run_gpu(c, max_size, use_grid_hint=False), want, "gpu (no grid hint) vs oracle")
    assert_same(run_gpu(c, max_size, use_grid_hint="no_keys"), want, "gpu (no keys) vs oracle")


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_small_configurations_equal_oracle(oracle_lib, seed):
    """Randomised shapes (point count, polygon count and size, depth, max_size, distribution,
    dtype, out-of-bbox points, duplicates): small inputs walk through the code paths that big
    benchmarks never touch (few tiles, tiny levels, warp-kernel tree levels, empty quadrants)."""
    rng = np.random.default_rng(1000 + seed)
    for it in range(8):
        n = int(rng.choice([1, 2, 33, 257, 1000, 4097, 20000, 60000]))
        n_poly = int(rng.choice([1, 2, 7, 31, 64]))
        depth = int(rng.integers(1, 16))
        max_size = int(rng.choice([1, 2, 5, 32, 64, 200, 512]))
        dtype = [np.float32, np.float64][int(rng.integers(0, 2))]
        kind = "uc"[int(rng.integers(0, 2))]
        oob = int(rng.integers(0, min(n, 20) + 1)) if rng.random() < 0.5 else 0
        dups = int(rng.integers(0, min(n - oob, 300) + 1)) if rng.random() < 0.5 else 0
        mv = int(rng.choice([4, 12, 40, 150]))
        c = make_case(n, n_poly, depth, kind, dtype, seed=int(rng.integers(1, 10**6)),
                      median_vertices=mv, oob=oob, dups=dups)
        want = run_host(oracle_lib, c, max_size)
        tag = "fuzz seed=%d it=%d n=%d polys=%d depth=%d max_size=%d %s %s oob=%d dups=%d mv=%d" % (
            seed, it, n, n_poly, depth, max_size, dtype.__name__, kind, oob, dups, mv)
        assert_same(run_gpu(c, max_size), want, tag)
        if it % 2 == 0:
            assert_same(run_gpu(c, max_size, use_grid_hint=False), want, tag + " (no hint)")


@pytest.mark.parametrize("n_poly,n,dtype,kind", [(10_000, 1_000_000, np.float64, "c"),
                                                 (50_000, 500_000, np.float32, "u")])
def test_many_polygons_equal_oracle(oracle_lib, n_poly, n, dtype, kind):
    """The polygon counts of BASELINE.json configs[3] and configs[4] (10 k and 50 k polygons):
    per-polygon index build, seeded traversal queue and pair bookkeeping at that scale."""
    c = make_case(n, n_poly, 15, kind, dtype, seed=123, median_vertices=24, oob=50)
    want = run_host(oracle_lib, c, 128)
    assert len(want["pairs"][0]) > n_poly
    assert_same(run_gpu(c, 128), want, "gpu vs oracle (%d polygons)" % n_poly)


def test_config1_1M_uniform_263_polygons_equals_oracle_and_reference(oracle_lib):
    """BASELINE.json configs[0]: 1M uniform fp64 points x 263 taxi-zone-like polygons."""
    from oracle import hostlib

    c = make_case(1_000_000, 263, 15, "u", np.float64, seed=20251017, median_vertices=200)
    g = run_gpu(c, 512)
    assert_same(g, run_host(oracle_lib, c, 512), "gpu vs oracle (config 1)")
    if hostlib.reference_available():
        assert_same(g, run_host(hostlib.reference(), c, 512), "gpu vs reference host build")
    assert len(g["hits"][0]) > 500_000


@pytest.mark.parametrize("dtype,n,kind", [(np.float64, 5_000_000, "u"), (np.float32, 5_000_000, "u"),
                                          (np.float64, 3_000_000, "c")])
def test_equals_reference_cuda_build(dtype, n, kind):
    """GPU vs GPU: the reference's own Thrust/CUB implementation (real nvcc FMA contraction,
    real device float->u16 conversion) compiled in place from its headers, on the same B200.
    Sizes the CPU checkers cannot reach in seconds; duplicates and out-of-bbox points included."""
    from oracle import cudalib
    from util import run_ref_cuda

    if not cudalib.available():
        pytest.skip("oracle/_ref/libcuspatial_ref_cuda.so not built (needs /root/reference)")
    c = make_case(n, 263, 15, kind, dtype, seed=77, median_vertices=120, oob=1000, dups=5000)
    assert_same(run_gpu(c, 512), run_ref_cuda(c, 512), "gpu vs reference CUDA build")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_near_edge_points_equal_oracle(oracle_lib, dtype):
    """The 4-ULP on-edge rule, vertical-edge quirk and exact edge skipping near boundaries."""
    c = make_case(10, 12, 6, "u", dtype, seed=9, median_vertices=16)
    vx, vy, ro = c["vx"], c["vy"], c["ro"]
    sq = np.array([[0.3, 0.3], [0.6, 0.3], [0.6, 0.6], [0.3, 0.6], [0.3, 0.3]], dtype=dtype)
    vx = np.concatenate([vx, sq[:, 0]]); vy = np.concatenate([vy, sq[:, 1]])
    ro = np.concatenate([ro, [ro[-1] + 5]]).astype(np.uint32)
    po = np.concatenate([c["po"], [c["po"][-1] + 1]]).astype(np.uint32)
    xs, ys = [], []
    for i in range(len(vx)):
        j = i + 1 if i + 1 < len(vx) else i
        for t in (0.0, 0.25, 0.5, 1.0):
            px, py = vx[i] + t * (vx[j] - vx[i]), vy[i] + t * (vy[j] - vy[i])
            for k in range(-6, 7, 3):
                xs.append(px); ys.append(py + dtype(k) * np.spacing(py))
                xs.append(px + dtype(k) * np.spacing(px)); ys.append(py)
    # plus points exactly on the vertical edges' x at arbitrary y (rejected regardless of y)
    for yy in np.linspace(0.05, 0.95, 40):
        xs += [0.3, 0.6]; ys += [yy, yy]
    x = np.array(xs, dtype=dtype); y = np.array(ys, dtype=dtype)
    keep = (x > c["ext"][0]) & (x < c["ext"][1]) & (y > c["ext"][2]) & (y < c["ext"][3])
    c2 = dict(c, x=x[keep], y=y[keep], po=po, ro=ro, vx=vx, vy=vy)
    for max_size in (8, 100000):
        want = run_host(oracle_lib, c2, max_size)
        assert_same(run_gpu(c2, max_size), want, "near-edge")
        assert_same(run_gpu(c2, max_size, use_grid_hint=False), want, "near-edge (no hint)")
        assert_same(run_gpu(c2, max_size, use_grid_hint="no_keys"), want, "near-edge (no keys)")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("depth,max_size", [(6, 4), (4, 50), (8, 1)])
def test_grid_aligned_polygons_and_lattice_points(oracle_lib, dtype, depth, max_size):
    """Polygon edges lying exactly on quadtree cell boundaries and points on a lattice that hits
    cell borders, polygon vertices and vertical edges: the whole-quadrant shortcut must classify
    every such quadrant as 'boundary' (vertical-edge rule, on-edge rule)."""
    ext = (0.0, 64.0, 0.0, 64.0)
    scale = 64.0 / (1 << depth)
    rings = [
        [(8, 8), (24, 8), (24, 24), (8, 24), (8, 8)],                       # cell-aligned square
        [(32, 4), (60, 4), (60, 30), (46, 30), (46, 18), (32, 18), (32, 4)],  # L-shape
        [(4, 40), (28, 40), (16, 60), (4, 40)],                              # triangle
        [(36, 36), (60, 36), (60, 60), (36, 60), (36, 36)],                  # square with a hole
        [(44, 44), (44, 52), (52, 52), (52, 44), (44, 44)],
    ]
    po = np.array([0, 1, 2, 3, 5], dtype=np.uint32)
    ro = np.cumsum([0] + [len(r) for r in rings]).astype(np.uint32)
    v = np.array([p for r in rings for p in r], dtype=dtype)
    g = np.arange(0, 64, 0.5, dtype=dtype)
    gx, gy = np.meshgrid(g, g)
    rng = np.random.default_rng(5)
    x = np.concatenate([gx.ravel(), rng.uniform(0, 64, 20000).astype(dtype)])
    y = np.concatenate([gy.ravel(), rng.uniform(0, 64, 20000).astype(dtype)])
    c = dict(x=x, y=y, po=po, ro=ro, vx=v[:, 0].copy(), vy=v[:, 1].copy(), ext=ext, scale=scale,
             depth=depth)
    want = run_host(oracle_lib, c, max_size)
    assert_same(run_gpu(c, max_size), want, "lattice")
    assert_same(run_gpu(c, max_size, use_grid_hint=False), want, "lattice (no hint)")
    assert_same(run_gpu(c, max_size, use_grid_hint="no_keys"), want, "lattice (no keys)")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_long_edges_reaching_far_outside_the_area(oracle_lib, dtype):
    """Polygons whose edges start hundreds of extents away: the products of the crossing test are
    large there, so the 4-ULP on-edge band of the reference is wide in absolute terms -- the
    cell-centre shortcut has to leave those points to the exact arithmetic (float32 above all)."""
    c = make_case(150000, 6, 15, "u", dtype, seed=21, median_vertices=12)
    x0, x1, y0, y1 = c["ext"]
    w, h = x1 - x0, y1 - y0
    far = 400.0
    rings = [
        [(x0 - far * w, y0 + 0.31 * h), (x1 + far * w, y0 + 0.52 * h), (x0 + 0.4 * w, y1 + far * h),
         (x0 - far * w, y0 + 0.31 * h)],
        [(x0 + 0.13 * w, y0 - far * h), (x0 + 0.77 * w, y1 + far * h), (x1 + far * w, y0 - far * h),
         (x0 + 0.13 * w, y0 - far * h)],
        [(x0 - far * w, y0 - far * h), (x1 + far * w, y1 + far * h), (x1 + far * w, y0 - far * h),
         (x0 - far * w, y0 - far * h)],
    ]
    vx = np.concatenate([c["vx"]] + [np.array([p[0] for p in r], dtype=dtype) for r in rings])
    vy = np.concatenate([c["vy"]] + [np.array([p[1] for p in r], dtype=dtype) for r in rings])
    ro = np.concatenate([c["ro"], c["ro"][-1] + np.cumsum([len(r) for r in rings])]).astype(np.uint32)
    po = np.concatenate([c["po"], c["po"][-1] + 1 + np.arange(len(rings))]).astype(np.uint32)
    c2 = dict(c, po=po, ro=ro, vx=vx, vy=vy)
    for max_size in (64, 4096):
        want = run_host(oracle_lib, c2, max_size)
        assert len(want["hits"][0]) > 1000
        assert_same(run_gpu(c2, max_size), want, "long edges")
        assert_same(run_gpu(c2, max_size, use_grid_hint=False), want, "long edges (no hint)")
        assert_same(run_gpu(c2, max_size, use_grid_hint="no_keys"), want, "long edges (no keys)")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_special_values_fall_back_to_reference_loop(oracle_lib, dtype):
    """NaN / Inf / denormal coordinates: the exact-skipping path must not be taken."""
    c = make_case(5000, 8, 6, "u", dtype, seed=4, median_vertices=16)
    x, y = c["x"].copy(), c["y"].copy()
    x[:20] = np.nan; y[20:40] = np.nan; x[40:50] = np.inf; y[50:60] = -np.inf
    x[60:70] = dtype(1e-310) if dtype == np.float64 else dtype(1e-42)
    c2 = dict(c, x=x, y=y)
    assert_same(run_gpu(c2, 64), run_host(oracle_lib, c2, 64), "special values")


@pytest.mark.parametrize("grid_log2", ["0", "4", "9", "11"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bitmask_equals_oracle(oracle_lib, dtype, grid_log2, monkeypatch):
    """Uniform points, points far outside the polygons' box, NaN / Inf / denormal coordinates;
    with the cell-class grid off and at three resolutions (0 = off)."""
    import cuspatial_b200 as cs

    monkeypatch.setenv("BSJ_BITMASK_GRID_LOG2", grid_log2)
    c = make_case(300000, 31, 8, "u", dtype, seed=3, median_vertices=60)
    x, y = c["x"].copy(), c["y"].copy()
    x[:20] = np.nan; y[20:40] = np.nan; x[40:50] = np.inf; y[50:60] = -np.inf
    x[60:70] = dtype(1e-310) if dtype == np.float64 else dtype(1e-42)
    x[100:200] += dtype(5.0); y[200:300] -= dtype(7.0); x[300:400] = -x[300:400]
    c = dict(c, x=x, y=y)
    po, ro = c["po"].astype(np.int32), c["ro"].astype(np.int32)
    want = oracle_lib.point_in_polygon(c["x"], c["y"], po, ro, c["vx"], c["vy"])
    got = cs.point_in_polygon_bitmask((_t(c["x"]), _t(c["y"])),
                                      (_t(po), _t(ro), _t(c["vx"]), _t(c["vy"])))
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_hint_is_bound_to_the_buffers_it_describes(oracle_lib, dtype):
    """The cell-geometry hint rides on the quadtree Frame but describes ONE set of point buffers.
    Passing other points of the same length (a permuted copy, or the same tensors modified in
    place) or another point_indices must give what the reference gives for THOSE arguments --
    the rows of the hint-free path -- not rows decided from the stale keys."""
    import torch

    import cuspatial_b200 as cs

    c = make_case(50000, 30, 10, "u", dtype, seed=21, median_vertices=40)
    ext = c["ext"]
    x, y = _t(c["x"]), _t(c["y"])
    polys = tuple(_t(a) for a in (c["po"], c["ro"], c["vx"], c["vy"]))
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], c["scale"],
                                       c["depth"], 32)
    bb = cs.polygon_bounding_boxes(polys)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3],
                                                c["scale"], c["depth"])

    def rows(points, indices, drop_hint=False):
        t = tree
        if drop_hint:
            t = cs.Frame([(k, tree[k]) for k in tree.columns])
        h = cs.quadtree_point_in_polygon(pairs, t, indices, points, polys)
        return h["polygon_index"].cpu().numpy(), h["point_index"].cpu().numpy()

    base = rows((x, y), pidx)
    np.testing.assert_array_equal(base[0], rows((x, y), pidx, drop_hint=True)[0])
    perm = torch.randperm(x.shape[0], device=x.device)
    x2, y2 = x[perm].contiguous(), y[perm].contiguous()
    for got, want in ((rows((x2, y2), pidx), rows((x2, y2), pidx, drop_hint=True)),
                      (rows((x, y), pidx.flip(0).contiguous()),
                       rows((x, y), pidx.flip(0).contiguous(), drop_hint=True))):
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
    assert not (len(base[0]) == len(rows((x2, y2), pidx)[0]) and
                np.array_equal(base[1], rows((x2, y2), pidx)[1])), "permutation changed nothing?"
    # in-place modification of the very tensors the tree was built from
    x.copy_(x2)
    y.copy_(y2)
    got, want = rows((x, y), pidx), rows((x2, y2), pidx, drop_hint=True)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])


@pytest.mark.parametrize("grid_log2", ["0", "6", "9"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_bitmask_concentric_ngons_overflow_the_exact_test_queue(oracle_lib, dtype, grid_log2,
                                                                monkeypatch):
    """The reference benchmark's shape (31 regular n-gons sharing one centroid,
    cpp/benchmarks/point_in_polygon/point_in_polygon.cu:41-102) with the points concentrated on
    the common outline: every point there is undecided for all 31 polygons, which overflows the
    kernel's shared-memory queue of exact tests (its evaluate-on-the-spot path) in every chunk."""
    import cuspatial_b200 as cs
    from cuspatial_b200 import datagen as D

    monkeypatch.setenv("BSJ_BITMASK_GRID_LOG2", grid_log2)
    po, ro, vx, vy = D.regular_ngons(31, 10, 10.0, centroid=(3.0, -2.0), dtype=dtype)
    rng = np.random.default_rng(5)
    n = 120000
    th = rng.uniform(0, 2 * np.pi, n)
    r = 10.0 * np.cos(np.pi / 10) / np.cos((th % (2 * np.pi / 10)) - np.pi / 10)  # on the outline
    r = r + rng.choice([0.0, 1e-9, -1e-9, 1e-4, -1e-4, 0.3, -0.3, 5.0], n)
    x = (3.0 + r * np.cos(th)).astype(dtype)
    y = (-2.0 + r * np.sin(th)).astype(dtype)
    x[:64] = vx[:64]; y[:64] = vy[:64]                          # the vertices themselves
    want = oracle_lib.point_in_polygon(x, y, po.astype(np.int32), ro.astype(np.int32), vx, vy)
    got = cs.point_in_polygon_bitmask((_t(x), _t(y)), (_t(po.astype(np.int32)),
                                                        _t(ro.astype(np.int32)), _t(vx), _t(vy)))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert 0 < int((want != 0).sum()) < n


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_pairwise_cases(golden, dtype):
    """pairwise_point_in_polygon_test.cu known answers (point i vs polygon i)."""
    import cuspatial_b200 as cs

    for c in golden["pairwise_cases"]:
        v = np.array(c["vertices"], dtype=dtype)
        for call in c["calls"]:
            p = np.array(call["points"], dtype=dtype)
            k = len(p)
            got = cs.pairwise_point_in_polygon(
                (_t(p[:, 0]), _t(p[:, 1])),
                (_t(np.array(c["part_offsets"][:k + 1], np.int32)),
                 _t(np.array(c["ring_offsets"], np.int32)), _t(v[:, 0]), _t(v[:, 1])))
            assert got.dtype.is_floating_point is False and got.element_size() == 1
            assert got.cpu().numpy().tolist() == call["expected"], (c["name"], dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("degenerate", [False, True])
def test_pairwise_equals_oracle(oracle_lib, dtype, degenerate):
    """Random pairs incl. points on vertices / edges / a few ulps off, and rings with
    zero-length segments (the reference keeps `b` over those: lane-parallel form must fall back)."""
    import cuspatial_b200 as cs
    from util import pairwise_case

    px, py, po, ro, vx, vy = pairwise_case(20000, dtype, 13, degenerate)
    want = oracle_lib.pairwise_point_in_polygon(px, py, po, ro, vx, vy)
    got = cs.pairwise_point_in_polygon((_t(px), _t(py)), (_t(po), _t(ro), _t(vx), _t(vy)))
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    pairs = cs.contains_properly((_t(po), _t(ro), _t(vx), _t(vy)), (_t(px), _t(py)), mode="pairwise")
    np.testing.assert_array_equal(pairs["point_index"].cpu().numpy(), np.nonzero(want)[0])
    with pytest.raises(RuntimeError, match="same number of points as polygons"):
        cs.pairwise_point_in_polygon((_t(px[:5]), _t(py[:5])), (_t(po), _t(ro), _t(vx), _t(vy)))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_contains_properly_quadtree_mode_equals_oracle_composition(oracle_lib, dtype):
    """contains.py:19-74: depth 15, max_size ceil(sqrt(N)), minimum scale, vertex extent; rows
    carry ORIGINAL point ids.  Also the <=31-polygon brute-force mode gives the same set."""
    import math

    import cuspatial_b200 as cs
    from cuspatial_b200 import datagen as D

    n = 30000
    po, ro, vx, vy = D.taxi_zone_like_polygons(25, seed=3, dtype=dtype, median_vertices=40)
    ext = (float(vx.min()), float(vx.max()), float(vy.min()), float(vy.max()))
    x, y = D.uniform_points(n, ext, seed=4, dtype=dtype)
    scale = max(ext[1] - ext[0], ext[3] - ext[2]) / ((1 << 15) + 2)
    ms = math.ceil(math.sqrt(n))
    tree = oracle_lib.quadtree_on_points(x, y, ext[0], ext[1], ext[2], ext[3], scale, 15, ms)
    bb = oracle_lib.polygon_bounding_boxes(po, ro, vx, vy)
    pp, pq = oracle_lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], scale, 15)
    hp, hq = oracle_lib.quadtree_point_in_polygon(pp, pq, tree, tree["point_indices"], x, y, po, ro,
                                                  vx, vy)
    polys = (_t(po), _t(ro), _t(vx), _t(vy))
    got = cs.contains_properly(polys, (_t(x), _t(y)), mode="quadtree")
    assert got.columns == ["point_index", "part_index"]
    np.testing.assert_array_equal(got["part_index"].cpu().numpy(), hp)
    np.testing.assert_array_equal(got["point_index"].cpu().numpy(), tree["point_indices"][hq])
    brute = cs.contains_properly(polys, (_t(x), _t(y)), mode="byte")
    a = set(zip(got["point_index"].cpu().numpy().tolist(), got["part_index"].cpu().numpy().tolist()))
    b = set(zip(brute["point_index"].cpu().numpy().tolist(), brute["part_index"].cpu().numpy().tolist()))
    assert a == b and len(a) > 1000


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_golden_nearest_linestring(golden, dtype):
    """quadtree_point_to_nearest_linestring_test_small.cu: pairs, indices and every distance bit
    for bit (the expected values come from the reference's CUDA build)."""
    import cuspatial_b200 as cs

    n = golden["nearest_linestring"]
    p = np.array(n["points"], dtype=dtype)
    v = np.array(n["vertices"], dtype=dtype)
    b = n["bbox"]
    pidx, tree = cs.quadtree_on_points((_t(p[:, 0]), _t(p[:, 1])), b[0], b[1], b[2], b[3],
                                       n["scale"], n["max_depth"], n["max_size"])
    ls = (_t(np.array(n["line_offsets"], np.uint32)), _t(v[:, 0]), _t(v[:, 1]))
    bb = cs.linestring_bounding_boxes(ls, n["expansion_radius"])
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, b[0], b[1], b[2], b[3], n["scale"],
                                                n["max_depth"])
    assert pairs["bbox_offset"].cpu().numpy().tolist() == n["pair_line"]
    assert pairs["quad_offset"].cpu().numpy().tolist() == n["pair_quad"]
    out = cs.quadtree_point_to_nearest_linestring(pairs, tree, pidx, (_t(p[:, 0]), _t(p[:, 1])), ls)
    assert out["point_index"].cpu().numpy().tolist() == n["point_index"]
    assert out["linestring_index"].cpu().numpy().tolist() == n["linestring_index"]
    want = np.array(n["distance_f32" if dtype == np.float32 else "distance_f64"], dtype=dtype)
    np.testing.assert_array_equal(out["distance"].cpu().numpy(), want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("full_cover", [True, False])
def test_nearest_linestring_equals_oracle(oracle_lib, dtype, full_cover):
    c = make_case(30000, 5, 12, "c" if full_cover else "u", dtype, seed=41, dups=50)
    lines = make_linestrings(60, c["ext"], 17, dtype)
    w = c["ext"][1] - c["ext"][0] + c["ext"][3] - c["ext"][2]
    radius = w if full_cover else 0.01 * w
    got = run_gpu_nearest(c, lines, 40, radius)
    want = run_host_nearest(oracle_lib, c, lines, 40, radius)
    m = covered_positions(want["tree"], want["pairs"][1], 30000)
    assert m.all() if full_cover else 0 < m.sum() < 30000
    assert_same_nearest(got, want, "gpu vs oracle (nearest linestring)")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nearest_linestring_equals_reference_cuda_build(dtype):
    """GPU vs GPU, every output bit, on a fully covered table (see the coverage note in
    test_oracle.py: the reference scatters in place, so partial tables are undefined there)."""
    from oracle import cudalib
    from util import run_ref_cuda_nearest

    if not cudalib.available():
        pytest.skip("oracle/_ref/libcuspatial_ref_cuda.so not built (needs /root/reference)")
    c = make_case(200_000, 5, 15, "u", dtype, seed=43)
    lines = make_linestrings(30, c["ext"], 19, dtype, median_vertices=30)
    w = c["ext"][1] - c["ext"][0] + c["ext"][3] - c["ext"][2]
    a, b = run_gpu_nearest(c, lines, 256, w), run_ref_cuda_nearest(c, lines, 256, w)
    # A point within rounding of a segment can get d0 - r < 0, i.e. a NaN distance, for that
    # linestring (point_linestring_distance.cuh:46-52).  The reference's selection rule is not
    # associative once a NaN is involved, so ITS answer for such a point depends on the shape of
    # CUB's reduction tree; the port folds the candidates in order (what a sequential
    # reduce_by_key gives).  Rows may differ only there.
    bad = np.nonzero((a["nearest"][1] != b["nearest"][1]) |
                     (a["nearest"][2].view(np.uint32 if dtype == np.float32 else np.uint64) !=
                      b["nearest"][2].view(np.uint32 if dtype == np.float32 else np.uint64)))[0]
    assert len(bad) <= 20
    from oracle import hostlib
    lo, lx, ly = lines
    one = {"key": np.zeros(1, np.uint32), "level": np.zeros(1, np.uint8),
           "is_internal_node": np.zeros(1, np.uint8), "length": np.ones(1, np.uint32),
           "offset": np.zeros(1, np.uint32)}
    for pos in bad:
        pid = a["tree"]["point_indices"][pos]
        ds = [hostlib.oracle().quadtree_point_to_nearest_linestring(
            np.array([k], np.uint32), np.zeros(1, np.uint32), one, np.zeros(1, np.uint32),
            c["x"][pid:pid + 1], c["y"][pid:pid + 1], lo, lx, ly)[2][0] for k in range(len(lo) - 1)]
        assert np.isnan(ds).any(), (pos, ds)
        a["nearest"][1][pos], a["nearest"][2][pos] = b["nearest"][1][pos], b["nearest"][2][pos]
    assert_same_nearest(a, b, "gpu vs reference CUDA build (nearest linestring)")


def test_nearest_linestring_empty_and_errors():
    import torch

    import cuspatial_b200 as cs

    c = make_case(100, 3, 4, "u", np.float64, seed=1)
    x, y = _t(c["x"]), _t(c["y"])
    pidx, tree = cs.quadtree_on_points((x, y), *c["ext"], c["scale"], 4, 10)
    e32 = torch.empty(0, dtype=torch.uint32, device="cuda")
    ls = (_t(np.array([0, 2], np.uint32)), _t(np.array([0.0, 1.0])), _t(np.array([0.0, 1.0])))
    out = cs.quadtree_point_to_nearest_linestring((e32, e32), tree, pidx, (x, y), ls)
    assert len(out) == 0 and out["distance"].dtype == torch.float64
    with pytest.raises(RuntimeError, match="same data type"):
        cs.quadtree_point_to_nearest_linestring((e32, e32), tree, pidx, (x, y),
                                                (ls[0], ls[1].float(), ls[2].float()))
    with pytest.raises(RuntimeError, match="expansion radius"):
        cs.linestring_bounding_boxes(ls, -1.0)
    with pytest.raises(RuntimeError, match="at least 2 vertices"):
        cs.linestring_bounding_boxes((_t(np.array([0, 1, 2], np.uint32)), ls[1], ls[2]), 0.0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("world", [1, 3, 8])
def test_sharding_kernels_on_one_gpu(dtype, world):
    """The multi-GPU kernels with every "rank" living on this GPU: keys equal the quadtree
    builder's keys; the device plan (two-level splitters, send counts, offsets) equals its host
    restatement in multi_gpu.py; the fused partition is stable and complete (bulk-copy and plain
    store variants); quadtree_on_keys equals quadtree_on_points; the refinement with coordinate
    segments equals the plain one."""
    import ctypes as C

    import torch

    import cuspatial_b200 as cs
    from cuspatial_b200 import _lib, multi_gpu as mg
    from cuspatial_b200.api import _ptr, _stream

    c = make_case(300_000, 40, 15, "c", dtype, seed=5, oob=100, dups=1000, median_vertices=30)
    n = len(c["x"])
    ext = c["ext"]
    cuts = [r * n // world for r in range(world + 1)]
    xs = [_t(c["x"][a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    ys = [_t(c["y"][a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    sizes = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
    x, y = _t(c["x"]), _t(c["y"])
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], c["scale"], 15, 64)
    shift = mg.hist_shift_for(15)
    n_bins = 1 << mg.HIST_BITS
    sub_bits = mg.sub_bits_for(world, shift)
    sub_shift, n_sub = shift - sub_bits, 1 << sub_bits
    # ---- keys + histogram + flags, per virtual rank
    keys, exts = [], []
    for r in range(world):
        k, e = mg.cuda_keys_and_histogram(xs[r], ys[r], ext, c["scale"], 15, shift, n_bins)
        keys.append(k); exts.append(e)
    kall = np.concatenate([k.cpu().numpy().view(np.uint32) for k in keys])
    np.testing.assert_array_equal(np.sort(kall, kind="stable"), tree._sorted_keys.cpu().numpy())
    np.testing.assert_array_equal(np.argsort(kall, kind="stable").astype(np.uint32),
                                  pidx.cpu().numpy())
    for r in range(world):
        kr = keys[r].cpu().numpy().view(np.uint32)
        np.testing.assert_array_equal(exts[r][:n_bins].cpu().numpy(),
                                      np.bincount(kr >> shift, minlength=n_bins))
    assert int(sum(e[n_bins] for e in exts)) >= 1       # out-of-box points were flagged
    gext = torch.stack(exts).sum(0).to(torch.int32)
    ghist = gext[:n_bins].cpu().numpy().astype(np.int64)
    targets, bounds = mg.refine_splitters(ghist, world, shift)
    # ---- plan on the device, per virtual rank; compare with the host restatement
    plans, subs = [], []
    for r in range(world):
        p = mg.cuda_plan_level1(gext[:n_bins].contiguous(), sizes, world, r, shift, sub_shift, n_sub)
        plans.append(p)
        subs.append(mg.cuda_sub_histogram(keys[r], p, world, n_sub))
    gsub = torch.stack(subs).sum(0).to(torch.int32)
    sub_np = gsub.cpu().numpy().astype(np.int64).reshape(-1, n_sub)[:len(targets)]
    want_split = mg.splitters_from_subhist(bounds, targets, sub_np, shift, sub_shift) \
        if world > 1 else np.empty(0, np.uint32)
    send = []
    for r in range(world):
        sc = mg.cuda_plan_level2(plans[r], exts[r][:n_bins].contiguous(), subs[r], gsub, world)
        send.append(sc.clone())
    M = torch.stack(send).contiguous()                   # [src][dst]
    dest_all = np.searchsorted(want_split.astype(np.int64), kall.astype(np.int64), side="right")
    for r in range(world):
        kr = keys[r].cpu().numpy().view(np.uint32).astype(np.int64)
        d = np.searchsorted(want_split.astype(np.int64), kr, side="right")
        np.testing.assert_array_equal(M[r].cpu().numpy(), np.bincount(d, minlength=world))
    recv_tot = np.bincount(dest_all, minlength=world)
    assert recv_tot.max() < 1.05 * n / world + 2000      # balanced (two-level splitters)
    # ---- fused partition into per-destination buffers (all local here), both store variants
    for bulk in (1, 0):
        cap = int(recv_tot.max()) + 64
        bk = [torch.full((cap,), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
        bg = [torch.full((cap,), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
        pk, pg = (C.c_void_p * world)(), (C.c_void_p * world)()
        for d in range(world):
            pk[d], pg[d] = bk[d].data_ptr(), bg[d].data_ptr()
        for r in range(world):
            L = _lib.lib()
            _lib.check(L.bsj_shard_plan_finalize(_ptr(M), cap, plans[r].ptr(), _stream(x.device)))
            _lib.check(L.bsj_partition_keys(_ptr(keys[r]), keys[r].shape[0], plans[r].ptr(), world,
                                            pk, pg, bulk, _stream(x.device)))
            h = plans[r].read_back()
            assert h.status == 0 and list(h.splitter[:world - 1]) == want_split.tolist()
            assert list(h.recv_total[:world]) == recv_tot.tolist()
            assert h.gid_base[r] == cuts[r] and h.n_targets == len(targets)
        torch.cuda.synchronize()
        for d in range(world):
            ids = np.nonzero(dest_all == d)[0]             # ascending global id == stable order
            np.testing.assert_array_equal(bg[d][:len(ids)].cpu().numpy().view(np.uint32), ids)
            np.testing.assert_array_equal(bk[d][:len(ids)].cpu().numpy().view(np.uint32), kall[ids])
            assert int((bg[d][len(ids):] != -1).sum()) == 0   # nothing written past the bucket
    # too small a receive capacity is reported, and the partition then writes nothing
    _lib.check(_lib.lib().bsj_shard_plan_finalize(_ptr(M), 10, plans[0].ptr(), _stream(x.device)))
    assert plans[0].read_back().status == 1
    # ---- owner side: quadtree_on_keys + refinement through coordinate segments, for rank 0's range
    ids = np.nonzero(dest_all == 0)[0]
    rk, rg = bk[0][:len(ids)].clone(), bg[0][:len(ids)].clone()
    pts = mg.ShardedPoints(xs[0], ys[0], sizes, None, [t.data_ptr() for t in xs],
                           [t.data_ptr() for t in ys])
    polys = tuple(_t(a) for a in (c["po"], c["ro"], c["vx"], c["vy"]))
    flags = [int(gext[n_bins]), int(gext[n_bins + 1])]
    pg_, comp, n_hits = mg.cuda_local_compact(rk, rg, pts, flags, polys, ext, c["scale"], 15, 64)
    # the same key range on the plain single-GPU path
    xr, yr = _t(c["x"][ids]), _t(c["y"][ids])
    pidx_r, tree_r = cs.quadtree_on_points((xr, yr), ext[0], ext[1], ext[2], ext[3], c["scale"],
                                           15, 64)
    np.testing.assert_array_equal(pg_.cpu().numpy().view(np.uint32),
                                  ids[pidx_r.cpu().numpy().astype(np.int64)])
    bb = cs.polygon_bounding_boxes(polys)
    pairs_r = cs.join_quadtree_and_bounding_boxes(tree_r, bb, ext[0], ext[1], ext[2], ext[3],
                                                  c["scale"], 15)
    hits_r = cs.quadtree_point_in_polygon(pairs_r, tree_r, pidx_r, (xr, yr), polys)
    op = torch.empty(n_hits, dtype=torch.int32, device="cuda")
    oq = torch.empty(n_hits, dtype=torch.int32, device="cuda")
    mg.cuda_expand(comp, n_hits, 0, op, oq)
    np.testing.assert_array_equal(op.cpu().numpy().view(np.uint32),
                                  hits_r["polygon_index"].cpu().numpy())
    np.testing.assert_array_equal(oq.cpu().numpy().view(np.uint32),
                                  hits_r["point_index"].cpu().numpy())
    assert n_hits > 0


def test_error_conditions_match_reference():
    """cpp/tests/join/join_quadtree_and_bounding_boxes_test.cpp:35-86."""
    import cuspatial_b200 as cs

    c = make_case(1000, 5, 5, "u", np.float64, seed=1)
    ext = c["ext"]
    pidx, tree = cs.quadtree_on_points((_t(c["x"]), _t(c["y"])), *ext, c["scale"], 5, 16)
    bb = cs.polygon_bounding_boxes(tuple(_t(a) for a in (c["po"], c["ro"], c["vx"], c["vy"])))
    from cuspatial_b200 import _lib
    import ctypes as C

    def raw(scale, x_min, x_max, max_depth):
        out = _lib.bsj_pairs()
        cols = [tree[k] for k in TREE_COLS]
        rc = _lib.lib().bsj_join_quadtree_and_bounding_boxes(
            *[C.c_void_p(t.data_ptr()) for t in cols], len(tree),
            *[C.c_void_p(bb[k].data_ptr()) for k in ("minx", "miny", "maxx", "maxy")], 1, len(bb),
            x_min, x_max, 0.0, 1.0, scale, max_depth, None, None, C.byref(out))
        return rc, _lib.lib().bsj_last_error().decode()

    rc, msg = raw(0.0, 0.0, 1.0, 5)
    assert rc == _lib.BSJ_INVALID_ARGUMENT and "scale must be positive" in msg
    rc, msg = raw(1.0, 1.0, 0.0, 5)
    assert rc == _lib.BSJ_INVALID_ARGUMENT and "invalid bounding box" in msg
    rc, msg = raw(1.0, 0.0, 1.0, 16)
    assert rc == _lib.BSJ_INVALID_ARGUMENT and "maximum depth must be positive and less than 16" in msg
    rc, msg = raw(1.0, 0.0, 1.0, 0)
    assert rc == _lib.BSJ_INVALID_ARGUMENT
    with pytest.raises(RuntimeError):
        cs.join_quadtree_and_bounding_boxes(tree, bb, 0, 1, 0, 1, 1.0, 16)


def test_empty_inputs_return_empty_outputs():
    import torch

    import cuspatial_b200 as cs

    e = torch.empty(0, dtype=torch.float64, device="cuda")
    pidx, tree = cs.quadtree_on_points((e, e), 0, 1, 0, 1, 1, 3, 2)
    assert pidx.numel() == 0 and len(tree) == 0 and tree.columns == list(TREE_COLS)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, (e, e, e, e), 0, 1, 0, 1, 1, 3)
    assert len(pairs) == 0 and pairs.columns == ["bbox_offset", "quad_offset"]
    eo = torch.zeros(1, dtype=torch.int32, device="cuda").to(torch.uint32)
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (e, e), (eo, eo, e, e))
    assert len(hits) == 0 and hits.columns == ["polygon_index", "point_index"]


def test_scale_clamp_warning_like_reference():
    import cuspatial_b200 as cs

    c = make_case(1000, 5, 5, "u", np.float64, seed=1)
    with pytest.warns(UserWarning, match="is less than required minimum"):
        cs.quadtree_on_points((_t(c["x"]), _t(c["y"])), *c["ext"], -1, 5, 16)


def test_full_size_properties_20M(oracle_lib):
    """Size-independent invariants at a size the oracle cannot check directly (20M points):
    point_indices is a permutation, keys along it are sorted with ties in index order (stable),
    leaf ranges tile [0, N), pair rows are ordered, every emitted row verifies under the oracle
    predicate on a sample, and indexed hits == brute-force bitmask hits on a sample."""
    import torch

    import cuspatial_b200 as cs

    n = 20_000_000
    from cuspatial_b200 import datagen as D

    po, ro, vx, vy = D.taxi_zone_like_polygons(263, seed=20251017)
    ext = D.polygon_extent(vx, vy)
    scale = D.quadtree_params(ext, 15)
    x, y = D.uniform_points_torch(n, ext, 1, torch.float64, "cuda")
    polys = tuple(_t(a) for a in (po, ro, vx, vy))
    pidx, tree = cs.quadtree_on_points((x, y), *ext, scale, 15, 512)
    pi = pidx.to(torch.int64)
    assert torch.equal(torch.sort(pi).values, torch.arange(n, device="cuda"))
    t = {k: tree[k].to(torch.int64) for k in TREE_COLS}
    leaf = t["is_internal_node"] == 0
    lo, ll = t["offset"][leaf], t["length"][leaf]
    order = torch.argsort(lo)
    assert int(lo[order][0]) == 0 and int((lo[order] + ll[order])[-1]) == n
    assert torch.equal((lo[order] + ll[order])[:-1], lo[order][1:])
    inner = ~leaf
    assert torch.all(t["length"][inner] >= 1) and torch.all(t["length"][inner] <= 4)
    bb = cs.polygon_bounding_boxes(polys)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, *ext, scale, 15)
    po_ = t["offset"][pairs["quad_offset"].to(torch.int64)]
    key = po_ * 1024 + pairs["bbox_offset"].to(torch.int64)
    assert torch.all(key[1:] > key[:-1])           # ordered by (leaf offset, bbox), no duplicates
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polys)
    hp, hq = hits["polygon_index"].to(torch.int64), hits["point_index"].to(torch.int64)
    assert len(hits) > n // 2
    # sample 200k points: brute-force bitmask over the first 31 polygons must agree
    sub = torch.randperm(n, device="cuda")[:200_000]
    polys31 = (polys[0][:32].to(torch.int32), polys[1].to(torch.int32), polys[2], polys[3])
    mask = cs.point_in_polygon_bitmask((x[sub], y[sub]), polys31).cpu().numpy()
    want = oracle_lib.point_in_polygon(x[sub].cpu().numpy(), y[sub].cpu().numpy(),
                                       po[:32].astype(np.int32), ro.astype(np.int32), vx, vy)
    np.testing.assert_array_equal(mask, want)
    orig = pi[hq]                                   # original point id of every hit row
    sel = hp < 31
    got = torch.zeros(n, dtype=torch.int64, device="cuda")
    got.scatter_add_(0, orig[sel], (1 << hp[sel]))
    np.testing.assert_array_equal(got[sub].cpu().numpy(), want.astype(np.int64))

"""Shared helpers for the parity tests: run the whole path on a host checker or on the GPU."""
import numpy as np

from cuspatial_b200 import datagen as D

TREE_COLS = ("key", "level", "is_internal_node", "length", "offset")


def in_contract(x, y, ext, scale, depth, dtype):
    """Mask of the points whose cell index, computed the way the reference does in `dtype`
    (phase_1.cuh:78-85: subtract, IEEE divide, truncate), stays below 2^depth -- the regime in
    which the reference's tree construction is well defined.  Out-of-bbox points
    (which take the fixed key 4^depth - 1) are in contract."""
    T = dtype
    x0, x1, y0, y1 = (T(v) for v in ext)
    lo_x, hi_x, lo_y, hi_y = min(x0, x1), max(x0, x1), min(y0, y1), max(y0, y1)
    s = max(T(scale), max(hi_x - lo_x, hi_y - lo_y) / T((1 << depth) + 2))
    inside = ~((x < lo_x) | (x > hi_x) | (y < lo_y) | (y > hi_y))
    with np.errstate(invalid="ignore"):
        ix = np.nan_to_num(((x - lo_x) / s).astype(T), nan=0.0)
        iy = np.nan_to_num(((y - lo_y) / s).astype(T), nan=0.0)
    lim = T(1 << depth)
    return ~inside | ((ix < lim) & (iy < lim))


def make_case(n, n_poly, depth, kind="u", dtype=np.float64, seed=0, median_vertices=40,
              oob=0, dups=0, extent=None):
    """`extent` = (x0, x1, y0, y1) places polygons and points there (default: the unit square);
    points that the cast to `dtype` pushes out of the reference's well-defined key range are
    dropped (only happens for coordinates far from the origin in float32)."""
    po, ro, vx, vy = D.taxi_zone_like_polygons(n_poly, seed=seed + 11, dtype=dtype,
                                               median_vertices=median_vertices,
                                               **({"extent": extent} if extent else {}))
    ext = D.polygon_extent(vx, vy)
    scale = D.quadtree_params(ext, depth)
    gen = D.uniform_points if kind == "u" else D.clustered_points
    x, y = gen(n, ext, seed=seed + 5, dtype=dtype)
    if extent is not None:
        keep = in_contract(x, y, ext, scale, depth, dtype)
        x, y = x[keep], y[keep]
    if oob:
        x[:oob] = dtype(ext[1] + (ext[1] - ext[0]))  # outside the area of interest
    if dups:
        x[oob:oob + dups] = x[oob]
        y[oob:oob + dups] = y[oob]
    return dict(x=x, y=y, po=po, ro=ro, vx=vx, vy=vy, ext=ext, scale=scale, depth=depth)


def pairwise_case(n, dtype, seed, degenerate=False):
    """n (point, polygon) pairs: polygon i from the taxi-zone generator, point i near it."""
    rng = np.random.default_rng(seed)
    po, ro, vx, vy = D.taxi_zone_like_polygons(n, seed=seed, dtype=dtype, median_vertices=24)
    if degenerate:  # repeat a vertex inside some rings: zero-length segments (b is not advanced)
        for r in rng.choice(len(ro) - 1, size=max(1, (len(ro) - 1) // 3), replace=False):
            a, b = int(ro[r]), int(ro[r + 1])
            if b - a > 4:
                j = int(rng.integers(a + 1, b - 2))
                vx[j + 1], vy[j + 1] = vx[j], vy[j]
    px, py = np.empty(n, dtype), np.empty(n, dtype)
    for i in range(n):
        a, b = int(ro[po[i]]), int(ro[po[i] + 1])
        kind = i % 4
        if kind == 0:      # bbox-uniform
            px[i] = rng.uniform(vx[a:b].min(), vx[a:b].max())
            py[i] = rng.uniform(vy[a:b].min(), vy[a:b].max())
        elif kind == 1:    # exactly a vertex (on the boundary => 0)
            j = int(rng.integers(a, b))
            px[i], py[i] = vx[j], vy[j]
        elif kind == 2:    # on / a few ulps off an edge
            j = int(rng.integers(a, b - 1))
            t = dtype(rng.uniform())
            px[i] = vx[j] + t * (vx[j + 1] - vx[j])
            py[i] = vy[j] + t * (vy[j + 1] - vy[j])
            if i % 8 == 2:
                px[i] = np.nextafter(px[i], dtype(np.inf))
        else:              # centroid-ish
            px[i], py[i] = vx[a:b].mean(), vy[a:b].mean()
    return px, py, po, ro, vx, vy


def make_linestrings(n_lines, ext, seed, dtype=np.float64, median_vertices=12):
    """Random-walk polylines inside the extent -> (part_offset u32[n+1], x, y)."""
    rng = np.random.default_rng(seed)
    w, h = ext[1] - ext[0], ext[3] - ext[2]
    counts = np.clip(rng.lognormal(np.log(median_vertices), 0.6, n_lines).astype(np.int64), 2, 400)
    lo = np.zeros(n_lines + 1, dtype=np.uint32)
    lo[1:] = np.cumsum(counts)
    xs, ys = np.empty(lo[-1], dtype), np.empty(lo[-1], dtype)
    for i in range(n_lines):
        n = int(counts[i])
        step = 0.02 * min(w, h)
        x0, y0 = rng.uniform(ext[0], ext[1]), rng.uniform(ext[2], ext[3])
        dx = np.cumsum(rng.normal(0, step, n))
        dy = np.cumsum(rng.normal(0, step, n))
        xs[lo[i]:lo[i + 1]] = np.clip(x0 + dx, ext[0], ext[1])
        ys[lo[i]:lo[i + 1]] = np.clip(y0 + dy, ext[2], ext[3])
        if i % 7 == 3 and n > 3:      # a zero-length segment
            xs[lo[i] + 2], ys[lo[i] + 2] = xs[lo[i] + 1], ys[lo[i] + 1]
    return lo, xs, ys


def covered_positions(tree, pair_quad, n_points):
    """Sorted positions of the points whose quadrant appears in the pair table."""
    m = np.zeros(n_points, dtype=bool)
    off, ln = np.asarray(tree["offset"]), np.asarray(tree["length"])
    for q in np.unique(np.asarray(pair_quad)):
        m[off[q]:off[q] + ln[q]] = True
    return m


def run_host_nearest(lib, c, lines, max_size, radius):
    ext = c["ext"]
    lo, lx, ly = lines
    tree = lib.quadtree_on_points(c["x"], c["y"], ext[0], ext[1], ext[2], ext[3], c["scale"],
                                  c["depth"], max_size)
    bb = lib.linestring_bounding_boxes(lo, lx, ly, radius)
    pairs = lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], c["scale"], c["depth"])
    out = lib.quadtree_point_to_nearest_linestring(pairs[0], pairs[1], tree, tree["point_indices"],
                                                   c["x"], c["y"], lo, lx, ly)
    return dict(tree=tree, bbox=bb, pairs=pairs, nearest=out)


def run_gpu_nearest(c, lines, max_size, radius):
    import torch

    import cuspatial_b200 as cs

    dev = "cuda"
    ext = c["ext"]
    x, y = torch.as_tensor(c["x"], device=dev), torch.as_tensor(c["y"], device=dev)
    ls = tuple(torch.as_tensor(a, device=dev) for a in lines)
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], c["scale"],
                                       c["depth"], max_size)
    bb = cs.linestring_bounding_boxes(ls, radius)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3],
                                                c["scale"], c["depth"])
    out = cs.quadtree_point_to_nearest_linestring(pairs, tree, pidx, (x, y), ls)
    assert out.columns == ["point_index", "linestring_index", "distance"]
    t = {k: tree[k].cpu().numpy().astype(np.uint8 if k in ("level", "is_internal_node")
                                         else np.uint32) for k in TREE_COLS}
    t["point_indices"] = pidx.cpu().numpy()
    return dict(tree=t, bbox=tuple(bb[k].cpu().numpy() for k in ("minx", "miny", "maxx", "maxy")),
                pairs=(pairs["bbox_offset"].cpu().numpy(), pairs["quad_offset"].cpu().numpy()),
                nearest=tuple(out[k].cpu().numpy() for k in out.columns))


def run_ref_cuda_nearest(c, lines, max_size, radius):
    """The same flow on the reference's own CUDA build (oracle/_ref/libcuspatial_ref_cuda.so)."""
    import torch

    from oracle import cudalib

    lib = cudalib.reference_cuda()
    dev = "cuda"
    ext = c["ext"]
    x, y = torch.as_tensor(c["x"], device=dev), torch.as_tensor(c["y"], device=dev)
    lo, lx, ly = (torch.as_tensor(a, device=dev) for a in lines)
    tree, _ = lib.quadtree_on_points(x, y, ext[0], ext[1], ext[2], ext[3], c["scale"], c["depth"],
                                     max_size)
    bb = lib.linestring_bounding_boxes(lo, lx, ly, radius)
    pairs, _ = lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], c["scale"],
                                                    c["depth"])
    out, _ = lib.quadtree_point_to_nearest_linestring(pairs[0], pairs[1], tree,
                                                      tree["point_indices"], x, y, lo, lx, ly)
    return dict(tree={k: v.cpu().numpy() for k, v in tree.items()},
                bbox=tuple(b.cpu().numpy() for b in bb),
                pairs=tuple(p.cpu().numpy() for p in pairs),
                nearest=tuple(o.cpu().numpy() for o in out))


def assert_same_nearest(a, b, what="", covered_only=False):
    for k in TREE_COLS + ("point_indices",):
        np.testing.assert_array_equal(a["tree"][k], b["tree"][k], err_msg="%s tree.%s" % (what, k))
    for i in range(4):
        np.testing.assert_array_equal(a["bbox"][i], b["bbox"][i], err_msg="%s bbox[%d]" % (what, i))
    for i in range(2):
        np.testing.assert_array_equal(a["pairs"][i], b["pairs"][i], err_msg="%s pairs[%d]" % (what, i))
    m = slice(None)
    if covered_only:  # the reference leaves the index columns of uncovered points uninitialised
        m = covered_positions(a["tree"], a["pairs"][1], len(a["tree"]["point_indices"]))
        np.testing.assert_array_equal(a["nearest"][2][~m], 0, err_msg=what + " uncovered distance")
        np.testing.assert_array_equal(b["nearest"][2][~m], 0, err_msg=what + " uncovered distance")
    for i, name in enumerate(("point_index", "linestring_index", "distance")):
        np.testing.assert_array_equal(a["nearest"][i][m], b["nearest"][i][m],
                                      err_msg="%s nearest.%s" % (what, name))


def run_host(lib, c, max_size):
    """Full path on a HostLib (oracle or reference host build)."""
    ext = c["ext"]
    tree = lib.quadtree_on_points(c["x"], c["y"], ext[0], ext[1], ext[2], ext[3], c["scale"],
                                  c["depth"], max_size)
    bb = lib.polygon_bounding_boxes(c["po"], c["ro"], c["vx"], c["vy"])
    if c["depth"] >= 1:
        pairs = lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], c["scale"],
                                                     c["depth"])
        hits = lib.quadtree_point_in_polygon(pairs[0], pairs[1], tree, tree["point_indices"],
                                             c["x"], c["y"], c["po"], c["ro"], c["vx"], c["vy"])
    else:
        pairs = hits = (np.empty(0, np.uint32), np.empty(0, np.uint32))
    return dict(tree=tree, bbox=bb, pairs=pairs, hits=hits)


def run_gpu(c, max_size, use_grid_hint=True):
    """Full path through the product's Python API (-> C ABI -> CUDA kernels).
    use_grid_hint=False drops the cell-geometry hint so that every quadrant goes through the
    per-point refinement (the two must give identical rows)."""
    import torch

    import cuspatial_b200 as cs

    dev = "cuda"
    ext = c["ext"]
    x, y = torch.as_tensor(c["x"], device=dev), torch.as_tensor(c["y"], device=dev)
    polys = tuple(torch.as_tensor(a, device=dev) for a in (c["po"], c["ro"], c["vx"], c["vy"]))
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], c["scale"],
                                       c["depth"], max_size)
    bb = cs.polygon_bounding_boxes(polys)
    t = {k: tree[k].cpu().numpy().astype(np.uint8 if k in ("level", "is_internal_node")
                                         else np.uint32) for k in TREE_COLS}
    t["point_indices"] = pidx.cpu().numpy()
    out = dict(tree=t, bbox=tuple(bb[k].cpu().numpy() for k in ("minx", "miny", "maxx", "maxy")))
    if c["depth"] >= 1:
        pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3],
                                                    c["scale"], c["depth"])
        if use_grid_hint == "no_keys":      # cell rectangles only, no per-point cell test
            tree._grid.sorted_keys = None
            tree._grid.n_sorted_keys = 0
        elif not use_grid_hint:
            tree._grid = None
        hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polys)
        out["pairs"] = (pairs["bbox_offset"].cpu().numpy(), pairs["quad_offset"].cpu().numpy())
        out["hits"] = (hits["polygon_index"].cpu().numpy(), hits["point_index"].cpu().numpy())
    else:
        out["pairs"] = out["hits"] = (np.empty(0, np.uint32), np.empty(0, np.uint32))
    return out


def run_ref_cuda(c, max_size):
    """Full path on the reference's own CUDA build (oracle/_ref/libcuspatial_ref_cuda.so)."""
    import torch

    from oracle import cudalib

    lib = cudalib.reference_cuda()
    dev = "cuda"
    ext = c["ext"]
    x, y = torch.as_tensor(c["x"], device=dev), torch.as_tensor(c["y"], device=dev)
    po, ro, vx, vy = (torch.as_tensor(a, device=dev) for a in (c["po"], c["ro"], c["vx"], c["vy"]))
    tree, _ = lib.quadtree_on_points(x, y, ext[0], ext[1], ext[2], ext[3], c["scale"], c["depth"],
                                     max_size)
    bb = lib.polygon_bounding_boxes(po, ro, vx, vy)
    pairs, _ = lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], c["scale"],
                                                    c["depth"])
    hits, _ = lib.quadtree_point_in_polygon(pairs[0], pairs[1], tree, tree["point_indices"], x, y,
                                            po, ro, vx, vy)
    return dict(tree={k: v.cpu().numpy() for k, v in tree.items()},
                bbox=tuple(b.cpu().numpy() for b in bb),
                pairs=tuple(p.cpu().numpy() for p in pairs),
                hits=tuple(h.cpu().numpy() for h in hits))


def assert_same(a, b, what=""):
    for k in TREE_COLS + ("point_indices",):
        np.testing.assert_array_equal(a["tree"][k], b["tree"][k], err_msg="%s tree.%s" % (what, k))
    for i in range(4):
        np.testing.assert_array_equal(a["bbox"][i], b["bbox"][i], err_msg="%s bbox[%d]" % (what, i))
    for name in ("pairs", "hits"):
        for i in range(2):
            np.testing.assert_array_equal(a[name][i], b[name][i],
                                          err_msg="%s %s[%d]" % (what, name, i))

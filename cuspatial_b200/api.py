"""Python drop-in for the reference's quadtree point-in-polygon join entry points.

Same names, positional order, pre-processing, warnings, column names and dtypes as
  cuspatial.quadtree_on_points               (python/cuspatial/cuspatial/core/spatial/indexing.py:15-199)
  cuspatial.join_quadtree_and_bounding_boxes (core/spatial/join.py:105-175)
  cuspatial.quadtree_point_in_polygon        (core/spatial/join.py:178-262)
  cuspatial.point_in_polygon                 (core/spatial/join.py:23-102)
  cuspatial.polygon_bounding_boxes           (core/spatial/bounding.py:19-80)
but over plain device tensors through the C ABI -- no cuDF, RMM or GeoSeries on the hot path:
  points   : (x, y) tensors, an (N, 2) tensor, or a flat interleaved xy tensor (GeoArrow)
  polygons : (part_offset, ring_offset, x, y) tensors (GeoArrow: one polygon per geometry)
Inputs must live on a CUDA device (torch tensors; anything exposing __cuda_array_interface__ is
wrapped with torch.as_tensor).  Outputs are torch tensors allocated through torch's caching
allocator (handed to the library as its output allocator, the analogue of the reference's `mr`).
"""
import ctypes as C
import threading
import warnings

import torch

from . import _lib
from .frame import Frame

_DTYPE_CODE = {torch.float32: 0, torch.float64: 1}


class _TorchAllocator:
    """bsj_allocator backed by torch.empty: output columns become ordinary torch tensors."""

    def __init__(self, device):
        self.device = device
        self.live = {}
        self._alloc = _lib.ALLOC_FN(self._allocate)
        self._free = _lib.FREE_FN(self._deallocate)
        self.struct = _lib.bsj_allocator(self._alloc, self._free, None)

    def _allocate(self, nbytes, stream, ctx):
        try:
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        except Exception:
            return None
        self.live[t.data_ptr()] = t
        return t.data_ptr()

    def _deallocate(self, ptr, nbytes, stream, ctx):
        self.live.pop(ptr, None)

    def take(self, ptr, n, dtype):
        """The tensor behind `ptr`, viewed as n elements of dtype."""
        if n == 0 or not ptr:
            return torch.empty(0, dtype=dtype, device=self.device)
        v = self.live.pop(ptr).view(dtype)
        return v if v.shape[0] == n else v[:n]


_ALLOCATORS = threading.local()


def _allocator_for(device):
    """One _TorchAllocator per (thread, device): building the two ctypes callbacks costs ~20 us per
    API call, which is GPU idle time when it happens right after a call's synchronisation."""
    cache = getattr(_ALLOCATORS, "by_device", None)
    if cache is None:
        cache = _ALLOCATORS.by_device = {}
    key = (device.type, device.index)
    a = cache.get(key)
    if a is None:
        a = cache[key] = _TorchAllocator(device)
    elif a.live:   # a previous call failed half-way: drop what it left behind
        a.live.clear()
    return a


def _as_cuda(t, dtype=None, name="input"):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, device="cuda") if not hasattr(t, "__cuda_array_interface__") \
            else torch.as_tensor(t, device="cuda")
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (no CPU fallback on this path)" % name)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _accessor(obj, kind):
    """A geoarrow.GeoArrowSeries -> its .points / .polygons / .lines accessor; anything else as is."""
    from .geoarrow import GeoArrowSeries

    return obj._get(kind) if isinstance(obj, GeoArrowSeries) else obj


def _reject_multi(obj, what):
    from .geoarrow import GeoArrowSeries, is_multi

    if isinstance(obj, GeoArrowSeries) and is_multi(obj):
        raise ValueError("GeoSeries cannot contain %s." % what)  # join.py:75-78, 323-326


def _split_points(points):
    """(x, y) | (N,2) | flat interleaved xy | GeoArrowSeries of points  ->  contiguous x, y
    (reference: geoseries .x/.y make strided copies of the interleaved buffer,
    core/geoseries.py:238-243)."""
    points = _accessor(points, "points")
    if hasattr(points, "xy") and hasattr(points, "x"):
        x, y = _as_cuda(points.x, name="points.x"), _as_cuda(points.y, name="points.y")
    elif isinstance(points, (tuple, list)) and len(points) == 2:
        x, y = _as_cuda(points[0], name="points.x"), _as_cuda(points[1], name="points.y")
    else:
        p = _as_cuda(points, name="points")
        if p.dim() == 2 and p.shape[1] == 2:
            x, y = p[:, 0].contiguous(), p[:, 1].contiguous()
        elif p.dim() == 1 and p.numel() % 2 == 0:
            x, y = p[0::2].contiguous(), p[1::2].contiguous()
        else:
            raise ValueError("points must be (x, y), an (N, 2) tensor or interleaved xy")
    if x.dtype != y.dtype or x.dtype not in _DTYPE_CODE:
        raise TypeError("point coordinates must both be float32 or float64")
    if x.shape[0] != y.shape[0]:
        raise RuntimeError("x and y columns must have the same length")
    return x, y


class _Binding:
    """What a quadtree's cell-geometry hint (bsj_grid + sorted keys) describes: the point buffers
    the tree was built from and its point_indices.  The source tensors are kept alive (so their
    addresses cannot be recycled for other data) together with their in-place modification
    counters; quadtree_point_in_polygon forwards the hint only for exactly these buffers,
    unmodified -- any other `points` / `point_indices` take the hint-free path, whose result is
    the reference's by construction."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.versions = [t._version for t in self.tensors]

    @staticmethod
    def sources(points):
        pts = _accessor(points, "points")
        if hasattr(pts, "xy") and hasattr(pts, "x"):
            srcs = [pts.xy]
        elif isinstance(pts, (tuple, list)):
            srcs = list(pts)
        else:
            srcs = [pts]
        return srcs if all(isinstance(t, torch.Tensor) for t in srcs) else None

    def matches(self, tensors):
        if tensors is None or len(tensors) != len(self.tensors):
            return False
        for a, b, v in zip(self.tensors, tensors, self.versions):
            same = a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and
                              a.stride() == b.stride() and a.dtype == b.dtype and
                              a._version == b._version)
            if not same or a._version != v:
                return False
        return True


def _split_polygons(polygons):
    polygons = _accessor(polygons, "polygons")
    if hasattr(polygons, "part_offset"):
        polygons = (polygons.part_offset, polygons.ring_offset, polygons.x, polygons.y)
    if not (isinstance(polygons, (tuple, list)) and len(polygons) == 4):
        raise ValueError("polygons must be (part_offset, ring_offset, x, y)")
    po, ro, vx, vy = polygons
    vx, vy = _as_cuda(vx, name="polygons.x"), _as_cuda(vy, name="polygons.y")
    return po, ro, vx, vy


def _ptr(t):
    return C.c_void_p(t.data_ptr() if t.numel() else 0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _quadtree_columns(quadtree):
    cols = []
    for name, dt in (("key", torch.uint32), ("level", torch.uint8),
                     ("is_internal_node", torch.bool), ("length", torch.uint32),
                     ("offset", torch.uint32)):
        t = _as_cuda(quadtree[name], name="quadtree." + name)
        if t.dtype != dt:
            t = t.to(dt)
        cols.append(t)
    return cols


def _clamp_scale(x_min, x_max, y_min, y_max, scale, max_depth):
    # indexing.py:165-186 / join.py:144-166 -- float64 arithmetic in Python, before the cast to T
    x_min, x_max, y_min, y_max = (min(x_min, x_max), max(x_min, x_max),
                                  min(y_min, y_max), max(y_min, y_max))
    min_scale = max(x_max - x_min, y_max - y_min) / ((1 << max_depth) + 2)
    if scale < min_scale:
        warnings.warn("scale {} is less than required minimum ".format(scale)
                      + "scale {}. Clamping to minimum scale".format(min_scale))
    return x_min, x_max, y_min, y_max, max(scale, min_scale)


def quadtree_on_points(points, x_min, x_max, y_min, y_max, scale, max_depth, max_size):
    """Construct a quadtree from a set of points for a given area-of-interest bounding box.

    Returns (point_indices uint32 tensor, Frame[key u32, level u8, is_internal_node bool,
    length u32, offset u32]) -- reference: indexing.py:15-199, dtypes per test_indexing.py:27-31.
    """
    x, y = _split_points(points)
    x_min, x_max, y_min, y_max, scale = _clamp_scale(x_min, x_max, y_min, y_max, scale, max_depth)
    dev = x.device
    with torch.cuda.device(dev):
        alloc = _allocator_for(dev)
        out = _lib.bsj_quadtree()
        rc = _lib.lib().bsj_quadtree_on_points(
            _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], float(x_min), float(x_max),
            float(y_min), float(y_max), float(scale), int(max_depth), int(max_size),
            C.byref(alloc.struct), _stream(dev), C.byref(out))
        _lib.check(rc)
    n, q = int(out.num_points), int(out.num_nodes)
    point_indices = alloc.take(out.point_indices, n, torch.uint32)
    tree = Frame([
        ("key", alloc.take(out.key, q, torch.uint32)),
        ("level", alloc.take(out.level, q, torch.uint8)),
        ("is_internal_node", alloc.take(out.is_internal_node, q, torch.bool)),
        ("length", alloc.take(out.length, q, torch.uint32)),
        ("offset", alloc.take(out.offset, q, torch.uint32)),
    ])
    # cell geometry of this tree: lets quadtree_point_in_polygon settle whole quadrants without
    # reading their points (include/cuspatial_b200.h, bsj_grid).  Rides along on the Frame only.
    g = _lib.bsj_grid()
    C.memmove(C.byref(g), C.byref(out.grid), C.sizeof(_lib.bsj_grid))
    tree._grid = g
    # the hint points at the sorted Morton keys: keep that buffer alive with the Frame
    tree._sorted_keys = alloc.take(out.sorted_keys, n, torch.uint32)
    # ... and is only valid for these very buffers (see _Binding)
    srcs = _Binding.sources(points)
    tree._hint_points = _Binding(srcs) if srcs is not None else None
    tree._hint_indices = _Binding([point_indices])
    return point_indices, tree


def _hint_for(quadtree, points, point_indices):
    """The tree's bsj_grid if `points` / `point_indices` are the buffers it was computed from."""
    grid = getattr(quadtree, "_grid", None)
    if grid is None:
        return None
    bp, bi = getattr(quadtree, "_hint_points", None), getattr(quadtree, "_hint_indices", None)
    if bp is None or bi is None:
        return None
    if not bp.matches(_Binding.sources(points)):
        return None
    if not (isinstance(point_indices, torch.Tensor) and bi.matches([point_indices])):
        return None
    return grid


def join_quadtree_and_bounding_boxes(quadtree, bounding_boxes, x_min, x_max, y_min, y_max, scale,
                                     max_depth):
    """Search a quadtree for polygon or linestring bounding box intersections.

    Returns Frame[bbox_offset u32, quad_offset u32] -- reference: join.py:105-175.
    `bounding_boxes`: Frame/dict with minx, miny, maxx, maxy or a 4-tuple in that order.
    """
    x_min, x_max, y_min, y_max, scale = _clamp_scale(x_min, x_max, y_min, y_max, scale, max_depth)
    if isinstance(bounding_boxes, (tuple, list)):
        bcols = list(bounding_boxes)
    else:
        bcols = [bounding_boxes[c] for c in ("minx", "miny", "maxx", "maxy")]
    if len(bcols) != 4:
        raise RuntimeError("bbox table must have 4 columns")
    bcols = [_as_cuda(b, name="bounding_boxes") for b in bcols]
    if bcols[0].dtype not in _DTYPE_CODE or any(b.dtype != bcols[0].dtype for b in bcols):
        raise TypeError("bounding box columns must all be float32 or float64")
    tcols = _quadtree_columns(quadtree)
    dev = bcols[0].device
    with torch.cuda.device(dev):
        alloc = _allocator_for(dev)
        out = _lib.bsj_pairs()
        rc = _lib.lib().bsj_join_quadtree_and_bounding_boxes(
            *[_ptr(t) for t in tcols], tcols[0].shape[0], *[_ptr(b) for b in bcols],
            _DTYPE_CODE[bcols[0].dtype], bcols[0].shape[0], float(x_min), float(x_max),
            float(y_min), float(y_max), float(scale), int(max_depth), C.byref(alloc.struct),
            _stream(dev), C.byref(out))
        _lib.check(rc)
    p = int(out.size)
    return Frame([("bbox_offset", alloc.take(out.first, p, torch.uint32)),
                  ("quad_offset", alloc.take(out.second, p, torch.uint32))])


def quadtree_point_in_polygon(poly_quad_pairs, quadtree, point_indices, points, polygons):
    """Test whether the specified points are inside any of the specified polygons.

    Returns Frame[polygon_index u32, point_index u32]; point_index indexes `point_indices`
    ("an index to an index") -- reference: join.py:178-262.
    """
    x, y = _split_points(points)
    po, ro, vx, vy = _split_polygons(polygons)
    if vx.dtype != vy.dtype:
        raise RuntimeError("polygon columns must have the same data type")
    if x.dtype != vx.dtype:
        raise RuntimeError("points and polygons must have the same data type")
    po = _as_cuda(po, torch.uint32 if getattr(po, "dtype", None) != torch.int32 else None)
    ro = _as_cuda(ro, torch.uint32 if getattr(ro, "dtype", None) != torch.int32 else None)
    if isinstance(poly_quad_pairs, (tuple, list)):
        pp, pq = poly_quad_pairs
    else:
        names = list(poly_quad_pairs.columns) if hasattr(poly_quad_pairs, "columns") \
            else list(poly_quad_pairs.keys())
        if len(names) != 2:
            raise RuntimeError("a quadrant-polygon table must have 2 columns")
        pp, pq = poly_quad_pairs[names[0]], poly_quad_pairs[names[1]]
    pp = _as_cuda(pp, torch.uint32)
    pq = _as_cuda(pq, torch.uint32)
    pi = _as_cuda(point_indices, torch.uint32)
    if pi.shape[0] != x.shape[0]:
        raise RuntimeError("number of points must be the same for both x and y columns")
    tcols = _quadtree_columns(quadtree)
    dev = x.device
    with torch.cuda.device(dev):
        alloc = _allocator_for(dev)
        out = _lib.bsj_pairs()
        grid = _hint_for(quadtree, points, point_indices)
        rc = _lib.lib().bsj_quadtree_point_in_polygon_ex(
            _ptr(pp), _ptr(pq), pp.shape[0], *[_ptr(t) for t in tcols], tcols[0].shape[0],
            _ptr(pi), _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], _ptr(po), po.shape[0],
            _ptr(ro), ro.shape[0], _ptr(vx), _ptr(vy), vx.shape[0],
            C.byref(grid) if grid is not None else None, C.byref(alloc.struct),
            _stream(dev), C.byref(out))
        _lib.check(rc)
    h = int(out.size)
    return Frame([("polygon_index", alloc.take(out.first, h, torch.uint32)),
                  ("point_index", alloc.take(out.second, h, torch.uint32))])


_COMPACT_FIELDS = (("pair_offset", torch.uint32, "n_pairs"), ("pair_length", torch.uint32, "n_pairs"),
                   ("pair_hits", torch.uint32, "n_pairs"), ("pair_class", torch.uint8, "n_pairs"),
                   ("pair_word_base", torch.int64, "n_pairs"),
                   ("pair_row_base", torch.int64, "n_pairs"), ("mask_words", torch.uint32, "n_words"))


def quadtree_point_in_polygon_compact(poly_quad_pairs, quadtree, point_indices, points, polygons):
    """quadtree_point_in_polygon stopped before the rows are written: returns the compact result
    (dict of device tensors + 'n_hits', see bsj_pip_compact) for `expand_pip_compact`.  Used by
    the multi-GPU merge, which all-gathers this form instead of the expanded table."""
    x, y = _split_points(points)
    po, ro, vx, vy = _split_polygons(polygons)
    po = _as_cuda(po, torch.uint32 if getattr(po, "dtype", None) != torch.int32 else None)
    ro = _as_cuda(ro, torch.uint32 if getattr(ro, "dtype", None) != torch.int32 else None)
    names = list(poly_quad_pairs.columns)
    pp = _as_cuda(poly_quad_pairs[names[0]], torch.uint32)
    pq = _as_cuda(poly_quad_pairs[names[1]], torch.uint32)
    pi = _as_cuda(point_indices, torch.uint32)
    tcols = _quadtree_columns(quadtree)
    dev = x.device
    with torch.cuda.device(dev):
        alloc = _allocator_for(dev)
        out = _lib.bsj_pip_compact()
        grid = _hint_for(quadtree, points, point_indices)
        rc = _lib.lib().bsj_quadtree_point_in_polygon_compact(
            _ptr(pp), _ptr(pq), pp.shape[0], *[_ptr(t) for t in tcols], tcols[0].shape[0],
            _ptr(pi), _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], _ptr(po), po.shape[0],
            _ptr(ro), ro.shape[0], _ptr(vx), _ptr(vy), vx.shape[0],
            C.byref(grid) if grid is not None else None, C.byref(alloc.struct), _stream(dev),
            C.byref(out))
        _lib.check(rc)
    res = {"pair_poly": pp, "n_hits": int(out.n_hits)}
    for name, dt, cnt in _COMPACT_FIELDS:
        n = int(getattr(out, cnt))
        if name == "mask_words":
            n = max(n, 1) if getattr(out, name) else 0
        res[name] = alloc.take(getattr(out, name), n, dt)
    return res


def expand_pip_compact(compact, position_base, out_polygon_index, out_point_index):
    """Write the rows of a compact result into caller-provided uint32/int32 tensors
    (point_index = sorted position + position_base)."""
    c = _lib.bsj_pip_compact()
    keep = []
    for name, dt, _ in _COMPACT_FIELDS:
        t = compact[name]
        if t.dtype != dt:
            t = t.view(dt) if t.element_size() == torch.empty(0, dtype=dt).element_size() else t.to(dt)
        t = t.contiguous()
        keep.append(t)
        setattr(c, name, t.data_ptr() if t.numel() else None)
    c.n_pairs = compact["pair_hits"].shape[0]
    c.n_words = compact["mask_words"].shape[0]
    c.n_hits = int(compact["n_hits"])
    pp = compact["pair_poly"].contiguous()
    dev = pp.device
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bsj_expand_pip_compact(
            _ptr(pp), C.byref(c), int(position_base) & 0xFFFFFFFF, _stream(dev),
            _ptr(out_polygon_index), _ptr(out_point_index)))


def point_in_polygon_bitmask(points, polygons):
    """The libcuspatial result of point_in_polygon: one INT32 per point, bit i = inside polygon i
    (cpp/include/cuspatial/point_in_polygon.hpp:75-82)."""
    x, y = _split_points(points)
    po, ro, vx, vy = _split_polygons(polygons)
    if x.dtype != vx.dtype or vx.dtype != vy.dtype:
        raise RuntimeError("All points much have the same type for both x and y")
    po = _as_cuda(po, torch.int32)
    ro = _as_cuda(ro, torch.int32)
    out = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().bsj_point_in_polygon(
            _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], _ptr(po), po.shape[0], _ptr(ro),
            ro.shape[0], _ptr(vx), _ptr(vy), vx.shape[0], _stream(x.device), _ptr(out))
        _lib.check(rc)
    return out


def point_in_polygon(points, polygons):
    """Compute from a set of points and a set of polygons which points fall within which polygons.

    Returns a Frame with one bool column per polygon (column i <-> polygon i), reference:
    join.py:23-102 (bitmask unpacked as utils/join_utils.py:12-46 does).
    """
    _reject_multi(polygons, "multipolygon")
    po = polygons[0] if isinstance(polygons, (tuple, list)) else polygons.part_offset
    n_poly = max(int(po.shape[0]) - 1, 0)
    if n_poly == 0:
        return Frame([])
    mask = point_in_polygon_bitmask(points, polygons)
    return Frame([(i, ((mask >> i) & 1).to(torch.bool)) for i in range(n_poly)])


def _split_linestrings(linestrings):
    """(part_offset, x, y) arrays in GeoArrow layout, or an object with .part_offset/.x/.y
    (the reference takes a GeoSeries and reads linestrings.lines.part_offset/x/y)."""
    linestrings = _accessor(linestrings, "linestrings")
    if hasattr(linestrings, "part_offset"):
        linestrings = (linestrings.part_offset, linestrings.x, linestrings.y)
    if not (isinstance(linestrings, (tuple, list)) and len(linestrings) == 3):
        raise ValueError("linestrings must be (part_offset, x, y)")
    lo, lx, ly = linestrings
    lx, ly = _as_cuda(lx, name="linestrings.x"), _as_cuda(ly, name="linestrings.y")
    lo = _as_cuda(lo, torch.uint32 if getattr(lo, "dtype", None) != torch.int32 else None)
    return lo, lx, ly


def linestring_bounding_boxes(linestrings, expansion_radius):
    """Axis-aligned bounding box of every linestring, grown by `expansion_radius`
    -> Frame[minx, miny, maxx, maxy] (reference: core/spatial/bounding.py:83-140)."""
    acc = _accessor(linestrings, "linestrings")
    lo, lx, ly = _split_linestrings(linestrings)
    if getattr(acc, "geometry_offset", None) is not None:
        # bounding.py:123-125: one box per geometry, over all of its parts
        go = _as_cuda(acc.geometry_offset, torch.int64)
        lo = lo.view(torch.int32)[go].view(lo.dtype) if lo.dtype == torch.uint32 else lo[go]
    if lx.dtype != ly.dtype:
        raise RuntimeError("Data type mismatch")
    if lx.shape[0] != ly.shape[0]:
        raise RuntimeError("x and y must be the same size")
    n = max(int(lo.shape[0]) - 1, 0)
    outs = [torch.empty(n, dtype=lx.dtype, device=lx.device) for _ in range(4)]
    with torch.cuda.device(lx.device):
        rc = _lib.lib().bsj_linestring_bounding_boxes(
            _ptr(lo), lo.shape[0], _ptr(lx), _ptr(ly), _DTYPE_CODE[lx.dtype], lx.shape[0],
            float(expansion_radius), _stream(lx.device), *[_ptr(o) for o in outs])
        _lib.check(rc)
    return Frame(zip(("minx", "miny", "maxx", "maxy"), outs))


def quadtree_point_to_nearest_linestring(linestring_quad_pairs, quadtree, point_indices, points,
                                         linestrings):
    """Finds the nearest linestring to each point in a quadrant, and the distance between them.

    Returns Frame[point_index u32, linestring_index u32, distance T], one row per point in
    quadtree order; point_index indexes `point_indices` -- reference: join.py:265-352.
    """
    _reject_multi(linestrings, "multilinestrings")
    x, y = _split_points(points)
    lo, lx, ly = _split_linestrings(linestrings)
    if lx.dtype != ly.dtype:
        raise RuntimeError("linestring columns must have the same data type")
    if x.dtype != lx.dtype:
        raise RuntimeError("points and linestrings must have the same data type")
    if lx.shape[0] != ly.shape[0]:
        raise RuntimeError("numbers of vertices must be the same for both x and y columns")
    if isinstance(linestring_quad_pairs, (tuple, list)):
        pl, pq = linestring_quad_pairs
    else:
        names = list(linestring_quad_pairs.columns) if hasattr(linestring_quad_pairs, "columns") \
            else list(linestring_quad_pairs.keys())
        if len(names) != 2:
            raise RuntimeError("a quadrant-linestring table must have 2 columns")
        pl, pq = linestring_quad_pairs[names[0]], linestring_quad_pairs[names[1]]
    pl = _as_cuda(pl, torch.uint32)
    pq = _as_cuda(pq, torch.uint32)
    pi = _as_cuda(point_indices, torch.uint32)
    if pi.shape[0] != x.shape[0]:
        raise RuntimeError("number of points must be the same for both x and y columns")
    tcols = _quadtree_columns(quadtree)
    dev = x.device
    n = x.shape[0]
    out_p = torch.empty(n, dtype=torch.uint32, device=dev)
    out_l = torch.empty(n, dtype=torch.uint32, device=dev)
    out_d = torch.empty(n, dtype=x.dtype, device=dev)
    rows = C.c_uint64(0)
    with torch.cuda.device(dev):
        rc = _lib.lib().bsj_quadtree_point_to_nearest_linestring(
            _ptr(pl), _ptr(pq), pl.shape[0], *[_ptr(t) for t in tcols], tcols[0].shape[0],
            _ptr(pi), _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], n, _ptr(lo), lo.shape[0], _ptr(lx),
            _ptr(ly), lx.shape[0], _stream(dev), _ptr(out_p), _ptr(out_l), _ptr(out_d),
            C.byref(rows))
        _lib.check(rc)
    r = int(rows.value)
    return Frame([("point_index", out_p[:r]), ("linestring_index", out_l[:r]),
                  ("distance", out_d[:r])])


def pairwise_point_in_polygon(points, polygons):
    """Point i against polygon i: one UINT8 per pair, 1 = strictly inside
    (reference: _lib/pairwise_point_in_polygon.pyx:15-44 over
    cpp/src/point_in_polygon/point_in_polygon.cu:172-190)."""
    x, y = _split_points(points)
    po, ro, vx, vy = _split_polygons(polygons)
    if x.dtype != vx.dtype or vx.dtype != vy.dtype:
        raise RuntimeError("All points much have the same type for both x and y")
    po = _as_cuda(po, torch.int32)
    ro = _as_cuda(ro, torch.int32)
    out = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().bsj_pairwise_point_in_polygon(
            _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], _ptr(po), po.shape[0], _ptr(ro),
            ro.shape[0], _ptr(vx), _ptr(vy), vx.shape[0], _stream(x.device), _ptr(out))
        _lib.check(rc)
    return out


def _quadtree_contains_properly(points, polygons):
    """core/binpreds/contains.py:19-74: the GeoPandas-compatible `contains_properly` driver of the
    hot path (max_depth 15, max_size ceil(sqrt(N)), minimum scale, extent of the polygon
    vertices).  Returns Frame[point_index (ORIGINAL point id), part_index (polygon id)]."""
    from math import ceil, sqrt

    x, y = _split_points(points)
    po, ro, vx, vy = _split_polygons(polygons)
    max_depth = 15
    min_size = ceil(sqrt(x.shape[0])) if x.shape[0] else 1
    if max(int(po.shape[0]) - 1, 0) == 0:
        return Frame([])
    x_min, x_max = float(vx.min()), float(vx.max())
    y_min, y_max = float(vy.min()), float(vy.max())
    scale = max(x_max - x_min, y_max - y_min) / ((1 << max_depth) + 2)
    point_indices, quadtree = quadtree_on_points((x, y), x_min, x_max, y_min, y_max, scale,
                                                 max_depth, min_size)
    poly_bboxes = polygon_bounding_boxes((po, ro, vx, vy))
    intersections = join_quadtree_and_bounding_boxes(quadtree, poly_bboxes, x_min, x_max, y_min,
                                                     y_max, scale, max_depth)
    rows = quadtree_point_in_polygon(intersections, quadtree, point_indices, (x, y),
                                     (po, ro, vx, vy))
    # (torch has no uint32 gather: same bits through an int32 view)
    original = point_indices.view(torch.int32)[rows["point_index"].to(torch.int64)].view(torch.uint32)
    return Frame([("point_index", original), ("part_index", rows["polygon_index"])])


def _pairwise_contains_properly(points, polygons):
    """contains.py:123-170 -> Frame[pairwise_index, point_index, result] of the true pairs."""
    flags = pairwise_point_in_polygon(points, polygons).to(torch.bool)
    trues = torch.nonzero(flags).reshape(-1)
    return Frame([("pairwise_index", trues), ("point_index", trues),
                  ("result", torch.ones_like(trues, dtype=torch.bool))])


def _brute_force_contains_properly(points, polygons):
    """contains.py:77-120 -> Frame with one bool column per polygon."""
    return point_in_polygon(points, polygons)


def contains_properly(polygons, points, mode="pairwise"):
    """core/binpreds/contains.py:173-185: which points are properly contained (boundary excluded)
    by which polygons; `mode` = "quadtree" (the indexed hot path), "pairwise" or anything else
    for the <= 31-polygon all-pairs bitmask."""
    if mode == "quadtree":
        return _quadtree_contains_properly(points, polygons)
    if mode == "pairwise":
        return _pairwise_contains_properly(points, polygons)
    table = _brute_force_contains_properly(points, polygons)
    cols = [table[c] for c in table.columns]
    if not cols:
        return Frame([])
    grid = torch.stack(cols, dim=1)                     # [point, polygon], like DataFrame.stack()
    nz = torch.nonzero(grid)
    return Frame([("point_index", nz[:, 0]), ("part_index", nz[:, 1]),
                  ("result", torch.ones(nz.shape[0], dtype=torch.bool, device=grid.device))])


def polygon_bounding_boxes(polygons, expansion_radius=0.0):
    """Axis-aligned bounding box of every polygon -> Frame[minx, miny, maxx, maxy]
    (reference: core/spatial/bounding.py:19-80)."""
    po, ro, vx, vy = _split_polygons(polygons)
    po = _as_cuda(po, torch.uint32 if getattr(po, "dtype", None) != torch.int32 else None)
    ro = _as_cuda(ro, torch.uint32 if getattr(ro, "dtype", None) != torch.int32 else None)
    n = max(int(po.shape[0]) - 1, 0)
    outs = [torch.empty(n, dtype=vx.dtype, device=vx.device) for _ in range(4)]
    with torch.cuda.device(vx.device):
        rc = _lib.lib().bsj_polygon_bounding_boxes(
            _ptr(po), po.shape[0], _ptr(ro), ro.shape[0], _ptr(vx), _ptr(vy),
            _DTYPE_CODE[vx.dtype], vx.shape[0], float(expansion_radius), _stream(vx.device),
            *[_ptr(o) for o in outs])
        _lib.check(rc)
    return Frame(zip(("minx", "miny", "maxx", "maxy"), outs))

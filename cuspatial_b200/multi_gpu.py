"""Multi-GPU quadtree point-in-polygon join: one process per GPU, torch.distributed (NCCL).

The reference is single-GPU; this is the sharded form of the same path.
Points shard by contiguous Morton-key range, so that global sorted position = rank base + local
sorted position.  Per call:

  1. the polygon table (offsets + vertices) is replicated by an NCCL broadcast from rank 0;
  2. every rank computes the Morton keys of its points and a histogram of their leading bits
     (CUDA, partition.cu); the histograms are summed with ONE all-reduce; the splitters are derived
     ON THE DEVICE in two levels (first-level bin of each rank boundary, then a second histogram
     of the next key bits inside those bins, one more all-reduce), as are this rank's send counts;
     one all-gather of the send counts gives every rank its write offsets.  Everything lives in a
     device struct (bsj_shard_plan): no host round trip up to here;
  3. the partition kernel stably partitions (key, global id) by destination rank and its stores go
     straight into the destination GPUs' receive buffers (symmetric memory, peer pointers over
     NVLink; per-destination runs leave the SM as bulk copies) -- partition and all-to-all in one
     kernel, 8 bytes per point on the wire; a device-side barrier follows.  The ONE host
     synchronisation of the exchange reads the plan back (receive counts);
  4. every rank sorts the keys it received (payload = global ids), builds its sub-quadtree with the
     GLOBAL area of interest / scale / depth / max_size, filters the polygon boxes and refines.
     Coordinates never moved: the refinement reads them through peer pointers into the owning
     ranks' point columns (ShardedPoints, symmetric memory) for the few points whose finest cell
     is touched by a polygon edge -- everything else is decided from the keys;
  5. what is merged is the COMPACT result (per-pair records + ballot words, tens of MB), one
     all-gather; every rank expands every rank's rows at HBM speed with
     point_index = rank base + local sorted position.

The merged pair SET equals a single-GPU run's bit for bit (the per-rank sub-quadtrees are not the
global quadtree, so row order inside the table differs; compare sorted rows), and the concatenated
point_indices equal the single-GPU point_indices.

The device steps are injected as callables so that the host-side logic (plan, index fix-up, merge)
is testable on CPU with the gloo backend (tests/test_multi_gpu_cpu.py); the functions
refine_splitters / splitters_from_subhist / send_counts_for below are the host restatement of the
plan kernels, and a GPU test checks the kernels against them.
"""
import os

import numpy as np
import torch

HIST_BITS = 13   # 8192 first-level bins: privatised in shared memory by the histogram kernel
MAX_RANKS = 32
LAST_PROFILE = {}


# ------------------------------------------------------------------------------------------------
# host restatement of the sharding plan (CPU tests; checker of the plan kernels)
# ------------------------------------------------------------------------------------------------
def choose_splitters(global_hist, n_ranks, shift):
    """One-level variant: R-1 ascending uint32 key splitters on bin boundaries balancing the
    point counts.  Rank r owns keys in [splitter[r-1], splitter[r])."""
    h = np.asarray(global_hist, dtype=np.int64)
    csum = np.cumsum(h)
    total = int(csum[-1]) if len(csum) else 0
    out = []
    for r in range(1, n_ranks):
        target = (total * r + n_ranks - 1) // n_ranks
        b = int(np.searchsorted(csum, target, side="left")) + 1  # first bin of the next rank
        b = min(max(b, (out[-1] >> shift) if out else 0), len(h))
        out.append(min(b << shift, 0xFFFFFFFF))
    return np.asarray(out, dtype=np.uint32)


def sub_bits_for(n_ranks, shift):
    """Width of the second-level histogram: (n_ranks - 1) * 2^bits counters must fit the
    sub-histogram kernel's 48 KB of shared memory (12288 counters)."""
    bits = 10
    while bits > 0 and max(n_ranks - 1, 1) * (1 << bits) > 12288:
        bits -= 1
    return min(bits, shift)


def refine_splitters(global_hist, n_ranks, shift):
    """Level 1: for every rank boundary the first-level bin in which the target cumulative count
    is reached, and how many points are still missing when that bin starts.
    Returns (target_bins sorted unique, [(bin, missing)] per boundary)."""
    h = np.asarray(global_hist, dtype=np.int64)
    csum = np.cumsum(h)
    total = int(csum[-1]) if len(csum) else 0
    bounds = []
    for r in range(1, n_ranks):
        target = (total * r + n_ranks - 1) // n_ranks
        b = int(np.searchsorted(csum, target, side="left"))
        b = min(b, len(h) - 1)
        before = int(csum[b] - h[b])
        bounds.append((b, max(target - before, 0)))
    return sorted(set(b for b, _ in bounds)), bounds


def splitters_from_subhist(bounds, targets, sub_global, shift, shift2):
    """Level 2: inside each target bin, the sub-bin boundary where the missing count is met."""
    n_sub = sub_global.shape[1]
    out = []
    for b, missing in bounds:
        t = targets.index(b)
        cs = np.cumsum(sub_global[t].astype(np.int64))
        j = int(np.searchsorted(cs, missing, side="left")) if missing > 0 else -1
        j = min(j, n_sub - 1)
        v = (b << shift) + ((j + 1) << shift2)  # first key of the next rank
        if out and v < out[-1]:
            v = out[-1]
        out.append(min(v, 0xFFFFFFFF))
    return np.asarray(out, dtype=np.uint32)


def send_counts_for(splitters, local_hist, targets, sub_local, shift, shift2, world):
    """Points this rank sends to every destination, from its local (sub-)histograms."""
    sp = splitters.astype(np.int64)
    n_bins = len(local_hist)
    owner = np.searchsorted(sp, np.arange(n_bins, dtype=np.int64) << shift, side="right")
    lh = np.asarray(local_hist, dtype=np.int64).copy()
    counts = np.zeros(world, dtype=np.int64)
    for t, b in enumerate(targets):  # bins that may be cut by a splitter: count by sub-bin
        lh[b] = 0
        n_sub = sub_local.shape[1]
        sub_owner = np.searchsorted(sp, (b << shift) + (np.arange(n_sub, dtype=np.int64) << shift2),
                                    side="right")
        counts += np.bincount(sub_owner, weights=sub_local[t], minlength=world).astype(np.int64)
    counts += np.bincount(owner, weights=lh, minlength=world).astype(np.int64)
    return counts


def hist_shift_for(max_depth):
    key_bits = min(32, 2 * (max(0, min(15, int(max_depth))) + 2))
    return max(0, key_bits - HIST_BITS)


class HostPlan:
    """The plan as plain numpy (CPU tests).  Same fields as bsj_shard_plan."""

    def __init__(self, ghist, sizes, n_ranks, rank, shift, sub_shift, n_sub):
        self.n_ranks, self.rank, self.shift, self.sub_shift, self.n_sub = (n_ranks, rank, shift,
                                                                          sub_shift, n_sub)
        self.sizes = list(sizes)
        self.gid_base = [int(v) for v in np.concatenate([[0], np.cumsum(self.sizes)])]
        self.targets, self.bounds = refine_splitters(ghist, n_ranks, shift)


# ------------------------------------------------------------------------------------------------
# points registered for peer access
# ------------------------------------------------------------------------------------------------
_SYMM = {}


def _group_key(group):
    import torch.distributed as dist

    return tuple(dist.get_process_group_ranks(group if group is not None else dist.group.WORLD))


def _symm_alloc(dev, dtype, capacity, group, tag):
    """A symmetric-memory buffer (same capacity on every rank, mapped into every peer), cached per
    (device, process group, purpose, dtype) and grown collectively."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm

    key = (dev.index, _group_key(group), tag, dtype)
    ent = _SYMM.get(key)
    if ent is None or ent["cap"] < capacity:
        t = symm.empty(int(capacity), dtype=dtype, device=dev)
        hdl = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
        ent = {"cap": int(capacity), "buf": t, "hdl": hdl}
        _SYMM[key] = ent
    return ent


class ShardedPoints:
    """This rank's point columns, placed where every peer GPU can read them (symmetric memory),
    plus what every rank needs to address them: all ranks' sizes and the peer pointers."""

    def __init__(self, x, y, sizes, group, peers_x=None, peers_y=None):
        self.x, self.y, self.sizes, self.group = x, y, list(sizes), group
        self.peers_x, self.peers_y = peers_x, peers_y   # data_ptr of every rank's columns

    @property
    def dtype(self):
        return self.x.dtype

    def __len__(self):
        return int(self.x.shape[0])


def allocate_points(n, dtype, device, group=None):
    """Point columns of length `n` in symmetric memory (uninitialised; fill `.x` / `.y` in place),
    with the sizes and peer pointers of every rank.  A setup step (one host-synchronising
    all-gather): do it once per point set, outside any timed region."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device(device)
    nt = torch.tensor([n], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(nt) for _ in range(world)]
    dist.all_gather(all_n, nt, group=group)
    sizes = [int(v.item()) for v in all_n]
    if sum(sizes) >= 2 ** 32 - 2 ** 14:
        raise ValueError("total number of points must fit uint32 global indices")
    cap = max(max(sizes), 1)
    ex = _symm_alloc(dev, dtype, cap, group, "points_x")
    ey = _symm_alloc(dev, dtype, cap, group, "points_y")
    px = [ex["hdl"].get_buffer(r, (ex["cap"],), dtype).data_ptr() for r in range(world)]
    py = [ey["hdl"].get_buffer(r, (ey["cap"],), dtype).data_ptr() for r in range(world)]
    return ShardedPoints(ex["buf"][:n], ey["buf"][:n], sizes, group, px, py)


def register_points(x, y, group=None):
    """Copy (x, y) into symmetric memory (allocate_points + one device copy).  When the point set
    changes but its size does not, write the new coordinates into `.x` / `.y` in place."""
    import torch.distributed as dist

    pts = allocate_points(x.shape[0], x.dtype, x.device, group)
    pts.x.copy_(x)
    pts.y.copy_(y)
    torch.cuda.synchronize(x.device)
    dist.barrier(group=group)
    return pts


def _host_register(x, y, group=None):
    """CPU stand-in of register_points (gloo tests): sizes only."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64)
    all_n = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(all_n, n, group=group)
    return ShardedPoints(x, y, [int(v.item()) for v in all_n], group)


# ------------------------------------------------------------------------------------------------
# device steps (CUDA through the C ABI); replaced by host callables in the CPU tests
# ------------------------------------------------------------------------------------------------
def cuda_keys_and_histogram(x, y, bbox, scale, max_depth, shift, n_bins):
    """keys (int32 view of uint32) and [histogram | #out-of-box flag | #NaN flag] (int32)."""
    from . import _lib
    from .api import _DTYPE_CODE, _ptr, _stream

    keys = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    ext = torch.zeros(n_bins + 2, dtype=torch.int32, device=x.device)
    flags = torch.zeros(1, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().bsj_point_keys_histogram(
            _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], float(bbox[0]), float(bbox[1]),
            float(bbox[2]), float(bbox[3]), float(scale), int(max_depth), int(shift), _ptr(keys),
            _ptr(ext), n_bins, _ptr(flags), _stream(x.device)))
    ext[n_bins] = flags[0] & 1          # summed by the all-reduce: > 0 means "somewhere"
    ext[n_bins + 1] = (flags[0] >> 1) & 1
    return keys, ext


class CudaPlan:
    """bsj_shard_plan in device memory (an int32 tensor) + a pinned host mirror."""

    def __init__(self, dev):
        import ctypes as C

        from . import _lib

        self.words = C.sizeof(_lib.bsj_shard_plan) // 4
        self.dev_t = torch.zeros(self.words, dtype=torch.int32, device=dev)
        self.host_t = torch.zeros(self.words, dtype=torch.int32).pin_memory()
        self._struct = _lib.bsj_shard_plan
        f = _lib.bsj_shard_plan
        self.off_send_count = f.send_count.offset // 4

    def ptr(self):
        import ctypes as C

        return C.c_void_p(self.dev_t.data_ptr())

    def read_back(self):
        """Device -> pinned host copy + the host synchronisation; returns the ctypes struct."""
        self.host_t.copy_(self.dev_t, non_blocking=True)
        torch.cuda.current_stream(self.dev_t.device).synchronize()
        return self._struct.from_buffer_copy(self.host_t.numpy().tobytes())


def cuda_plan_level1(ghist, sizes, world, rank, shift, sub_shift, n_sub):
    import ctypes as C

    from . import _lib
    from .api import _ptr, _stream

    plan = CudaPlan(ghist.device)
    sz = (C.c_uint32 * world)(*sizes)
    with torch.cuda.device(ghist.device):
        _lib.check(_lib.lib().bsj_shard_plan_level1(
            _ptr(ghist), ghist.shape[0], sz, world, rank, int(shift), int(sub_shift), int(n_sub),
            plan.ptr(), _stream(ghist.device)))
    return plan


def cuda_sub_histogram(keys, plan, world, n_sub):
    from . import _lib
    from .api import _ptr, _stream

    bins = torch.zeros(max(world - 1, 1) * n_sub, dtype=torch.int32, device=keys.device)
    with torch.cuda.device(keys.device):
        _lib.check(_lib.lib().bsj_shard_subhistogram(
            _ptr(keys), keys.shape[0], plan.ptr(), world, int(n_sub), _ptr(bins),
            _stream(keys.device)))
    return bins


def cuda_plan_level2(plan, local_hist, local_sub, global_sub, world):
    """Fills splitters + send counts; returns the send counts as a VIEW of the device plan (the
    all-gather reads it in place)."""
    from . import _lib
    from .api import _ptr, _stream

    with torch.cuda.device(local_hist.device):
        _lib.check(_lib.lib().bsj_shard_plan_level2(
            _ptr(local_hist), local_hist.shape[0], _ptr(local_sub), _ptr(global_sub), plan.ptr(),
            _stream(local_hist.device)))
    return plan.dev_t[plan.off_send_count: plan.off_send_count + world]


def cuda_exchange(keys, plan, counts_matrix, points, world, rank, group):
    """Plan finalisation + fused partition/all-to-all of (key, global id) + device barrier, then
    the one host synchronisation of the exchange (the plan comes back with the receive counts).
    Returns (received keys, received global ids, host plan)."""
    import ctypes as C

    from . import _lib
    from .api import _ptr, _stream

    dev = keys.device
    total = sum(points.sizes)
    use_bulk = 0 if os.environ.get("BSJ_MG_BULK_COPY") == "0" else 1
    cap = int(total / world * 1.25) + 16384
    for attempt in range(2):
        ek = _symm_alloc(dev, torch.int32, cap, group, "recv_key")
        eg = _symm_alloc(dev, torch.int32, cap, group, "recv_gid")
        cap = min(ek["cap"], eg["cap"])
        pk, pg = (C.c_void_p * world)(), (C.c_void_p * world)()
        for d in range(world):
            pk[d] = ek["hdl"].get_buffer(d, (ek["cap"],), torch.int32).data_ptr()
            pg[d] = eg["hdl"].get_buffer(d, (eg["cap"],), torch.int32).data_ptr()
        with torch.cuda.device(dev):
            L = _lib.lib()
            _lib.check(L.bsj_shard_plan_finalize(_ptr(counts_matrix), cap, plan.ptr(), _stream(dev)))
            _lib.check(L.bsj_partition_keys(_ptr(keys), keys.shape[0], plan.ptr(), world, pk, pg,
                                            use_bulk, _stream(dev)))
        # every rank's stores must have landed before anyone reads its receive buffers: a
        # device-side barrier over the symmetric-memory signal pads, enqueued after the partition
        # kernel on the same stream (kernel completion makes its peer stores visible system-wide).
        # BSJ_MG_HOST_BARRIER=1 swaps in a host synchronisation + NCCL barrier.
        if os.environ.get("BSJ_MG_HOST_BARRIER") == "1":
            import torch.distributed as dist

            torch.cuda.synchronize(dev)
            dist.barrier(group=group)
        else:
            ek["hdl"].barrier(channel=0)
        h = plan.read_back()
        if h.status == 0:
            n_recv = int(h.recv_total[rank])
            return ek["buf"][:n_recv], eg["buf"][:n_recv], h
        # a receive total exceeds the buffers (extreme skew: one key holds most points): every
        # rank sees the same plan, so all of them grow the buffers and repeat the exchange
        cap = int(max(h.recv_total[:world]) * 1.1) + 16384
    raise RuntimeError("sharded exchange: receive buffers could not be sized")


def _grid_struct(bbox, scale, max_depth, dtype, has_oob, has_nan):
    """The key geometry exactly as the encode kernel used it (values of the coordinate type)."""
    from . import _lib

    T = np.float32 if dtype == torch.float32 else np.float64
    x0, x1, y0, y1 = (T(v) for v in bbox)
    d = max(0, min(15, int(max_depth)))
    sc = max(T(scale), max(x1 - x0, y1 - y0) / T((1 << d) + 2))
    g = _lib.bsj_grid()
    g.valid, g.max_depth = 1, d
    g.min_x, g.min_y, g.max_x, g.max_y = float(x0), float(y0), float(x1), float(y1)
    g.scale = float(sc)
    g.has_nan, g.has_out_of_bbox = int(bool(has_nan)), int(bool(has_oob))
    return g


def cuda_local_compact(rkeys, rgids, points, flags, polygons, bbox, scale, max_depth, max_size):
    """Sort the received (key, global id) pairs, build the sub-quadtree, filter, refine up to the
    COMPACT PIP result.  Coordinates are read through peer pointers (segments)."""
    import ctypes as C

    from . import _lib, api
    from .api import _DTYPE_CODE, _allocator_for, _ptr, _stream
    from .frame import Frame

    dev = rkeys.device
    n = rkeys.shape[0]
    grid = _grid_struct(bbox, scale, max_depth, points.dtype, flags[0], flags[1])
    with torch.cuda.device(dev):
        alloc = _allocator_for(dev)
        out = _lib.bsj_quadtree()
        _lib.check(_lib.lib().bsj_quadtree_on_keys(
            _ptr(rkeys), _ptr(rgids), n, C.byref(grid), int(max_size), C.byref(alloc.struct),
            _stream(dev), C.byref(out)))
    q = int(out.num_nodes)
    pidx = alloc.take(out.point_indices, n, torch.uint32)
    tree = Frame([
        ("key", alloc.take(out.key, q, torch.uint32)),
        ("level", alloc.take(out.level, q, torch.uint8)),
        ("is_internal_node", alloc.take(out.is_internal_node, q, torch.bool)),
        ("length", alloc.take(out.length, q, torch.uint32)),
        ("offset", alloc.take(out.offset, q, torch.uint32)),
    ])
    g = _lib.bsj_grid()
    C.memmove(C.byref(g), C.byref(out.grid), C.sizeof(_lib.bsj_grid))
    sorted_keys = alloc.take(out.sorted_keys, n, torch.uint32)
    bb = api.polygon_bounding_boxes(polygons)
    pairs = api.join_quadtree_and_bounding_boxes(tree, bb, bbox[0], bbox[1], bbox[2], bbox[3],
                                                 scale, max_depth)
    # refinement with segmented coordinates
    po, ro, vx, vy = api._split_polygons(polygons)
    po = api._as_cuda(po, torch.uint32 if po.dtype != torch.int32 else None)
    ro = api._as_cuda(ro, torch.uint32 if ro.dtype != torch.int32 else None)
    segs = _lib.bsj_coord_segments()
    world = len(points.sizes)
    segs.n_segments = world
    first = 0
    for r in range(world):
        segs.first_id[r] = first
        segs.x[r], segs.y[r] = points.peers_x[r], points.peers_y[r]
        first += points.sizes[r]
    segs.first_id[world] = first
    pp, pq = pairs["bbox_offset"], pairs["quad_offset"]
    tcols = api._quadtree_columns(tree)
    with torch.cuda.device(dev):
        alloc = _allocator_for(dev)
        c = _lib.bsj_pip_compact()
        _lib.check(_lib.lib().bsj_quadtree_point_in_polygon_compact_seg(
            _ptr(pp), _ptr(pq), pp.shape[0], *[_ptr(t) for t in tcols], tcols[0].shape[0],
            _ptr(pidx), C.byref(segs), _DTYPE_CODE[points.dtype], n, _ptr(po), po.shape[0],
            _ptr(ro), ro.shape[0], _ptr(vx), _ptr(vy), vx.shape[0], C.byref(g),
            C.byref(alloc.struct), _stream(dev), C.byref(c)))
    comp = {"pair_poly": pp.view(torch.int32)}
    for name, dt, cnt in api._COMPACT_FIELDS:
        m = int(getattr(c, cnt))
        if name == "mask_words":
            m = max(m, 1) if getattr(c, name) else 0
        t = alloc.take(getattr(c, name), m, dt)
        comp[name] = t.view(torch.int32) if t.dtype == torch.uint32 else t
    del sorted_keys
    return pidx.view(torch.int32), comp, int(c.n_hits)


def cuda_expand(comp, n_hits, position_base, out_poly, out_point):
    from . import api

    api.expand_pip_compact(dict(comp, n_hits=n_hits), position_base, out_poly, out_point)


# ------------------------------------------------------------------------------------------------
def _all_gather_varlen(t, dist, group, sizes=None):
    """all-gather of 1-D tensors of different lengths straight into the merged buffer: one
    broadcast per source rank into its slice (no padding, no concatenation copy)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if sizes is None:
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sz = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sz, n, group=group)
        sizes = [int(s.item()) for s in sz]
    out = torch.empty(sum(sizes), dtype=t.dtype, device=t.device)
    works, o = [], 0
    for r in range(world):
        sl = out[o: o + sizes[r]]
        o += sizes[r]
        if sizes[r] == 0:
            continue
        if r == rank:
            sl.copy_(t)
        src = dist.get_global_rank(group, r) if group is not None else r
        works.append(dist.broadcast(sl, src=src, group=group, async_op=True))
    for w in works:
        w.wait()
    return out, sizes


class _Phases:
    """Phase timing with CUDA events on the launching stream (no host synchronisation while the
    step runs; the events are read after the step's last synchronisation)."""

    def __init__(self, dev, on):
        self.on = bool(on) and dev.type == "cuda"
        self.dev, self.marks = dev, []
        if self.on:
            self.mark("begin")

    def mark(self, name):
        if self.on:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.dev))
            self.marks.append((name, e))

    def finish(self):
        if not self.on:
            return
        torch.cuda.current_stream(self.dev).synchronize()
        LAST_PROFILE.clear()
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            LAST_PROFILE[name] = LAST_PROFILE.get(name, 0.0) + a.elapsed_time(b)


def sharded_quadtree_point_in_polygon(points, polygons, x_min, x_max, y_min, y_max, scale,
                                      max_depth, max_size, group=None, gather_pairs=True,
                                      gather_point_indices=False, steps=None, profile=False):
    """Distributed quadtree PIP join over the ranks of `group`.

    points    : this rank's shard -- a ShardedPoints (register_points: zero-copy, the fast path)
                or an (x, y) pair (registered on the fly: one extra copy + host synchronisation);
                global point id = (sum of earlier ranks' sizes) + i
    polygons  : (part_offset, ring_offset, x, y); only rank 0's content is used (broadcast)
    Returns a dict with
      polygon_index, point_index : the merged pair table (every rank, if gather_pairs) -- or this
                                   rank's rows with GLOBAL point_index if not
      local_point_indices        : global sorted position -> global point id for this rank's key
                                   range (a fresh tensor owned by the result)
      point_indices              : the whole map (if gather_point_indices)
      base, counts               : first global sorted position / number of points per rank
    """
    import torch.distributed as dist

    steps = steps or {}
    register = steps.get("register", register_points)
    keys_hist = steps.get("keys_hist", cuda_keys_and_histogram)
    plan_level1 = steps.get("plan_level1", cuda_plan_level1)
    sub_hist = steps.get("sub_hist", cuda_sub_histogram)
    plan_level2 = steps.get("plan_level2", cuda_plan_level2)
    exchange = steps.get("exchange", cuda_exchange)
    local_compact = steps.get("local_compact", cuda_local_compact)
    expand = steps.get("expand", cuda_expand)

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world > MAX_RANKS:
        raise ValueError("at most %d ranks" % MAX_RANKS)
    if not isinstance(points, ShardedPoints):
        points = register(points[0], points[1], group)
    x, y = points.x, points.y
    dev = x.device
    if sum(points.sizes) >= 2 ** 32 - 2 ** 14:
        raise ValueError("total number of points must fit uint32 global indices")
    bbox = (min(x_min, x_max), max(x_min, x_max), min(y_min, y_max), max(y_min, y_max))
    min_scale = max(bbox[1] - bbox[0], bbox[3] - bbox[2]) / ((1 << max_depth) + 2)
    scale = max(scale, min_scale)

    prof = _Phases(dev, profile)
    # 1. replicate the polygon table -- two NCCL broadcasts: the four sizes, then ONE packed byte
    # buffer (each array padded to 16 bytes) that the receivers slice into typed views
    src0 = dist.get_global_rank(group, 0) if group else 0
    meta = torch.tensor([t.shape[0] for t in polygons], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=src0, group=group)
    sizes = meta.tolist()
    nbytes = [n * t.element_size() for n, t in zip(sizes, polygons)]
    padded = [b + (-b) % 16 for b in nbytes]
    if rank == 0:
        parts = []
        for t, b, pb in zip(polygons, nbytes, padded):
            parts.append(t.contiguous().view(torch.uint8))
            if pb > b:
                parts.append(torch.zeros(pb - b, dtype=torch.uint8, device=dev))
        packed_polys = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device=dev)
    else:
        packed_polys = torch.empty(sum(padded), dtype=torch.uint8, device=dev)
    # the broadcast runs on NCCL's stream while this stream computes keys and histograms; the
    # polygons are first needed by the local join
    poly_work = dist.broadcast(packed_polys, src=src0, group=group, async_op=True) \
        if packed_polys.numel() else None
    polys, o = [], 0
    for t, n, b, pb in zip(polygons, sizes, nbytes, padded):
        polys.append(packed_polys[o: o + b].view(t.dtype) if n else
                     torch.empty(0, dtype=t.dtype, device=dev))
        o += pb
    polys = tuple(polys)
    prof.mark("broadcast_polygons")

    # 2. keys + leading-bit histogram; the plan is derived on the device from the SUMMED histograms
    shift = hist_shift_for(max_depth)
    n_bins = 1 << min(HIST_BITS, 32 - shift) if shift < 32 else 1
    sub_bits = sub_bits_for(world, shift)
    sub_shift, n_sub = shift - sub_bits, 1 << sub_bits
    keys, local_ext = keys_hist(x, y, bbox, scale, max_depth, shift, n_bins)
    prof.mark("keys_hist")
    global_ext = local_ext.clone()
    dist.all_reduce(global_ext, op=dist.ReduceOp.SUM, group=group)
    plan = plan_level1(global_ext[:n_bins], points.sizes, world, rank, shift, sub_shift, n_sub)
    local_sub = sub_hist(keys, plan, world, n_sub)
    global_sub = local_sub.clone()
    dist.all_reduce(global_sub, op=dist.ReduceOp.SUM, group=group)
    send_counts = plan_level2(plan, local_ext[:n_bins], local_sub, global_sub, world)
    counts_matrix = torch.empty(world * world, dtype=send_counts.dtype, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(counts_matrix, send_counts.contiguous(), group=group)
    else:  # gloo (CPU tests)
        dist.all_gather(list(counts_matrix.view(world, world).unbind(0)), send_counts.contiguous(),
                        group=group)
    prof.mark("plan_collectives")
    # 3. fused partition + all-to-all of (key, global id), device barrier, ONE host sync
    rkeys, rgids, hplan = exchange(keys, plan, counts_matrix, points, world, rank, group)
    flags = [int(v) for v in global_ext[n_bins: n_bins + 2].tolist()]
    n_recv = int(rkeys.shape[0])
    prof.mark("partition_exchange")
    if poly_work is not None:
        poly_work.wait()   # stream-ordered for NCCL (no host synchronisation)
    # 4. the single-GPU path on this rank's key range, stopped at the compact result
    pidx_global, comp, n_hits = local_compact(rkeys, rgids, points, flags, polys, bbox, scale,
                                              max_depth, max_size)
    prof.mark("local_join")

    # 5. global indices and merge.  What crosses NVLink is the COMPACT result (per-pair records +
    # ballot words, tens of MB); every rank then expands every rank's rows itself, at HBM speed,
    # with point_index = rank base + local sorted position.
    names = sorted(comp)
    stat = torch.tensor([n_recv, n_hits] + [comp[k].shape[0] for k in names], dtype=torch.int64,
                        device=dev)
    if dev.type == "cuda":  # one collective, one host read
        allstat = torch.empty(world * stat.shape[0], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allstat, stat, group=group)
        stats = allstat.view(world, -1).tolist()
    else:
        stats = [torch.zeros_like(stat) for _ in range(world)]
        dist.all_gather(stats, stat, group=group)
        stats = [s_.tolist() for s_ in stats]
    counts = [s_[0] for s_ in stats]
    hits = [s_[1] for s_ in stats]
    if sum(hits) >= 2 ** 32 and gather_pairs:
        raise ValueError("merged pair table must have fewer than 2^32 rows")
    base = sum(counts[:rank])
    out = {"base": base, "counts": counts, "rows_per_rank": hits, "splitters":
           [int(v) for v in hplan.splitter[: world - 1]] if hasattr(hplan, "splitter") else None}
    prof.mark("stats")
    if gather_pairs:
        # one packed byte buffer per rank -> R broadcasts in total (not R per array)
        def _bytes(t):
            return t.contiguous().view(torch.uint8)

        sizes_b = [[s_[2 + i] * comp[k].element_size() for i, k in enumerate(names)]
                   for s_ in stats]
        pad = [[(-b) % 16 for b in row] for row in sizes_b]  # keep every array 16-byte aligned
        tot_b = [sum(b + p_ for b, p_ in zip(sizes_b[r], pad[r])) for r in range(world)]
        peer_pull = dev.type == "cuda" and os.environ.get("BSJ_MG_PEER_EXPAND") != "0"
        total = sum(hits)
        out_poly = torch.empty(total, dtype=torch.int32, device=dev)
        out_point = torch.empty(total, dtype=torch.int32, device=dev)
        row_of = [sum(hits[:r]) for r in range(world)]

        def expand_block(r, src, o):
            part = {}
            for i, k in enumerate(names):
                nb = sizes_b[r][i]
                part[k] = src[o: o + nb].view(comp[k].dtype)
                o += nb + pad[r][i]
            if hits[r]:
                expand(part, hits[r], sum(counts[:r]), out_poly[row_of[r]: row_of[r] + hits[r]],
                       out_point[row_of[r]: row_of[r] + hits[r]])

        if peer_pull:
            # Fused all-gather + expansion over peer memory: every rank parks its packed compact
            # result in symmetric memory; after a device-side barrier each rank PULLS its peers'
            # blocks with peer-to-peer copies on a side stream (copy engines over NVLink, ring
            # order so that no GPU is everybody's first source) while the main stream already
            # expands this rank's own rows, then each peer's rows as soon as its block has landed.
            # (Expanding straight out of peer memory was measured too: the kernel's record and
            # ballot-word loads at NVLink latency cost more than the copies, 4.7 vs 3.2 ms.)
            cap = max(tot_b)
            cap = cap + cap // 8 + (1 << 20)
            key = (dev.index, _group_key(group), "compact", torch.uint8)
            if key in _SYMM and _SYMM[key]["cap"] >= max(tot_b):
                cap = _SYMM[key]["cap"]  # same decision on every rank: all see the same sizes
            ent = _symm_alloc(dev, torch.uint8, cap, group, "compact")
            o = 0
            for i, k in enumerate(names):
                nb = sizes_b[rank][i]
                if nb:
                    ent["buf"][o: o + nb].copy_(_bytes(comp[k]))
                o += nb + pad[rank][i]
            # every rank's block is complete before anyone reads it.  The buffer is rewritten in
            # the next call only after that call's exchange barrier, which no rank passes before
            # its own copies below have finished (the main stream waits for them): no trailing
            # barrier needed
            ent["hdl"].barrier(channel=0)
            if "peers" not in ent:
                ent["peers"] = [ent["buf"] if r == rank else
                                ent["hdl"].get_buffer(r, (ent["cap"],), torch.uint8)
                                for r in range(world)]
                ent["side"] = torch.cuda.Stream(dev)
            main, side = torch.cuda.current_stream(dev), ent["side"]
            order = [(rank + k) % world for k in range(1, world)]
            landed = {}
            local = {r: torch.empty(tot_b[r], dtype=torch.uint8, device=dev) for r in order}
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                for r in order:
                    if tot_b[r]:
                        local[r].copy_(ent["peers"][r][: tot_b[r]], non_blocking=True)
                    landed[r] = torch.cuda.Event()
                    landed[r].record(side)
            prof.mark("gather_compact")
            expand_block(rank, ent["buf"], 0)
            for r in order:
                main.wait_event(landed[r])
                expand_block(r, local[r], 0)
        else:
            packed = torch.cat([torch.cat([_bytes(comp[k]),
                                           torch.zeros(pad[rank][i], dtype=torch.uint8, device=dev)])
                                for i, k in enumerate(names)]) if names else \
                torch.empty(0, dtype=torch.uint8, device=dev)
            allb, _ = _all_gather_varlen(packed, dist, group, tot_b)
            prof.mark("gather_compact")
            for r in range(world):
                expand_block(r, allb, sum(tot_b[:r]))
        out["polygon_index"], out["point_index"] = out_poly, out_point
        prof.mark("expand_rows")
    else:
        out_poly = torch.empty(n_hits, dtype=torch.int32, device=dev)
        out_point = torch.empty(n_hits, dtype=torch.int32, device=dev)
        if n_hits:
            expand(comp, n_hits, base, out_poly, out_point)
        out["polygon_index"], out["point_index"] = out_poly, out_point
        prof.mark("expand_rows")
    # global sorted position -> global point id for this rank's key range
    out["local_point_indices"] = pidx_global
    if gather_point_indices:
        out["point_indices"], _ = _all_gather_varlen(pidx_global, dist, group, counts)
        prof.mark("gather_point_indices")
    prof.finish()
    return out

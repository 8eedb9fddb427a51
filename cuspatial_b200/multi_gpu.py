"""Multi-GPU quadtree point-in-polygon join: one process per GPU, torch.distributed (NCCL).

The reference is single-GPU; this is the sharded form of the same path:

  1. the polygon table (offsets + vertices) is replicated by an NCCL broadcast from rank 0;
  2. every rank computes the Morton keys of its local points and a histogram of their leading
     16 bits (CUDA, partition.cu); the histograms are summed with ONE all-reduce and every rank
     derives the same R-1 key splitters, placed on histogram-bin boundaries so equal keys never
     straddle two ranks;
  3. every rank stably partitions (x, y, global id) by destination rank (CUDA, partition.cu) and
     the buckets are exchanged with all-to-all (NVLink); received points arrive ordered by global
     id, so the tie order of the reference's stable sort is preserved;
  4. every rank runs the unchanged single-GPU path on its key range with the GLOBAL area of
     interest / scale / depth / max_size;
  5. local results are lifted to global indices -- global sorted position = rank base + local
     position, global point_indices = the received global ids permuted by the local sort -- and
     the (polygon_index, point_index) tables are merged with an all-gather.

The merged pair SET equals a single-GPU run's bit for bit (the per-rank sub-quadtrees are not the
global quadtree, so row order inside the table differs; compare sorted rows).

The three device steps are injected as callables so that the host-side logic (splitters, index
fix-up, merge) is testable on CPU with the gloo backend (tests/test_multi_gpu_cpu.py).
"""
import os
import time

import numpy as np
import torch

HIST_BITS = 13   # 8192 bins: privatised in shared memory by the histogram kernel
LAST_PROFILE = {}


class _Prof:
    """Optional wall-clock phase timing (BSJ_MG_PROFILE=1): synchronises, so never on by default."""

    def __init__(self, dev):
        self.on = os.environ.get("BSJ_MG_PROFILE") == "1" and dev.type == "cuda"
        self.dev, self.t, self.out = dev, None, {}
        if self.on:
            torch.cuda.synchronize(dev)
            self.t = time.perf_counter()

    def mark(self, name):
        if self.on:
            torch.cuda.synchronize(self.dev)
            now = time.perf_counter()
            self.out[name] = self.out.get(name, 0.0) + 1e3 * (now - self.t)
            self.t = now


def choose_splitters(global_hist, n_ranks, shift):
    """R-1 ascending uint32 key splitters on bin boundaries balancing the point counts.
    Rank r owns keys in [splitter[r-1], splitter[r])."""
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


SUB_BITS = 10


def refine_splitters(global_hist, n_ranks, shift):
    """First step of the two-level splitter search: for every rank boundary the first-level bin
    in which the target cumulative count is reached, and how many points are still missing when
    that bin starts.  Returns (target_bins sorted unique, [(bin, missing)] per boundary)."""
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
    """Second step: inside each target bin, the sub-bin boundary where the missing count is met."""
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


def cuda_sub_histogram(keys, shift, targets, shift2, n_sub):
    import ctypes as C

    from . import _lib
    from .api import _ptr, _stream

    bins = torch.zeros(len(targets) * n_sub, dtype=torch.int32, device=keys.device)
    tg = np.ascontiguousarray(targets, dtype=np.uint32)
    with torch.cuda.device(keys.device):
        _lib.check(_lib.lib().bsj_key_subhistogram(
            _ptr(keys), keys.shape[0], int(shift), tg.ctypes.data_as(C.c_void_p), len(targets),
            int(shift2), int(n_sub), _ptr(bins), _stream(keys.device)))
    return bins.to(torch.int64).view(len(targets), n_sub)


def hist_shift_for(max_depth):
    key_bits = min(32, 2 * (max(0, min(15, int(max_depth))) + 2))
    return max(0, key_bits - HIST_BITS)


# ------------------------------------------------------------------------------------------------
# device steps (CUDA through the C ABI); replaced by host callables in the CPU tests
# ------------------------------------------------------------------------------------------------
def cuda_keys_and_histogram(x, y, bbox, scale, max_depth, shift, n_bins):
    import ctypes as C

    from . import _lib
    from .api import _DTYPE_CODE, _ptr, _stream

    keys = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
    bins = torch.zeros(n_bins, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().bsj_point_keys_histogram(
            _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], float(bbox[0]), float(bbox[1]),
            float(bbox[2]), float(bbox[3]), float(scale), int(max_depth), int(shift), _ptr(keys),
            _ptr(bins), n_bins, _stream(x.device)))
    return keys, bins.to(torch.int64)


_SYMM = {}


def _symmetric_buffers(dev, dtype, capacity, group):
    """Receive buffers (x, y, gid) in symmetric memory, mapped into every peer: grown
    collectively (all ranks see the same `capacity`), cached across calls."""
    import torch.distributed._symmetric_memory as symm

    key = (dev.index, dtype)
    ent = _SYMM.get(key)
    if ent is None or ent["cap"] < capacity:
        cap = int(capacity * 1.25) + 4096
        bufs, hdls = [], []
        for dt in (dtype, dtype, torch.int32):
            t = symm.empty(cap, dtype=dt, device=dev)
            hdls.append(symm.rendezvous(t, group if group is not None else
                                        torch.distributed.group.WORLD))
            bufs.append(t)
        ent = {"cap": cap, "bufs": bufs, "hdls": hdls}
        _SYMM[key] = ent
    return ent


def cuda_partition_exchange(keys, x, y, gid_base, splitters, counts_matrix, rank, group):
    """Stable partition by destination rank whose stores go straight into the destination GPUs'
    receive buffers (peer memory over NVLink): partition and all-to-all in ONE kernel.
    counts_matrix[src][dst] = points rank src sends to rank dst (identical on every rank)."""
    import ctypes as C

    import torch.distributed as dist

    from . import _lib
    from .api import _DTYPE_CODE, _ptr, _stream

    R = len(counts_matrix)
    dev = x.device
    recv_tot = [sum(counts_matrix[s][d] for s in range(R)) for d in range(R)]
    ent = _symmetric_buffers(dev, x.dtype, max(recv_tot), group)
    cap = ent["cap"]
    esz = x.element_size()
    ptr_x, ptr_y, ptr_g = (C.c_void_p * R)(), (C.c_void_p * R)(), (C.c_void_p * R)()
    for d in range(R):
        off = sum(counts_matrix[s][d] for s in range(rank))  # where my bucket starts in rank d
        bx = ent["hdls"][0].get_buffer(d, (cap,), x.dtype)
        by = ent["hdls"][1].get_buffer(d, (cap,), x.dtype)
        bg = ent["hdls"][2].get_buffer(d, (cap,), torch.int32)
        ptr_x[d] = bx.data_ptr() + off * esz
        ptr_y[d] = by.data_ptr() + off * esz
        ptr_g[d] = bg.data_ptr() + off * 4
    sp = np.ascontiguousarray(splitters, dtype=np.uint32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bsj_partition_points(
            _ptr(keys), _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], x.shape[0], int(gid_base),
            sp.ctypes.data_as(C.c_void_p), R, ptr_x, ptr_y, ptr_g, _stream(dev)))
    # every rank's stores must have landed before anyone reads its receive buffers: a device-side
    # barrier over the symmetric-memory signal pads, enqueued after the partition kernel on the
    # same stream (kernel completion makes its peer stores visible system-wide) -- no host
    # synchronisation and no NCCL round trip.  BSJ_MG_HOST_BARRIER=1 restores the host barrier.
    # The device barrier is verified at 2 and 4 GPUs; larger groups keep the host barrier (the
    # form verified at 8 GPUs) until it has been re-run there -- BSJ_MG_DEVICE_BARRIER=1 forces it.
    use_device = (R <= 4 or os.environ.get("BSJ_MG_DEVICE_BARRIER") == "1") and \
        os.environ.get("BSJ_MG_HOST_BARRIER") != "1"
    if use_device:
        ent["hdls"][0].barrier(channel=0)
    else:
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
    n_recv = recv_tot[rank]
    rx, ry, rg = (b[:n_recv] for b in ent["bufs"])
    return rx, ry, rg


def cuda_local_compact(x, y, polygons, bbox, scale, max_depth, max_size):
    """The single-GPU path on this rank's points, stopped at the COMPACT PIP result."""
    from . import api

    pidx, tree = api.quadtree_on_points((x, y), bbox[0], bbox[1], bbox[2], bbox[3], scale,
                                        max_depth, max_size)
    bb = api.polygon_bounding_boxes(polygons)
    pairs = api.join_quadtree_and_bounding_boxes(tree, bb, bbox[0], bbox[1], bbox[2], bbox[3],
                                                 scale, max_depth)
    comp = api.quadtree_point_in_polygon_compact(pairs, tree, pidx, (x, y), polygons)
    n_hits = comp.pop("n_hits")
    comp = {k: (v.view(torch.int32) if v.dtype == torch.uint32 else v) for k, v in comp.items()}
    return pidx.view(torch.int32), comp, n_hits


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


def sharded_quadtree_point_in_polygon(points, polygons, x_min, x_max, y_min, y_max, scale,
                                      max_depth, max_size, group=None, gather_pairs=True,
                                      gather_point_indices=False, steps=None):
    """Distributed quadtree PIP join over the ranks of `group`.

    points    : this rank's (x, y) shard; global point id = (sum of earlier ranks' sizes) + i
    polygons  : (part_offset, ring_offset, x, y); only rank 0's content is used (broadcast)
    Returns a dict with
      polygon_index, point_index : the merged pair table (every rank, if gather_pairs) -- or this
                                   rank's rows with GLOBAL point_index if not
      point_indices              : global sorted-position -> global point id map (this rank's key
                                   range, or the whole array if gather_point_indices)
      base, counts               : first global sorted position / number of points per rank
    """
    import torch.distributed as dist

    steps = steps or {}
    keys_hist = steps.get("keys_hist", cuda_keys_and_histogram)
    partition = steps.get("partition")  # host stand-in for the CPU tests
    local_compact = steps.get("local_compact", cuda_local_compact)
    expand = steps.get("expand", cuda_expand)

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    x, y = points
    dev = x.device
    bbox = (min(x_min, x_max), max(x_min, x_max), min(y_min, y_max), max(y_min, y_max))
    min_scale = max(bbox[1] - bbox[0], bbox[3] - bbox[2]) / ((1 << max_depth) + 2)
    scale = max(scale, min_scale)

    prof = _Prof(dev)
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
    if packed_polys.numel():
        dist.broadcast(packed_polys, src=src0, group=group)
    polys, o = [], 0
    for t, n, b, pb in zip(polygons, sizes, nbytes, padded):
        polys.append(packed_polys[o: o + b].view(t.dtype) if n else
                     torch.empty(0, dtype=t.dtype, device=dev))
        o += pb
    polys = tuple(polys)

    prof.mark("broadcast_polygons")
    # 2. keys + leading-bit histogram, one all-reduce, identical splitters everywhere.  The
    # per-rank point counts (global ids are rank-major) ride in `world` extra slots of the same
    # all-reduce, and local + global histogram come back to the host in one copy.
    shift = hist_shift_for(max_depth)
    n_bins = 1 << min(HIST_BITS, 32 - shift) if shift < 32 else 1
    keys, hist = keys_hist(x, y, bbox, scale, max_depth, shift, n_bins)
    prof.mark("keys_hist")
    slots = torch.zeros(world, dtype=hist.dtype, device=dev)
    slots[rank] = x.shape[0]
    both = torch.stack([torch.cat([hist, slots])] * 2)   # row 0 stays local, row 1 is reduced
    dist.all_reduce(both[1], op=dist.ReduceOp.SUM, group=group)
    both_h = both.cpu().numpy()
    local_hist, hist_h = both_h[0, :n_bins].copy(), both_h[1, :n_bins].copy()
    all_n = [int(v) for v in both_h[1, n_bins:]]
    gid_base = sum(all_n[:rank])
    # two-level splitters: a first-level bin can hold a whole cluster, so the boundary is placed
    # inside it with a second histogram of the next SUB_BITS key bits (one more all-reduce)
    sub_hist = steps.get("sub_hist", cuda_sub_histogram)
    shift2 = max(0, shift - SUB_BITS)
    n_sub = 1 << (shift - shift2)
    targets, bounds = refine_splitters(hist_h, world, shift)
    if n_sub > 1 and targets:
        sub = sub_hist(keys, shift, targets, shift2, n_sub)
        sub_both = torch.stack([sub, sub])
        dist.all_reduce(sub_both[1], op=dist.ReduceOp.SUM, group=group)
        sub_h = sub_both.cpu().numpy()
        sub_local = sub_h[0].copy()
        splitters = splitters_from_subhist(bounds, targets, sub_h[1], shift, shift2)
    else:
        targets, sub_local = [], np.zeros((0, 1), dtype=np.int64)
        splitters = choose_splitters(hist_h, world, shift)
    prof.mark("allreduce_splitters")
    # 3. stable partition by destination + all-to-all
    send_counts = send_counts_for(splitters, local_hist, targets, sub_local, shift, shift2, world)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rows = [torch.zeros_like(sc) for _ in range(world)]
    dist.all_gather(rows, sc, group=group)       # also orders this step after every rank's
    counts_matrix = [r_.tolist() for r_ in rows]  # previous step (receive buffers are reused)
    recv_counts = [counts_matrix[s_][rank] for s_ in range(world)]
    n_recv = int(sum(recv_counts))
    if partition is None:
        # fused partition + exchange: the kernel writes into the peers' receive buffers
        rx, ry, rgid = cuda_partition_exchange(keys, x, y, gid_base, splitters, counts_matrix,
                                               rank, group)
        prof.mark("partition_exchange")
    else:
        sx, sy, sgid = partition(keys, x, y, gid_base, splitters, send_counts.tolist())
        prof.mark("partition")
        rx = torch.empty(n_recv, dtype=x.dtype, device=dev)
        ry = torch.empty(n_recv, dtype=y.dtype, device=dev)
        rgid = torch.empty(n_recv, dtype=torch.int32, device=dev)
        for dst, src in ((rx, sx), (ry, sy), (rgid, sgid)):
            dist.all_to_all_single(dst, src, output_split_sizes=recv_counts,
                                   input_split_sizes=send_counts.tolist(), group=group)
        prof.mark("all_to_all")
    # 4. the unchanged single-GPU path on this rank's key range, stopped at the compact result
    pidx_local, comp, n_hits = local_compact(rx, ry, polys, bbox, scale, max_depth, max_size)
    prof.mark("local_join")

    # 5. global indices and merge.  What crosses NVLink is the COMPACT result (per-pair records +
    # ballot words, tens of MB); every rank then expands every rank's rows itself, at HBM speed,
    # with point_index = rank base + local sorted position.
    names = sorted(comp)
    stat = torch.tensor([n_recv, n_hits] + [comp[k].shape[0] for k in names], dtype=torch.int64,
                        device=dev)
    stats = [torch.zeros_like(stat) for _ in range(world)]
    dist.all_gather(stats, stat, group=group)
    stats = [s_.tolist() for s_ in stats]
    counts = [s_[0] for s_ in stats]
    hits = [s_[1] for s_ in stats]
    base = sum(counts[:rank])
    out = {"base": base, "counts": counts, "rows_per_rank": hits}
    prof.mark("global_indices")
    if gather_pairs:
        # one packed byte buffer per rank -> R broadcasts in total (not R per array)
        def _bytes(t):
            return t.contiguous().view(torch.uint8)

        sizes_b = [[s_[2 + i] * comp[k].element_size() for i, k in enumerate(names)]
                   for s_ in stats]
        pad = [[(-b) % 16 for b in row] for row in sizes_b]  # keep every array 16-byte aligned
        packed = torch.cat([torch.cat([_bytes(comp[k]),
                                       torch.zeros(pad[rank][i], dtype=torch.uint8, device=dev)])
                            for i, k in enumerate(names)]) if names else \
            torch.empty(0, dtype=torch.uint8, device=dev)
        tot_b = [sum(b + p_ for b, p_ in zip(sizes_b[r], pad[r])) for r in range(world)]
        allb, _ = _all_gather_varlen(packed, dist, group, tot_b)
        total = sum(hits)
        out_poly = torch.empty(total, dtype=torch.int32, device=dev)
        out_point = torch.empty(total, dtype=torch.int32, device=dev)
        row, boff = 0, 0
        for r in range(world):
            part, o = {}, boff
            for i, k in enumerate(names):
                nb = sizes_b[r][i]
                part[k] = allb[o: o + nb].view(comp[k].dtype)
                o += nb + pad[r][i]
            boff += tot_b[r]
            if hits[r]:
                expand(part, hits[r], sum(counts[:r]), out_poly[row: row + hits[r]],
                       out_point[row: row + hits[r]])
            row += hits[r]
        out["polygon_index"], out["point_index"] = out_poly, out_point
    else:
        out_poly = torch.empty(n_hits, dtype=torch.int32, device=dev)
        out_point = torch.empty(n_hits, dtype=torch.int32, device=dev)
        if n_hits:
            expand(comp, n_hits, base, out_poly, out_point)
        out["polygon_index"], out["point_index"] = out_poly, out_point
    # global sorted position -> global point id, for this rank's key range (lazy: a 100M-element
    # gather that most callers of a join do not need)
    out["local_point_indices"], out["received_global_ids"] = pidx_local, rgid
    if gather_point_indices:
        point_indices = rgid[pidx_local.to(torch.int64)] if n_recv else rgid
        out["point_indices"], _ = _all_gather_varlen(point_indices, dist, group, counts)
    prof.mark("all_gather")
    if prof.on:
        LAST_PROFILE.clear()
        LAST_PROFILE.update(prof.out)
    return out

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
import numpy as np
import torch

HIST_BITS = 16


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


def cuda_partition(keys, x, y, gid_base, splitters, counts):
    import ctypes as C

    from . import _lib
    from .api import _DTYPE_CODE, _ptr, _stream

    n, R = x.shape[0], len(counts)
    base = torch.zeros(R, dtype=torch.int64)
    base[1:] = torch.cumsum(torch.as_tensor(counts[:-1], dtype=torch.int64), 0)
    d_base = base.to(torch.int32).to(x.device)
    ox, oy = torch.empty_like(x), torch.empty_like(y)
    ogid = torch.empty(n, dtype=torch.int32, device=x.device)
    sp = np.ascontiguousarray(splitters, dtype=np.uint32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().bsj_partition_points(
            _ptr(keys), _ptr(x), _ptr(y), _DTYPE_CODE[x.dtype], n, int(gid_base),
            sp.ctypes.data_as(C.c_void_p), R, _ptr(d_base), _ptr(ox), _ptr(oy), _ptr(ogid),
            _stream(x.device)))
    return ox, oy, ogid


def cuda_local_join(x, y, polygons, bbox, scale, max_depth, max_size):
    from . import api

    pidx, tree = api.quadtree_on_points((x, y), bbox[0], bbox[1], bbox[2], bbox[3], scale,
                                        max_depth, max_size)
    bb = api.polygon_bounding_boxes(polygons)
    pairs = api.join_quadtree_and_bounding_boxes(tree, bb, bbox[0], bbox[1], bbox[2], bbox[3],
                                                 scale, max_depth)
    hits = api.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polygons)
    return (pidx.view(torch.int32), hits["polygon_index"].view(torch.int32),
            hits["point_index"].view(torch.int32))


# ------------------------------------------------------------------------------------------------
def _all_gather_varlen(t, dist, group):
    """all-gather of 1-D tensors of different lengths (padded all_gather_into_tensor)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes) if sizes else 0
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = torch.empty(m * world, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m: r * m + sizes[r]] for r in range(world)]), sizes


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
    partition = steps.get("partition", cuda_partition)
    local_join = steps.get("local_join", cuda_local_join)

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    x, y = points
    dev = x.device
    bbox = (min(x_min, x_max), max(x_min, x_max), min(y_min, y_max), max(y_min, y_max))
    min_scale = max(bbox[1] - bbox[0], bbox[3] - bbox[2]) / ((1 << max_depth) + 2)
    scale = max(scale, min_scale)

    # 1. replicate the polygon table (sizes first, then payload) -- NCCL broadcast
    meta = torch.tensor([t.shape[0] for t in polygons], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=dist.get_global_rank(group, 0) if group else 0, group=group)
    polys = []
    for t, n in zip(polygons, meta.tolist()):
        if rank != 0:
            t = torch.empty(n, dtype=t.dtype, device=dev)
        t = t.contiguous()
        view = t.view(torch.int32) if t.dtype == torch.uint32 else t
        dist.broadcast(view, src=dist.get_global_rank(group, 0) if group else 0, group=group)
        polys.append(t)
    polys = tuple(polys)

    # global ids are rank-major
    n_local = torch.tensor([x.shape[0]], dtype=torch.int64, device=dev)
    all_n = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(all_n, n_local, group=group)
    all_n = [int(v.item()) for v in all_n]
    gid_base = sum(all_n[:rank])

    # 2. keys + leading-bit histogram, one all-reduce, identical splitters everywhere
    shift = hist_shift_for(max_depth)
    n_bins = 1 << min(HIST_BITS, 32 - shift) if shift < 32 else 1
    keys, hist = keys_hist(x, y, bbox, scale, max_depth, shift, n_bins)
    local_hist = hist.detach().clone().cpu().numpy()
    dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    hist_h = hist.cpu().numpy()
    splitters = choose_splitters(hist_h, world, shift)

    # 3. stable partition by destination + all-to-all
    bin_owner = np.searchsorted(splitters.astype(np.int64) >> shift,
                                np.arange(n_bins, dtype=np.int64), side="right")
    send_counts = np.bincount(bin_owner, weights=local_hist, minlength=world).astype(np.int64)
    sx, sy, sgid = partition(keys, x, y, gid_base, splitters, send_counts.tolist())
    sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = rc.tolist()
    n_recv = int(sum(recv_counts))
    rx = torch.empty(n_recv, dtype=x.dtype, device=dev)
    ry = torch.empty(n_recv, dtype=y.dtype, device=dev)
    rgid = torch.empty(n_recv, dtype=torch.int32, device=dev)
    for dst, src in ((rx, sx), (ry, sy), (rgid, sgid)):
        dist.all_to_all_single(dst, src, output_split_sizes=recv_counts,
                               input_split_sizes=send_counts.tolist(), group=group)

    # 4. the unchanged single-GPU path on this rank's key range
    pidx_local, poly_idx, pos_local = local_join(rx, ry, polys, bbox, scale, max_depth, max_size)

    # 5. global indices and merge
    cnt = torch.tensor([n_recv], dtype=torch.int64, device=dev)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    counts = [int(c.item()) for c in cnts]
    base = sum(counts[:rank])
    point_indices = rgid[pidx_local.to(torch.int64)] if n_recv else rgid
    pos_global = (pos_local.to(torch.int64) + base).to(torch.int32)
    out = {"base": base, "counts": counts}
    if gather_pairs:
        out["polygon_index"], _ = _all_gather_varlen(poly_idx, dist, group)
        out["point_index"], _ = _all_gather_varlen(pos_global, dist, group)
    else:
        out["polygon_index"], out["point_index"] = poly_idx, pos_global
    if gather_point_indices:
        out["point_indices"], _ = _all_gather_varlen(point_indices, dist, group)
    else:
        out["point_indices"] = point_indices
    return out

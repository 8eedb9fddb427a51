"""CPU test (gloo, world_size 2) of the multi-GPU host logic: splitters, stable exchange, global
index fix-up and merge.  The three device steps are replaced by host callables (numpy + the CPU
oracle); the merged result must equal a single-process oracle run on all points."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _dilate(v):
    v = v.astype(np.uint32) & 0xFFFF
    v = (v | (v << 8)) & 0x00FF00FF
    v = (v | (v << 4)) & 0x0F0F0F0F
    v = (v | (v << 2)) & 0x33333333
    v = (v | (v << 1)) & 0x55555555
    return v


def _host_steps(oracle):
    """Host stand-ins (numpy + the CPU oracle) of the device steps of multi_gpu.py."""
    import torch
    import torch.distributed as dist

    from cuspatial_b200 import multi_gpu as mg

    def keys_hist(x, y, bbox, scale, max_depth, shift, n_bins):
        xn, yn = x.numpy(), y.numpy()
        T = xn.dtype.type
        mnx, mxx, mny, mxy = (T(b) for b in bbox)
        sc = max(T(scale), max(mxx - mnx, mxy - mny) / T((1 << max_depth) + 2))
        oob = (xn < mnx) | (xn > mxx) | (yn < mny) | (yn > mxy)
        ix = ((xn - mnx) / sc).astype(np.uint32) & 0xFFFF
        iy = ((yn - mny) / sc).astype(np.uint32) & 0xFFFF
        k = (_dilate(iy) << 1) | _dilate(ix)
        k[oob] = (1 << (2 * max_depth)) - 1
        ext = np.zeros(n_bins + 2, dtype=np.int32)
        ext[:n_bins] = np.bincount(k >> shift, minlength=n_bins)
        ext[n_bins] = int(oob.any())
        ext[n_bins + 1] = int(np.isnan(xn).any() or np.isnan(yn).any())
        return torch.from_numpy(k.astype(np.int32)), torch.from_numpy(ext)

    def plan_level1(ghist, sizes, world, rank, shift, sub_shift, n_sub):
        return mg.HostPlan(ghist.numpy().view(np.uint32).astype(np.int64), sizes, world, rank,
                           shift, sub_shift, n_sub)

    def sub_hist(keys, plan, world, n_sub):
        k = keys.numpy().view(np.uint32).astype(np.int64)
        out = np.zeros((max(world - 1, 1), n_sub), dtype=np.int32)
        for t, b in enumerate(plan.targets):
            sel = k[(k >> plan.shift) == b]
            out[t] = np.bincount((sel >> plan.sub_shift) & (n_sub - 1), minlength=n_sub)
        return torch.from_numpy(out.reshape(-1))

    def plan_level2(plan, local_hist, local_sub, global_sub, world):
        nt = len(plan.targets)
        gs = global_sub.numpy().view(np.uint32).astype(np.int64).reshape(-1, plan.n_sub)[:nt]
        ls = local_sub.numpy().view(np.uint32).astype(np.int64).reshape(-1, plan.n_sub)[:nt]
        plan.splitter = mg.splitters_from_subhist(plan.bounds, plan.targets, gs, plan.shift,
                                                  plan.sub_shift)
        lh = local_hist.numpy().view(np.uint32).astype(np.int64)
        plan.send_count = mg.send_counts_for(plan.splitter, lh, plan.targets, ls, plan.shift,
                                             plan.sub_shift, world)
        return torch.from_numpy(plan.send_count.astype(np.int32))

    def exchange(keys, plan, counts_matrix, points, world, rank, group):
        k = keys.numpy().view(np.uint32)
        dest = np.searchsorted(plan.splitter.astype(np.int64), k.astype(np.int64), side="right")
        order = np.argsort(dest, kind="stable")
        gid = (plan.gid_base[rank] + np.arange(len(k))).astype(np.int32)
        M = counts_matrix.numpy().reshape(world, world)
        assert np.bincount(dest, minlength=world).tolist() == M[rank].tolist()
        send, recv = M[rank].tolist(), M[:, rank].tolist()
        rk = torch.empty(sum(recv), dtype=torch.int32)
        rg = torch.empty(sum(recv), dtype=torch.int32)
        for dst_t, src in ((rk, k.view(np.int32)[order]), (rg, gid[order])):
            dist.all_to_all_single(dst_t, torch.from_numpy(np.ascontiguousarray(src)),
                                   output_split_sizes=recv, input_split_sizes=send, group=group)
        return rk, rg, plan

    def local_compact(rkeys, rgids, points, flags, polys, bbox, scale, max_depth, max_size):
        # coordinates of the received ids: the CPU stand-in simply gathers all points
        world = len(points.sizes)
        cap = max(points.sizes)
        parts = []
        for src in (points.x, points.y):
            pad = torch.zeros(cap, dtype=src.dtype)
            pad[: src.shape[0]] = src
            got = [torch.zeros_like(pad) for _ in range(world)]
            dist.all_gather(got, pad, group=points.group)
            parts.append(np.concatenate([g.numpy()[:n] for g, n in zip(got, points.sizes)]))
        gid = rgids.numpy().view(np.uint32).astype(np.int64)
        xn, yn = parts[0][gid], parts[1][gid]
        po, ro, vx, vy = (p.numpy() for p in polys)
        po, ro = po.view(np.uint32), ro.view(np.uint32)
        t = oracle.quadtree_on_points(xn, yn, bbox[0], bbox[1], bbox[2], bbox[3], scale,
                                      max_depth, max_size)
        bb = oracle.polygon_bounding_boxes(po, ro, vx, vy)
        pairs = oracle.join_quadtree_and_bounding_boxes(t, *bb, bbox[0], bbox[2], scale, max_depth)
        hp, hq = oracle.quadtree_point_in_polygon(pairs[0], pairs[1], t, t["point_indices"], xn,
                                                  yn, po, ro, vx, vy)
        comp = {"poly": torch.from_numpy(hp.view(np.int32).copy()),
                "pos": torch.from_numpy(hq.view(np.int32).copy())}
        pidx_global = gid[t["point_indices"].astype(np.int64)].astype(np.uint32)
        return torch.from_numpy(pidx_global.view(np.int32).copy()), comp, len(hp)

    def expand(comp, n_hits, position_base, out_poly, out_point):
        out_poly.copy_(comp["poly"])
        out_point.copy_((comp["pos"].to(torch.int64) + position_base).to(torch.int32))

    return {"register": mg._host_register, "keys_hist": keys_hist, "plan_level1": plan_level1,
            "sub_hist": sub_hist, "plan_level2": plan_level2, "exchange": exchange,
            "local_compact": local_compact, "expand": expand}


def _worker(rank, world, port, kind, dtype_name, q, gather=True):
    import torch
    import torch.distributed as dist

    from cuspatial_b200 import multi_gpu as mg
    from oracle import hostlib
    from util import make_case

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dtype = np.dtype(dtype_name).type
        c = make_case(30000, 25, 10, kind, dtype, seed=77, oob=20, dups=200, median_vertices=24)
        n = len(c["x"])
        lo, hi = rank * n // world, (rank + 1) * n // world
        x, y = torch.from_numpy(c["x"][lo:hi].copy()), torch.from_numpy(c["y"][lo:hi].copy())
        polys = tuple(torch.from_numpy(a.view(np.int32).copy() if a.dtype == np.uint32 else a)
                      for a in (c["po"], c["ro"], c["vx"], c["vy"]))
        if rank != 0:  # only rank 0's polygon content may be used
            polys = tuple(torch.zeros_like(p) for p in polys)
        ext = c["ext"]
        out = mg.sharded_quadtree_point_in_polygon(
            (x, y), polys, ext[0], ext[1], ext[2], ext[3], c["scale"], c["depth"], 32,
            gather_pairs=gather, gather_point_indices=True, steps=_host_steps(hostlib.oracle()))
        q.put((rank, out["polygon_index"].numpy(), out["point_index"].numpy(),
               out["point_indices"].numpy(), out["counts"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind,dtype_name", [("u", "float64"), ("c", "float64"), ("c", "float32")])
def test_sharded_join_world2_equals_single_process_oracle(oracle_lib, kind, dtype_name):
    import torch.multiprocessing as mp

    from util import make_case, run_host

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, dtype_name, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    dtype = np.dtype(dtype_name).type
    c = make_case(30000, 25, 10, kind, dtype, seed=77, oob=20, dups=200, median_vertices=24)
    ref = run_host(oracle_lib, c, 32)
    want = np.stack([ref["hits"][0].astype(np.int64), ref["hits"][1].astype(np.int64)], 1)
    want = want[np.lexsort((want[:, 1], want[:, 0]))]
    for rank, hp, hq, pidx, counts in results:
        got = np.stack([hp.view(np.uint32).astype(np.int64), hq.view(np.uint32).astype(np.int64)], 1)
        got = got[np.lexsort((got[:, 1], got[:, 0]))]
        np.testing.assert_array_equal(got, want)               # same pair set on every rank
        np.testing.assert_array_equal(pidx.view(np.uint32), ref["tree"]["point_indices"])
        # two-level splitters: the ranks are balanced to within a sub-bin even for clustered data
        assert sum(counts) == len(c["x"]) and min(counts) > 0.9 * len(c["x"]) / world


def test_sharded_join_world3_partitioned_rows_union_equals_oracle(oracle_lib):
    """Odd world size and gather_pairs=False: every rank keeps only the rows of its own key range
    (with GLOBAL point_index); their disjoint union is the single-process pair set."""
    import torch.multiprocessing as mp

    from util import make_case, run_host

    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, "c", "float64", q, False))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = make_case(30000, 25, 10, "c", np.float64, seed=77, oob=20, dups=200, median_vertices=24)
    ref = run_host(oracle_lib, c, 32)
    want = np.stack([ref["hits"][0].astype(np.int64), ref["hits"][1].astype(np.int64)], 1)
    want = want[np.lexsort((want[:, 1], want[:, 0]))]
    parts = [np.stack([hp.view(np.uint32).astype(np.int64), hq.view(np.uint32).astype(np.int64)], 1)
             for _, hp, hq, _, _ in results]
    assert sum(len(p) for p in parts) == len(want)            # disjoint: no row twice
    got = np.concatenate(parts)
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    np.testing.assert_array_equal(got, want)
    for _, _, _, pidx, counts in results:
        np.testing.assert_array_equal(pidx.view(np.uint32), ref["tree"]["point_indices"])
        assert sum(counts) == len(c["x"]) and min(counts) > 0.85 * len(c["x"]) / world


def test_two_level_splitters_balance_a_heavy_bin():
    from cuspatial_b200.multi_gpu import (refine_splitters, send_counts_for,
                                          splitters_from_subhist)

    rng = np.random.default_rng(1)
    shift, shift2, n_sub = 19, 9, 1024  # sub-histogram covers key bits [9, 19)
    keys = np.concatenate([rng.integers(0, 1 << 30, 50_000),
                           (777 << shift) + rng.integers(0, 1 << shift, 150_000)])  # one heavy bin
    hist = np.bincount(keys >> shift, minlength=1 << 11)
    for R in (2, 4, 8):
        targets, bounds = refine_splitters(hist, R, shift)
        sub = np.stack([np.bincount((keys[(keys >> shift) == b] >> shift2) & (n_sub - 1),
                                    minlength=n_sub) for b in targets])
        sp = splitters_from_subhist(bounds, targets, sub, shift, shift2)
        dest = np.searchsorted(sp.astype(np.int64), keys, side="right")
        loads = np.bincount(dest, minlength=R)
        assert loads.max() < 1.02 * len(keys) / R + 400, (R, loads)
        counts = send_counts_for(sp, hist, targets, sub, shift, shift2, R)
        np.testing.assert_array_equal(counts, loads)


def test_choose_splitters_balances_and_stays_on_bin_boundaries():
    from cuspatial_b200.multi_gpu import choose_splitters

    rng = np.random.default_rng(0)
    hist = rng.integers(0, 1000, size=1 << 16)
    for R in (2, 4, 8):
        sp = choose_splitters(hist, R, 14).astype(np.int64)
        assert len(sp) == R - 1 and np.all(np.diff(sp) >= 0) and np.all(sp % (1 << 14) == 0)
        owner = np.searchsorted(sp >> 14, np.arange(1 << 16), side="right")
        loads = np.bincount(owner, weights=hist, minlength=R)
        assert loads.max() < 1.05 * hist.sum() / R

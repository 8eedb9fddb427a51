#!/usr/bin/env python
"""Parity check of the sharded join on real GPUs (launch with torchrun, one rank per GPU):
the merged pair set and the global point_indices must equal a single-process CPU-oracle run.
Run by tests/test_multi_gpu_gpu.py (pytest -m gpu) with 1 rank and with every visible GPU.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi_gpu_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cuspatial_b200 import multi_gpu as mg  # noqa: E402
from util import make_case, run_host  # noqa: E402


def _sorted_rows(a, b):
    rows = np.stack([a.astype(np.int64), b.astype(np.int64)], 1)
    return rows[np.lexsort((rows[:, 1], rows[:, 0]))]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = (("u", np.float64, 400_000, None), ("c", np.float64, 600_000, None),
             ("c", np.float32, 300_000, None),
             ("c", np.float64, 200_000, (-74.3, -73.6, 40.4, 41.0)))
    for kind, dtype, n, extent in cases:
        c = make_case(n, 60, 15, kind, dtype, seed=5 + n, oob=100, dups=500, median_vertices=50,
                      extent=extent)
        n = len(c["x"])
        # uneven shards (rank r gets a share proportional to r + 1)
        w = np.cumsum([0] + [r + 1 for r in range(world)]) / (world * (world + 1) / 2)
        lo, hi = int(w[rank] * n), int(w[rank + 1] * n)
        x = torch.as_tensor(c["x"][lo:hi], device=dev)
        y = torch.as_tensor(c["y"][lo:hi], device=dev)
        polys = tuple(torch.as_tensor(a, device=dev) for a in (c["po"], c["ro"], c["vx"], c["vy"]))
        if rank != 0:
            polys = tuple(torch.zeros_like(p) for p in polys)
        ext = c["ext"]
        pts = mg.register_points(x, y)                   # the zero-copy path, reused below
        out = mg.sharded_quadtree_point_in_polygon(pts, polys, ext[0], ext[1], ext[2], ext[3],
                                                   c["scale"], c["depth"], 128, gather_pairs=True,
                                                   gather_point_indices=True)
        part = mg.sharded_quadtree_point_in_polygon((x, y), polys, ext[0], ext[1], ext[2], ext[3],
                                                    c["scale"], c["depth"], 128,
                                                    gather_pairs=False)
        # partitioned rows of all ranks, gathered for the check only
        pp, _ = mg._all_gather_varlen(part["polygon_index"], dist, None)
        pq, _ = mg._all_gather_varlen(part["point_index"], dist, None)
        if rank == 0:
            from oracle import hostlib

            ref = run_host(hostlib.oracle(), c, 128)
            want = _sorted_rows(ref["hits"][0], ref["hits"][1])
            got = _sorted_rows(out["polygon_index"].cpu().numpy().view(np.uint32),
                               out["point_index"].cpu().numpy().view(np.uint32))
            same = got.shape == want.shape and np.array_equal(got, want)
            got2 = _sorted_rows(pp.cpu().numpy().view(np.uint32), pq.cpu().numpy().view(np.uint32))
            same2 = got2.shape == want.shape and np.array_equal(got2, want)
            pi = out["point_indices"].cpu().numpy().view(np.uint32)
            same_pi = np.array_equal(pi, ref["tree"]["point_indices"])
            counts = out["counts"]
            balanced = max(counts) < 1.1 * n / world + 4096
            print("case", kind, dtype.__name__, n, "world", world, "pairs", len(want),
                  "merged set equal:", same, "partitioned union equal:", same2,
                  "point_indices equal:", same_pi, "counts", counts, flush=True)
            ok = ok and same and same2 and same_pi and balanced
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

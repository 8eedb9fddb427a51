#!/usr/bin/env python
"""Parity check of the sharded join on real GPUs (launch with torchrun, one rank per GPU):
the merged pair set and the global point_indices must equal a single-process CPU-oracle run.

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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for kind, dtype, n in (("u", np.float64, 400_000), ("c", np.float64, 600_000),
                           ("c", np.float32, 300_000)):
        c = make_case(n, 60, 15, kind, dtype, seed=5 + n, oob=100, dups=500, median_vertices=50)
        lo, hi = rank * n // world, (rank + 1) * n // world
        x = torch.as_tensor(c["x"][lo:hi], device=dev)
        y = torch.as_tensor(c["y"][lo:hi], device=dev)
        polys = tuple(torch.as_tensor(a, device=dev) for a in (c["po"], c["ro"], c["vx"], c["vy"]))
        if rank != 0:
            polys = tuple(torch.zeros_like(p) for p in polys)
        ext = c["ext"]
        out = mg.sharded_quadtree_point_in_polygon((x, y), polys, ext[0], ext[1], ext[2], ext[3],
                                                   c["scale"], c["depth"], 128, gather_pairs=True,
                                                   gather_point_indices=True)
        if rank == 0:
            from oracle import hostlib

            ref = run_host(hostlib.oracle(), c, 128)
            want = np.stack([ref["hits"][0].astype(np.int64), ref["hits"][1].astype(np.int64)], 1)
            want = want[np.lexsort((want[:, 1], want[:, 0]))]
            hp = out["polygon_index"].cpu().numpy().view(np.uint32).astype(np.int64)
            hq = out["point_index"].cpu().numpy().view(np.uint32).astype(np.int64)
            got = np.stack([hp, hq], 1)
            got = got[np.lexsort((got[:, 1], got[:, 0]))]
            same = got.shape == want.shape and np.array_equal(got, want)
            pi = out["point_indices"].cpu().numpy().view(np.uint32)
            same_pi = np.array_equal(pi, ref["tree"]["point_indices"])
            print("case", kind, dtype.__name__, n, "pairs", len(want), "pair set equal:", same,
                  "point_indices equal:", same_pi, "counts", out["counts"], flush=True)
            ok = ok and same and same_pi
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED", flush=True)
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

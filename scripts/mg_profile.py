"""Per-rank phase and kernel-stage profile of the sharded join (torchrun script, NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/mg_profile.py [--workload config4|configs1] [--points TOTAL]

Every rank prints its own phases (CUDA events around the collectives) and the library's stage
times of the local join, so that skew between ranks is visible (bench.py reports the max).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

import bench
from cuspatial_b200 import _lib
from cuspatial_b200 import multi_gpu as mg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config4")
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--gather", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    w = bench.WORKLOADS[a.workload]
    tdt = torch.float64 if w["dtype"] == "f64" else torch.float32
    total = a.points or w["points"]
    per = total // world if w["scaling"] == "strong" else total
    polys_np, ext, scale = bench.make_polygons(w["n_poly"])
    if tdt == torch.float32:
        polys_np = (polys_np[0], polys_np[1], polys_np[2].astype("float32"),
                    polys_np[3].astype("float32"))
    polys = tuple(torch.as_tensor(p, device=dev) for p in polys_np)
    pts = mg.allocate_points(per, tdt, dev)
    bench.gen_points(w["kind"], per, ext, bench.SEED + rank, tdt, dev, out=(pts.x, pts.y))
    torch.cuda.synchronize(dev)
    dist.barrier()

    def step(profile=False):
        return mg.sharded_quadtree_point_in_polygon(
            pts, polys, ext[0], ext[1], ext[2], ext[3], scale, bench.MAX_DEPTH, bench.MAX_SIZE,
            gather_pairs=bool(a.gather), profile=profile)

    for _ in range(2):
        out = step()
        del out
    torch.cuda.synchronize(dev)
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
        del out
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / a.steps
    dist.barrier()
    out = step(profile=True)
    rows = int(out["polygon_index"].shape[0])
    counts = out["counts"]
    del out
    phases = {k: round(v, 3) for k, v in mg.LAST_PROFILE.items()}
    _lib.set_profiling(True)
    _lib.get_profile()
    out = step()
    del out
    torch.cuda.synchronize(dev)
    stages = {}
    for name, t in _lib.get_profile():
        stages[name] = round(stages.get(name, 0.0) + t, 3)
    _lib.set_profiling(False)
    for r in range(world):
        dist.barrier()
        if r == rank:
            print(json.dumps({"rank": rank, "ms_per_step": round(ms, 3), "rows": rows,
                              "points_here": counts[rank], "phases": phases,
                              "stages": stages}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

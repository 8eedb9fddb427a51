"""Quick single-GPU timing of the join on configs[1] (or --workload config4) with the library's
stage profile: python scripts/quick_join.py [--points N] [--workload configs1|config4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import cuspatial_b200 as cs
from cuspatial_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="configs1")
ap.add_argument("--points", type=int, default=0)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
w = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
polys_np, ext, scale = bench.make_polygons(w["n_poly"])
tdt = torch.float64 if w["dtype"] == "f64" else torch.float32
if tdt == torch.float32:
    polys_np = (polys_np[0], polys_np[1], polys_np[2].astype("float32"), polys_np[3].astype("float32"))
polys = tuple(torch.as_tensor(p, device=dev) for p in polys_np)
n = a.points or w["points"]
x, y = bench.gen_points(w["kind"], n, ext, bench.SEED, tdt, dev)
for _ in range(2):
    r = bench.join_step(cs, x, y, polys, ext, scale)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    r = bench.join_step(cs, x, y, polys, ext, scale)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
_lib.set_profiling(True)
_lib.get_profile()
for _ in range(a.steps):
    r = bench.join_step(cs, x, y, polys, ext, scale)
torch.cuda.synchronize()
st = {}
for k, v in _lib.get_profile():
    st[k] = st.get(k, 0.0) + v / a.steps
print(json.dumps({"nodes": len(r[1]), "pairs": len(r[2]), "hits": len(r[3]), "ms_per_step": round(ms, 4), "stages": {k: round(v, 4) for k, v in st.items()}}))

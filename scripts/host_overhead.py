"""Host-side cost of the three API calls: with a small point set the GPU work is negligible and
the wall time per call is launch / allocation / synchronisation / Python overhead.
    python scripts/host_overhead.py [n_points]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import cuspatial_b200 as cs
from cuspatial_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dev = torch.device("cuda", 0)
polys_np, ext, scale = bench.make_polygons(263)
polys = tuple(torch.as_tensor(p, device=dev) for p in polys_np)
x, y = bench.gen_points("uniform", n, ext, bench.SEED, torch.float64, dev)


def once():
    t = [time.perf_counter()]
    bb = cs.polygon_bounding_boxes(polys); torch.cuda.synchronize(); t.append(time.perf_counter())
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], scale, 15, 512)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3], scale, 15)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polys)
    torch.cuda.synchronize(); t.append(time.perf_counter())
    return [1e6 * (b - a) for a, b in zip(t[:-1], t[1:])]


for _ in range(5):
    once()
l0 = _lib.kernel_launch_count()
rows = [once() for _ in range(50)]
launches = (_lib.kernel_launch_count() - l0) / 50
med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(4)]
print("n=%d  median wall us per call: bbox %.0f  quadtree %.0f  join %.0f  pip %.0f  total %.0f"
      "  (%.0f kernel launches per step)" % (n, *med, sum(med), launches))
# raw driver-call costs on this host
s = torch.cuda.current_stream().cuda_stream
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
p = ctypes.c_void_p()
t0 = time.perf_counter()
for _ in range(2000):
    rt.cudaMallocAsync(ctypes.byref(p), ctypes.c_size_t(1 << 20), ctypes.c_void_p(s))
    rt.cudaFreeAsync(p, ctypes.c_void_p(s))
torch.cuda.synchronize()
print("cudaMallocAsync+cudaFreeAsync pair: %.2f us" % (1e6 * (time.perf_counter() - t0) / 2000))
t0 = time.perf_counter()
for _ in range(2000):
    torch.empty(1 << 20, dtype=torch.uint8, device=dev)
print("torch.empty: %.2f us" % (1e6 * (time.perf_counter() - t0) / 2000))
z = torch.zeros(1, device=dev)
t0 = time.perf_counter()
for _ in range(2000):
    z.add_(1)
torch.cuda.synchronize()
print("tiny torch kernel launch: %.2f us" % (1e6 * (time.perf_counter() - t0) / 2000))
t0 = time.perf_counter()
for _ in range(500):
    z.add_(1)
    torch.cuda.synchronize()
print("launch + synchronize: %.2f us" % (1e6 * (time.perf_counter() - t0) / 500))

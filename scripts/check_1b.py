#!/usr/bin/env python
"""Maximum-size run on one B200: 1 G uniform fp64 points x 263 polygons (BASELINE.json's target
size on a single GPU), size-independent properties only (no CPU checker reaches this size).

  python scripts/check_1b.py [n_points] [n_polygons]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cuspatial_b200 as cs  # noqa: E402
from cuspatial_b200 import datagen as D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
n_poly = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_POLY
dev = torch.device("cuda", 0)
if n_poly == bench.N_POLY:
    (po, ro, vx, vy), ext, scale = bench.make_polygons()
else:
    po, ro, vx, vy = D.taxi_zone_like_polygons(n_poly, seed=bench.SEED)
    ext = D.polygon_extent(vx, vy)
    scale = D.quadtree_params(ext, bench.MAX_DEPTH)
polys = tuple(torch.as_tensor(a, device=dev) for a in (po, ro, vx, vy))
x, y = D.uniform_points_torch(n, ext, bench.SEED, torch.float64, dev)
bb = cs.polygon_bounding_boxes(polys)
for it in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], scale, 15, 512)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3], scale, 15)
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polys)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if it == 0:
        del pidx, tree, pairs, hits
print("n=%d polygons=%d (%d vertices)  step %.1f ms  %.2f G points/s"
      % (n, n_poly, len(vx), dt * 1e3, n / dt / 1e9))
keys = tree._sorted_keys
ok_sorted = bool((keys[1:].view(torch.int32) >= keys[:-1].view(torch.int32)).all())  # keys < 2^31
leaf = ~tree["is_internal_node"]
ok_cover = int(tree["length"].to(torch.int64)[leaf].sum()) == n
s = int(pidx.view(torch.int32).to(torch.int64).sum()) if n < 2**31 else None
ok_perm = s == n * (n - 1) // 2 if s is not None else None
h = len(hits)
print("nodes %d pairs %d hits %d (%.5f per point)" % (len(tree), len(pairs), h, h / n))
# uniform points: the 100 M x 263 run has 1.43444 hits per point; other polygon sets: just sane
ok_hits = abs(h / n - 1.43444) < 2e-3 if n_poly == bench.N_POLY else 0.5 < h / n < 4.0
mx = int(hits["point_index"].view(torch.int32).max())
print("sorted keys:", ok_sorted, " leaves cover N:", ok_cover, " index sum:", ok_perm,
      " hit ratio:", ok_hits, " max point_index < n:", mx < n)
print("peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
print("CHECK_1B", "OK" if (ok_sorted and ok_cover and ok_perm in (True, None) and ok_hits and mx < n)
      else "FAILED")

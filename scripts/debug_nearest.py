"""Debug helper: rows where the product and the reference CUDA build pick different linestrings."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import make_case, make_linestrings, run_gpu_nearest, run_ref_cuda_nearest  # noqa: E402

from oracle import hostlib  # noqa: E402

dtype = np.float32
c = make_case(200_000, 5, 15, "u", dtype, seed=43)
lines = make_linestrings(30, c["ext"], 19, dtype, median_vertices=30)
w = c["ext"][1] - c["ext"][0] + c["ext"][3] - c["ext"][2]
a = run_gpu_nearest(c, lines, 256, w)
b = run_ref_cuda_nearest(c, lines, 256, w)
bad = np.nonzero((a["nearest"][1] != b["nearest"][1]) | (a["nearest"][2] != b["nearest"][2]))[0]
print("mismatching rows:", bad)
orc = hostlib.oracle()
lo, lx, ly = lines
for pos in bad[:5]:
    pid = a["tree"]["point_indices"][pos]
    print("pos", pos, "pid", pid, "xy", repr(c["x"][pid]), repr(c["y"][pid]))
    print(" ours", a["nearest"][1][pos], repr(a["nearest"][2][pos]), " ref", b["nearest"][1][pos],
          repr(b["nearest"][2][pos]))
    for line in sorted({int(a["nearest"][1][pos]), int(b["nearest"][1][pos])}):
        # distance of this point to this line alone, through the oracle on a 1-point problem
        t = {"key": np.zeros(1, np.uint32), "level": np.zeros(1, np.uint8),
             "is_internal_node": np.zeros(1, np.uint8), "length": np.ones(1, np.uint32),
             "offset": np.zeros(1, np.uint32)}
        r = orc.quadtree_point_to_nearest_linestring(
            np.array([line], np.uint32), np.zeros(1, np.uint32), t, np.zeros(1, np.uint32),
            c["x"][pid:pid + 1], c["y"][pid:pid + 1], lo, lx, ly)
        print("  oracle distance to line", line, repr(r[2][0]), "verts", lo[line], lo[line + 1])

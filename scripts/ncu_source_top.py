#!/usr/bin/env python
"""Top source lines of one kernel by executed warp instructions, from an ncu report captured with
--import-source on (source page, CUDA view).

  python scripts/ncu_source_top.py gpurun_out/prof.ncu-rep <kernel substring> [launch index] [n]
"""
import csv
import io
import subprocess
import sys


def _i(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    # the report is a sequence of per-kernel blocks: "Kernel Name",... then per-file tables
    blocks, cur, fpath = [], None, None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1]
        elif row[0] == "Function Name":
            # one block per (file, kernel launch); launches of one kernel repeat in capture order
            cur = {"name": row[1], "rows": [], "hdr": None, "file": fpath}
            blocks.append(cur)
        elif cur is None:
            continue
        elif row[0] == "Line No":
            cur["hdr"] = row
        elif cur["hdr"] and row[0].isdigit():
            cur["rows"].append((cur["file"], row))
    sel = [b for b in blocks if pat in b["name"]]
    if not sel:
        print("kernels:", sorted(set(b["name"][:80] for b in blocks)))
        return
    # merge the per-file blocks of the chosen launch (a launch = a run of consecutive blocks)
    launches, prev = [], None
    for b in sel:
        if prev is None or (b["file"] in prev["files"]):
            prev = {"name": b["name"], "files": set(), "rows": [], "hdr": b["hdr"]}
            launches.append(prev)
        prev["files"].add(b["file"])
        prev["rows"] += b["rows"]
        prev["hdr"] = prev["hdr"] or b["hdr"]
    b = launches[min(which, len(launches) - 1)]
    h = b["hdr"]
    ci, cs, ct = h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
    total = sum(_i(r[ci]) for _, r in b["rows"])
    tots = sum(_i(r[cs]) for _, r in b["rows"])
    print("%s\n  %d warp instructions, %d samples" % (b["name"][:100], total, tots))
    import os
    key = cs if os.environ.get("BY_SAMPLES") else ci
    rows = sorted(b["rows"], key=lambda fr: -_i(fr[1][key]))[:top]
    for f, r in rows:
        print("%5.1f%% inst %5.1f%% smp thr %4s  %s:%s  %s" % (
            100.0 * _i(r[ci]) / max(total, 1), 100.0 * _i(r[cs]) / max(tots, 1), r[ct],
            f.split("/")[-1], r[0], r[1].strip()[:100]))


if __name__ == "__main__":
    main()

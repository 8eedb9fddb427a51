#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the tracked text summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_r1_full.ncu-rep profiles/r1_ncu_full.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "derived__memory_l1_wavefronts_shared_excessive",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "dram__sectors_read.sum",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r".*::", "", name).replace("void ", "")


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e6 if d["Metric Unit"] == "ns" else v / 1e3 if d["Metric Unit"].startswith("us") else v
        agg.setdefault(short(d["Kernel Name"]), []).append(v)
        order.append((short(d["Kernel Name"]), d["Grid Size"], v))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Source: `%s` (cold-cache, serialised launches: compare SHARES, not absolutes).\n"
                "Includes the torch kernels that generate the synthetic input.\n\n" % src)
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.3f | %.4f | %.1f%% |\n" % (k[:60], len(v), sum(v),
                                                               sum(v) / len(v), 100 * sum(v) / tot))
        f.write("\n## Launches in order (one bench step shown after warm-up)\n\n```\n")
        for k, g, v in order[-140:]:
            f.write("%-48s grid %-16s %9.4f ms\n" % (k[:48], g, v))
        f.write("```\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    with open(dst, "w") as f:
        f.write("# ncu --set full summary\n\nSource: `%s` (`ncu --set full --clock-control none "
                "--import-source on`).\n\n" % src)
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            grid = r[hdr.index("Grid Size")]
            if (name, grid) in seen:
                continue
            seen.add((name, grid))
            f.write("## `%s` grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n"
                    % (name, grid, r[hdr.index("Block Size")]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write("| %s | %s | %s |\n" % (k, r[i], units[i]))
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

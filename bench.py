#!/usr/bin/env python
"""bench.py -- quadtree point-in-polygon join throughput (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch of synthetic input:
    polygon_bounding_boxes -> quadtree_on_points -> join_quadtree_and_bounding_boxes
    -> quadtree_point_in_polygon

Workloads (BASELINE.json `configs`):
  configs1 : 100 M uniform fp64 points x 263 taxi-zone-like polygons, max_depth 15, max_size 512
             -- the headline at N = 1 (weak-scaled, 100 M points per GPU, when forced at N > 1)
  config2  : non-indexed bitmask point_in_polygon, 100 M points x 31 polygons (1 GPU)
  config4  : 1 G clustered (64-component Gaussian mixture) fp64 points x 10 000 polygons, STRONG
             scaling (1 G points in total, Morton-range sharded) -- the default at N >= 2, and an
             extra key of the N = 1 line (the 1-GPU time the strong-scaling speed-up refers to)
  config5  : 2 G fp32 uniform points x 50 000 polygons on 8 GPUs (HBM sizing)
With no --workload flag: configs1 at N = 1 (plus extras: config2, config4 on one GPU, the
reference's own benchmark shape families), config4 at N >= 2 (plus configs1 weak as an extra).

  value : points/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e   : the same through the public Python API with HOST (pinned) buffers: the H2D copy of the
          point columns and of the polygon table, polygon_bounding_boxes and the D2H read of the
          (polygon_index, point_index) table are inside the timed region; copies of step i+1
          overlap the read-back of step i (two copy streams, pinned double buffers)
  parity : "equal" when an order-independent 64-bit checksum over all (polygon_index, original
           point id) rows equals the same checksum of an independent implementation on the same
           input: at N = 1 the reference's own CUDA build (`gpu_reference` leg), at N > 1 a
           single-GPU run of the same global point set (check step outside the timed region)
  roofline     : the dominant kernel against the measured HBM peak in MEASURED_PEAKS.json; its
                 duration comes from CUDA events the library records between its kernels on the
                 launching stream, live in this run, over a second pass of the same K steps
  cpu_baseline : the reference's own header-only implementation compiled for the host
                 (oracle/_ref, Thrust OpenMP; kind "reference") or, if absent, the repo's CPU
                 restatement (kind "port"), on a bounded sample of the same workload
  gpu_reference: the reference's own CUDA implementation (its header-only Thrust/CUB path,
                 oracle/_ref/libcuspatial_ref_cuda.so) on the same B200 and the same inputs, run
                 in a child process (`--impl reference-cuda`); null when that library is absent

`--impl reference` times the CPU implementation only (rank 0; other ranks exit), on a bounded
sample of the workload this N would run, with all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAX_DEPTH = 15
MAX_SIZE = 512
SEED = 20251017
WORKLOADS = {
    "configs1": dict(points=100_000_000, n_poly=263, kind="uniform", dtype="f64", scaling="weak",
                     label="configs[1]"),
    "config4": dict(points=1_000_000_000, n_poly=10_000, kind="clustered", dtype="f64",
                    scaling="strong", label="configs[3]"),
    "config5": dict(points=2_000_000_000, n_poly=50_000, kind="uniform", dtype="f32",
                    scaling="strong", label="configs[4]"),
}
# CPU legs: a bounded sample of the same workload (about 3 s per pass on 16 host cores)
CPU_SAMPLE = 10_000_000
CPU_SAMPLES = {"configs1": 10_000_000, "config4": 5_000_000, "config5": 5_000_000}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (measured)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled during the timed region.

    NVML through pynvml when it loads (in-process, ~20 us per sample, every 20 ms); otherwise the
    `nvidia-smi` query of the profiling recipe (a fork + driver query per sample, every 200 ms --
    measurably slows a launch-heavy step, hence the preference for NVML)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        flag = lambda bit: "Active" if (r & bit) else "Not Active"  # noqa: E731
        return [str(sm), str(mx), flag(n.nvmlClocksEventReasonHwSlowdown),
                flag(n.nvmlClocksEventReasonHwThermalSlowdown),
                flag(n.nvmlClocksEventReasonSwThermalSlowdown),
                flag(n.nvmlClocksEventReasonSwPowerCap)]

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index),
                                          "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.02 if self.nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------------------
def make_polygons(n_poly):
    from cuspatial_b200 import datagen as D

    po, ro, vx, vy = D.taxi_zone_like_polygons(n_poly, seed=SEED)
    ext = D.polygon_extent(vx, vy)
    return (po, ro, vx, vy), ext, D.quadtree_params(ext, MAX_DEPTH)


def gen_points(kind, n, ext, seed, tdt, dev, out=None):
    """Synthetic point cloud of the workload, generated in device memory."""
    from cuspatial_b200 import datagen as D

    if kind == "uniform":
        x, y = D.uniform_points_torch(n, ext, seed, tdt, dev)
        if out is not None:
            out[0].copy_(x)
            out[1].copy_(y)
            return out
        return x, y
    return D.clustered_points_torch(n, ext, seed, tdt, dev, mixture_seed=SEED, out=out)


def checksum_rows(polygon_index, original_id, chunk=1 << 27):
    """Order-independent 64-bit checksum of (polygon_index, original point id) rows: the sum
    (mod 2^64) of a 64-bit mix of every row, computed on the device in chunks."""
    import torch

    total = 0
    n = polygon_index.shape[0]
    for a in range(0, n, chunk):
        p = polygon_index[a: a + chunk].to(torch.int64) & 0xFFFFFFFF
        o = original_id[a: a + chunk].to(torch.int64) & 0xFFFFFFFF
        h = (p << 32) | o
        h = h * -7046029254386353131          # 0x9E3779B97F4A7C15 as int64, wraps mod 2^64
        h = h ^ ((h >> 29) & ((1 << 35) - 1))  # logical shift
        h = h * -4658895280553007687          # 0xBF58476D1CE4E5B9
        h = h ^ ((h >> 32) & 0xFFFFFFFF)
        total = (total + int(h.sum().item())) & 0xFFFFFFFFFFFFFFFF
    return total


def result_checksum(pidx, hits, chunk=1 << 27):
    """Checksum of a single-GPU result: point_index -> original id through point_indices."""
    import torch

    total = 0
    pi32 = pidx.view(torch.int32)
    hp, hq = hits["polygon_index"].view(torch.int32), hits["point_index"].view(torch.int32)
    for a in range(0, hp.shape[0], chunk):
        q = hq[a: a + chunk].to(torch.int64) & 0xFFFFFFFF
        total = (total + checksum_rows(hp[a: a + chunk], pi32[q])) & 0xFFFFFFFFFFFFFFFF
    return total


# ------------------------------------------------------------------------------------------------
# CPU legs
# ------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs use every host core."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except Exception:
        pass
    return n


def cpu_baseline(workload, sample_points):
    """Time the CPU implementation of the whole path on `sample_points` points of the workload."""
    import numpy as np

    from cuspatial_b200 import datagen as D
    from oracle import hostlib

    cores = use_all_host_threads()
    lib = hostlib.reference() if hostlib.reference_available() else hostlib.oracle()
    w = WORKLOADS[workload]
    (po, ro, vx, vy), ext, scale = make_polygons(w["n_poly"])
    dt = np.float64 if w["dtype"] == "f64" else np.float32
    if w["kind"] == "uniform":
        x, y = D.uniform_points(sample_points, ext, seed=1, dtype=dt)
    else:
        x, y = D.clustered_points(sample_points, ext, seed=SEED, dtype=dt)
    vx, vy = vx.astype(dt), vy.astype(dt)
    t0 = time.perf_counter()
    tree = lib.quadtree_on_points(x, y, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH, MAX_SIZE)
    bb = lib.polygon_bounding_boxes(po, ro, vx, vy)
    pairs = lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], scale, MAX_DEPTH)
    hits = lib.quadtree_point_in_polygon(pairs[0], pairs[1], tree, tree["point_indices"], x, y,
                                         po, ro, vx, vy)
    dt_s = time.perf_counter() - t0
    return {
        "value": sample_points / dt_s, "unit": "points/s", "cores": cores, "kind": lib.kind,
        "sample": "%d %s %s points of the same workload (%d polygons, max_depth %d, max_size %d), "
                  "whole path, %.2f s, %d hit rows" % (sample_points, w["kind"], w["dtype"],
                                                      w["n_poly"], MAX_DEPTH, MAX_SIZE, dt_s,
                                                      len(hits[0])),
        "seconds": dt_s,
    }


def default_workload(n_gpus):
    return "configs1" if n_gpus <= 1 else "config4"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload if args.workload in WORKLOADS else default_workload(args.gpus)
    w = WORKLOADS[workload]
    # bounded: the whole --steps/--warmup run ends within ~2 minutes on 16 cores
    n_calls = max(args.warmup + args.steps, 1)
    sample = CPU_SAMPLES[workload]
    if n_calls > 30:
        sample = max(1_000_000, sample * 30 // n_calls)
    vals, info = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_baseline(workload, sample)
        if i >= args.warmup:
            vals.append(info["seconds"])
    ms = 1e3 * sum(vals) / max(len(vals), 1)
    v = sample / (ms / 1e3)
    info = dict(info, value=v)
    info.pop("seconds", None)
    print(json.dumps({
        "impl": "reference", "metric": "quadtree PIP join points/sec", "value": v,
        "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None,
        "dtype": w["dtype"], "data": "synthetic",
        "config": {"workload": "%s: %s %s points x %d polygons, max_depth=%d max_size=%d (each "
                               "step a %d-point sample on the host cores)"
                               % (w["label"], w["kind"], w["dtype"], w["n_poly"], MAX_DEPTH,
                                  MAX_SIZE, sample)},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# the reference's CUDA build on the same GPU (child process)
# ------------------------------------------------------------------------------------------------
def run_reference_cuda(args):
    """The reference's CUDA build (oracle/_ref/libcuspatial_ref_cuda.so) on configs[1], one GPU.

    Same inputs as the main arm (same generator and seed), inputs resident in HBM, each of the
    three reference calls timed around its own synchronous call (it ends in a stream sync).
    """
    import torch

    from oracle import cudalib

    if not cudalib.available():
        print(json.dumps({"impl": "reference-cuda", "unavailable":
                          "oracle/_ref/libcuspatial_ref_cuda.so not built or no GPU"}))
        return
    lib = cudalib.reference_cuda()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    n = args.points or WORKLOADS["configs1"]["points"]
    (po, ro, vx, vy), ext, scale = make_polygons(WORKLOADS["configs1"]["n_poly"])
    po, ro, vx, vy = (torch.as_tensor(a, device=dev) for a in (po, ro, vx, vy))
    x, y = gen_points("uniform", n, ext, SEED, torch.float64, dev)
    bb = lib.polygon_bounding_boxes(po, ro, vx, vy)
    stages, rows, csum = [], None, None
    for i in range(args.warmup + args.steps):
        tree, t1 = lib.quadtree_on_points(x, y, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH,
                                          MAX_SIZE)
        pairs, t2 = lib.join_quadtree_and_bounding_boxes(tree, *bb, ext[0], ext[2], scale,
                                                         MAX_DEPTH)
        hits, t3 = lib.quadtree_point_in_polygon(pairs[0], pairs[1], tree, tree["point_indices"],
                                                 x, y, po, ro, vx, vy)
        rows = (tree["key"].numel(), pairs[0].numel(), hits[0].numel())
        if i == args.warmup + args.steps - 1:
            csum = result_checksum(tree["point_indices"], {"polygon_index": hits[0],
                                                           "point_index": hits[1]})
        del tree, pairs, hits
        if i >= args.warmup:
            stages.append((t1, t2, t3))
    k = len(stages)
    avg = [1e3 * sum(s[j] for s in stages) / k for j in range(3)]
    ms = sum(avg)
    print(json.dumps({
        "impl": "reference-cuda", "metric": "quadtree PIP join points/sec",
        "value": n / (ms / 1e3), "unit": "points/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "points": n,
        "stage_ms": {"quadtree_on_points": avg[0], "join_quadtree_and_bounding_boxes": avg[1],
                     "quadtree_point_in_polygon": avg[2]},
        "nodes": rows[0], "pairs": rows[1], "hits": rows[2], "checksum": "%016x" % csum,
        "peak_mem_GB": torch.cuda.mem_get_info(dev)[1] / 1e9 - torch.cuda.mem_get_info(dev)[0] / 1e9,
        "what": "rapidsai/cuspatial header-only path (Thrust/CUB), nvcc sm_100a, default FP "
                "flags, stream-ordered pool allocator; inputs resident in HBM",
    }))


def gpu_reference_leg(points, timeout_s=600):
    """Run `--impl reference-cuda` in a child process (a crash or OOM there cannot take the
    main bench line down) and return its JSON, or a dict saying why there is none."""
    from oracle import cudalib

    if not os.path.exists(cudalib.REF_CUDA_PATH):
        return None
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference-cuda",
                              "--points", str(points), "--steps", "3", "--warmup", "1"],
                             capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
        for line in reversed(out.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"failed": (out.stderr or out.stdout).strip()[-300:], "rc": out.returncode}
    except Exception as e:  # timeout etc.
        return {"failed": repr(e)[:300]}


# ------------------------------------------------------------------------------------------------
# single-GPU measurements
# ------------------------------------------------------------------------------------------------
def join_step(cs, x, y, polys, ext, scale):
    bb = cs.polygon_bounding_boxes(polys)
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], scale,
                                       MAX_DEPTH, MAX_SIZE)
    pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3],
                                                scale, MAX_DEPTH)
    hits = cs.quadtree_point_in_polygon(pairs, tree, pidx, (x, y), polys)
    return pidx, tree, pairs, hits


def time_join(dev, x, y, polys, ext, scale, steps, warmup, with_checksum=False):
    """K timed steps of the join on device-resident inputs (CUDA events around the K steps)."""
    import torch

    import cuspatial_b200 as cs
    from cuspatial_b200 import _lib

    info = {}
    for _ in range(max(warmup, 1)):
        out = join_step(cs, x, y, polys, ext, scale)
    info["nodes"], info["pairs"], info["hits"] = len(out[1]), len(out[2]), len(out[3])
    lengths = out[1]["length"].to(torch.int64)[out[2]["quad_offset"].to(torch.int64)]
    info["candidates"] = int(lengths.sum())
    if with_checksum:
        info["checksum"] = result_checksum(out[0], out[3])
    del out, lengths
    torch.cuda.synchronize(dev)
    l0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = join_step(cs, x, y, polys, ext, scale)
        del out
    e1.record()
    torch.cuda.synchronize(dev)
    info["ms_per_step"] = e0.elapsed_time(e1) / steps
    info["gpu_launches"] = int(_lib.kernel_launch_count() - l0)
    return info


def bitmask_measurement(dev, n, steps, warmup, tdt=None, polys_np=None, ext=None, label=None):
    """configs[2]: non-indexed point_in_polygon, n points x 31 polygons through the bitmask API."""
    import torch

    import cuspatial_b200 as cs
    from cuspatial_b200 import _lib

    tdt = tdt or torch.float64
    if polys_np is None:
        (po, ro, vx, vy), ext, _ = make_polygons(WORKLOADS["configs1"]["n_poly"])
        po = po[:32]
    else:
        po, ro, vx, vy = polys_np
    polys = (torch.as_tensor(po.astype("int32"), device=dev),
             torch.as_tensor(ro.astype("int32"), device=dev),
             torch.as_tensor(vx, device=dev).to(tdt), torch.as_tensor(vy, device=dev).to(tdt))
    x, y = gen_points("uniform", n, ext, SEED, tdt, dev)
    for _ in range(max(warmup, 1)):
        m = cs.point_in_polygon_bitmask((x, y), polys)
    inside = int((m != 0).sum())
    torch.cuda.synchronize(dev)
    l0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m = cs.point_in_polygon_bitmask((x, y), polys)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    peak, src = read_peaks()
    T = x.element_size()
    gbs = n * (2 * T + 4) / (ms / 1e3) / 1e9
    return {
        "metric": "bitmask point_in_polygon points/sec", "value": n / (ms / 1e3),
        "unit": "points/s", "ms_per_step": ms, "steps": steps, "warmup": warmup,
        "dtype": "f64" if T == 8 else "f32",
        "workload": label or ("configs[2]: %d uniform fp64 points x 31 polygons, bitmask API" % n),
        "points_in_some_polygon": inside,
        "roofline": {"bound": "hbm", "kernel": "pip_bitmask", "achieved": gbs, "peak": peak,
                     "unit": "GB/s", "frac": gbs / peak, "peak_source": src,
                     "alg_bytes_per_point": 2 * T + 4},
        "gpu_launches": int(_lib.kernel_launch_count() - l0),
    }


def shape_family_measurements(dev, steps, warmup):
    """The reference's own benchmark shape families:
    cpp/benchmarks/point_in_polygon/point_in_polygon.cu:41-102 (31 regular n-gons of radius 10
    around the origin, points uniform in [-20, 20]^2, 10 M points, 4/10/100 sides) and
    cpp/benchmarks/indexing/quadtree_on_points.cu:33-130 (golden-ratio nested rectangles,
    max_size = N / 4^4, scale -1, depth 15)."""
    import numpy as np
    import torch

    import cuspatial_b200 as cs
    from cuspatial_b200 import datagen as D

    out = {}
    for sides in (4, 10, 100):
        polys = D.regular_ngons(31, sides, 10.0)
        r = bitmask_measurement(dev, 10_000_000, steps, warmup, polys_np=polys,
                                ext=(-20.0, 20.0, -20.0, 20.0),
                                label="reference PIP benchmark shape: 10000000 points x 31 regular "
                                      "%d-gons" % sides)
        out["pip_%dgon" % sides] = {"points_per_s": r["value"], "ms": r["ms_per_step"],
                                    "hbm_frac": r["roofline"]["frac"]}
    x, y = D.nested_rectangle_points(10_000)
    n = len(x)
    dx, dy = torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)
    x0, x1, y0, y1 = float(x.min()), float(x.max()), float(y.min()), float(y.max())
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(max(warmup, 1)):
            cs.quadtree_on_points((dx, dy), x0, x1, y0, y1, -1, 15, n // 256)
        torch.cuda.synchronize(dev)
        # per-step events and the median: this extra follows workloads of very different sizes in
        # the same process, and a single step that has to regrow the memory pool (tens of ms) would
        # otherwise dominate a five-step mean
        per_step = []
        for _ in range(max(steps, 3)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, tree = cs.quadtree_on_points((dx, dy), x0, x1, y0, y1, -1, 15, n // 256)
            e1.record()
            torch.cuda.synchronize(dev)
            per_step.append(e0.elapsed_time(e1))
    ms = sorted(per_step)[len(per_step) // 2]
    out["quadtree_nested_rectangles"] = {"points": n, "ms": ms, "ms_max": max(per_step),
                                         "points_per_s": n / (ms / 1e3),
                                         "nodes": len(tree), "bounding_box_size": 10_000,
                                         "dtype": str(np.dtype(x.dtype))}
    return out


def e2e_single(dev, n, tdt, polys_np, ext, scale, x, y, n_hits, steps):
    """End to end through the public API with HOST buffers.  Per step: H2D of the point columns
    and of the polygon table from pinned memory, polygon_bounding_boxes, the three calls, D2H of
    the full pair table into pinned memory.  Double-buffered: the upload of step i+1 runs on a copy
    stream while step i computes and step i-1's table is read back on a second copy stream."""
    import torch

    import cuspatial_b200 as cs

    T = x.element_size()
    hx = torch.empty(n, dtype=tdt).pin_memory()
    hy = torch.empty(n, dtype=tdt).pin_memory()
    hx.copy_(x)
    hy.copy_(y)
    hp = [torch.as_tensor(a).pin_memory() for a in polys_np]
    cap = int(n_hits * 1.05) + 1024
    ha = [torch.empty(cap, dtype=torch.uint32).pin_memory() for _ in range(2)]
    hb = [torch.empty(cap, dtype=torch.uint32).pin_memory() for _ in range(2)]
    dx = [torch.empty(n, dtype=tdt, device=dev) for _ in range(2)]
    dy = [torch.empty(n, dtype=tdt, device=dev) for _ in range(2)]
    main = torch.cuda.current_stream(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]    # compute finished reading dx[b], dy[b]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    keep = [None, None]
    poly_bytes = sum(t.numel() * t.element_size() for t in hp)

    def upload(i):
        b = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[b])
            dx[b].copy_(hx, non_blocking=True)
            dy[b].copy_(hy, non_blocking=True)
            ev_in[b].record(s_in)

    def run(k):
        d2h = 0
        upload(0)
        for i in range(k):
            b = i % 2
            if i + 1 < k:
                upload(i + 1)                        # overlaps this step's compute and read-back
            main.wait_event(ev_in[b])
            polys = tuple(t.to(dev, non_blocking=True) for t in hp)
            _, _, _, hits = join_step(cs, dx[b], dy[b], polys, ext, scale)
            ev_free[b].record(main)
            h = len(hits)
            ev_out[b].synchronize()                  # pinned slot b free again (step i-2 read back)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                ha[b][:h].copy_(hits["polygon_index"], non_blocking=True)
                hb[b][:h].copy_(hits["point_index"], non_blocking=True)
                ev_out[b].record(s_out)
            keep[b] = hits                           # alive until its read-back has finished
            d2h = 2 * h * 4
        torch.cuda.synchronize(dev)
        return d2h

    for e in ev_free + ev_out:
        e.record(main)
    run(2)
    torch.cuda.synchronize(dev)
    # a pipeline has a fill (first upload) and a drain (last read-back) that no step overlaps:
    # run enough steps for the steady state to dominate (20 by default, ~0.6 s)
    k = max(8, min(4 * steps, 24))
    t0 = time.perf_counter()
    d2h = run(k)
    dt = (time.perf_counter() - t0) / k
    return {"value": n / dt, "unit": "points/s",
            "h2d_bytes_per_step": 2 * n * T + poly_bytes, "d2h_bytes_per_step": int(d2h),
            "ms_per_step": 1e3 * dt, "steps": k,
            "note": "pinned host x, y and polygon table -> device, polygon_bounding_boxes + 3 API "
                    "calls, full (polygon_index, point_index) table read back to pinned host "
                    "memory; uploads of step i+1 overlap compute and read-back of step i"}


def run_single(args, workload):
    import torch

    from cuspatial_b200 import _lib

    w = WORKLOADS[workload]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    n = args.points or w["points"]
    dtype = args.dtype or w["dtype"]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    T = 8 if dtype == "f64" else 4
    polys_np, ext, scale = make_polygons(w["n_poly"])
    polys = [torch.as_tensor(a, device=dev) for a in polys_np]
    if dtype == "f32":
        polys[2], polys[3] = polys[2].float(), polys[3].float()
        polys_np = (polys_np[0], polys_np[1], polys_np[2].astype("float32"),
                    polys_np[3].astype("float32"))
    polys = tuple(polys)
    x, y = gen_points(w["kind"], n, ext, SEED, tdt, dev)

    sampler = ClockSampler(dev.index)
    sampler.start()
    head = time_join(dev, x, y, polys, ext, scale, args.steps, args.warmup, with_checksum=True)
    ms_step = head["ms_per_step"]
    # second pass of the same K steps with the library's own CUDA events between kernels (on the
    # launching stream): per-kernel durations for the roofline.  The event records cost bubbles,
    # so they are kept out of the headline region above.
    import cuspatial_b200 as cs

    _lib.set_profiling(True)
    _lib.get_profile()
    torch.cuda.synchronize(dev)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        out = join_step(cs, x, y, polys, ext, scale)
        del out
    p1.record()
    torch.cuda.synchronize(dev)
    ms_instrumented = p0.elapsed_time(p1) / args.steps
    profile = _lib.get_profile()
    _lib.set_profiling(False)
    clocks = sampler.stop()
    value = n / (ms_step / 1e3)

    stage_ms = {}
    for name, ms in profile:
        stage_ms.setdefault(name, []).append(ms)
    stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}          # per launch
    stage_per_step = {k: sum(v) / args.steps for k, v in stage_ms.items()}  # per step

    peak, peak_src = read_peaks()
    c, h = head["candidates"] / n, head["hits"] / n
    passes = 4
    # algorithmic bytes per launch of the candidate dominant kernels
    alg_bytes = {
        "onesweep_pass": n * (12 + 16 * (passes - 1)) / passes,
        "encode_hist": n * (2 * T + 4),
        "pip_eval": n * c * (4 + 2 * T),   # index + gathered coords per candidate
        "pip_emit": n * h * 8,
    }
    dom = max(alg_bytes, key=lambda k: stage_per_step.get(k, 0.0))
    dom_ms = stage_avg.get(dom, float("nan"))
    achieved = alg_bytes[dom] / (dom_ms / 1e3) / 1e9
    b_alg = (2 * T + 4) + (12 + 16 * (passes - 1)) + 4 + c * (4 + 2 * T) + 8 * h
    traffic = None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:  # per-launch DRAM bytes of that kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", name)) as f:
                traffic = json.load(f).get(dom)
            break
        except Exception:
            pass
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "kernel_ms_per_launch": dom_ms,
        "kernel_share_of_step": stage_per_step.get(dom, 0) / ms_instrumented,
        "instrumented_ms_per_step": ms_instrumented,
        "timing": "kernel durations: CUDA events recorded by the library between its kernels on "
                  "the launching stream, over a second pass of the same K steps",
        "pipeline": {"alg_bytes_per_point": b_alg, "candidates_per_point": c,
                     "hits_per_point": h,
                     "achieved_GBs": n * b_alg / (ms_step / 1e3) / 1e9,
                     "frac": n * b_alg / (ms_step / 1e3) / 1e9 / peak},
    }

    e2e = None
    if not args.no_e2e:
        e2e = e2e_single(dev, n, tdt, polys_np, ext, scale, x, y, head["hits"], args.steps)

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(workload, CPU_SAMPLE)
        cpu.pop("seconds", None)

    extra = {}
    del x, y
    torch.cuda.empty_cache()
    gpu_ref, parity = None, None
    if not args.no_gpu_reference and workload == "configs1" and dtype == "f64":
        gpu_ref = gpu_reference_leg(n)
        if gpu_ref and "checksum" in gpu_ref:
            same = (gpu_ref["checksum"] == "%016x" % head["checksum"]
                    and gpu_ref["hits"] == head["hits"] and gpu_ref["nodes"] == head["nodes"]
                    and gpu_ref["pairs"] == head["pairs"])
            parity = "equal" if same else "DIFFERENT"
    if not args.no_extras and workload == "configs1":
        ks, kw = max(2, min(args.steps, 5)), max(1, min(args.warmup, 3))
        try:
            extra["configs[2]"] = bitmask_measurement(dev, 100_000_000, ks, kw)
            torch.cuda.empty_cache()
        except Exception as e:  # an extra must not take the headline down
            extra["configs[2]"] = {"failed": repr(e)[:300]}
        try:
            extra["reference_benchmark_shapes"] = shape_family_measurements(dev, ks, kw)
            torch.cuda.empty_cache()
        except Exception as e:
            extra["reference_benchmark_shapes"] = {"failed": repr(e)[:300]}
        try:
            w4 = WORKLOADS["config4"]
            p4, ext4, scale4 = make_polygons(w4["n_poly"])
            polys4 = tuple(torch.as_tensor(a, device=dev) for a in p4)
            x4, y4 = gen_points("clustered", w4["points"], ext4, SEED, torch.float64, dev)
            r4 = time_join(dev, x4, y4, polys4, ext4, scale4, 2, 1, with_checksum=True)
            del x4, y4
            torch.cuda.empty_cache()
            c4, h4 = r4["candidates"] / w4["points"], r4["hits"] / w4["points"]
            b4 = 20 + 60 + 4 + c4 * 20 + 8 * h4
            extra["configs[3]_on_1_gpu"] = {
                "workload": "%d clustered (64-component Gaussian mixture) fp64 points x %d "
                            "polygons on ONE GPU: the time the strong-scaling speed-up of the "
                            "N >= 2 lines refers to" % (w4["points"], w4["n_poly"]),
                "value": w4["points"] / (r4["ms_per_step"] / 1e3), "unit": "points/s",
                "ms_per_step": r4["ms_per_step"], "steps": 2, "warmup": 1,
                "nodes": r4["nodes"], "pairs": r4["pairs"], "candidates": r4["candidates"],
                "hits": r4["hits"], "checksum": "%016x" % r4["checksum"],
                "pipeline_roofline_frac": w4["points"] * b4 / (r4["ms_per_step"] / 1e3) / 1e9 / peak,
                "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 1e9}
        except Exception as e:
            extra["configs[3]_on_1_gpu"] = {"failed": repr(e)[:300]}

    print(json.dumps({
        "metric": "quadtree PIP join points/sec", "value": value, "unit": "points/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": dtype,
        "data": "synthetic",
        "config": {"workload": "%s: %d %s %s points x %d taxi-zone-like polygons, quadtree "
                               "max_depth=%d max_size=%d, single GPU"
                               % (w["label"], n, w["kind"], "fp64" if T == 8 else "fp32",
                                  w["n_poly"], MAX_DEPTH, MAX_SIZE),
                   "l2": "inputs (%.1f GB) and every intermediate exceed the 126 MB L2"
                         % (2 * n * T / 1e9),
                   "nodes": head["nodes"], "pairs": head["pairs"],
                   "candidates": head["candidates"], "hits": head["hits"],
                   "parallelism": "single GPU"},
        "parity": parity, "checksum": "%016x" % head["checksum"],
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": head["gpu_launches"], "clocks": clocks, "gpu_reference": gpu_ref,
        "stage_ms_per_step": {k: round(v, 4) for k, v in stage_per_step.items()},
        "extra": extra or None,
    }))


# ------------------------------------------------------------------------------------------------
# N > 1: the distributed join
# ------------------------------------------------------------------------------------------------
def sharded_parity_check(dist, dev, rank, world, w, ext, scale, polys, tdt):
    """Outside the timed region: the sharded join on a <= 20 M-point global set against a
    single-GPU run of the same global point set (rank 0 regenerates every rank's shard from its
    seed), compared by the order-independent checksum of (polygon_index, original point id)."""
    import torch

    import cuspatial_b200 as cs
    from cuspatial_b200 import multi_gpu as mg

    total = min(20_000_000, w["points"])
    per = total // world
    x, y = gen_points(w["kind"], per, ext, SEED + 1000 + rank, tdt, dev)
    out = mg.sharded_quadtree_point_in_polygon(
        (x, y), polys, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH, MAX_SIZE,
        gather_pairs=True, gather_point_indices=True)
    q = out["point_index"].to(torch.int64) & 0xFFFFFFFF
    got = checksum_rows(out["polygon_index"], out["point_indices"][q])
    rows = int(out["polygon_index"].shape[0])
    del out, q, x, y
    res = None
    if rank == 0:
        xs, ys = [], []
        for r in range(world):
            a, b = gen_points(w["kind"], per, ext, SEED + 1000 + r, tdt, dev)
            xs.append(a)
            ys.append(b)
        gx, gy = torch.cat(xs), torch.cat(ys)
        del xs, ys
        pidx, tree, pairs, hits = join_step(cs, gx, gy, polys, ext, scale)
        want = result_checksum(pidx, hits)
        res = {"parity": "equal" if (want == got and len(hits) == rows) else "DIFFERENT",
               "check_points": per * world, "check_rows": rows, "checksum": "%016x" % got,
               "single_gpu_checksum": "%016x" % want}
        del gx, gy, pidx, tree, pairs, hits
    torch.cuda.empty_cache()
    dist.barrier()
    return res


def time_sharded(dist, dev, pts, polys, ext, scale, steps, warmup, gather_pairs):
    import torch

    from cuspatial_b200 import _lib
    from cuspatial_b200 import multi_gpu as mg

    def step(profile=False):
        return mg.sharded_quadtree_point_in_polygon(
            pts, polys, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH, MAX_SIZE,
            gather_pairs=gather_pairs, gather_point_indices=False, profile=profile)

    def barrier():
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(warmup, 1)):
        out = step()
    info = {"rows": int(out["polygon_index"].shape[0]), "counts": out["counts"],
            "rows_per_rank": out["rows_per_rank"]}
    del out
    l0 = _lib.kernel_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
        del out
    e1.record()
    barrier()
    info["gpu_launches"] = int(_lib.kernel_launch_count() - l0)
    tmax = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    info["ms_per_step"] = float(tmax.item()) / steps
    # one more step with CUDA-event phase marks (no host synchronisation inside the step)
    out = step(profile=True)
    del out
    ph = dict(mg.LAST_PROFILE)
    names = sorted(ph)
    t = torch.tensor([ph[k] for k in names], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    info["phase_ms_last_step"] = {k: round(float(v), 3) for k, v in zip(names, t.tolist())}
    barrier()
    return info


def e2e_sharded(dist, dev, pts, polys_np, ext, scale, steps):
    """Host-buffer end-to-end at N > 1: every rank uploads its shard (pinned -> its registered
    device columns) and the polygon table, runs the sharded join keeping the rows partitioned
    (gather_pairs=False: the union over ranks is the table) and reads its rows back."""
    import torch

    from cuspatial_b200 import multi_gpu as mg

    n = len(pts)
    hx = torch.empty(n, dtype=pts.dtype).pin_memory()
    hy = torch.empty(n, dtype=pts.dtype).pin_memory()
    hx.copy_(pts.x)
    hy.copy_(pts.y)
    hp = [torch.as_tensor(a).pin_memory() for a in polys_np]

    def step():
        pts.x.copy_(hx, non_blocking=True)
        pts.y.copy_(hy, non_blocking=True)
        polys = tuple(t.to(dev, non_blocking=True) for t in hp)
        out = mg.sharded_quadtree_point_in_polygon(
            pts, polys, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH, MAX_SIZE,
            gather_pairs=False)
        a = out["polygon_index"].to("cpu", non_blocking=True)
        b = out["point_index"].to("cpu", non_blocking=True)
        torch.cuda.synchronize(dev)
        return 4 * (int(a.shape[0]) + int(b.shape[0]))

    step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    k = max(1, min(steps, 3))
    t0 = time.perf_counter()
    for _ in range(k):
        d2h = step()
    torch.cuda.synchronize(dev)
    dist.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / k, float(d2h)], device=dev, dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    poly_bytes = sum(t.numel() * t.element_size() for t in hp)
    return float(dt[0].item()), 2 * n * pts.x.element_size() + poly_bytes, int(dt[1].item())


def run_sharded(args, workload):
    import torch
    import torch.distributed as dist

    from cuspatial_b200 import multi_gpu as mg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)

    def build(name, points_override=None):
        w = WORKLOADS[name]
        dtype = args.dtype or w["dtype"]
        tdt = torch.float64 if dtype == "f64" else torch.float32
        total = points_override or w["points"]
        per = total // world if w["scaling"] == "strong" else total
        polys_np, ext, scale = make_polygons(w["n_poly"])
        if dtype == "f32":
            polys_np = (polys_np[0], polys_np[1], polys_np[2].astype("float32"),
                        polys_np[3].astype("float32"))
        polys = tuple(torch.as_tensor(a, device=dev) for a in polys_np)
        return w, dtype, tdt, per, polys_np, polys, ext, scale

    w, dtype, tdt, per, polys_np, polys, ext, scale = build(workload, args.points)
    T = 8 if dtype == "f64" else 4
    parity = None
    if not args.no_parity_check:
        parity = sharded_parity_check(dist, dev, rank, world, w, ext, scale, polys, tdt)
    # this rank's shard lives in symmetric memory (peers read coordinates on demand)
    pts = mg.allocate_points(per, tdt, dev)
    gen_points(w["kind"], per, ext, SEED + rank, tdt, dev, out=(pts.x, pts.y))
    torch.cuda.synchronize(dev)
    dist.barrier()

    sampler = ClockSampler(dev.index)
    sampler.start()
    head = time_sharded(dist, dev, pts, polys, ext, scale, args.steps, args.warmup, True)
    clocks = sampler.stop()
    ms_step = head["ms_per_step"]
    total = per * world
    value = total / (ms_step / 1e3)
    part = time_sharded(dist, dev, pts, polys, ext, scale, max(2, min(args.steps, 5)), 1, False)
    # roofline of the dominant kernel on rank 0: one more step with the library's own CUDA events
    # between its kernels (outside the timed region above)
    roofline = None
    try:
        from cuspatial_b200 import _lib

        _lib.set_profiling(True)
        _lib.get_profile()
        out = mg.sharded_quadtree_point_in_polygon(
            pts, polys, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH, MAX_SIZE,
            gather_pairs=True)
        n_local = int(out["counts"][rank])
        del out
        torch.cuda.synchronize(dev)
        prof = _lib.get_profile()
        _lib.set_profiling(False)
        st = {}
        for name, ms in prof:
            st.setdefault(name, []).append(ms)
        if rank == 0 and st.get("onesweep_pass"):
            dom = max(st, key=lambda k: sum(st[k]))
            peak, peak_src = read_peaks()
            sort_ms = sum(st["onesweep_pass"]) / len(st["onesweep_pass"])
            alg = n_local * 16.0   # per pass: key + global id read, key + global id written
            roofline = {
                "bound": "hbm", "kernel": "onesweep_pass", "achieved": alg / (sort_ms / 1e3) / 1e9,
                "peak": peak, "unit": "GB/s", "frac": alg / (sort_ms / 1e3) / 1e9 / peak,
                "traffic": None, "peak_source": peak_src, "kernel_ms_per_launch": sort_ms,
                "kernel_share_of_step": sum(st["onesweep_pass"]) / ms_step,
                "largest_stage_on_rank0": dom,
                "stage_ms_rank0": {k: round(sum(v), 4) for k, v in st.items()},
                "timing": "rank 0, one instrumented step after the timed region: CUDA events "
                          "recorded by the library between its kernels; algorithmic bytes = 16 B "
                          "per received key per pass (all four passes carry the global ids)"}
    except Exception as e:  # the line is still valid without it
        roofline = {"failed": repr(e)[:200]} if rank == 0 else None
        try:
            _lib.set_profiling(False)
        except Exception:
            pass
    e2e = None
    if not args.no_e2e:
        dt, h2d, d2h = e2e_sharded(dist, dev, pts, polys_np, ext, scale, args.steps)
        e2e = {"value": total / dt, "unit": "points/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * dt,
               "note": "per rank: pinned host shard + polygon table -> device, sharded join with "
                       "the rows kept partitioned by key range (their union is the table), this "
                       "rank's rows read back to the host; bytes are per rank (max over ranks)"}
    del pts
    torch.cuda.empty_cache()
    extra = {
        "rows_partitioned_not_merged": {
            "ms_per_step": part["ms_per_step"], "value": total / (part["ms_per_step"] / 1e3),
            "unit": "points/s", "phase_ms_last_step": part["phase_ms_last_step"],
            "what": "same step with gather_pairs=False: every rank keeps the rows of its own key "
                    "range (global point_index) instead of expanding every rank's rows"}}
    if not args.no_extras and workload == "config4":
        try:
            w1, dtype1, tdt1, per1, pnp1, polys1, ext1, scale1 = build("configs1")
            p1 = mg.allocate_points(per1, tdt1, dev)
            gen_points("uniform", per1, ext1, SEED + rank, tdt1, dev, out=(p1.x, p1.y))
            ks = max(2, min(args.steps, 5))
            r1 = time_sharded(dist, dev, p1, polys1, ext1, scale1, ks, 2, True)
            r1p = time_sharded(dist, dev, p1, polys1, ext1, scale1, ks, 1, False)
            extra["configs[1]_weak"] = {
                "workload": "%d uniform fp64 points PER GPU x %d polygons (weak scaling)"
                            % (per1, w1["n_poly"]),
                "value": per1 * world / (r1["ms_per_step"] / 1e3), "unit": "points/s",
                "ms_per_step": r1["ms_per_step"], "merged_rows": r1["rows"],
                "phase_ms_last_step": r1["phase_ms_last_step"],
                "rows_partitioned_value": per1 * world / (r1p["ms_per_step"] / 1e3),
                "rows_partitioned_ms_per_step": r1p["ms_per_step"]}
            del p1
        except Exception as e:
            extra["configs[1]_weak"] = {"failed": repr(e)[:300]}
    if rank == 0:
        print(json.dumps({
            "metric": "quadtree PIP join points/sec", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": dtype,
            "data": "synthetic",
            "config": {"workload": "%s: %d %s %s points in total (%d per GPU) x %d polygons, "
                                   "max_depth=%d max_size=%d; points Morton-range sharded over %d "
                                   "GPUs (fused partition + all-to-all of 8-byte (key, id) records "
                                   "over NVLink, coordinates read through peer pointers on "
                                   "demand), polygons broadcast, pair table merged on every rank"
                                   % (w["label"], total, w["kind"], "fp64" if T == 8 else "fp32",
                                      per, w["n_poly"], MAX_DEPTH, MAX_SIZE, world),
                       "l2": "inputs and intermediates exceed the 126 MB L2",
                       "merged_rows": head["rows"], "points_per_rank_after_sharding": head["counts"],
                       "parallelism": "morton-range point shards x%d, replicated polygons" % world},
            "parity": parity["parity"] if parity else None, "parity_check": parity,
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e,
            "gpu_launches": head["gpu_launches"], "clocks": clocks,
            "phase_ms_last_step": head["phase_ms_last_step"], "extra": extra,
        }))
    dist.barrier()
    dist.destroy_process_group()


def run_nearest(args):
    """SURVEY 8f row 2 (informational): quadtree_point_to_nearest_linestring.  `--points` uniform
    fp64 points against the 263 polygon outlines taken as linestrings; the bounding boxes are
    grown by the whole extent so that every point is a candidate of every linestring (the regime
    where the reference's result is defined for every row).  A step = linestring boxes + bbox join
    + nearest-linestring refinement on a prebuilt quadtree; the reference's CUDA build runs the
    same three calls on the same inputs when it is present."""
    import torch

    import cuspatial_b200 as cs
    from cuspatial_b200 import _lib

    dev = torch.device("cuda", 0)
    (po, ro, vx, vy), ext, scale = make_polygons(263)
    lo = ro[po]                      # one linestring per polygon: its rings' vertices in order
    lines = (torch.as_tensor(lo.astype("uint32"), device=dev), torch.as_tensor(vx, device=dev),
             torch.as_tensor(vy, device=dev))
    n = args.points or 1_000_000
    x, y = gen_points("uniform", n, ext, SEED, torch.float64, dev)
    radius = (ext[1] - ext[0]) + (ext[3] - ext[2])
    pidx, tree = cs.quadtree_on_points((x, y), ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH,
                                       MAX_SIZE)

    def step():
        bb = cs.linestring_bounding_boxes(lines, radius)
        pairs = cs.join_quadtree_and_bounding_boxes(tree, bb, ext[0], ext[1], ext[2], ext[3],
                                                    scale, MAX_DEPTH)
        return pairs, cs.quadtree_point_to_nearest_linestring(pairs, tree, pidx, (x, y), lines)

    for _ in range(max(args.warmup, 1)):
        pairs, out = step()
    n_pairs = len(pairs)
    segs = int(lo[-1]) - (len(lo) - 1)
    checksum = float(out["distance"].sum())
    torch.cuda.synchronize()
    l0 = _lib.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = int(_lib.kernel_launch_count() - l0)

    ref = None
    from oracle import cudalib
    if cudalib.available() and not args.no_gpu_reference:
        lib = cudalib.reference_cuda()
        rt, _ = lib.quadtree_on_points(x, y, ext[0], ext[1], ext[2], ext[3], scale, MAX_DEPTH,
                                       MAX_SIZE)
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            bb = lib.linestring_bounding_boxes(lines[0], lines[1], lines[2], radius)
            t_bb = time.perf_counter() - t0
            rp, t1 = lib.join_quadtree_and_bounding_boxes(rt, *bb, ext[0], ext[2], scale, MAX_DEPTH)
            ro_, t2 = lib.quadtree_point_to_nearest_linestring(rp[0], rp[1], rt,
                                                               rt["point_indices"], x, y, *lines)
            ts.append((t_bb + t1 + t2) * 1e3)
        same = bool(torch.equal(ro_[1], out["linestring_index"]) and
                    torch.equal(ro_[2], out["distance"]))
        ref = {"ms_per_step": min(ts[1:]), "points_per_s": n / (min(ts[1:]) / 1e3),
               "identical_rows": same}
    print(json.dumps({
        "metric": "quadtree nearest-linestring points/sec", "value": n / (ms / 1e3),
        "unit": "points/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%d uniform fp64 points x %d linestrings (%d segments), every "
                               "point a candidate of every linestring, prebuilt quadtree "
                               "max_depth=%d max_size=%d" % (n, len(lo) - 1, segs, MAX_DEPTH,
                                                             MAX_SIZE),
                   "pairs": n_pairs, "point_segment_distances_per_step": n * segs,
                   "distance_checksum": checksum},
        "gpu_launches": launches, "gpu_reference": ref,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--points", type=int, default=0,
                    help="override the workload's point count (total for config4/config5, per "
                         "GPU for configs1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--dtype", default=None, choices=["f64", "f32"],
                    help="coordinate type (default: the workload's)")
    ap.add_argument("--workload", default="auto",
                    choices=["auto", "configs1", "join", "config2", "bitmask", "config4", "config5",
                             "nearest", "shapes"],
                    help="auto = configs1 at N = 1, config4 (strong scaling) at N >= 2")
    args = ap.parse_args()
    if args.workload == "join":
        args.workload = "configs1"
    if args.workload == "bitmask":
        args.workload = "config2"
    if args.impl == "reference":
        use_all_host_threads()   # before anything loads libgomp with torchrun's OMP_NUM_THREADS=1
        return run_reference(args)
    if args.impl == "reference-cuda":
        return run_reference_cuda(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "nearest":
        return run_nearest(args)
    if args.workload in ("config2", "shapes"):
        import torch

        dev = torch.device("cuda", 0)
        if args.workload == "config2":
            tdt = torch.float32 if args.dtype == "f32" else torch.float64
            r = bitmask_measurement(dev, args.points or 100_000_000, args.steps, args.warmup, tdt)
            r.update({"n_gpus": 1, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "data": "synthetic",
                      "config": {"workload": r.pop("workload")}})
        else:
            r = shape_family_measurements(dev, args.steps, args.warmup)
        print(json.dumps(r))
        return
    workload = args.workload if args.workload != "auto" else default_workload(world)
    if world > 1:
        return run_sharded(args, workload)
    return run_single(args, workload)


if __name__ == "__main__":
    main()

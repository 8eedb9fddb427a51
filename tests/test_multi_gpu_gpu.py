"""GPU tests of the sharded join over NCCL + symmetric memory (pytest -m gpu): torchrun launches
tests/multi_gpu_parity.py, which compares the merged pair set, the partitioned rows and the global
point_indices with a single-process CPU-oracle run.  One rank exercises the whole flow on a
single-GPU box (the device plan, the fused partition into this rank's own receive buffer, the sort
of received keys, the refinement through coordinate segments, the merge); with two or more GPUs
visible the same script runs on all of them (peer stores over NVLink)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(nproc, extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
           str(nproc), "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    log = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(log):
        with open(os.path.join(log, "multi_gpu_parity_n%d.log" % nproc), "w") as f:
            f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-4000:])
    assert r.returncode == 0 and "MULTI_GPU_PARITY OK" in r.stdout, (r.stdout[-3000:],
                                                                   r.stderr[-3000:])


def test_sharded_join_one_rank_nccl():
    _run(1)


def test_sharded_join_one_rank_plain_stores():
    _run(1, {"BSJ_MG_BULK_COPY": "0"})


def test_sharded_join_all_visible_gpus_nccl():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    _run(min(n, 8))


def test_second_device_in_the_same_process(oracle_lib):
    """Kernel attributes (dynamic shared memory above 48 KB) and the memory-pool set-up are per
    DEVICE: after a call on cuda:0 the same process must run unchanged on cuda:1
    (configure_once_per_device in csrc/api.cu; round 1 kept one flag per process)."""
    import numpy as np
    import torch

    from util import assert_same, make_case, run_gpu, run_host

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    c = make_case(300000, 40, 12, "c", np.float64, seed=5, median_vertices=40, oob=50)
    want = run_host(oracle_lib, c, 64)
    for dev in (0, 1, 0):
        with torch.cuda.device(dev):
            assert_same(run_gpu(c, 64), want, "cuda:%d" % dev)

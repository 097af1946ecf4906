"""The ABBA combinations over NCCL on two or more GPUs of one box (skipped below two):
scripts/check_abba_multigpu.py under torchrun, one rank per GPU -- mean (partial sums + reduce),
median and sigma clipping (row-slab exchange) must equal the single-process oracle bit for bit.
The host-side logic of the same functions runs on the CPU with gloo (test_distributed_gloo.py)."""

import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abba_combinations_over_nccl():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs two GPUs on one box")
    world = 2 if ngpu < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "check_abba_multigpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("[{")][-1]
    for case in json.loads(line):
        assert case["world"] == world
        assert case["mean_bit_exact"] and case["median_bit_exact"] and case["sigmaclip_bit_exact"]

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
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=200, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith("[{")][-1]
    for case in json.loads(line):
        assert case["world"] == world
        assert case["mean_bit_exact"] and case["median_bit_exact"] and case["sigmaclip_bit_exact"]
        assert case["fused_mean_bit_exact"]          # exchange over NVLink peer memory (peer.py)


def _load_nccl():
    import ctypes
    import glob
    import os
    names = []
    try:
        import nvidia.nccl
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            names += glob.glob(os.path.join(base, "lib", "libnccl.so*"))
    except ImportError:
        pass
    for name in names + ["libnccl.so.2"]:
        try:
            return ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
        except OSError:
            continue
    return None


def test_c_abi_exchanges_with_a_callers_communicator():
    """emirwv_gather / emirwv_abba_reduce with an ncclComm_t the CALLER owns (a non-Python
    caller's situation): one process, two GPUs, communicators from ncclCommInitAll; the library
    finds NCCL in the process.  Gather in rank order (unequal counts, a rank without frames),
    in-place float64 reduce of the A and B sums, then emirwv_abba_finish."""
    import ctypes

    import numpy as np
    import torch

    from pyemir_b200 import _cabi
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    nccl = _load_nccl()
    if nccl is None:
        pytest.skip("no libnccl.so.2 to create a communicator with")
    lib = _cabi.load()
    vp = ctypes.c_void_p
    comms = (vp * 2)()
    devs = (ctypes.c_int * 2)(0, 1)
    assert nccl.ncclCommInitAll(comms, 2, devs) == 0
    try:
        rng = np.random.default_rng(17)
        h, w = 37, 129
        for counts in ([3, 5], [4, 0], [0, 2]):
            stacks = [rng.normal(size=(n, h, w)).astype(np.float32) for n in counts]
            local = [torch.from_numpy(s).to(f"cuda:{r}") for r, s in enumerate(stacks)]
            root_buf = torch.full((sum(counts), h, w), float("nan"), device="cuda:0")
            carr = (ctypes.c_int32 * 2)(*counts)
            streams = [torch.cuda.Stream(device=r) for r in range(2)]
            assert nccl.ncclGroupStart() == 0
            for r in range(2):
                with torch.cuda.device(r):
                    _cabi.check(lib.emirwv_gather(
                        vp(local[r].data_ptr()) if counts[r] else None, carr, 2, r, h * w * 4,
                        vp(root_buf.data_ptr()) if r == 0 else None, comms[r], 0,
                        vp(streams[r].cuda_stream)))
            assert nccl.ncclGroupEnd() == 0
            for r in range(2):
                streams[r].synchronize()
            assert np.array_equal(root_buf.cpu().numpy(), np.concatenate(stacks))
        # mean product: float64 sums of both ranks added on rank 0
        npix = h * w
        sums = [torch.from_numpy(rng.normal(size=(2, npix))).to(f"cuda:{r}") for r in range(2)]
        want = (sums[0].cpu() + sums[1].cpu()).numpy()
        streams = [torch.cuda.Stream(device=r) for r in range(2)]
        assert nccl.ncclGroupStart() == 0
        for r in range(2):
            with torch.cuda.device(r):
                _cabi.check(lib.emirwv_abba_reduce(vp(sums[r][0].data_ptr()), vp(sums[r][1].data_ptr()),
                                                   npix, comms[r], 0, vp(streams[r].cuda_stream)))
        assert nccl.ncclGroupEnd() == 0
        for r in range(2):
            streams[r].synchronize()
        assert np.array_equal(sums[0].cpu().numpy(), want)
        with pytest.raises(ValueError):
            _cabi.check(lib.emirwv_gather(None, None, 2, 0, 4, None, comms[0], 0, None))
    finally:
        for r in range(2):
            nccl.ncclCommDestroy(comms[r])

#!/usr/bin/env python
"""Device-resident throughput of median_slitlets_rectified (modes 0 and 1) on one GPU.

usage: python scripts/bench_median.py [--frames 64] [--naxis1 3400] [--steps 50]
Prints one JSON line per mode: frames/s and the HBM bytes the kernel has to move
(read the product once; write 55 spectra (mode 1) or the full image (mode 0)).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemir_b200.median_slitlets_rectified import median_frames_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--naxis1", type=int, default=3400)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--dtype", default="float32")
    args = ap.parse_args()
    dtype = torch.float32 if args.dtype == "float32" else torch.float64
    x = torch.randn((args.frames, 2090, args.naxis1), device="cuda", dtype=dtype)
    peak = 6452.8
    try:
        with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    for mode in (1, 0):
        out = median_frames_torch(x, mode=mode)
        for _ in range(5):
            median_frames_torch(x, mode=mode, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            median_frames_torch(x, mode=mode, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        nbytes = x.numel() * x.element_size() + out.numel() * out.element_size()
        print(json.dumps({
            "metric": "median_slitlets_rectified frames/sec", "mode": mode,
            "value": args.frames / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
            "frames": args.frames, "dtype": args.dtype,
            "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak}}))


if __name__ == "__main__":
    main()

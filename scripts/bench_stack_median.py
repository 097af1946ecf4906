#!/usr/bin/env python
"""Device-resident throughput of the ABBA median combination on one GPU.

usage: python scripts/bench_stack_median.py [--frames 64] [--naxis1 3400] [--steps 30]
One JSON line per stack size: the stack holds `frames` rectified products, half A and half
B; a step = median of the A images + median of the B images + their float32 difference.
HBM bytes a step has to move: read every product once, write the float32 difference
(emirwv_abba_median takes both medians of a pixel in registers).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemir_b200 import abba  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="+", default=[8, 16, 32, 64, 128])
    ap.add_argument("--naxis1", type=int, default=3400)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--dtype", default="float32")
    args = ap.parse_args()
    dtype = torch.float32 if args.dtype == "float32" else torch.float64
    peak = 6452.8
    try:
        with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    for n in args.frames:
        x = torch.randn((n, 2090, args.naxis1), device="cuda", dtype=dtype)
        labels = abba.labels_of(abba.build_full_set("ABBA", 1, n // 4))
        for _ in range(3):
            abba.combine_abba_median(x, labels, [n])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            abba.combine_abba_median(x, labels, [n])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        npix = 2090 * args.naxis1
        nbytes = x.numel() * x.element_size() + npix * 4
        print(json.dumps({
            "metric": "ABBA median combination, rectified frames/sec", "images_per_position": n // 2,
            "value": n / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
            "frames": n, "dtype": args.dtype,
            "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak}}), flush=True)
        del x


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Device-resident throughput of the pre-reduction kernel (bias, dark, flat) on one GPU.

usage: python scripts/bench_prereduce.py [--frames 64] [--steps 30]
HBM bytes a step has to move: read every frame once, write it once (float32), read the three
calibration images once (they stay in L2 for the other frames).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemir_b200.prereduce import prereduce_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    peak = 6452.8
    try:
        with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as fh:
            peak = float(json.load(fh)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    cal = [torch.rand((2048, 2048), device="cuda") + 0.5 for _ in range(3)]
    for name, dtype in (("float32", torch.float32), ("uint16", torch.uint16)):
        if dtype == torch.uint16:
            x = torch.randint(0, 65535, (args.frames, 2048, 2048), device="cuda",
                              dtype=torch.int32).to(torch.uint16)
        else:
            x = torch.rand((args.frames, 2048, 2048), device="cuda") * 1000.0
        out = torch.empty((args.frames, 2048, 2048), device="cuda", dtype=torch.float32)
        for _ in range(3):
            prereduce_torch(x, *cal, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            prereduce_torch(x, *cal, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        nbytes = x.numel() * x.element_size() + out.numel() * 4 + 3 * 2048 * 2048 * 4
        print(json.dumps({
            "metric": "pre-reduction (bias, dark, flat) frames/sec", "in_dtype": name,
            "value": args.frames / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
            "frames": args.frames,
            "roofline": {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak}}), flush=True)
        del x, out


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Device-resident throughput of the kernels behind the ABBA product, one GPU:
per-frame vertical offsets (k_shift_rows), sigma-clip combination (k_abba_sigmaclip), the
pre-reduction with and without the bad-pixel stage (k_prereduce), and combine_abba as a whole.

One JSON line per kernel / stack size with the HBM roofline of what it has to move
(algorithmic bytes: every input element read once, every output element written once).
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemir_b200 import abba  # noqa: E402
from pyemir_b200.prereduce import prereduce_torch  # noqa: E402

H, W = 2090, 3400


def peak_gbs():
    try:
        with open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        return 6650.0


def timed(fn, steps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def line(kernel, frames, ms, nbytes, **extra):
    peak = peak_gbs()
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(kernel=kernel, frames=frames, ms_per_step=ms,
                          frames_per_s=frames / (ms * 1e-3),
                          roofline={"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s",
                                    "frac": gbs / peak, "bytes_per_launch": nbytes}, **extra)),
          flush=True)


def main():
    n = 64
    x = torch.randn((n, H, W), device="cuda") * 20 + 500
    full_set = abba.build_full_set("ABBA", 1, n // 4)
    # every frame shifted / half of them (the other half is copied)
    for name, offs in (("all frames shifted", [0.3127 + 0.01 * i for i in range(n)]),
                       ("half of the frames shifted", [0.0 if i % 2 else 0.41 for i in range(n)])):
        out = torch.empty_like(x)
        ms = timed(lambda: abba.shift_frames(x, offs, out=out))
        line("k_shift_rows<float,float>", n, ms, 2 * x.numel() * 4, what=name)
    del out
    big = torch.cat([x, x * 1.01])               # 128 products: 64 images per nodding position
    for m in (8, 16, 32, 64, 128):
        fs = abba.build_full_set("ABBA", 1, m // 4)
        sub = big[:m]
        for method in ("sigmaclip", "median", "mean"):
            ms = timed(lambda: abba.combine_abba(sub, fs, None, method))
            line("combine_abba " + method, m, ms, m * H * W * 4 + H * W * 4,
                 images_per_position=m // 2)
    del big
    offs = [0.0 if i % 4 in (0, 3) else 0.3127 for i in range(n)]
    for method in ("mean", "sigmaclip"):
        ms = timed(lambda: abba.combine_abba(x, full_set, offs, method), steps=10)
        line("combine_abba with offsets, " + method, n, ms, 3 * x.numel() * 4 + H * W * 4,
             what="shift (read + write) then combination (read): 3 passes over the stack")
    del x
    # pre-reduction of raw uint16 frames, with and without the bad-pixel stage
    rng = np.random.default_rng(1)
    raw = torch.from_numpy(rng.integers(0, 60000, size=(n, 2048, 2048)).astype(np.uint16)).cuda()
    bias = rng.normal(1000, 2, size=(2048, 2048)).astype(np.float32)
    flat = rng.normal(1, 0.05, size=(2048, 2048)).astype(np.float32)
    bpm = (rng.random((2048, 2048)) < 0.005).astype(np.uint8)
    out = torch.empty((n, 2048, 2048), device="cuda", dtype=torch.float32)
    cals = dict(bias=torch.from_numpy(bias).cuda(), dark=torch.from_numpy(bias * 0.001).cuda(),
                flat=torch.from_numpy(flat).cuda())
    mask = torch.from_numpy(bpm).cuda()
    for name, kw in (("bias, dark, flat", {}), ("bad pixels (0.5 %), bias, dark, flat", {"bpm": mask})):
        ms = timed(lambda: prereduce_torch(raw, out=out, **cals, **kw))
        line("k_prereduce<uint16,float,4>", n, ms, raw.numel() * 2 + out.numel() * 4, what=name)


if __name__ == "__main__":
    main()

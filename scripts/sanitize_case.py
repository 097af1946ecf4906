#!/usr/bin/env python
"""Small run of every kernel on the path, meant to be executed under compute-sanitizer
(scripts/sanitize.sh): BASELINE config 2 geometry restricted to a few slitlets (the first and
the last one touch the detector edge: zero-filled window columns and rows), 4 + 3 frames
(a full frame group and an odd tail), float32 and float64, nearest and area resampling,
device and host entries (reduced and raw frames, dense and sparse copy-back, the ABBA product),
and the side kernels (median, ABBA mean / median / sigma clipping, vertical offsets,
pre-reduction with the bad-pixel stage)."""

import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from pyemir_b200 import abba, synthetic
    from pyemir_b200.median_slitlets_rectified import median_frames_torch
    from pyemir_b200.plan import RectWavePlan

    keep = tuple(int(s) for s in os.environ.get("SANITIZE_SLITLETS", "1,28,55").split(","))
    coeff = copy.deepcopy(synthetic.make_config(2))
    coeff.missing_slitlets = [i for i in range(1, 56) if i not in keep]
    frames = synthetic.make_frames(coeff, 7, seed=9)
    checks = []
    for dtype, mode in (("float32", 2), ("float64", 2), ("float32", 1)):
        plan = RectWavePlan(coeff, mode, dtype)
        d_in = torch.from_numpy(frames.astype(plan.np_dtype)).cuda()
        out, jmm = plan.run_torch(d_in)
        torch.cuda.synchronize()
        host_out, host_jmm = plan.run_host(frames.astype(plan.np_dtype))
        same = np.array_equal(host_out, out.cpu().numpy()) and np.array_equal(host_jmm, jmm.cpu().numpy())
        checks.append((dtype, mode, bool(same), int(np.count_nonzero(host_out))))
        if dtype == "float32" and mode == 2:
            med = median_frames_torch(out, mode=1)
            diff = abba.combine_abba_mean(out[:4].contiguous(), abba.labels_of("ABBA"), 2, 2)
            dmed = abba.combine_abba_median(out[:4].contiguous(), abba.labels_of("ABBA"), [4])
            torch.cuda.synchronize()
            checks.append(("median", tuple(med.shape), "abba", tuple(diff.shape), tuple(dmed.shape)))
            # the rest of the ABBA product and the flow in front of the path
            offs = [0.0, 0.37, -1.2, 0.0]
            for method in ("mean", "median", "sigmaclip"):
                abba.combine_abba(out[:4].contiguous(), "ABBA", offs, method)
            rng = np.random.default_rng(2)
            raw = np.clip(frames * 8.0 + 1000.0, 0, 65535).astype(np.uint16)
            bpm = (rng.random(raw.shape[1:]) < 0.003).astype(np.uint8)
            bpm[0, 0] = bpm[2047, 2047] = 1
            flat = rng.normal(8.0, 0.2, size=raw.shape[1:]).astype(np.float32)
            plan.set_calibration(bpm=bpm, bias=flat * 100.0, flat=flat)
            raw_out, _ = plan.run_host_raw(raw)
            zeroed = np.zeros_like(raw_out)
            plan.run_host_raw(raw, out=zeroed, out_zeroed=True)
            img = plan.abba_host(raw[:4], abba.labels_of("ABBA"), offs, "sigmaclip", raw=True)
            torch.cuda.synchronize()
            checks.append(("raw", bool(np.array_equal(raw_out, zeroed)), tuple(img.shape)))
        plan.close()
    print("sanitize_case ok:", checks, flush=True)
    assert all(c[2] for c in checks if isinstance(c[2], bool))
    assert all(c[1] for c in checks if c[0] == "raw")


if __name__ == "__main__":
    main()

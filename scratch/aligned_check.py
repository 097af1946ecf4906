"""One-process check of the experimental lane mapping of k_rectwv (EMIRWV_ALIGNED=1):
bit-identity against the default kernel and device-resident throughput of both, config 2.

usage (GPU box): python scratch/aligned_check.py [frames] [steps]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyemir_b200 import synthetic  # noqa: E402
from pyemir_b200.plan import RectWavePlan  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    t0 = time.perf_counter()
    coeff = synthetic.make_config(2)
    fr = synthetic.make_frames(coeff, frames, seed=1002)
    d_in = torch.from_numpy(fr).cuda()
    res = {}
    plans = {}
    for name, flag in (("default", "0"), ("aligned", "1")):
        os.environ["EMIRWV_ALIGNED"] = flag
        plans[name] = RectWavePlan(coeff, 2, "float32", max_order=5)
        out, jmm = plans[name].run_torch(d_in)
        torch.cuda.synchronize()
        res[name] = (out, jmm)
    same = bool(torch.equal(res["default"][0], res["aligned"][0]) and
                torch.equal(res["default"][1], res["aligned"][1]))
    report = {"frames": frames, "bit_identical": same,
              "nonzero": int(torch.count_nonzero(res["default"][0][0]).item()),
              "setup_s": time.perf_counter() - t0}
    if not same:
        diff = (res["default"][0] != res["aligned"][0])
        report["differing_elements"] = int(diff.sum().item())
    for name in ("default", "aligned", "default", "aligned"):
        plan = plans[name]
        out, jmm = res[name]
        for _ in range(5):
            plan.run_torch(d_in, out, jmm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            plan.run_torch(d_in, out, jmm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        report.setdefault(name + "_frames_per_s", []).append(frames / (ms * 1e-3))
    print(json.dumps(report), flush=True)


if __name__ == "__main__":
    main()

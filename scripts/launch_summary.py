#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.

usage: python scripts/launch_summary.py launches.csv > profiles/r01_launches_summary.txt
"""
import collections
import csv
import statistics
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ik].replace("<unnamed>::", ""), []).append(float(r[iv]) / 1e3)
    total = sum(sum(v) for v in agg.values())
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)")
    print("# command: python bench.py --steps 3 --warmup 3 --cpu-frames 0 --e2e-steps 0"
          "   (64 frames per step, config 2)")
    print("# kernel | launches | median us | total us | share of GPU time")
    for k, v in agg.items():
        print(f"{k} | {len(v)} | {statistics.median(v):.1f} | {sum(v):.1f} | {100 * sum(v) / total:.1f}%")
    step = {k: statistics.median(v) for k, v in agg.items()
            if "k_init_jmm" in k or "k_zero_fill" in k or "k_rectwv" in k}
    if step:
        hot = [v for k, v in step.items() if "rectwv" in k][0]
        parts = " + ".join(f"{k.split('(')[0].split('<')[0].replace('void ', '')} {v:.1f} us"
                           for k, v in step.items())
        print(f"# one step (64 frames) = {parts} = {sum(step.values()):.1f} us serialised; "
              f"k_rectwv share {100 * hot / sum(step.values()):.1f}% "
              "(k_zero_fill overlaps with it on a side stream in normal runs)")


if __name__ == "__main__":
    main()

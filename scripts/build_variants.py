#!/usr/bin/env python
"""Build experiment variants of the library next to the product build (CPU, nvcc only).

    python scripts/build_variants.py x3            # one variant
    python scripts/build_variants.py               # all of them

A process picks a variant with EMIRWV_LIB=libemirwv_<name>.so (pyemir_b200/_cabi.py).
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SLIM = ("EMIRWV_SLIM=1",)         # float32 / K = 3 / G = 4 only: compiles in half a minute
VARIANTS = {
    "slim": SLIM,                                        # the product configuration, slim
    # round-1 occupancy: two CTAs per SM (96 registers), 12-row window of 288 columns
    "c2": SLIM + ("EMIRWV_MINBLOCKS=2", "EMIRWV_WIN_ROWS=12", "EMIRWV_WIN_PITCH=288"),
    "pf2": SLIM + ("EMIRWV_PREFETCH=2",), "pf4": SLIM + ("EMIRWV_PREFETCH=4",),
    "pf8": SLIM + ("EMIRWV_PREFETCH=8",),
    "recpf": SLIM + ("EMIRWV_RECPF=1",),
    "c3s3": SLIM + ("EMIRWV_NSTAGE=3", "EMIRWV_WIN_ROWS=10"),
    # float64 kernels compiled for two CTAs per SM (full build: every instantiation)
    "f64c2": ("EMIRWV_MINBLOCKS_F64=2",),
    # timing diagnostics (WRONG results): one gather tap instead of four / no knot traffic
    "diag4": SLIM + ("EMIRWV_DIAG=4",),     # every source-row copy as two requests (same bytes)
    "diag1": SLIM + ("EMIRWV_DIAG=1",), "diag2": SLIM + ("EMIRWV_DIAG=2",),
}


def main():
    from pyemir_b200 import _build
    names = sys.argv[1:] or sorted(VARIANTS)
    for name in names:
        path = _build.build_library(force=True, defines=VARIANTS[name],
                                    out=f"libemirwv_{name}.so")
        print(path, VARIANTS[name], flush=True)


if __name__ == "__main__":
    main()

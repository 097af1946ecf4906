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
    # three CTAs per SM: 72 registers, four rows in flight, ten-row source window (73 KB)
    "c3s4": SLIM + ("EMIRWV_MINBLOCKS=3", "EMIRWV_NSTAGE=4", "EMIRWV_WIN_ROWS=10"),
    "c3s3": SLIM + ("EMIRWV_MINBLOCKS=3", "EMIRWV_NSTAGE=3", "EMIRWV_WIN_ROWS=10"),
    # two CTAs per SM, deeper pipeline
    "c2s6": SLIM + ("EMIRWV_NSTAGE=6", "EMIRWV_WIN_ROWS=13"),
    "c2s5": SLIM + ("EMIRWV_NSTAGE=5", "EMIRWV_WIN_ROWS=12"),
    "c2s8": SLIM + ("EMIRWV_NSTAGE=8", "EMIRWV_WIN_ROWS=15"),
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

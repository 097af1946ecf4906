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

VARIANTS = {
    # three CTAs per SM: 72 registers, two rows in flight, nine-row source window (73 KB)
    "x3": ("EMIRWV_MINBLOCKS=3", "EMIRWV_NSTAGE=2", "EMIRWV_WIN_ROWS=9"),
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

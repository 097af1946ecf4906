"""Build recipe of the CUDA library (nvcc, sm_100a only, in-tree output)."""

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libemirwv.so")
HOSTCHECK_PATH = os.path.join(_HERE, "libemirwv_hostcheck.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-cudart", "static",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    mtime = os.path.getmtime(target)
    return any(os.path.getmtime(s) > mtime for s in sources)


def build_library(force=False, verbose=False, defines=(), out=None):
    """nvcc -gencode arch=compute_100a,code=sm_100a … -> pyemir_b200/libemirwv.so

    ``defines`` / ``out`` build an experiment variant next to the product library
    (e.g. ``defines=("EMIRWV_MINBLOCKS=3",), out="libemirwv_x.so"``); a process picks one
    with the environment variable ``EMIRWV_LIB`` (see ``_cabi.library_path``).
    """
    sources = [os.path.join(CSRC, name) for name in sorted(os.listdir(CSRC))
               if name.endswith((".cu", ".cuh", ".inl", ".inc"))]
    main = os.path.join(CSRC, "emirwv.cu")
    sources.append(os.path.join(ROOT, "include", "emirwv.h"))
    target = LIB_PATH if out is None else os.path.join(_HERE, out)
    if not force and not _stale(target, sources):
        return target
    cmd = ["nvcc"] + NVCC_FLAGS + ["-D" + d for d in defines] + [
        "-I", os.path.join(ROOT, "include"), "-o", target, main, "-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return target


def build_hostcheck(force=False):
    """g++ build of the shared geometry header, for CPU-side tests of the math."""
    sources = [os.path.join(CSRC, "hostcheck.cpp"), os.path.join(CSRC, "geom.cuh"),
               os.path.join(CSRC, "pixel_ops.cuh"), os.path.join(CSRC, "stack_median_nets.inc")]
    if not force and not _stale(HOSTCHECK_PATH, sources):
        return HOSTCHECK_PATH
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
           "-o", HOSTCHECK_PATH, sources[0]]
    subprocess.run(cmd, check=True)
    return HOSTCHECK_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_hostcheck(force=True))

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
    "--shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-cudart", "static",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    mtime = os.path.getmtime(target)
    return any(os.path.getmtime(s) > mtime for s in sources)


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a … -> pyemir_b200/libemirwv.so"""
    sources = [os.path.join(CSRC, name) for name in (
        "emirwv.cu", "geom.cuh", "pixel_ops.cuh", "plan_kernels.cuh", "device_utils.cuh", "side_kernels.cuh",
        "rectwv_kernel.cuh", "plan_build.inl", "median_net38.inc", "stack_median_nets.inc")]
    sources.append(os.path.join(ROOT, "include", "emirwv.h"))
    if not force and not _stale(LIB_PATH, sources):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"),
                                   "-o", LIB_PATH, sources[0]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


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

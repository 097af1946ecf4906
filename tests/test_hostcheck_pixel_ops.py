"""The per-pixel arithmetic of the side kernels, compiled for the host.

pyemir_b200/csrc/pixel_ops.cuh holds what k_stack_median / k_abba_median and k_prereduce do to
one pixel (padding rule + selection network; bias, dark and flat stages).  The CUDA kernels
include it; here the same functions (g++ build, libemirwv_hostcheck.so) are compared on the CPU
with numpy.median and with the oracle of the pre-reduction, bit for bit.
"""

import ctypes

import numpy as np
import pytest

from oracle import pipeline as oracle
from pyemir_b200 import _build

vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def hc():
    lib = ctypes.CDLL(_build.build_hostcheck())
    lib.hostcheck_stack_median.restype = ctypes.c_int
    lib.hostcheck_stack_median.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64, vp]
    lib.hostcheck_prereduce.restype = ctypes.c_int
    lib.hostcheck_prereduce.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int64,
                                        vp, vp, vp, ctypes.c_int]
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stack_median_every_stack_size(hc, dtype):
    """n = 1 .. 64 (every bucket, every padding split), ties, signed zeros, infinities, NaNs."""
    rng = np.random.default_rng(7)
    npix = 257
    for n in range(1, 65):
        x = rng.normal(100.0, 30.0, size=(n, npix)).astype(dtype)
        x[:, 0:40] = np.round(x[:, 0:40] / 10.0)                 # many ties
        x[::2, 40:50] = 0.0
        x[1::2, 40:50] = -0.0
        x[0, 50] = np.nan
        x[n - 1, 51] = np.nan
        x[0, 52] = np.inf
        x[n - 1, 53] = -np.inf
        x[:, 54] = np.inf
        x[:, 55] = -np.inf
        out = np.empty(npix, dtype=np.float64)
        assert hc.hostcheck_stack_median(_p(x), int(dtype is np.float64), n, npix, _p(out)) == 0
        with np.errstate(invalid="ignore"):
            want = np.median(x.astype(np.float64), axis=0)
        assert np.array_equal(out, want, equal_nan=True), n
    assert hc.hostcheck_stack_median(_p(x), 0, 65, npix, _p(out)) == -1


@pytest.mark.parametrize("img_dtype,cal_dtype", [("float32", "float32"), ("uint16", "float32"),
                                                 ("float32", "float64"), ("float64", "float32"),
                                                 ("uint16", "float64"), ("float64", "float64")])
def test_prereduce_stages_bit_exact(hc, img_dtype, cal_dtype):
    rng = np.random.default_rng(3)
    shape = (3, 37, 53)
    if img_dtype == "uint16":
        frames = rng.integers(0, 65535, size=shape).astype(np.uint16)
    else:
        frames = rng.normal(2000.0, 500.0, size=shape).astype(img_dtype)
        frames[0, 0, 0] = np.nan
        frames[1, 0, 1] = np.inf
    bias = rng.normal(100.0, 3.0, size=shape[1:]).astype(cal_dtype)
    dark = rng.normal(5.0, 1.0, size=shape[1:]).astype(cal_dtype)
    flat = rng.normal(1.0, 0.3, size=shape[1:]).astype(cal_dtype)
    flat[3, :20] = 0.0
    flat[4, :20] = -1.0
    flat[5, 0] = np.nan
    code = {"float32": 0, "float64": 1, "uint16": 2}
    npix = shape[1] * shape[2]
    for use in [(1, 1, 1), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]:
        kw = {k: v for k, v, u in zip(("bias", "dark", "flat"), (bias, dark, flat), use) if u}
        out = np.empty(shape, dtype=np.float32)
        rc = hc.hostcheck_prereduce(_p(frames), code[img_dtype], _p(out), shape[0], npix,
                                    _p(kw.get("bias")), _p(kw.get("dark")), _p(kw.get("flat")),
                                    code[cal_dtype])
        assert rc == 0
        want = oracle.prereduce(frames, **kw)
        assert np.array_equal(out, want, equal_nan=True), use

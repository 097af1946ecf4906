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
    lib.hostcheck_stack_sigmaclip.restype = ctypes.c_int
    lib.hostcheck_stack_sigmaclip.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int64,
                                              ctypes.c_double, ctypes.c_double, vp]
    lib.hostcheck_shift_rows.restype = None
    lib.hostcheck_shift_rows.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, vp]
    lib.hostcheck_prereduce_bpm.restype = ctypes.c_int
    lib.hostcheck_prereduce_bpm.argtypes = [vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, vp, vp, vp, vp, ctypes.c_int]
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


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sigmaclip_every_stack_size(hc, dtype):
    """n = 1 .. 64 with outliers that take one, two and three passes to clear; equal samples
    (std 0), NaN and infinite samples; narrow limits that would reject everything."""
    rng = np.random.default_rng(11)
    npix = 193
    for n in list(range(1, 20)) + [31, 32, 33, 63, 64]:
        x = rng.normal(100.0, 5.0, size=(n, npix)).astype(dtype)
        x[rng.integers(0, n, 40), rng.integers(0, 100, 40)] += 200.0     # cosmic rays
        x[rng.integers(0, n, 20), rng.integers(0, 100, 20)] -= 60.0
        if n > 12:
            x[0, 100:120] = 1.0e4                                    # several passes
            x[1, 100:120] = 1.0e3
            x[2, 100:120] = 4.0e2
        x[:, 120:130] = 7.0                                          # std = 0
        x[0, 130] = np.nan
        x[n - 1, 131] = np.inf
        for low, high in ((3.0, 3.0), (2.0, 1.5), (0.5, 0.5)):
            out = np.empty(npix, dtype=np.float64)
            assert hc.hostcheck_stack_sigmaclip(_p(x), int(dtype is np.float64), n, npix, low,
                                                high, _p(out)) == 0
            want = oracle.sigmaclip_stack(x, low, high)
            assert np.array_equal(out, want, equal_nan=True), (n, low, high)
    # a clipped mean really drops the outlier
    x = np.full((16, 4), 10.0, dtype=dtype)
    x += rng.normal(0, 0.1, size=x.shape).astype(dtype)
    x[3, :] = 1000.0
    out = np.empty(4)
    hc.hostcheck_stack_sigmaclip(_p(x), int(dtype is np.float64), 16, 4, 3.0, 3.0, _p(out))
    assert np.all(np.abs(out - 10.0) < 0.2)


@pytest.mark.parametrize("yoffset", [0.3712, -0.3712, 1.0, -2.5, 0.5, 17.0001, -3.9999, 1e-4,
                                     2090.5, -5000.0])
def test_shift_rows_has_the_bits_of_the_general_clipping(hc, yoffset):
    """shift_row / shift_area (what k_shift_rows evaluates) against the oracle's
    shift_image2d = rectify2d with a translation map through the general polygon clipping."""
    from oracle import numina_restated as nr
    rng = np.random.default_rng(5)
    img = rng.normal(1000.0, 300.0, size=(75, 131))
    img[10, 10] = 0.0
    out = np.empty_like(img)
    hc.hostcheck_shift_rows(_p(img), img.shape[0], img.shape[1], yoffset, _p(out))
    want = nr.shift_image2d(img, yoffset=yoffset)
    assert np.array_equal(out, want)
    if abs(yoffset) < 70:
        # flux is conserved away from the edges, and a unit shift is a copy
        k = int(np.ceil(abs(yoffset))) + 1
        assert abs(out[k:-k].sum() - img.sum()) < img.sum()
    if yoffset == 1.0:
        assert np.array_equal(out[1:], img[:-1]) and not out[0].any()


def test_shift_rows_far_columns_and_rows(hc):
    """the products x * y of the shoelace sum round differently at large coordinates: the last
    rows and columns of a full-size product (2090 x 3400)."""
    from oracle import numina_restated as nr
    rng = np.random.default_rng(6)
    img = np.zeros((2090, 3400))
    img[2000:, 3300:] = rng.normal(1000.0, 300.0, size=(90, 100))
    out = np.empty_like(img)
    hc.hostcheck_shift_rows(_p(img), 2090, 3400, 0.4321, _p(out))
    sub = np.zeros((2090, 3400))
    # the oracle on the same corner only (rows=...: the other rows stay zero)
    want = nr.rectify2d(img, [-0.0, 1.0, 0.0], [-0.4321, 0.0, 1.0], 2, rows=range(1995, 2090))
    assert np.array_equal(out[1995:], want[1995:])


@pytest.mark.parametrize("img_dtype,cal_dtype", [("float32", "float32"), ("uint16", "float32"),
                                                 ("float64", "float32"), ("uint16", "float64")])
def test_prereduce_with_bad_pixel_stage_bit_exact(hc, img_dtype, cal_dtype):
    rng = np.random.default_rng(4)
    shape = (2, 41, 47)
    if img_dtype == "uint16":
        frames = rng.integers(0, 65535, size=shape).astype(np.uint16)
    else:
        frames = rng.normal(2000.0, 500.0, size=shape).astype(img_dtype)
        frames[0, 20, 21] = np.nan               # a NaN neighbour of a bad pixel
    mask = (rng.random(shape[1:]) < 0.05).astype(np.uint8)
    mask[0, 0] = mask[0, 1] = mask[40, 46] = 1   # corners and edges
    mask[20, 20] = 3                             # any non-zero value is bad
    mask[28:35, 28:35] = 1                       # a block: its centre has no good neighbour
    bias = rng.normal(100.0, 3.0, size=shape[1:]).astype(cal_dtype)
    flat = rng.normal(1.0, 0.1, size=shape[1:]).astype(cal_dtype)
    code = {"float32": 0, "float64": 1, "uint16": 2}
    for kw in ({}, {"bias": bias, "flat": flat}):
        out = np.empty(shape, dtype=np.float32)
        rc = hc.hostcheck_prereduce_bpm(_p(frames), code[img_dtype], _p(out), shape[0], shape[1],
                                        shape[2], _p(mask), _p(kw.get("bias")), None,
                                        _p(kw.get("flat")), code[cal_dtype])
        assert rc == 0
        want = oracle.prereduce_bpm(frames, mask, **kw)
        assert np.array_equal(out, want, equal_nan=True)
        assert out[0, 31, 31] == (0.0 if not kw else np.float32(
            np.float32(np.float32(0.0) - bias[31, 31]) / flat[31, 31]))

"""The geometry header shared with the CUDA kernels, compiled for the host.

pyemir_b200/csrc/geom.cuh is used by the plan-building kernels; here the same
functions (g++ build, libemirwv_hostcheck.so) are compared with the oracle on the
CPU: polynomial map bit-identical to numpy's evaluation order, clipping areas
bit-identical to oracle/clippix.c, Steffen pieces against the interpolator.
"""

import ctypes

import numpy as np
import pytest

from oracle import numina_restated as nr
from oracle.build import load_clippix
from pyemir_b200 import _build, synthetic

dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hc():
    lib = ctypes.CDLL(_build.build_hostcheck())
    lib.hostcheck_quad_box_area.restype = ctypes.c_double
    lib.hostcheck_quad_box_area.argtypes = [dp, dp] + [ctypes.c_double] * 4
    lib.hostcheck_fmap.restype = None
    lib.hostcheck_fmap.argtypes = [ctypes.c_int, dp, dp, dp, dp, ctypes.c_int64, dp, dp]
    lib.hostcheck_steffen_slope.restype = ctypes.c_double
    lib.hostcheck_steffen_slope.argtypes = [ctypes.c_double] * 3
    lib.hostcheck_steffen_partial.restype = ctypes.c_double
    lib.hostcheck_steffen_partial.argtypes = [ctypes.c_double] * 5
    lib.hostcheck_quad_extent.restype = None
    lib.hostcheck_quad_extent.argtypes = [dp, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int)]
    return lib


def _ptr(a):
    return a.ctypes.data_as(dp)


@pytest.mark.parametrize("number,order", [(1, 2), (3, 5)])
def test_fmap_bit_identical_to_numpy(hc, number, order):
    coeff = synthetic.make_config(number)
    entry = coeff.contents[40]
    a = np.array(entry["ttd_aij"], dtype=float)
    b = np.array(entry["ttd_bij"], dtype=float)
    x = np.concatenate([np.arange(2048.0), np.arange(2048.0) + 0.5, np.arange(2048.0) - 0.5])
    y = np.resize(np.array([3.0, 17.5, 38.5, 22.0, 0.5]), x.size).astype(float)
    u_ref, v_ref = nr.fmap(order, a, b, x, y)
    u = np.empty_like(x)
    v = np.empty_like(x)
    hc.hostcheck_fmap(order, _ptr(a), _ptr(b), _ptr(x), _ptr(y), x.size, _ptr(u), _ptr(v))
    assert np.array_equal(u, u_ref) and np.array_equal(v, v_ref)


def test_clip_area_bit_identical_to_oracle(hc, rng):
    ora = load_clippix()
    for _ in range(4000):
        cx, cy = rng.uniform(0, 2048, 2)
        ang = rng.uniform(-0.1, 0.1)
        sx, sy = rng.uniform(0.8, 1.6, 2)
        corners = np.array([[-.5, -.5], [-.5, .5], [.5, .5], [.5, -.5]]) * [sx, sy]
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        q = corners @ rot.T + [cx, cy]
        qx = np.ascontiguousarray(q[:, 0])
        qy = np.ascontiguousarray(q[:, 1])
        j = np.floor(cx + rng.integers(-1, 2))
        i = np.floor(cy + rng.integers(-1, 2))
        box = (j - 0.5, j + 0.5, i - 0.5, i + 0.5)
        got = hc.hostcheck_quad_box_area(_ptr(qx), _ptr(qy), *box)
        want = ora.oracle_poly_box_area(_ptr(qx), _ptr(qy), 4, *box)
        assert got == want


def test_quad_extent(hc):
    lo, hi = ctypes.c_int(), ctypes.c_int()
    for q, want in (([4.5, 4.5, 5.5, 5.5], (5, 5)), ([4.4, 4.4, 5.6, 5.6], (4, 6)),
                    ([4.6, 4.6, 5.4, 5.4], (5, 5)), ([-0.7, -0.7, 0.2, 0.2], (-1, 0))):
        arr = np.array(q, dtype=float)
        hc.hostcheck_quad_extent(_ptr(arr), ctypes.byref(lo), ctypes.byref(hi))
        assert (lo.value, hi.value) == want


def test_steffen_pieces_match_interpolator(hc, rng):
    x = np.cumsum(rng.uniform(0.5, 1.5, 40))
    flux = rng.normal(5.0, 4.0, 39)              # both signs: the limiter must engage
    y = np.concatenate([[0.0], np.cumsum(flux)])
    interp = nr.SteffenInterpolator(x, y, extrapolate="border")
    h = np.diff(x)
    s = flux / h
    yp = np.zeros(40)
    for i in range(1, 39):
        yp[i] = hc.hostcheck_steffen_slope(s[i - 1], s[i], h[i] / (h[i - 1] + h[i]))
    assert np.allclose(yp, interp._yp, rtol=1e-13, atol=1e-13)
    for i in rng.integers(0, 39, 200):
        t = rng.uniform(0, h[i])
        got = y[i] + hc.hostcheck_steffen_partial(s[i], yp[i], yp[i + 1], t, t / h[i])
        assert got == pytest.approx(float(interp(np.array([x[i] + t]))[0]), rel=1e-12, abs=1e-12)


def test_flux_unit_form_equals_slope_form(hc, rng):
    """The kernels' flux-unit Steffen piece (only tau needed) is the slope form."""
    hc.hostcheck_steffen_partial_flux.restype = ctypes.c_double
    hc.hostcheck_steffen_partial_flux.argtypes = [ctypes.c_double] * 4
    h = rng.uniform(0.5, 1.5, 3)                       # widths of pixels i-1, i, i+1
    for _ in range(500):
        flux = rng.normal(3.0, 5.0, 3)
        s = flux / h
        yp0 = hc.hostcheck_steffen_slope(s[0], s[1], h[1] / (h[0] + h[1]))
        yp1 = hc.hostcheck_steffen_slope(s[1], s[2], h[2] / (h[1] + h[2]))
        # homogeneity: h_i * steffen(s_l, s_r) = steffen(h_i s_l, h_i s_r)
        ya = hc.hostcheck_steffen_slope(flux[0] * (h[1] / h[0]), flux[1], h[1] / (h[0] + h[1]))
        yb = hc.hostcheck_steffen_slope(flux[1], flux[2] * (h[1] / h[2]), h[2] / (h[1] + h[2]))
        assert ya == pytest.approx(h[1] * yp0, rel=1e-13, abs=1e-13)
        assert yb == pytest.approx(h[1] * yp1, rel=1e-13, abs=1e-13)
        tau = rng.uniform(0, 1)
        want = hc.hostcheck_steffen_partial(s[1], yp0, yp1, tau * h[1], tau)
        got = hc.hostcheck_steffen_partial_flux(flux[1], ya, yb, tau)
        assert got == pytest.approx(want, rel=1e-12, abs=1e-12)
    assert hc.hostcheck_steffen_partial_flux(7.0, 1.0, 2.0, 0.0) == 0.0
    assert hc.hostcheck_steffen_partial_flux(7.0, 1.0, 2.0, 1.0) == pytest.approx(7.0, rel=1e-15)

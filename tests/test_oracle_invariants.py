"""numina-independent invariants of the oracle.

The oracle is a restatement of third-party code that cannot be imported here
(PARITY UNPINNED), so these analytic properties are what vouches for it.
"""

import ctypes

import numpy as np
import pytest

from oracle import numina_restated as nr
from oracle.build import load_clippix

IDENT_A = [0.0, 1.0, 0.0, 0.0, 0.0, 0.0]
IDENT_B = [0.0, 0.0, 1.0, 0.0, 0.0, 0.0]


def _area(xs, ys, x0, x1, y0, y1):
    lib = load_clippix()
    dp = ctypes.POINTER(ctypes.c_double)
    xs = np.ascontiguousarray(xs, dtype=float)
    ys = np.ascontiguousarray(ys, dtype=float)
    return lib.oracle_poly_box_area(xs.ctypes.data_as(dp), ys.ctypes.data_as(dp),
                                    len(xs), x0, x1, y0, y1)


def test_clip_known_areas():
    sq = ([0.0, 0.0, 1.0, 1.0], [0.0, 1.0, 1.0, 0.0])
    assert _area(*sq, 0.0, 1.0, 0.0, 1.0) == 1.0
    assert _area(*sq, 0.5, 2.0, 0.25, 2.0) == pytest.approx(0.5 * 0.75, abs=1e-15)
    assert _area(*sq, 1.0, 2.0, 0.0, 1.0) == 0.0
    # diamond |x|+|y| <= 1 against the unit box [0,1]^2: a triangle of area 1/2
    assert _area([1.0, 0.0, -1.0, 0.0], [0.0, 1.0, 0.0, -1.0], 0.0, 1.0, 0.0, 1.0) == \
        pytest.approx(0.5, abs=1e-15)
    # orientation does not matter
    assert _area([1.0, 0.0, -1.0, 0.0][::-1], [0.0, 1.0, 0.0, -1.0][::-1],
                 0.0, 1.0, 0.0, 1.0) == pytest.approx(0.5, abs=1e-15)


def test_order_fmap():
    assert [nr.order_fmap(n) for n in (3, 6, 10, 15)] == [1, 2, 3, 4]
    with pytest.raises(ValueError):
        nr.order_fmap(21)
    assert nr.order_fmap(21, nmax_order=5) == 5
    with pytest.raises(ValueError):
        nr.order_fmap(7)


def test_fmap_term_order():
    x = np.array([2.0])
    y = np.array([3.0])
    a = [1, 10, 100, 1000, 10000, 100000]        # 1, x, y, x², xy, y²
    u, _ = nr.fmap(2, a, a, x, y)
    assert u[0] == 1 + 10 * 2 + 100 * 3 + 1000 * 4 + 10000 * 6 + 100000 * 9


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_identity_map_returns_input(rng, mode):
    img = rng.normal(100.0, 20.0, (12, 40))
    out = nr.rectify2d(img, IDENT_A, IDENT_B, mode)
    assert np.array_equal(out, img) if mode != 2 else np.allclose(out, img, rtol=0, atol=1e-13)


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_integer_translation(rng, mode):
    img = rng.normal(100.0, 20.0, (12, 40))
    a = [3.0, 1.0, 0.0, 0.0, 0.0, 0.0]           # u = x + 3
    b = [-2.0, 0.0, 1.0, 0.0, 0.0, 0.0]          # v = y - 2
    out = nr.rectify2d(img, a, b, mode)
    expect = np.zeros_like(img)
    expect[2:, :-3] = img[:-2, 3:]
    assert np.allclose(out, expect, rtol=0, atol=1e-12)


def test_mode1_rounds_half_to_even():
    img = np.arange(6, dtype=float).reshape(1, 6) + 1.0
    a = [0.5, 1.0, 0.0, 0.0, 0.0, 0.0]           # u = x + 0.5 -> rint: 0,2,2,4,4,6
    out, iv, ju, ok = nr.rectify2d(img, a, IDENT_B, 1, return_maps=True)
    assert ju.tolist() == [[0, 2, 2, 4, 4, 6]]
    assert ok.tolist() == [[True, True, True, True, True, False]]
    assert out.tolist() == [[1.0, 3.0, 3.0, 5.0, 5.0, 0.0]]


def test_mode2_constant_image_gives_quad_area(rng):
    # affine map with shear and stretch: area of every mapped pixel = |det|
    a = [2.3, 1.03, 0.07, 0.0, 0.0, 0.0]
    b = [1.7, -0.02, 0.96, 0.0, 0.0, 0.0]
    det = 1.03 * 0.96 - 0.07 * (-0.02)
    img = np.full((30, 60), 5.0)
    out = nr.rectify2d(img, a, b, 2)
    inner = out[6:20, 6:40]                       # quads fully inside the image
    assert np.allclose(inner, 5.0 * det, rtol=1e-13)


def test_mode2_flux_conservation_shear(rng):
    # |det| = 1 shear: total flux of a blob well inside the image is preserved
    a = [0.0, 1.0, 0.25, 0.0, 0.0, 0.0]
    b = [0.0, 0.0, 1.0, 0.0, 0.0, 0.0]
    img = np.zeros((40, 80))
    img[15:25, 30:50] = rng.uniform(1.0, 2.0, (10, 20))
    out = nr.rectify2d(img, a, b, 2)
    assert out.sum() == pytest.approx(img.sum(), rel=1e-12)


def test_steffen_reproduces_knots_and_is_monotone(rng):
    x = np.cumsum(rng.uniform(0.5, 1.5, 50))
    y = np.cumsum(rng.uniform(0.0, 3.0, 50))
    s = nr.SteffenInterpolator(x, y, extrapolate="border")
    assert np.allclose(s(x), y, rtol=1e-14)
    fine = np.linspace(x[0] - 2, x[-1] + 2, 5000)
    vals = s(fine)
    assert np.all(np.diff(vals) >= -1e-12)
    assert vals[0] == y[0] and vals[-1] == y[-1]


def test_rebin_identity_when_grids_coincide(rng):
    nchan = 64
    img = rng.uniform(10.0, 20.0, (3, nchan))
    crval1, cdelt1 = 5000.0, 2.0
    # wavelength of 1-based pixel p: crval1 + cdelt1*(p-1)
    out = nr.resample_image2d_flux(img, nchan, cdelt1, crval1, 1.0,
                                   [crval1 - cdelt1, cdelt1])
    assert np.allclose(out, img, rtol=1e-12)


def test_rebin_flux_conservation_and_exact_zeros(rng):
    nchan = 200
    img = rng.uniform(10.0, 20.0, (2, nchan))
    coeff = [5000.0, 1.9, 2.0e-4]
    out = nr.resample_image2d_flux(img, 400, 1.5, 4950.0, 1.0, coeff)
    assert np.allclose(out.sum(axis=1), img.sum(axis=1), rtol=1e-12)
    assert np.all(out >= 0.0)
    wl_lo = np.polynomial.polynomial.polyval(0.5, coeff)
    wl_hi = np.polynomial.polynomial.polyval(nchan + 0.5, coeff)
    right_edges = 4950.0 + 1.5 * (np.arange(400) + 0.5)
    left_edges = right_edges - 1.5
    assert np.all(out[:, right_edges < wl_lo] == 0.0)
    assert np.all(out[:, left_edges > wl_hi] == 0.0)
    assert np.all(out[:, (left_edges > wl_lo) & (right_edges < wl_hi)] > 0.0)


def test_find_pix_borders():
    assert nr.find_pix_borders(np.zeros(5), 0) == (-1, 5)
    assert nr.find_pix_borders(np.array([0, 0, 3.0, 0, 1.0, 0]), 0) == (2, 4)
    assert nr.find_pix_borders(np.array([1.0, 0, 0]), 0) == (0, 0)


def test_apply_integer_offsets():
    img = np.arange(12.0).reshape(3, 4)
    out = nr.apply_integer_offsets(img, 1, -1)
    expect = np.zeros((3, 4))
    expect[0:2, 1:4] = img[1:3, 0:3]
    assert np.array_equal(out, expect) and out.dtype == np.float64
    assert np.array_equal(nr.apply_integer_offsets(img, 0, 0), img)
    with pytest.raises(ValueError):
        nr.apply_integer_offsets(img, 1.0, 0)
    with pytest.raises(ValueError):
        nr.apply_integer_offsets(img, np.int64(1), 0)


# ------------------------------------------------------------------ independent area check
def _random_map(rng, order):
    """Near-identity polynomial map of the given order (term order 1, x, y, x², xy, …) with
    shear, stretch and curvature of the size a strongly distorted slitlet shows."""
    ncoef = (order + 1) * (order + 2) // 2
    a = np.zeros(ncoef)
    b = np.zeros(ncoef)
    a[0], a[1], a[2] = rng.uniform(-3, 3), rng.uniform(0.8, 1.3), rng.uniform(-0.2, 0.2)
    b[0], b[1], b[2] = rng.uniform(-3, 3), rng.uniform(-0.1, 0.1), rng.uniform(0.8, 1.3)
    k = 3
    for i in range(2, order + 1):
        for _ in range(i + 1):
            a[k] = rng.uniform(-1, 1) * 0.3 / 200.0 ** i       # sub-pixel over 200 px
            b[k] = rng.uniform(-1, 1) * 0.3 / 200.0 ** i
            k += 1
    return a.tolist(), b.tolist()


@pytest.mark.parametrize("order", [2, 3, 4, 5])
def test_overlap_areas_match_scanline_integration(order):
    """oracle/clippix.c (Sutherland-Hodgman + shoelace) against tests/area_scanline.py
    (integration of the chord length, no clipping) on random maps of order 2 … 5: every
    overlap area, and their sum against the shoelace area of the mapped pixel, to the rounding
    of the clipper's shoelace sum in absolute coordinates (a few eps x |x| x |y|: products
    of coordinates ~1e4 are added and cancel down to an area below 1)."""
    import area_scanline as scan
    rng = np.random.default_rng(900 + order)
    lib = load_clippix()
    dp = ctypes.POINTER(ctypes.c_double)
    worst = 0.0
    for _ in range(6):
        aij, bij = _random_map(rng, order)
        for _ in range(25):
            j, i = rng.integers(5, 195), rng.integers(5, 35)
            xs = np.array([j - .5, j - .5, j + .5, j + .5])
            ys = np.array([i - .5, i + .5, i + .5, i - .5])
            qx, qy = nr.fmap(order, aij, bij, xs, ys)
            j0 = int(np.floor(qx.min() + 0.5))
            i0 = int(np.floor(qy.min() + 0.5))
            nx = int(np.floor(qx.max() + 0.5)) - j0 + 1
            ny = int(np.floor(qy.max() + 0.5)) - i0 + 1
            got = np.zeros((ny, nx))
            qx = np.ascontiguousarray(qx)
            qy = np.ascontiguousarray(qy)
            lib.oracle_quad_weights(qx.ctypes.data_as(dp), qy.ctypes.data_as(dp), i0, j0, ny, nx,
                                    got.ctypes.data_as(dp))
            want = scan.quad_weights(qx, qy, i0, j0, ny, nx)
            tol = 8 * np.finfo(float).eps * np.abs(qx).max() * np.abs(qy).max()
            worst = max(worst, float(np.abs(got - want).max()) / tol)
            assert np.allclose(got, want, rtol=0, atol=tol)
            # (the reference sum: shoelace about the quad's own centre, no cancellation)
            area = scan.shoelace(qx - qx.mean(), qy - qy.mean())
            assert got.sum() == pytest.approx(area, abs=4 * tol)
            assert want.sum() == pytest.approx(area, abs=1e-13)
    assert 0.0 < worst < 1.0


def test_scanline_area_known_answers():
    import area_scanline as scan
    sq = ([0.0, 0.0, 1.0, 1.0], [0.0, 1.0, 1.0, 0.0])
    assert scan.quad_box_area(*sq, 0.0, 1.0, 0.0, 1.0) == 1.0
    assert scan.quad_box_area(*sq, 0.5, 2.0, 0.25, 2.0) == pytest.approx(0.375, abs=1e-15)
    assert scan.quad_box_area(*sq, 1.0, 2.0, 0.0, 1.0) == 0.0
    diamond = ([1.0, 0.0, -1.0, 0.0], [0.0, 1.0, 0.0, -1.0])
    assert scan.quad_box_area(*diamond, 0.0, 1.0, 0.0, 1.0) == pytest.approx(0.5, abs=1e-15)
    assert scan.quad_box_area(*diamond, -1.0, 1.0, -1.0, 1.0) == pytest.approx(2.0, abs=1e-15)
    # a sheared parallelogram of area 1 spread over three unit boxes
    para = ([0.0, 0.5, 1.5, 1.0], [0.0, 1.0, 1.0, 0.0])
    parts = [scan.quad_box_area(*para, k, k + 1.0, 0.0, 1.0) for k in (-1.0, 0.0, 1.0, 2.0)]
    assert parts[0] == 0.0 and parts[3] == 0.0
    assert parts[1] == pytest.approx(0.75, abs=1e-15) and parts[2] == pytest.approx(0.25, abs=1e-15)

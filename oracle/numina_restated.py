"""ORACLE — numpy restatement of the numina array kernels on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); PARITY UNPINNED.

numina is a third-party dependency of the reference (``pyproject.toml:32``,
``numina>=0.36.0``) that is absent from ``/root/reference`` and from this image.
Each function below restates the published algorithm of the numina function
named in its docstring and cites the reference call site it serves.  float64,
numpy operation order kept (one rounding per ufunc, no fused multiply-add).
"""

import ctypes

import numpy as np
from numpy.polynomial.polynomial import polyval

from .build import load_clippix

# numina caps the order at 4; BASELINE config 3 asks for order 5, which the same
# formula covers.  The cap is therefore an argument.
NMAX_ORDER = 4


# ---------------------------------------------------------------------------
# numina.array.wavecalib.apply_integer_offsets  (apply_rectwv_coeff.py:74-78)
# ---------------------------------------------------------------------------
def apply_integer_offsets(image2d, offx, offy):
    """Shift by whole pixels with zero fill; result is float64."""
    if type(offx) is not int or type(offy) is not int:
        raise ValueError("Invalid non-integer offsets")
    naxis2, naxis1 = image2d.shape
    shifted = np.zeros((naxis2, naxis1))

    def lo(s):
        return max(0, s)

    def hi(s):
        return s if s < 0 else None

    shifted[lo(offy):hi(offy), lo(offx):hi(offx)] = \
        image2d[lo(-offy):hi(-offy), lo(-offx):hi(-offx)]
    return shifted


# ---------------------------------------------------------------------------
# numina.array.distortion: ncoef_fmap / order_fmap / fmap  (slitlet2d.py:211-212)
# ---------------------------------------------------------------------------
def ncoef_fmap(order):
    ncoef = 0
    for i in range(order + 1):
        for _ in range(i + 1):
            ncoef += 1
    return ncoef


def order_fmap(ncoef, nmax_order=NMAX_ORDER):
    order = 1
    while ncoef != ncoef_fmap(order):
        order += 1
        if order > nmax_order:
            raise ValueError("order > " + str(nmax_order) + " not implemented")
    return order


def fmap(order, aij, bij, x, y):
    """2-D polynomial map (x, y) -> (u, v); term order 1, x, y, x², xy, y², …"""
    u = np.zeros_like(x)
    v = np.zeros_like(y)
    k = 0
    for i in range(order + 1):
        for j in range(i + 1):
            u += aij[k] * (x ** (i - j)) * (y ** j)
            v += bij[k] * (x ** (i - j)) * (y ** j)
            k += 1
    return u, v


# ---------------------------------------------------------------------------
# numina.array.distortion.rectify2d  (slitlet2d.py:421-423)
# ---------------------------------------------------------------------------
def rectify2d(image2d, aij, bij, resampling, naxis1out=None, naxis2out=None,
              nmax_order=NMAX_ORDER, rows=None, return_maps=False):
    """Rectify ``image2d`` with the polynomial map (rectified -> original).

    resampling 1: nearest neighbour; 2: flux-preserving area overlap (the Cython
    ``_clippix._resample`` in numina, ``oracle/clippix.c`` here).
    resampling 3 is NOT in the reference: true bilinear interpolation at the
    mapped pixel centre (BASELINE.json names it; kept as a labelled extension).

    ``rows`` restricts the computation to some output rows (the others stay 0);
    the reference always computes all of them — this only saves test time.
    ``return_maps`` (mode 1) also returns the integer index maps (iv, ju, ok).
    """
    ncoef = len(aij)
    if len(bij) != ncoef:
        raise ValueError("Unexpected different lengths of aij and bij")
    order = order_fmap(ncoef, nmax_order)
    aij = [float(c) for c in aij]
    bij = [float(c) for c in bij]

    naxis2, naxis1 = image2d.shape
    if naxis1out is None:
        naxis1out = naxis1
    if naxis2out is None:
        naxis2out = naxis2
    image2d = np.ascontiguousarray(image2d, dtype=float)
    row_ids = np.arange(naxis2out) if rows is None else np.asarray(rows, dtype=int)

    image2d_rect = np.zeros((naxis2out, naxis1out), dtype=float)

    if resampling == 1:
        j = np.arange(0, naxis1out, dtype=float)
        i = row_ids.astype(float)
        xx = np.tile(j, (len(i),))
        yy = np.repeat(i, len(j))
        xxx, yyy = fmap(order, aij, bij, xx, yy)
        ixxx = np.rint(xxx).astype(int)
        iyyy = np.rint(yyy).astype(int)
        lxxx = np.logical_and(ixxx >= 0, ixxx < naxis1)
        lyyy = np.logical_and(iyyy >= 0, iyyy < naxis2)
        lok = np.logical_and(lxxx, lyyy)
        ixx = xx.astype(int)[lok]
        iyy = yy.astype(int)[lok]
        image2d_rect[iyy, ixx] = image2d[iyyy[lok], ixxx[lok]]
        if return_maps:
            shape = (len(i), naxis1out)
            return (image2d_rect, iyyy.reshape(shape), ixxx.reshape(shape),
                    lok.reshape(shape))
        return image2d_rect

    if resampling == 2:
        # four corners of every output pixel, anticlockwise:
        # (j-.5,i-.5) (j-.5,i+.5) (j+.5,i+.5) (j+.5,i-.5)
        dj = np.array([-0.5, -0.5, 0.5, 0.5])
        di = np.array([-0.5, 0.5, 0.5, -0.5])
        jj, ii = np.meshgrid(np.arange(naxis1out), row_ids)
        jj = jj.ravel()
        ii = ii.ravel()
        xx = (jj[:, None].astype(float) + dj[None, :]).ravel()
        yy = (ii[:, None].astype(float) + di[None, :]).ravel()
        xxx, yyy = fmap(order, aij, bij, xx, yy)
        xxx = np.ascontiguousarray(xxx)
        yyy = np.ascontiguousarray(yyy)
        ixx = np.ascontiguousarray(jj, dtype=np.int64)
        iyy = np.ascontiguousarray(ii, dtype=np.int64)
        lib = load_clippix()
        dp = ctypes.POINTER(ctypes.c_double)
        lp = ctypes.POINTER(ctypes.c_long)
        lib.oracle_resample(
            image2d.ctypes.data_as(dp), naxis2, naxis1,
            xxx.ctypes.data_as(dp), yyy.ctypes.data_as(dp), len(ixx),
            ixx.ctypes.data_as(lp), iyy.ctypes.data_as(lp),
            image2d_rect.ctypes.data_as(dp), naxis2out, naxis1out)
        return image2d_rect

    if resampling == 3:
        # extension, not in the reference: bilinear sample at the mapped centre;
        # samples outside the image count as zero
        j = np.arange(0, naxis1out, dtype=float)
        i = row_ids.astype(float)
        xx = np.tile(j, (len(i),))
        yy = np.repeat(i, len(j))
        u, v = fmap(order, aij, bij, xx, yy)
        u0 = np.floor(u)
        v0 = np.floor(v)
        fu = u - u0
        fv = v - v0
        acc = np.zeros_like(u)
        for dv, wv in ((0, 1.0 - fv), (1, fv)):
            for du, wu in ((0, 1.0 - fu), (1, fu)):
                cu = (u0 + du).astype(int)
                cv = (v0 + dv).astype(int)
                ok = (cu >= 0) & (cu < naxis1) & (cv >= 0) & (cv < naxis2)
                val = np.zeros_like(u)
                val[ok] = image2d[cv[ok], cu[ok]]
                acc += (wv * wu) * val
        image2d_rect[yy.astype(int), xx.astype(int)] = acc
        return image2d_rect

    raise ValueError("Sorry, resampling method must be 1 or 2")


# ---------------------------------------------------------------------------
# numina.array.interpolation.SteffenInterpolator
# ---------------------------------------------------------------------------
class SteffenInterpolator:
    """Monotone cubic Hermite interpolation (Steffen 1990, A&A 239, 443).

    End derivatives default to 0.0; ``extrapolate='border'`` returns the first /
    last ordinate outside the knots.
    """

    def __init__(self, x, y, yp_0=0.0, yp_N=0.0, extrapolate="raise",
                 fill_value=np.nan):
        self._x = np.asarray(x, dtype=float)
        self._y = np.asarray(y, dtype=float)
        if self._x.ndim != 1 or self._x.shape != self._y.shape:
            raise ValueError("x and y must be 1-D arrays of the same length")
        if self._x.size < 2:
            raise ValueError("at least two knots are needed")
        self.extrapolate = extrapolate
        self.fill_value = fill_value

        h = self._x[1:] - self._x[:-1]
        if np.any(h <= 0):
            raise ValueError("x must be strictly increasing")
        s = (self._y[1:] - self._y[:-1]) / h
        p = (s[:-1] * h[1:] + s[1:] * h[:-1]) / (h[1:] + h[:-1])
        yp = np.empty_like(self._x)
        yp[1:-1] = (np.sign(s[:-1]) + np.sign(s[1:])) * np.minimum(
            np.minimum(np.abs(s[:-1]), np.abs(s[1:])), 0.5 * np.abs(p))
        yp[0] = yp_0
        yp[-1] = yp_N
        self._yp = yp
        self._a = (yp[:-1] + yp[1:] - 2.0 * s) / (h * h)
        self._b = (3.0 * s - 2.0 * yp[:-1] - yp[1:]) / h
        self._c = yp[:-1]
        self._d = self._y[:-1]

    def __call__(self, x_new):
        x_new = np.asarray(x_new, dtype=float)
        below = x_new < self._x[0]
        above = x_new > self._x[-1]
        if self.extrapolate == "raise" and (below.any() or above.any()):
            raise ValueError("x_new outside the interpolation range")
        idx = np.searchsorted(self._x, x_new)
        idx = idx.clip(1, len(self._x) - 1) - 1
        t = x_new - self._x[idx]
        y_new = ((self._a[idx] * t + self._b[idx]) * t + self._c[idx]) * t + self._d[idx]
        if self.extrapolate == "border":
            y_new = np.where(below, self._y[0], y_new)
            y_new = np.where(above, self._y[-1], y_new)
        elif self.extrapolate == "zeros":
            y_new = np.where(below | above, 0.0, y_new)
        elif self.extrapolate == "const":
            y_new = np.where(below | above, self.fill_value, y_new)
        return y_new


# ---------------------------------------------------------------------------
# numina.array.wavecalib.resample  (apply_rectwv_coeff.py:152-159)
# ---------------------------------------------------------------------------
def map_borders(wls):
    """Pixel borders midway between consecutive centres, ends mirrored."""
    midpt_wl = 0.5 * (wls[1:] + wls[:-1])
    all_borders = np.zeros((wls.shape[0] + 1,))
    all_borders[1:-1] = midpt_wl
    all_borders[0] = 2 * wls[0] - midpt_wl[0]
    all_borders[-1] = 2 * wls[-1] - midpt_wl[-1]
    return all_borders


def resample_image2d_flux(image2d_orig, naxis1, cdelt1, crval1, crpix1, coeff,
                          rows=None):
    """Flux-conserving rebin of every row onto the linear wavelength grid.

    Per row: cumulative flux on the old pixel borders (wavelengths from the
    polynomial ``coeff`` evaluated at 1-based border positions), Steffen
    interpolation at the new borders, first difference.
    ``rows`` (test-time saving only) restricts the rows that are computed.
    """
    if image2d_orig.ndim == 1:
        image2d = np.zeros((1, image2d_orig.size))
        image2d[0, :] = image2d_orig
    elif image2d_orig.ndim == 2:
        image2d = np.copy(image2d_orig)
    else:
        raise ValueError("Unexpected number of dimensions: " + str(image2d_orig.ndim))
    nscan, nchan = image2d.shape

    new_x = np.arange(naxis1)
    new_wl = crval1 + cdelt1 * new_x

    old_x_borders = np.arange(-0.5, nchan)
    old_x_borders += crpix1  # following FITS criterium

    new_borders = map_borders(new_wl)

    accum_flux = np.empty((nscan, nchan + 1))
    accum_flux[:, 1:] = np.cumsum(image2d, axis=1)
    accum_flux[:, 0] = 0.0
    image2d_resampled = np.zeros((nscan, naxis1))

    old_wl_borders = polyval(old_x_borders, coeff)

    for iscan in (range(nscan) if rows is None else rows):
        interpolator = SteffenInterpolator(old_wl_borders, accum_flux[iscan],
                                           extrapolate="border")
        fl_borders = interpolator(new_borders)
        image2d_resampled[iscan] = fl_borders[1:] - fl_borders[:-1]

    if image2d_orig.ndim == 1:
        return image2d_resampled[0, :]
    return image2d_resampled


# ---------------------------------------------------------------------------
# numina.array.wavecalib.fix_pix_borders.find_pix_borders (apply_rectwv_coeff.py:190)
# ---------------------------------------------------------------------------
def find_pix_borders(sp, sought_value):
    """First and last index whose value differs from ``sought_value``.

    Returns ``(-1, naxis1)`` when every value equals ``sought_value``.
    """
    sp = np.asarray(sp)
    if sp.ndim != 1:
        raise ValueError("Unexpected number of dimensions:", sp.ndim)
    naxis1 = len(sp)
    jborder_min = -1
    jborder_max = naxis1
    if not np.all(sp == sought_value):
        while True:
            jborder_min += 1
            if sp[jborder_min] != sought_value:
                break
        while True:
            jborder_max -= 1
            if sp[jborder_max] != sought_value:
                break
    return jborder_min, jborder_max


# ---------------------------------------------------------------------------
# numina.array.wavecalib.fix_pix_borders.define_mask_borders
# (median_slitlets_rectified.py:19,118)  [numina, recalled]
# ---------------------------------------------------------------------------
def define_mask_borders(image2d, sought_value, nadditional=0):
    """Mask the borders of every row up to and including its first / last value that
    differs from ``sought_value`` (plus ``nadditional`` pixels); rows made only of
    ``sought_value`` stay unmasked.  Returns ``(mask2d, borders)``.
    """
    image2d = np.asarray(image2d)
    naxis2, naxis1 = image2d.shape
    mask2d = np.zeros((naxis2, naxis1), dtype=bool)
    borders = []
    for i in range(naxis2):
        jborder_min, jborder_max = find_pix_borders(image2d[i, :], sought_value=sought_value)
        borders.append((jborder_min, jborder_max))
        if (jborder_min, jborder_max) != (-1, naxis1):
            if jborder_min != -1:
                mask2d[i, 0:(jborder_min + nadditional + 1)] = True
            if jborder_max != naxis1:
                mask2d[i, (jborder_max - nadditional):naxis1] = True
    return mask2d, borders


# ---------------------------------------------------------------------------
# numina.array.distortion.shift_image2d (recipes/spec/abba.py:27,556)  [numina, recalled]
# ---------------------------------------------------------------------------
def shift_image2d(image2d, xoffset=0.0, yoffset=0.0, resampling=2):
    """Shift by (xoffset, yoffset) pixels with ``rectify2d`` and a pure translation map."""
    aij = np.array([-xoffset, 1.0, 0.0])
    bij = np.array([-yoffset, 0.0, 1.0])
    return rectify2d(image2d, aij, bij, resampling)

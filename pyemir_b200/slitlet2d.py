"""``Slitlet2D`` and ``rectify2d`` on the GPU (per-slitlet API of the reference).

Host-side mirror of ``emirdrp/processing/wavecal/slitlet2d.py:20-431``: the
constructor unpacks one entry of ``RectWaveCoeff.contents`` (every key the
reference reads is read here, so a missing one raises ``KeyError`` the same way),
``extract_slitlet2d`` crops the bounding box, ``rectify`` applies the direct
(``ttd_*``) or inverse (``tti_*``) polynomial map through ``emirwv_rectify2d_host``
(kernel 1 in its general form).  The plotting helpers of the reference are not part
of this path; ``debugplot`` is accepted and ignored.
"""

import ctypes

import numpy as np

from . import _cabi
from .core import EMIR_NAXIS1, EMIR_NAXIS2, EMIR_NPIXPERSLIT_RECTIFIED
from .plan import order_fmap

_SCALAR_KEYS = (
    "csu_bar_left", "csu_bar_right", "csu_bar_slit_center", "csu_bar_slit_width",
    "bb_nc1_orig", "bb_nc2_orig", "bb_ns1_orig", "bb_ns2_orig", "x0_reference",
    "y0_reference_lower", "y0_reference_middle", "y0_reference_upper",
    "y0_frontier_lower", "y0_frontier_upper", "y0_frontier_lower_expected",
    "y0_frontier_upper_expected", "corr_yrect_a", "corr_yrect_b",
    "min_row_rectified", "max_row_rectified",
)


def rectify2d(image2d, aij, bij, resampling, max_order=4):
    """numina ``rectify2d`` (``slitlet2d.py:421-423``) on the GPU; float64 in and out.

    resampling 1: nearest neighbour, 2: flux preserving (pixel-overlap areas);
    3: bilinear (extension, not in the reference).
    """
    if len(aij) != len(bij):
        raise ValueError("Unexpected different lengths of aij and bij")
    order_fmap(len(aij), int(max_order))
    image = np.ascontiguousarray(image2d, dtype=np.float64)
    if image.ndim != 2:
        raise ValueError("image2d must be two-dimensional")
    a = np.ascontiguousarray(aij, dtype=np.float64)
    b = np.ascontiguousarray(bij, dtype=np.float64)
    out = np.empty_like(image)
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_rectify2d_host(
        image.ctypes.data_as(vp), image.shape[0], image.shape[1], a.ctypes.data_as(vp),
        b.ctypes.data_as(vp), len(a), int(resampling), int(max_order), out.ctypes.data_as(vp)))
    return out


class Slitlet2D:
    """Slitlet2D definition (same attributes as the reference class)."""

    def __init__(self, islitlet, rectwv_coeff, debugplot=0, max_order=4):
        self.islitlet = islitlet
        tmpcontent = rectwv_coeff.contents[islitlet - 1]
        for key in _SCALAR_KEYS:
            setattr(self, key, tmpcontent[key])
        self.list_spectrails = [
            np.polynomial.Polynomial(tmpcontent["spectrail"]["poly_coef_" + name])
            for name in ("lower", "middle", "upper")]
        self.list_frontiers = [
            np.polynomial.Polynomial(tmpcontent["frontier"]["poly_coef_" + name])
            for name in ("lower", "upper")]
        self.iminslt = (islitlet - 1) * EMIR_NPIXPERSLIT_RECTIFIED + 1
        self.imaxslt = islitlet * EMIR_NPIXPERSLIT_RECTIFIED
        self.jminslt = 0
        self.jmaxslt = 0
        self.ttd_aij = tmpcontent["ttd_aij"]
        self.ttd_bij = tmpcontent["ttd_bij"]
        self.tti_aij = tmpcontent["tti_aij"]
        self.tti_bij = tmpcontent["tti_bij"]
        self._max_order = int(max_order)
        self.ttd_order = order_fmap(len(self.ttd_aij), self._max_order)
        self.wpoly = tmpcontent["wpoly_coeff"]
        self.debugplot = debugplot

    def extract_slitlet2d(self, image_2k2k, subtitle=None):
        """Bounding-box crop as float64 (``slitlet2d.py:340-379``)."""
        naxis2, naxis1 = image_2k2k.shape
        if naxis1 != EMIR_NAXIS1:
            raise ValueError("Unexpected naxis1")
        if naxis2 != EMIR_NAXIS2:
            raise ValueError("Unexpected naxis2")
        slitlet2d = image_2k2k[(self.bb_ns1_orig - 1):self.bb_ns2_orig,
                               (self.bb_nc1_orig - 1):self.bb_nc2_orig]
        return slitlet2d.astype(float)

    def rectify(self, slitlet2d, resampling, inverse=False, subtitle=None):
        """Rectify (or un-rectify with ``inverse=True``) one slitlet image."""
        if resampling not in [1, 2]:
            raise ValueError("Unexpected resampling value=" + str(resampling))
        naxis2, naxis1 = slitlet2d.shape
        if naxis1 != self.bb_nc2_orig - self.bb_nc1_orig + 1:
            raise ValueError("Unexpected slitlet2d_rect naxis1")
        if naxis2 != self.bb_ns2_orig - self.bb_ns1_orig + 1:
            raise ValueError("Unexpected slitlet2d_rect naxis2")
        if inverse:
            aij, bij = self.tti_aij, self.tti_bij
        else:
            aij, bij = self.ttd_aij, self.ttd_bij
        return rectify2d(slitlet2d, aij, bij, resampling, self._max_order)

"""Synthetic ``RectWaveCoeff`` products and EMIR-like frames (seeded, no I/O).

There is no network for real calibration products, so the tests, ``smoke()`` and
``bench.py`` work on products built here following SURVEY.md Appendix C.  The
conventions come from the reference code that writes real products:

* bounding boxes, ``corr_yrect_*``, ``min/max_row_rectified``:
  ``rectwv_coeff_from_mos_library.py:300-364``;
* ``ttd``/``tti`` are least-squares fits of the (rectified -> original) map and its
  inverse in bounding-box-shifted array coordinates: ``slitlet2darc.py:1089-1113``;
* ``wpoly_coeff`` gives wavelength vs 1-based pixel of the rectified slitlet
  (``slitlet2d.py:134-137``), built from the linear CRVAL1/CDELT1 polynomials of
  ``set_wv_parameters.py`` evaluated at a random ``csu_bar_slit_center``;
* DTU header values: ``testing/create_headers.py:45-56``.
"""

import numpy as np

from .core import (EMIR_NAXIS1, EMIR_NAXIS2, EMIR_NBARS,
                   EMIR_NPIXPERSLIT_RECTIFIED)
from .hdu import HDUList, Header, ImageHDU, PrimaryHDU
from .products import RectWaveCoeff
from .set_wv_parameters import set_wv_parameters

_PITCH = 37.0                              # px between slitlets (2048/55 ≈ 37.24, kept
                                           # a little smaller so slitlets 1 and 55 fit)
_X0 = EMIR_NAXIS1 / 2.0 + 0.5              # 1024.5, rectwv_coeff_from_mos_library.py:249

DTU_HEADER_EXAMPLE = {
    "XDTU_F": 0.922918, "YDTU_F": 0.971815, "ZDTU_F": 1.0,
    "XDTU": -205.679000854492, "YDTU": -24.4878005981445, "ZDTU": -463.765991210938,
    "XDTU_0": -205.651000976562, "YDTU_0": -24.4794006347656, "ZDTU_0": -463.821014404297,
}


def _ncoef(order):
    return (order + 1) * (order + 2) // 2


def _design(order, x, y, sx, sy):
    """Columns x^(i-j) y^j with scaled variables (well conditioned lstsq)."""
    xs = x / sx
    ys = y / sy
    cols = []
    scale = []
    for i in range(order + 1):
        for j in range(i + 1):
            cols.append(xs ** (i - j) * ys ** j)
            scale.append(sx ** (i - j) * sy ** j)
    return np.stack(cols, axis=1), np.array(scale)


def _fit_map(order, x, y, u, v):
    """Least-squares polynomial (x, y) -> (u, v); returns (aij, bij) lists."""
    sx = max(np.abs(x).max(), 1.0)
    sy = max(np.abs(y).max(), 1.0)
    design, scale = _design(order, x, y, sx, sy)
    sol_u = np.linalg.lstsq(design, u, rcond=None)[0] / scale
    sol_v = np.linalg.lstsq(design, v, rcond=None)[0] / scale
    return sol_u.tolist(), sol_v.tolist()


class _SlitGeometry:
    """Analytic slitlet geometry in full-frame 1-based pixel coordinates."""

    def __init__(self, islitlet, rng, sagitta_max):
        s = islitlet
        self.yb = 3.0 + (s - 1) * _PITCH
        self.kappa = sagitta_max * (s - 28) / 27.0
        self.theta = rng.uniform(-1.0e-3, 1.0e-3)
        self.height = _PITCH * rng.uniform(0.98, 1.02)
        self.curv = rng.uniform(1.0, 2.5) * (1 if rng.random() < 0.5 else -1)
        self.tilt = rng.uniform(-1.0, 1.0)

    def y_lower(self, x):
        z = (x - _X0) / 1024.0
        return self.yb + self.kappa * z * z + self.theta * (x - _X0)

    def y_upper(self, x):
        return self.y_lower(x) + self.height

    def delta(self, phi):
        q = 2.0 * phi - 1.0
        return self.curv * q * q + self.tilt * q


def make_rectwv_coeff(grism="J", filter_name="J", order=2, seed=1001,
                      missing_slitlets=(), offx=0, offy=0, wpoly_degree=3,
                      sagitta_max=12.0):
    """Build a synthetic ``RectWaveCoeff`` for one grism + filter combination."""
    rng = np.random.default_rng(seed)
    wvp = set_wv_parameters(filter_name, grism)
    coeff = RectWaveCoeff(instrument="EMIR")
    coeff.uuid = "%08x-0000-4000-8000-%012x" % (seed & 0xFFFFFFFF, seed)
    coeff.quality_control = "GOOD"
    coeff.tags = {"grism": grism, "filter": filter_name}
    coeff.meta_info["origin"]["bound_param"] = "uuid" + "b0b0b0b0-%04d" % (seed % 10000)
    coeff.meta_info["origin"]["master_rectwv"] = "uuid" + "a1a1a1a1-%04d" % (seed % 10000)
    coeff.meta_info["dtu_configuration"] = {k.lower(): v for k, v in DTU_HEADER_EXAMPLE.items()}
    coeff.global_integer_offset_x_pix = int(offx)
    coeff.global_integer_offset_y_pix = int(offy)
    coeff.total_slitlets = EMIR_NBARS
    missing = sorted(int(m) for m in missing_slitlets)
    coeff.missing_slitlets = list(missing)

    xgrid = np.linspace(1.0, float(EMIR_NAXIS1), EMIR_NAXIS1)
    for islitlet in range(1, EMIR_NBARS + 1):
        geo = _SlitGeometry(islitlet, rng, sagitta_max)
        center = rng.uniform(50.0, 290.0)
        quad = rng.uniform(-1.0, 1.0) * 1.0e-6
        cubic = rng.uniform(-1.0, 1.0) * 1.0e-10
        y0_exp_lo = (islitlet - 1) * EMIR_NPIXPERSLIT_RECTIFIED + 0.5
        y0_exp_hi = islitlet * EMIR_NPIXPERSLIT_RECTIFIED + 0.5
        entry = {
            "islitlet": islitlet,
            "csu_bar_left": center - 1.0,
            "csu_bar_right": 341.5 - center - 1.0,
            "csu_bar_slit_center": center,
            "csu_bar_slit_width": 2.0,
            "x0_reference": _X0,
            "y0_frontier_lower_expected": y0_exp_lo,
            "y0_frontier_upper_expected": y0_exp_hi,
        }
        if islitlet in missing:
            coeff.contents.append(entry)
            continue

        ylo = geo.y_lower(xgrid)
        yhi = geo.y_upper(xgrid)
        ymargin_bb = 2
        bb_ns1 = max(1, int(ylo.min() + 0.5) - ymargin_bb)
        bb_ns2 = min(EMIR_NAXIS2, int(yhi.max() + 0.5) + ymargin_bb)
        bb_nc1, bb_nc2 = 1, EMIR_NAXIS1

        y0_lo = float(geo.y_lower(_X0))
        y0_hi = float(geo.y_upper(_X0))
        corr_b = (y0_exp_hi - y0_exp_lo) / (y0_hi - y0_lo)
        corr_a = y0_exp_lo - corr_b * y0_lo
        ymid = (y0_exp_lo + y0_exp_hi) / 2
        ioffset = int(ymid - (bb_ns1 + bb_ns2) / 2.0)
        corr_a -= ioffset
        x1 = corr_a + corr_b * y0_lo
        x2 = corr_a + corr_b * y0_hi
        min_row = int((round(x1 * 10) + 5) / 10) - bb_ns1
        max_row = int((round(x2 * 10) - 5) / 10) - bb_ns1
        bbh = bb_ns2 - bb_ns1 + 1
        assert max_row - min_row + 1 == EMIR_NPIXPERSLIT_RECTIFIED
        assert min_row >= 0 and max_row + 1 < bbh, (islitlet, min_row, max_row, bbh)

        # exact (rectified -> original) map sampled on a grid, full-frame coords
        xr, yr = np.meshgrid(np.linspace(1.0, float(EMIR_NAXIS1), 33),
                             np.linspace(bb_ns1 + min_row - 2.0,
                                         bb_ns1 + max_row + 4.0, 9))
        xr = xr.ravel()
        yr = yr.ravel()
        phi = ((yr - corr_a) / corr_b - y0_lo) / geo.height
        xo = xr + geo.delta(phi)
        yo = geo.y_lower(xo) + phi * geo.height
        # bbox-shifted array coordinates (slitlet2darc.py:1089-1095)
        ttd_aij, ttd_bij = _fit_map(order, xr - bb_nc1, yr - bb_ns1,
                                    xo - bb_nc1, yo - bb_ns1)
        tti_aij, tti_bij = _fit_map(order, xo - bb_nc1, yo - bb_ns1,
                                    xr - bb_nc1, yr - bb_ns1)

        crval1_lin = float(wvp["poly_crval1_linear"](center))
        cdelt1_lin = float(wvp["poly_cdelt1_linear"](center))
        shifted = np.polynomial.Polynomial([-_X0, 1.0])
        wpoly = (np.polynomial.Polynomial([crval1_lin - cdelt1_lin, cdelt1_lin])
                 + (quad * cdelt1_lin) * shifted ** 2 + cubic * shifted ** 3)
        wcoef = list(wpoly.coef)
        while len(wcoef) < wpoly_degree + 1:
            wcoef.append(0.0)
        borders = np.arange(0.5, EMIR_NAXIS1 + 1.0)
        wl = np.polynomial.polynomial.polyval(borders, wcoef)
        assert np.all(np.diff(wl) > 0), "wavelength polynomial must be monotone"

        spectrail = {}
        for name, frac in (("lower", 0.1), ("middle", 0.5), ("upper", 0.9)):
            spectrail["poly_coef_" + name] = [
                float(geo.yb + frac * geo.height + geo.kappa * (_X0 / 1024.0) ** 2
                      - geo.theta * _X0),
                float(geo.theta - 2 * geo.kappa * _X0 / 1024.0 ** 2),
                float(geo.kappa / 1024.0 ** 2)]
        frontier = {
            "poly_coef_lower": [float(spectrail["poly_coef_lower"][0] - 0.1 * geo.height)]
            + spectrail["poly_coef_lower"][1:],
            "poly_coef_upper": [float(spectrail["poly_coef_upper"][0] + 0.1 * geo.height)]
            + spectrail["poly_coef_upper"][1:],
        }
        entry.update({
            "bb_nc1_orig": bb_nc1, "bb_nc2_orig": bb_nc2,
            "bb_ns1_orig": bb_ns1, "bb_ns2_orig": bb_ns2,
            "ymargin_bb": ymargin_bb,
            "spectrail": spectrail, "frontier": frontier,
            "y0_reference_lower": y0_lo + 0.1 * geo.height,
            "y0_reference_middle": y0_lo + 0.5 * geo.height,
            "y0_reference_upper": y0_lo + 0.9 * geo.height,
            "y0_frontier_lower": y0_lo, "y0_frontier_upper": y0_hi,
            "corr_yrect_a": corr_a, "corr_yrect_b": corr_b,
            "min_row_rectified": min_row, "max_row_rectified": max_row,
            "ttd_order": order,
            "ttd_aij": ttd_aij, "ttd_bij": ttd_bij,
            "tti_aij": tti_aij, "tti_bij": tti_bij,
            "wpoly_coeff": [float(c) for c in wcoef],
        })
        coeff.contents.append(entry)
    return coeff


def _slit_spectrum(rng, nlines=200, sigma=1.5, n=EMIR_NAXIS1 + 64):
    """1-D spectrum on an extended grid: continuum + Gaussian emission lines."""
    x = np.arange(n, dtype=np.float64)
    spec = np.full(n, rng.uniform(50.0, 150.0))
    centers = rng.uniform(0.0, n, nlines)
    amps = rng.uniform(20.0, 2000.0, nlines) * rng.random(nlines) ** 3
    for c, a in zip(centers, amps):
        lo = max(0, int(c - 6 * sigma))
        hi = min(n, int(c + 6 * sigma) + 1)
        spec[lo:hi] += a * np.exp(-0.5 * ((x[lo:hi] - c) / sigma) ** 2)
    return spec


def make_base_scene(rectwv_coeff, seed=1):
    """Noise-free float64 2048x2048 scene following the product's slitlet traces.

    Curved spectra between the frontiers, curved emission lines, zero in the
    inter-slit gaps.  Geometry is re-derived from the product itself (bounding
    boxes and the ``tti``-free frontier polynomials stored in it).
    """
    rng = np.random.default_rng(seed)
    scene = np.zeros((EMIR_NAXIS2, EMIR_NAXIS1))
    xpix = np.arange(1, EMIR_NAXIS1 + 1, dtype=np.float64)
    for entry in rectwv_coeff.contents:
        if "ttd_aij" not in entry:
            continue
        lower = np.polynomial.polynomial.polyval(xpix, entry["frontier"]["poly_coef_lower"])
        upper = np.polynomial.polynomial.polyval(xpix, entry["frontier"]["poly_coef_upper"])
        spec = _slit_spectrum(rng)
        curv = rng.uniform(-2.0, 2.0)
        ns1, ns2 = entry["bb_ns1_orig"], entry["bb_ns2_orig"]
        for row in range(ns1, ns2 + 1):
            phi = (row - lower) / (upper - lower)
            inside = (phi >= 0.0) & (phi <= 1.0)
            if not inside.any():
                continue
            q = 2.0 * phi - 1.0
            xs = xpix - curv * q * q + 32.0
            vals = np.interp(xs, np.arange(spec.size, dtype=np.float64), spec)
            illum = 1.0 - 0.2 * q * q
            scene[row - 1, inside] = (vals * illum)[inside]
    return scene


def make_frames(rectwv_coeff, nframes, seed=1, dtype=np.float32, read_noise=5.0,
                scene=None):
    """``nframes`` noisy realisations of the scene, shape (n, 2048, 2048)."""
    rng = np.random.default_rng(seed + 7919)
    if scene is None:
        scene = make_base_scene(rectwv_coeff, seed)
    frames = np.empty((nframes, EMIR_NAXIS2, EMIR_NAXIS1), dtype=dtype)
    for i in range(nframes):
        gain = rng.uniform(0.8, 1.2)
        noise = rng.standard_normal(scene.shape, dtype=np.float32) * np.float32(read_noise)
        frames[i] = (scene * gain + noise).astype(dtype)
    return frames


def make_header(rectwv_coeff, extra=None):
    """Primary header a reduced EMIR frame would carry (only what the path reads)."""
    header = Header()
    header["INSTRUME"] = "EMIR"
    header["OBSMODE"] = "STARE_SPECTRA"
    header["FILTER"] = rectwv_coeff.tags["filter"]
    header["GRISM"] = rectwv_coeff.tags["grism"]
    for key, value in DTU_HEADER_EXAMPLE.items():
        header[key] = value
    # stale WCS cards that the path must remove (apply_rectwv_coeff.py:238-264)
    header["CRPIX1"] = (1024.5, "stale")
    header["CRVAL1"] = (123.0, "stale")
    header["CRPIX2"] = (1024.5, "stale")
    header["CRVAL2"] = (45.0, "stale")
    header["CD1_1"] = -5.4e-5
    header["CD1_2"] = 0.0
    header["CD2_1"] = 0.0
    header["CD2_2"] = 5.4e-5
    header["PCD1_1"] = 1.0
    header["PCRPIX1"] = 1024.5
    if extra:
        for key, value in extra.items():
            header[key] = value
    return header


def make_hdulist(frame, rectwv_coeff, with_mecs=False, extra=None):
    """Wrap one 2048x2048 array into an HDUList like the recipes pass in."""
    hdus = [PrimaryHDU(data=frame, header=make_header(rectwv_coeff, extra))]
    if with_mecs:
        hdus.append(ImageHDU(data=None, header=Header(DTU_HEADER_EXAMPLE), name="MECS"))
    return HDUList(hdus)


# BASELINE.json configs: grism, filter, order, resampling, seed
BASELINE_CONFIGS = {
    1: dict(grism="J", filter_name="J", order=2, seed=1001, frames=1),
    2: dict(grism="H", filter_name="H", order=2, seed=1002, frames=64),
    3: dict(grism="K", filter_name="Ksp", order=5, seed=1003, frames=256),
    4: dict(grism="LR", filter_name="YJ", order=2, seed=1004, frames=512),
    "4b": dict(grism="LR", filter_name="HK", order=2, seed=1004, frames=512),
    5: dict(grism="J", filter_name="J", order=2, seed=1005, frames=4096),
}


def make_config(number, **overrides):
    """``RectWaveCoeff`` of one BASELINE config (frames are made separately)."""
    cfg = dict(BASELINE_CONFIGS[number])
    cfg.pop("frames")
    cfg.update(overrides)
    return make_rectwv_coeff(**cfg)

"""ORACLE — CPU restatement of ``apply_rectwv_coeff`` (orchestration + headers).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); PARITY UNPINNED.

Follows ``emirdrp/processing/wavecal/apply_rectwv_coeff.py:35-287`` statement by
statement and ``slitlet2d.py:145-218,340-431`` for the per-slitlet object, on top
of the numina restatement in ``oracle/numina_restated.py``.  Everything is float64
and single threaded, like the reference.
"""

from datetime import datetime

import numpy as np

from pyemir_b200.core import (EMIR_NAXIS1, EMIR_NAXIS2, EMIR_NBARS,
                              EMIR_NPIXPERSLIT_RECTIFIED)
from pyemir_b200.dtu import DtuConf
from pyemir_b200.hdu import copy_img
from pyemir_b200.set_wv_parameters import set_wv_parameters

from . import numina_restated as nr


class OracleSlitlet2D:
    """The part of ``Slitlet2D`` that the hot path uses (slitlet2d.py:145-431)."""

    def __init__(self, islitlet, rectwv_coeff, nmax_order=nr.NMAX_ORDER):
        self.islitlet = islitlet
        c = rectwv_coeff.contents[islitlet - 1]
        # every key is read in the reference's constructor: a missing one raises
        for key in ("csu_bar_left", "csu_bar_right", "csu_bar_slit_center",
                    "csu_bar_slit_width", "x0_reference", "y0_reference_lower",
                    "y0_reference_middle", "y0_reference_upper", "y0_frontier_lower",
                    "y0_frontier_upper", "y0_frontier_lower_expected",
                    "y0_frontier_upper_expected", "corr_yrect_a", "corr_yrect_b"):
            setattr(self, key, c[key])
        for name in ("lower", "middle", "upper"):
            c["spectrail"]["poly_coef_" + name]
        for name in ("lower", "upper"):
            c["frontier"]["poly_coef_" + name]
        self.bb_nc1_orig = c["bb_nc1_orig"]
        self.bb_nc2_orig = c["bb_nc2_orig"]
        self.bb_ns1_orig = c["bb_ns1_orig"]
        self.bb_ns2_orig = c["bb_ns2_orig"]
        self.min_row_rectified = c["min_row_rectified"]
        self.max_row_rectified = c["max_row_rectified"]
        self.iminslt = (islitlet - 1) * EMIR_NPIXPERSLIT_RECTIFIED + 1
        self.imaxslt = islitlet * EMIR_NPIXPERSLIT_RECTIFIED
        self.jminslt = 0
        self.jmaxslt = 0
        self.ttd_aij = c["ttd_aij"]
        self.ttd_bij = c["ttd_bij"]
        self.tti_aij = c["tti_aij"]
        self.tti_bij = c["tti_bij"]
        self.nmax_order = nmax_order
        self.ttd_order = nr.order_fmap(len(self.ttd_aij), nmax_order)
        self.wpoly = c["wpoly_coeff"]

    def extract_slitlet2d(self, image_2k2k):
        naxis2, naxis1 = image_2k2k.shape
        if naxis1 != EMIR_NAXIS1:
            raise ValueError("Unexpected naxis1")
        if naxis2 != EMIR_NAXIS2:
            raise ValueError("Unexpected naxis2")
        slitlet2d = image_2k2k[(self.bb_ns1_orig - 1):self.bb_ns2_orig,
                               (self.bb_nc1_orig - 1):self.bb_nc2_orig]
        return slitlet2d.astype(float)

    def rectify(self, slitlet2d, resampling, inverse=False, rows=None,
                allow_extensions=False):
        valid = [1, 2, 3] if allow_extensions else [1, 2]
        if resampling not in valid:
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
        return nr.rectify2d(image2d=slitlet2d, aij=aij, bij=bij,
                            resampling=resampling, nmax_order=self.nmax_order,
                            rows=rows)


# header cards written after the slitlet loop (apply_rectwv_coeff.py:238-250);
# a bare value keeps whatever comment the card already had
_STALE_WCS = ("crval1", "crpix1", "crval2", "crpix2")
_STALE_CD = ("cd1_1", "cd1_2", "cd2_1", "cd2_2", "PCD1_1", "PCD1_2", "PCD2_1",
             "PCD2_2", "PCRPIX1", "PCRPIX2")


def _rewrite_wcs(header, crpix1, crval1, cdelt1):
    for keyword in _STALE_WCS:
        if keyword in header:
            header.remove(keyword)
    cards = (
        ("crpix1", (crpix1, "reference pixel")),
        ("crval1", (crval1, "central wavelength at crpix1")),
        ("cdelt1", (cdelt1, "linear dispersion (Angstrom/pixel)")),
        ("cunit1", ("Angstrom", "units along axis1")),
        ("ctype1", "WAVE"),
        ("crpix2", (0.0, "reference pixel")),
        ("crval2", (0.0, "central value at crpix2")),
        ("cdelt2", (1.0, "increment")),
        ("ctype2", "PIXEL"),
        ("cunit2", ("Pixel", "units along axis2")),
    )
    for keyword, card in cards:
        header[keyword] = card
    for keyword in _STALE_CD:
        if keyword in header:
            header.remove(keyword)


def _append_history(header, rectwv_coeff):
    """HISTORY cards of apply_rectwv_coeff.py:267-279 (the last one is a timestamp)."""
    origin = rectwv_coeff.meta_info["origin"]
    header["history"] = "Boundary parameters uuid:" + origin["bound_param"][4:]
    if "master_rectwv" in origin:
        header["history"] = "MasterRectWave uuid:" + origin["master_rectwv"][4:]
    header["history"] = "RectWaveCoeff uuid:" + rectwv_coeff.uuid
    header["history"] = ("Rectification and wavelength calibration time "
                         + datetime.now().isoformat())



def apply_rectwv_coeff(reduced_image, rectwv_coeff, args_resampling=2,
                       args_ignore_dtu_configuration=True, debugplot=0,
                       nmax_order=nr.NMAX_ORDER, only_needed_rows=False,
                       allow_extensions=False, return_float64=False):
    """Oracle for the hot path; same signature as the reference plus test knobs.

    ``nmax_order``       numina's order cap (4); raise it for BASELINE config 3.
    ``only_needed_rows`` compute only rectified rows ``min_row .. max_row+1`` (the
                         ones that reach the output or the JMNSLT/JMXSLT scan);
                         the reference computes every row of the bounding box and
                         discards the rest, with identical results.
    ``allow_extensions`` accept ``args_resampling=3`` (bilinear, not in the reference).
    ``return_float64``   also return the float64 image before the float32 cast.
    """
    rectwv_image = copy_img(reduced_image)
    header = rectwv_image[0].header
    image2d = rectwv_image[0].data

    image2d = nr.apply_integer_offsets(
        image2d=image2d,
        offx=rectwv_coeff.global_integer_offset_x_pix,
        offy=rectwv_coeff.global_integer_offset_y_pix,
    )

    filter_name = header["filter"]
    if filter_name != rectwv_coeff.tags["filter"]:
        raise ValueError("Filter name does not match!")
    grism_name = header["grism"]
    if grism_name != rectwv_coeff.tags["grism"]:
        raise ValueError("Grism name does not match!")

    dtu_conf = DtuConf.from_img(reduced_image)
    dtu_conf_calib = DtuConf.from_header(rectwv_coeff.meta_info["dtu_configuration"])
    if dtu_conf != dtu_conf_calib:
        if not args_ignore_dtu_configuration:
            raise ValueError("DTU configurations do not match!")

    list_valid_islitlets = list(range(1, EMIR_NBARS + 1))
    for idel in rectwv_coeff.missing_slitlets:
        list_valid_islitlets.remove(idel)

    wv_parameters = set_wv_parameters(filter_name, grism_name)
    crpix1_enlarged = wv_parameters["crpix1_enlarged"]
    crval1_enlarged = wv_parameters["crval1_enlarged"]
    cdelt1_enlarged = wv_parameters["cdelt1_enlarged"]
    naxis1_enlarged = wv_parameters["naxis1_enlarged"]

    naxis2_enlarged = EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED
    image2d_rectwv = np.zeros((naxis2_enlarged, naxis1_enlarged), dtype="float32")
    image2d_f64 = np.zeros((naxis2_enlarged, naxis1_enlarged)) if return_float64 else None

    for islitlet in range(1, EMIR_NBARS + 1):
        cslit = str(islitlet).zfill(2)
        if islitlet in list_valid_islitlets:
            slt = OracleSlitlet2D(islitlet, rectwv_coeff, nmax_order)
            slitlet2d = slt.extract_slitlet2d(image2d)

            ii1 = slt.min_row_rectified
            ii2 = slt.max_row_rectified + 1
            rows = list(range(ii1, ii2 + 1)) if only_needed_rows else None
            if rows is not None and rows[-1] >= slitlet2d.shape[0]:
                raise IndexError("index %d is out of bounds for axis 0 with size %d"
                                 % (rows[-1], slitlet2d.shape[0]))

            slitlet2d_rect = slt.rectify(slitlet2d, resampling=args_resampling,
                                         rows=rows, allow_extensions=allow_extensions)
            slitlet2d_rect_wv = nr.resample_image2d_flux(
                image2d_orig=slitlet2d_rect,
                naxis1=naxis1_enlarged,
                cdelt1=cdelt1_enlarged,
                crval1=crval1_enlarged,
                crpix1=crpix1_enlarged,
                coeff=slt.wpoly,
                rows=rows,
            )

            i1 = slt.iminslt - 1
            i2 = slt.imaxslt
            image2d_rectwv[i1:i2, :] = slitlet2d_rect_wv[ii1:ii2, :]
            if image2d_f64 is not None:
                image2d_f64[i1:i2, :] = slitlet2d_rect_wv[ii1:ii2, :]

            header["imnslt" + cslit] = (slt.iminslt,
                                        "minimum Y pixel of useful slitlet region")
            header["imxslt" + cslit] = (slt.imaxslt,
                                        "maximum Y pixel of useful slitlet region")

            jminslt = []
            jmaxslt = []
            for idum in range(ii1, ii2 + 1):
                jminmax = nr.find_pix_borders(slitlet2d_rect_wv[idum, :], sought_value=0)
                if jminmax != (-1, naxis1_enlarged):
                    jminslt.append(jminmax[0])
                    jmaxslt.append(jminmax[1])
            if len(jminslt) > 0:
                slt.jminslt = min(jminslt) + 1
                slt.jmaxslt = max(jmaxslt) + 1
            header["jmnslt" + cslit] = (slt.jminslt,
                                        "minimum X pixel of useful slitlet region")
            header["jmxslt" + cslit] = (slt.jmaxslt,
                                        "maximum X pixel of useful slitlet region")
        else:
            header["imnslt" + cslit] = (0, "minimum Y pixel of useful slitlet region")
            header["imxslt" + cslit] = (0, "maximum Y pixel of useful slitlet region")
            header["jmnslt" + cslit] = (0, "minimum X pixel of useful slitlet region")
            header["jmxslt" + cslit] = (0, "maximum X pixel of useful slitlet region")

    _rewrite_wcs(header, crpix1_enlarged, crval1_enlarged, cdelt1_enlarged)
    _append_history(header, rectwv_coeff)

    rectwv_image[0].data = image2d_rectwv
    if return_float64:
        return rectwv_image, image2d_f64
    return rectwv_image


def median_slitlets_rectified(input_image, mode=0, list_useful_slitlets=None, debugplot=0):
    """Oracle for the consumer right behind the path: median spectrum of every slitlet.

    Follows ``emirdrp/processing/wavecal/median_slitlets_rectified.py:31-164``: modes 0 and
    1 are plain numpy (pinned by ``tests/golden``: the reference's own function was
    executed); mode 2 adds numina's ``define_mask_borders`` (restated, unpinned) and a
    masked median over the useful slitlets.
    """
    result = copy_img(input_image)
    image_header = input_image[0].header
    image2d = input_image[0].data
    naxis2_expected = EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED
    naxis2, naxis1 = image2d.shape
    if naxis2 != naxis2_expected:
        raise ValueError("NAXIS2={0} should be {1}".format(naxis2, naxis2_expected))
    if image_header["instrume"] != "EMIR":
        raise ValueError("INSTRUME keyword is not 'EMIR'!")
    nrows = naxis2 if mode == 0 else EMIR_NBARS
    image2d_median = np.zeros((nrows, naxis1), dtype="float32")
    for i in range(EMIR_NBARS):
        ns1 = i * EMIR_NPIXPERSLIT_RECTIFIED
        sp_median = np.median(image2d[ns1:ns1 + EMIR_NPIXPERSLIT_RECTIFIED, :], axis=0)
        if mode == 0:
            image2d_median[ns1:ns1 + EMIR_NPIXPERSLIT_RECTIFIED, :] = sp_median
        else:
            image2d_median[i] = sp_median
    if mode == 2:
        if list_useful_slitlets is None:
            list_not_useful = []
        else:
            list_not_useful = [i for i in range(1, EMIR_NBARS + 1)
                               if i not in list_useful_slitlets]
        mask2d, _ = nr.define_mask_borders(image2d_median, sought_value=0)
        for islitlet in list_not_useful:
            mask2d[islitlet - 1, :] = True
        masked = np.ma.masked_array(image2d_median, mask=mask2d)
        result[0].data = np.ma.median(masked, axis=0).data.astype("float32")
    else:
        result[0].data = image2d_median.astype("float32")
    return result


def combine_abba_mean(frames, full_set):
    """Oracle of the ABBA "mean" combination (``recipes/spec/abba.py:579-606``).

    ``combine_imgs(list, method=mean)`` is numina code [recalled]: the mean of the stack in
    float64; the recipe then casts each mean to float32 and subtracts (``abba.py:604-606``).
    Sums run in frame order, like numpy's reduction over the first axis.
    """
    frames = np.asarray(frames)
    is_a = np.array([c == "A" for c in full_set])
    is_b = np.array([c == "B" for c in full_set])
    if not (is_a | is_b).all():
        raise ValueError("Unexpected char value")
    data_a = frames[is_a].astype(np.float64).sum(axis=0) / float(is_a.sum())
    data_b = frames[is_b].astype(np.float64).sum(axis=0) / float(is_b.sum())
    return data_a.astype("float32") - data_b.astype("float32")


def combine_abba_median(frames, full_set):
    """Oracle of the ABBA "median" combination (``recipes/spec/abba.py:132-137,579-606``).

    ``combine_imgs(list, method=median)`` is numina code [recalled]: the per-pixel median of
    the stack in float64 (mean of the two middle samples for an even count); the recipe then
    casts both medians to float32 and subtracts (``abba.py:604-606``).
    """
    frames = np.asarray(frames)
    is_a = np.array([c == "A" for c in full_set])
    is_b = np.array([c == "B" for c in full_set])
    if not (is_a | is_b).all():
        raise ValueError("Unexpected char value")
    data_a = np.median(frames[is_a].astype(np.float64), axis=0)
    data_b = np.median(frames[is_b].astype(np.float64), axis=0)
    return data_a.astype("float32") - data_b.astype("float32")


def flatfield_correct(data, flatdata, dtype="float32"):
    """Oracle of pyemir's FlatFieldCorrector arithmetic (``processing/flatfield.py:43,58-59``).

    ``flatdata[flatdata <= 0] = 1.0`` (on a copy here), ``(data / flatdata).astype(dtype)``.
    Pinned by tests/golden/flatfield.json (produced with the reference's own class).
    """
    flat = np.array(flatdata, copy=True)
    flat[flat <= 0] = 1.0
    with np.errstate(invalid="ignore", divide="ignore"):
        return (np.asarray(data) / flat).astype(dtype)


def prereduce(frames, bias=None, dark=None, flat=None):
    """Oracle of the element-wise part of ``flow(newimg)`` (``recipes/spec/abba.py:371``).

    numina's BiasCorrector and DarkCorrector [recalled]: image minus calibration, result cast
    to the corrector's dtype (float32); then ``flatfield_correct``.  numpy type promotion.
    """
    x = np.asarray(frames)
    with np.errstate(invalid="ignore"):
        if bias is not None:
            x = (x - np.asarray(bias)).astype(np.float32)
        if dark is not None:
            x = (x - np.asarray(dark)).astype(np.float32)
    if flat is not None:
        x = flatfield_correct(x, flat, "float32")
    return x.astype(np.float32)


# ---------------------------------------------------------------------------
# ABBA combination: method "sigmaclip", per-frame vertical offsets, the whole product
# (recipes/spec/abba.py:132-137,545-610)                                [numina, recalled]
# ---------------------------------------------------------------------------
def sigmaclip_stack(stack, low=3.0, high=3.0):
    """Oracle of numina ``combine.sigmaclip`` (the third ``method`` of ``abba.py:132-137``).

    [numina, recalled — PARITY UNPINNED] numina's C++ ``sigmaclip`` method: per pixel, repeat
    { mean and unbiased (n-1) standard deviation of the samples still in; drop the samples
    outside [mean - low*std, mean + high*std] } until a pass drops nothing; the value is the
    mean of the last pass (``combine_imgs`` keeps ``out[0]``).  Stated here with the sums in
    frame order over the samples still in; a pass that would drop every sample ends the loop.
    Returns the float64 image.
    """
    v = np.asarray(stack, dtype=np.float64)
    v = v.reshape(v.shape[0], -1)
    n, npix = v.shape
    keep = np.ones((n, npix), dtype=bool)
    result = np.zeros(npix)
    active = np.ones(npix, dtype=bool)
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        while active.any():
            cnt = keep.sum(axis=0)
            s = np.zeros(npix)
            for f in range(n):
                s = np.where(keep[f], s + v[f], s)
            mean = s / cnt
            ss = np.zeros(npix)
            for f in range(n):
                d = v[f] - mean
                ss = np.where(keep[f], ss + d * d, ss)
            var = np.where(cnt > 1, ss / np.maximum(cnt - 1, 1), 0.0)
            std = np.sqrt(var)
            lo = mean - std * low
            hi = mean + std * high
            new = keep & (v >= lo) & (v <= hi)
            ncnt = new.sum(axis=0)
            result[active] = mean[active]
            changed = active & (ncnt != cnt) & (ncnt > 0)
            keep[:, changed] = new[:, changed]
            active = changed
    return result.reshape(np.asarray(stack).shape[1:])


def combine_abba(frames, full_set, offsets=None, method="mean", low=3.0, high=3.0):
    """Oracle of ``abba.py:545-610``: per-frame vertical offset, combination of the A and of
    the B images with ``method``, float32(A) - float32(B).

    ``shifted = shift_image2d(data, yoffset=-offset).astype("float32")`` for every frame whose
    offset is not zero (``abba.py:554-558``), then numina's ``combine`` method [recalled] in
    float64, the casts and the difference of ``abba.py:604-606``.
    """
    frames = np.asarray(frames)
    if offsets is None:
        offsets = [0.0] * len(frames)
    shifted = []
    for data, offset in zip(frames, offsets):
        if offset != 0:
            data = nr.shift_image2d(data, yoffset=-offset).astype("float32")
        shifted.append(np.asarray(data))
    is_a = [c == "A" for c in full_set]
    is_b = [c == "B" for c in full_set]
    if not all(a or b for a, b in zip(is_a, is_b)):
        raise ValueError("Unexpected char value")

    def combine(sel):
        stack = np.stack([s for s, use in zip(shifted, sel) if use]).astype(np.float64)
        if method == "mean":
            return stack.sum(axis=0) / float(len(stack))
        if method == "median":
            return np.median(stack, axis=0)
        if method == "sigmaclip":
            return sigmaclip_stack(stack, low, high)
        raise ValueError(f"Unexpected method: {method}")

    return combine(is_a).astype("float32") - combine(is_b).astype("float32")


# ---------------------------------------------------------------------------
# bad-pixel stage of the flow (core/correctors.py:23-42)                 [numina, recalled]
# ---------------------------------------------------------------------------
def process_bpm_median(arr, mask, hwin=2, wwin=2, fill=0.0):
    """Oracle of numina ``BadPixelCorrector`` -> ``array.bpm.process_bpm_median``.

    [numina, recalled — PARITY UNPINNED] every pixel with ``mask != 0`` becomes the median of
    the unmasked pixels of the input within +-hwin rows and +-wwin columns (window clipped at
    the edges), ``fill`` when there is none; the other pixels are kept.  float64 result.
    """
    arr = np.asarray(arr)
    mask = np.asarray(mask)
    out = arr.astype(np.float64).copy()
    naxis2, naxis1 = arr.shape
    for r, c in zip(*np.nonzero(mask)):
        r0, r1 = max(r - hwin, 0), min(r + hwin, naxis2 - 1)
        c0, c1 = max(c - wwin, 0), min(c + wwin, naxis1 - 1)
        win = arr[r0:r1 + 1, c0:c1 + 1]
        good = win[mask[r0:r1 + 1, c0:c1 + 1] == 0].astype(np.float64)
        with np.errstate(invalid="ignore"):
            out[r, c] = np.median(good) if good.size else fill
    return out


def prereduce_bpm(frames, bpm=None, bias=None, dark=None, flat=None):
    """``flow(newimg)`` with the bad-pixel stage in front (``core/recipe.py:26-41``): the
    corrected frame keeps float64 for float64 input and is float32 otherwise (the window
    medians of uint16 / float32 data are exact in float32), then ``prereduce``."""
    frames = np.asarray(frames)
    if bpm is None:
        return prereduce(frames, bias, dark, flat)
    fixed = np.stack([process_bpm_median(f, bpm) for f in frames])
    fixed = fixed if frames.dtype == np.float64 else fixed.astype(np.float32)
    return prereduce(fixed, bias, dark, flat)

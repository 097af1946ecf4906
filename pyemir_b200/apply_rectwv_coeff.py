"""Drop-in ``apply_rectwv_coeff`` on the B200 kernels.

Same call signature, HDUList/FITS-header output and error behaviour as the
reference function ``emirdrp.processing.wavecal.apply_rectwv_coeff``
(apply_rectwv_coeff.py:35-287).  The host does the checks and the header; the
pixels go through ``libemirwv`` (``include/emirwv.h``).  There is no CPU path: if
the library or a CUDA device is missing the call raises.
"""

import logging
from datetime import datetime

import numpy as np

from .core import EMIR_NAXIS1, EMIR_NAXIS2, EMIR_NBARS, EMIR_NPIXPERSLIT_RECTIFIED
from .dtu import DtuConf
from . import _pinned
from .hdu import ImageHDU, copy_img
from .plan import get_plan
from .set_wv_parameters import set_wv_parameters

_JNONE = np.iinfo(np.int32).max

_CARD_COMMENTS = {
    "imnslt": "minimum Y pixel of useful slitlet region",
    "imxslt": "maximum Y pixel of useful slitlet region",
    "jmnslt": "minimum X pixel of useful slitlet region",
    "jmxslt": "maximum X pixel of useful slitlet region",
}


def _auto_dtype(array):
    """float32 kernels when the input is exactly representable in float32."""
    kind, size = array.dtype.kind, array.dtype.itemsize
    if (kind == "f" and size <= 4) or (kind in "iu" and size <= 2) or kind == "b":
        return "float32"
    return "float64"


def _copy_for_result(img):
    """``copy_img`` (apply_rectwv_coeff.py:69) without duplicating pixels that are replaced.

    The reference deep-copies the input and later replaces the primary data by the rectified
    image; the primary pixels of the copy are never used.  For this package's own HDU classes
    the copy therefore shares the (never modified) primary array with the input -- headers and
    extension HDUs are copied as before; unknown HDU classes (astropy) take the deep copy.
    """
    primary = img[0]
    if not isinstance(primary, ImageHDU):
        return copy_img(img)
    light = type(primary)(data=primary.data, header=primary.header.copy())
    return type(img)([light] + [hdu.copy() for hdu in img[1:]])


def _check_tags(header, rectwv_coeff, logger):
    """Filter / grism / DTU checks of apply_rectwv_coeff.py:80-108."""
    filter_name = header["filter"]
    logger.info("Filter: " + filter_name)
    if filter_name != rectwv_coeff.tags["filter"]:
        raise ValueError("Filter name does not match!")
    grism_name = header["grism"]
    logger.info("Grism: " + grism_name)
    if grism_name != rectwv_coeff.tags["grism"]:
        raise ValueError("Grism name does not match!")
    return filter_name, grism_name


def _check_dtu(reduced_image, rectwv_coeff, ignore, logger):
    dtu_conf = DtuConf.from_img(reduced_image)
    dtu_conf_calib = DtuConf.from_header(rectwv_coeff.meta_info["dtu_configuration"])
    if dtu_conf != dtu_conf_calib:
        if ignore:
            logger.warning("DTU configuration differences found!")
        else:
            logger.warning("DTU configuration from image header:")
            logger.warning(dtu_conf)
            logger.warning("DTU configuration from master calibration:")
            logger.warning(dtu_conf_calib)
            raise ValueError("DTU configurations do not match!")
    else:
        logger.info("DTU configuration match!")


def slitlet_keywords(valid_islitlets, jminmax):
    """Per-slitlet (imnslt, imxslt, jmnslt, jmxslt), 1-based, 0 when not available.

    ``jminmax`` is the kernel's raw (55, 2) result: 0-based first/last non-zero
    channel over the 39 scanned rows, (INT32_MAX, -1) when all are zero
    (apply_rectwv_coeff.py:163-226, slitlet2d.py:198-203).
    """
    valid = set(valid_islitlets)
    rows = []
    for islitlet in range(1, EMIR_NBARS + 1):
        if islitlet not in valid:
            rows.append((0, 0, 0, 0))
            continue
        imin = (islitlet - 1) * EMIR_NPIXPERSLIT_RECTIFIED + 1
        imax = islitlet * EMIR_NPIXPERSLIT_RECTIFIED
        jlo, jhi = int(jminmax[islitlet - 1, 0]), int(jminmax[islitlet - 1, 1])
        if jhi < 0 or jlo == _JNONE:
            rows.append((imin, imax, 0, 0))
        else:
            rows.append((imin, imax, jlo + 1, jhi + 1))
    return rows


def write_header(header, rectwv_coeff, wv_parameters, keywords):
    """Header edits of apply_rectwv_coeff.py:176-226,238-279 (host side)."""
    for islitlet, values in enumerate(keywords, start=1):
        cslit = str(islitlet).zfill(2)
        for name, value in zip(("imnslt", "imxslt", "jmnslt", "jmxslt"), values):
            header[name + cslit] = (value, _CARD_COMMENTS[name])

    for keyword in ("crval1", "crpix1", "crval2", "crpix2"):
        if keyword in header:
            header.remove(keyword)
    header["crpix1"] = (wv_parameters["crpix1_enlarged"], "reference pixel")
    header["crval1"] = (wv_parameters["crval1_enlarged"], "central wavelength at crpix1")
    header["cdelt1"] = (wv_parameters["cdelt1_enlarged"], "linear dispersion (Angstrom/pixel)")
    header["cunit1"] = ("Angstrom", "units along axis1")
    header["ctype1"] = "WAVE"
    header["crpix2"] = (0.0, "reference pixel")
    header["crval2"] = (0.0, "central value at crpix2")
    header["cdelt2"] = (1.0, "increment")
    header["ctype2"] = "PIXEL"
    header["cunit2"] = ("Pixel", "units along axis2")
    for keyword in ("cd1_1", "cd1_2", "cd2_1", "cd2_2", "PCD1_1", "PCD1_2", "PCD2_1",
                    "PCD2_2", "PCRPIX1", "PCRPIX2"):
        if keyword in header:
            header.remove(keyword)

    origin = rectwv_coeff.meta_info["origin"]
    header["history"] = "Boundary parameters uuid:" + origin["bound_param"][4:]
    if "master_rectwv" in origin:
        header["history"] = "MasterRectWave uuid:" + origin["master_rectwv"][4:]
    header["history"] = "RectWaveCoeff uuid:" + rectwv_coeff.uuid
    header["history"] = ("Rectification and wavelength calibration time "
                         + datetime.now().isoformat())


def apply_rectwv_coeff(reduced_image, rectwv_coeff, args_resampling=2,
                       args_ignore_dtu_configuration=True, debugplot=0, *,
                       dtype=None, max_order=4, device=None, allow_extensions=False):
    """Rectify and wavelength-calibrate one reduced EMIR frame on the GPU.

    Parameters (first five as in the reference)
    ----------
    reduced_image : HDUList
        Image with preliminary basic reduction; not modified.
    rectwv_coeff : RectWaveCoeff
        Rectification and wavelength calibration coefficients.
    args_resampling : int
        1: nearest neighbour, 2: flux preserving interpolation.  Any other value raises
        ``ValueError("Unexpected resampling value=...")`` like the reference
        (slitlet2d.py:403-404) -- unless ``allow_extensions`` admits 3 (true bilinear
        sampling, not in the reference).
    args_ignore_dtu_configuration : bool
        If True, only warn about DTU configuration differences.
    debugplot : int
        Accepted for compatibility; plotting is not part of this path.
    dtype : {None, "float32", "float64"}, keyword only
        Arithmetic type of the kernels.  None: float32 when the input pixels are
        exactly representable in float32, else float64.  The returned image is
        float32 either way, like the reference (apply_rectwv_coeff.py:128).
    max_order : int, keyword only
        Cap on the distortion polynomial order (numina: 4).
    device : int, keyword only
        CUDA device ordinal (default: the current one).
    allow_extensions : bool, keyword only
        Admit ``args_resampling=3`` (bilinear; BASELINE.json names it, the reference has none).

    Known deviations from the reference ("Known deviations"): float32 input runs
    float32 kernels (the JMNSLT/JMXSLT scan then tests float32 values for != 0; pass
    ``dtype="float64"`` for the reference's arithmetic), and a NaN / Inf pixel stays local
    instead of poisoning every redder column of its row through the cumulative sum.

    Returns
    -------
    rectwv_image : HDUList
        Rectified and wavelength calibrated image (2090 x naxis1_enlarged).
    """
    logger = logging.getLogger(__name__)

    rectwv_image = _copy_for_result(reduced_image)
    header = rectwv_image[0].header
    image2d = rectwv_image[0].data

    # apply_integer_offsets raises on non-int offsets before anything else
    if (type(rectwv_coeff.global_integer_offset_x_pix) is not int
            or type(rectwv_coeff.global_integer_offset_y_pix) is not int):
        raise ValueError("Invalid non-integer offsets")

    filter_name, grism_name = _check_tags(header, rectwv_coeff, logger)
    _check_dtu(reduced_image, rectwv_coeff, args_ignore_dtu_configuration, logger)

    wv_parameters = set_wv_parameters(filter_name, grism_name)
    image2d = np.asarray(image2d)
    if dtype is None:
        dtype = _auto_dtype(image2d)

    logger.info("Applying rectification and wavelength calibration")
    logger.info("RectWaveCoeff uuid={}".format(rectwv_coeff.uuid))

    if args_resampling not in ((1, 2, 3) if allow_extensions else (1, 2)):
        raise ValueError("Unexpected resampling value=" + str(args_resampling))
    plan = get_plan(rectwv_coeff, args_resampling, dtype, max_order, device=device)
    naxis1_enlarged = wv_parameters["naxis1_enlarged"]
    if plan.valid_islitlets:
        if image2d.ndim != 2:
            raise ValueError("not enough values to unpack (expected 2, got %d)" % image2d.ndim)
        if image2d.shape[1] != EMIR_NAXIS1:
            raise ValueError("Unexpected naxis1")
        if image2d.shape[0] != EMIR_NAXIS2:
            raise ValueError("Unexpected naxis2")
        # the result lands in page-locked memory (a pooled buffer: _pinned.py): the copy back
        # runs at PCIe speed instead of at the speed of page faults
        out = _pinned.empty((1,) + plan.out_shape, plan.np_dtype)
        out, jminmax = plan.run_host(image2d, out=out)
        image2d_rectwv = out[0].astype(np.float32, copy=False)
        jminmax = jminmax[0]
    else:
        image2d_rectwv = np.zeros((EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED, naxis1_enlarged),
                                  dtype=np.float32)
        jminmax = np.tile(np.array([_JNONE, -1], dtype=np.int32), (EMIR_NBARS, 1))

    logger.info("Updating image header")
    write_header(header, rectwv_coeff, wv_parameters,
                 slitlet_keywords(plan.valid_islitlets, jminmax))

    logger.info("Generating rectified and wavelength calibrated image")
    rectwv_image[0].data = image2d_rectwv
    return rectwv_image


def rectify_frames(frames, rectwv_coeff, args_resampling=2, *, dtype=None, max_order=4,
                   device=None):
    """Array-level batch entry: (n, 2048, 2048) -> (n, 2090, naxis1_enlarged).

    All frames share one RectWaveCoeff (the ABBA pattern, abba.py:191-192).
    Returns ``(out, keywords)``: ``out`` in the kernels' dtype and, per frame, the 55
    (imnslt, imxslt, jmnslt, jmxslt) tuples that go to the header.
    """
    frames = np.asarray(frames)
    if frames.ndim == 2:
        frames = frames[None]
    if dtype is None:
        dtype = _auto_dtype(frames)
    plan = get_plan(rectwv_coeff, args_resampling, dtype, max_order, device=device)
    out, jminmax = plan.run_host(frames)
    keywords = [slitlet_keywords(plan.valid_islitlets, jm) for jm in jminmax]
    return out, keywords


def apply_rectwv_coeff_batch(reduced_images, rectwv_coeff, args_resampling=2,
                             args_ignore_dtu_configuration=True, debugplot=0, *,
                             dtype=None, max_order=4, device=None, allow_extensions=False):
    """``apply_rectwv_coeff`` for a list of HDULists sharing one RectWaveCoeff.

    One plan, one pipelined pass over all frames; returns a list of HDULists
    identical to calling :func:`apply_rectwv_coeff` on each image.
    """
    logger = logging.getLogger(__name__)
    if (type(rectwv_coeff.global_integer_offset_x_pix) is not int
            or type(rectwv_coeff.global_integer_offset_y_pix) is not int):
        raise ValueError("Invalid non-integer offsets")
    results, arrays, wvp = [], [], None
    for reduced_image in reduced_images:
        rectwv_image = copy_img(reduced_image)
        filter_name, grism_name = _check_tags(rectwv_image[0].header, rectwv_coeff, logger)
        _check_dtu(reduced_image, rectwv_coeff, args_ignore_dtu_configuration, logger)
        wvp = set_wv_parameters(filter_name, grism_name)
        arrays.append(np.asarray(rectwv_image[0].data))
        results.append(rectwv_image)
    if not results:
        return results
    if args_resampling not in ((1, 2, 3) if allow_extensions else (1, 2)):
        raise ValueError("Unexpected resampling value=" + str(args_resampling))
    if dtype is None:
        kinds = {_auto_dtype(a) for a in arrays}
        dtype = "float64" if "float64" in kinds else "float32"
    for a in arrays:
        if a.ndim != 2 or a.shape[1] != EMIR_NAXIS1:
            raise ValueError("Unexpected naxis1")
        if a.shape[0] != EMIR_NAXIS2:
            raise ValueError("Unexpected naxis2")
    out, keywords = rectify_frames(np.stack(arrays), rectwv_coeff, args_resampling,
                                   dtype=dtype, max_order=max_order, device=device)
    for rectwv_image, data, kw in zip(results, out, keywords):
        write_header(rectwv_image[0].header, rectwv_coeff, wvp, kw)
        rectwv_image[0].data = data.astype(np.float32, copy=False)
    return results

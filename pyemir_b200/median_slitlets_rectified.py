"""``median_slitlets_rectified`` on the GPU (the consumer right behind the path).

Host-side mirror of ``emirdrp/processing/wavecal/median_slitlets_rectified.py:31-164``
(called by ``stare.py:144-153`` and ``refine_rectwv_coeff.py:109-112`` on the product of
``apply_rectwv_coeff``): same signature, same checks and error messages, same float32
output.  The per-slitlet, per-column median over the 38 rows runs in the CUDA kernel
``k_median_slitlets`` through ``emirwv_median_slitlets[_host]``; there is no CPU fallback.

Mode 2 (one collapsed spectrum) is mode 1 on the GPU followed, as in the reference, by
numina's ``define_mask_borders`` and a masked median of at most 55 values per column; that
tail works on a 55 x NAXIS1 array and stays on the host.  ``define_mask_borders`` is
third-party code restated here from its published behaviour (parity unpinned for it).
"""

import ctypes

import numpy as np

from . import _cabi
from .core import EMIR_NBARS, EMIR_NPIXPERSLIT_RECTIFIED
from .hdu import copy_img


def _check_image(input_image):
    image_header = input_image[0].header
    image2d = input_image[0].data
    naxis2_expected = EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED
    naxis2, naxis1 = image2d.shape
    if naxis2 != naxis2_expected:
        raise ValueError("NAXIS2={0} should be {1}".format(naxis2, naxis2_expected))
    if image_header["instrume"] != "EMIR":
        raise ValueError("INSTRUME keyword is not 'EMIR'!")
    return image2d, naxis1


def median_frames(frames, mode=1):
    """Median spectrum of every slitlet for a stack of rectified products.

    ``frames``: host array ``(n, 2090, naxis1)`` or ``(2090, naxis1)``, float32 or float64
    (other dtypes are converted to float64, like ``numpy.median`` would).
    Returns ``(n, 55, naxis1)`` for mode 1 and ``(n, 2090, naxis1)`` for mode 0, in the
    dtype the arithmetic ran in.
    """
    if mode not in (0, 1):
        raise ValueError("mode must be 0 or 1")
    arr = np.asarray(frames)
    single = arr.ndim == 2
    if single:
        arr = arr[None]
    if arr.ndim != 3 or arr.shape[1] != EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED:
        raise ValueError("NAXIS2={0} should be {1}".format(
            arr.shape[-2], EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED))
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    arr = np.ascontiguousarray(arr)
    n, naxis2, naxis1 = arr.shape
    out_rows = naxis2 if mode == 0 else EMIR_NBARS
    out = np.empty((n, out_rows, naxis1), dtype=arr.dtype)
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_median_slitlets_host(
        arr.ctypes.data_as(vp), out.ctypes.data_as(vp), n, EMIR_NBARS,
        EMIR_NPIXPERSLIT_RECTIFIED, naxis1, int(mode),
        _cabi.F32 if arr.dtype == np.float32 else _cabi.F64))
    return out[0] if single else out


def median_frames_torch(frames, mode=1, out=None):
    """Same for CUDA tensors, on torch's current stream (no host copies)."""
    import torch
    if mode not in (0, 1):
        raise ValueError("mode must be 0 or 1")
    if frames.dim() == 2:
        frames = frames.unsqueeze(0)
    if frames.dtype not in (torch.float32, torch.float64) or not frames.is_cuda \
            or not frames.is_contiguous():
        raise ValueError("frames must be a contiguous float32/float64 CUDA tensor")
    n, naxis2, naxis1 = frames.shape
    if naxis2 != EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED:
        raise ValueError("NAXIS2={0} should be {1}".format(
            naxis2, EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED))
    rows = naxis2 if mode == 0 else EMIR_NBARS
    if out is None:
        out = torch.empty((n, rows, naxis1), dtype=frames.dtype, device=frames.device)
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    _cabi.check(_cabi.load().emirwv_median_slitlets(
        ctypes.c_void_p(frames.data_ptr()), ctypes.c_void_p(out.data_ptr()), n, EMIR_NBARS,
        EMIR_NPIXPERSLIT_RECTIFIED, naxis1, int(mode),
        _cabi.F32 if frames.dtype == torch.float32 else _cabi.F64,
        ctypes.c_void_p(stream) if stream else None))
    return out


def find_pix_borders(sp, sought_value):
    """numina ``find_pix_borders``: first / last index that differs from ``sought_value``,
    ``(-1, naxis1)`` when there is none."""
    sp = np.asarray(sp)
    differs = np.flatnonzero(sp != sought_value)
    if differs.size == 0:
        return -1, len(sp)
    return int(differs[0]), int(differs[-1])


def define_mask_borders(image2d, sought_value, nadditional=0):
    """numina ``define_mask_borders`` (``median_slitlets_rectified.py:118``): mask every
    row up to and including its first / last value that differs from ``sought_value``."""
    naxis2, naxis1 = image2d.shape
    mask2d = np.zeros((naxis2, naxis1), dtype=bool)
    borders = []
    for i in range(naxis2):
        jmin, jmax = find_pix_borders(image2d[i, :], sought_value)
        borders.append((jmin, jmax))
        if (jmin, jmax) != (-1, naxis1):
            if jmin != -1:
                mask2d[i, 0:(jmin + nadditional + 1)] = True
            if jmax != naxis1:
                mask2d[i, (jmax - nadditional):naxis1] = True
    return mask2d, borders


def median_slitlets_rectified(input_image, mode=0, list_useful_slitlets=None, debugplot=0):
    """Compute median spectrum for each slitlet (drop-in for the reference function).

    mode 0: image of the size of the input, the median spectrum of each slitlet spanning
            all its rows; mode 1: image with 55 spectra; mode 2: single collapsed median
            spectrum of the useful slitlets.  ``debugplot`` is accepted and ignored.
    """
    result = copy_img(input_image)
    image2d, naxis1 = _check_image(input_image)
    image2d_median = median_frames(image2d, mode=0 if mode == 0 else 1).astype("float32")
    if mode == 2:
        if list_useful_slitlets is None:
            list_not_useful_slitlets = []
        else:
            list_not_useful_slitlets = [i for i in range(1, EMIR_NBARS + 1)
                                        if i not in list_useful_slitlets]
        mask2d, _ = define_mask_borders(image2d_median, sought_value=0)
        for islitlet in list_not_useful_slitlets:
            mask2d[islitlet - 1, :] = True
        image2d_masked = np.ma.masked_array(image2d_median, mask=mask2d)
        result[0].data = np.ma.median(image2d_masked, axis=0).data.astype("float32")
    else:
        result[0].data = image2d_median
    return result

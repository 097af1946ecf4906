"""Element-wise pre-reduction in front of the path, on the GPU.

The recipes run ``flow(newimg)`` on every frame before ``apply_rectwv_coeff``
(``recipes/spec/abba.py:371``, ``core/recipe.py:25-61``): bad-pixel mask, bias, dark, flat
field.  Here:

* ``FlatFieldCorrector`` mirrors pyemir's own node (``processing/flatfield.py:27-70``): same
  constructor arguments, the same in-place fix of the flat (values <= 0 become 1.0), the same
  ``data / flat`` cast to ``dtype`` and the same header cards; the division runs in
  ``emirwv_prereduce``.  Pinned by golden vectors produced with the reference's class
  (``tests/golden/make_golden_flatfield.py``).
* ``prereduce_torch`` applies bias, dark and flat to a device-resident stack in one pass
  (bias / dark subtraction restate numina's BiasCorrector / DarkCorrector: image minus
  calibration, float32 result; parity unpinned for those two stages).

* ``bpm`` adds the first stage of the flow, numina's BadPixelCorrector
  (``core/correctors.py:23-42``), restated as ``process_bpm_median`` with its 5 x 5 window
  (parity unpinned): a masked pixel becomes the median of the unmasked raw pixels around it,
  in the same pass (``emirwv_prereduce_ex``).  ``BadPixelCorrector`` mirrors the node.
"""

import ctypes
import datetime

import numpy as np

from . import _cabi

_IN_DTYPES = {"float32": _cabi.F32, "float64": _cabi.F64, "uint16": _cabi.U16}


def _cal_tensor(cal, device, want64):
    import torch
    if cal is None:
        return None
    if not isinstance(cal, torch.Tensor):
        cal = torch.from_numpy(np.ascontiguousarray(cal))
    return cal.to(device=device, dtype=torch.float64 if want64 else torch.float32).contiguous()


def prereduce_torch(frames, bias=None, dark=None, flat=None, out=None, bpm=None):
    """``((frames - bias) - dark) / flat'`` as float32, ``flat' = where(flat <= 0, 1, flat)``.

    bpm: (H, W) mask, non-zero = bad pixel: such a pixel of every frame is first replaced by
    the median of the good raw pixels within +-2 rows and columns (0 when there is none).

    frames: (n, H, W) or (H, W) contiguous CUDA tensor, float32 / float64 / uint16.
    bias, dark, flat: (H, W) arrays or tensors, or None to skip the stage.  If any of them is
    float64 all are used as float64 (numpy's promotion: the stage then computes in float64);
    every stage rounds to float32.  ``out`` may be ``frames`` itself for float32 frames.
    """
    import torch
    squeeze = frames.dim() == 2
    if squeeze:
        frames = frames.unsqueeze(0)
    name = str(frames.dtype).replace("torch.", "")
    if frames.dim() != 3 or not frames.is_cuda or not frames.is_contiguous() or \
            name not in _IN_DTYPES:
        raise ValueError("frames must be a contiguous float32/float64/uint16 CUDA tensor")
    n, height, width = frames.shape

    def is64(cal):
        return cal is not None and (cal.dtype == torch.float64 if isinstance(cal, torch.Tensor)
                                    else np.asarray(cal).dtype == np.float64)
    want64 = is64(bias) or is64(dark) or is64(flat)
    cals = [_cal_tensor(c, frames.device, want64) for c in (bias, dark, flat)]
    for cal in cals:
        if cal is not None and tuple(cal.shape) != (height, width):
            raise ValueError("calibration images must have the shape of the frames")
    if out is None:
        out = torch.empty((n, height, width), dtype=torch.float32, device=frames.device)
    elif out.dtype != torch.float32 or tuple(out.shape) != (n, height, width) or \
            not out.is_contiguous():
        raise ValueError("out must be a contiguous float32 tensor of the shape of frames")
    mask = None
    if bpm is not None:
        if not isinstance(bpm, torch.Tensor):
            bpm = torch.from_numpy(np.ascontiguousarray(np.asarray(bpm) != 0))
        mask = (bpm != 0).to(device=frames.device, dtype=torch.uint8).contiguous()
        if tuple(mask.shape) != (height, width):
            raise ValueError("the bad-pixel mask must have the shape of the frames")
        if out.data_ptr() == frames.data_ptr():
            raise ValueError("out must not be frames when a bad-pixel mask is given")
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    vp = ctypes.c_void_p
    _cabi.check(_cabi.load().emirwv_prereduce_ex(
        vp(frames.data_ptr()), _IN_DTYPES[name], vp(out.data_ptr()), n, height, width,
        vp(mask.data_ptr()) if mask is not None else None,
        *[vp(c.data_ptr()) if c is not None else None for c in cals],
        _cabi.F64 if want64 else _cabi.F32, vp(stream) if stream else None))
    return out[0] if squeeze else out


def as_supported_frames(frames):
    """Frames as one of the kernel's input types without changing what numpy would compute.

    float32, float64 and uint16 pass through.  Other types are widened exactly to the type
    numpy's promotion with a float32 calibration image gives: integers of up to 16 bits (and
    float16) to float32, wider integers to float64.
    """
    arr = np.ascontiguousarray(frames)
    if arr.dtype.name in _IN_DTYPES:
        return arr
    return arr.astype(np.result_type(arr.dtype, np.float32))


def prereduce(frames, bias=None, dark=None, flat=None, bpm=None):
    """Host arrays in, float32 host array out (copies around :func:`prereduce_torch`)."""
    import torch
    arr = as_supported_frames(frames)
    dev = torch.device("cuda", torch.cuda.current_device())
    out = prereduce_torch(torch.from_numpy(arr).to(dev), bias, dark, flat, bpm=bpm)
    return out.cpu().numpy()


class BadPixelCorrector:
    """A node that replaces the masked pixels of a frame (numina ``BadPixelCorrector`` as built
    by ``get_corrector_p``, ``core/correctors.py:23-42``; restated, parity unpinned).

    ``BadPixelCorrector(hdul[0].data, datamodel=..., calibid=...)``; ``run(img)`` replaces the
    primary data by the corrected frame (``dtype``, float32 by default) and records the
    calibration in the header like numina's correctors do (``NUM-BPM`` + a HISTORY card).
    """

    def __init__(self, badpixelmask, mark=True, tagger=None, datamodel=None,
                 calibid="calibid-unknown", dtype="float32"):
        self.bpm = np.asarray(badpixelmask)
        self.mark = mark
        self.datamodel = datamodel
        self.calibid = calibid
        self.dtype = dtype

    def __call__(self, img):
        return self.run(img)

    def run(self, img):
        data = img["primary"].data if self.datamodel is None else self.datamodel.get_data(img)
        result = prereduce(data, bpm=self.bpm)
        img["primary"].data = result if np.dtype(self.dtype) == np.float32 else result.astype(self.dtype)
        if self.mark:
            hdr = img["primary"].header
            hdr["NUM-BPM"] = self.calibid
            hdr["history"] = "BPM correction with {}".format(self.calibid)
            hdr["history"] = "BPM correction time {}".format(
                datetime.datetime.now(datetime.timezone.utc).isoformat())
        return img


class FlatFieldCorrector:
    """A node that corrects a frame from flat-field (``processing/flatfield.py:27-70``).

    Same arguments and effects as the reference: ``flatdata`` is fixed IN PLACE (values <= 0
    become 1.0, ``:43``), ``flat_stats`` is its mean, ``run(img)`` replaces the primary data
    by ``(data / flatdata).astype(dtype)`` and appends ``NUM-FF`` and three HISTORY cards.
    """

    def __init__(self, flatdata, datamodel=None, calibid="calibid-unknown", dtype="float32"):
        self.update_variance = False
        self.datamodel = datamodel
        self.calibid = calibid
        self.dtype = dtype
        self.flatdata = flatdata
        self.flatdata[flatdata <= 0] = 1.0  # To avoid NaN
        self.flat_stats = flatdata.mean()

    def __call__(self, img):
        return self.run(img)

    def run(self, img):
        data = img["primary"].data if self.datamodel is None else self.datamodel.get_data(img)
        result = prereduce(data, flat=self.flatdata)
        # the kernel's float32 result IS (data / flat).astype("float32"); other dtypes are a
        # cast of it only when they are wider
        if np.dtype(self.dtype) != np.float32:
            if np.dtype(self.dtype) == np.float64 and \
                    np.result_type(np.asarray(data).dtype, self.flatdata.dtype) == np.float64:
                raise ValueError("dtype float64 needs the float64 quotient: not supported")
            result = result.astype(self.dtype)
        img["primary"].data = result
        hdr = img["primary"].header
        hdr["NUM-FF"] = self.calibid
        hdr["history"] = "Flat-field correction with {}".format(self.calibid)
        hdr["history"] = "Flat-field correction time {}".format(
            datetime.datetime.now(datetime.timezone.utc).isoformat())
        hdr["history"] = "Flat-field correction mean {}".format(self.flat_stats)
        return img

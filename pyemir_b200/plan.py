"""``RectWavePlan``: a RectWaveCoeff turned into device tables (host side).

Holds the opaque ``emirwv_plan`` of the C ABI.  Building a plan replaces what the
reference redoes for every frame: unpacking the 55 ``Slitlet2D`` objects
(slitlet2d.py:145-218), mapping pixel corners through the distortion polynomial
and clipping them (numina ``rectify2d``), and locating every output wavelength
border among the source borders (numina ``resample_image2d_flux``).  Plans are
cached per product/offset/resampling/dtype, so a loop such as the ABBA recipe's
(abba.py:351-381, one RectWaveCoeff for all frames) builds it once.
"""

import ctypes
import hashlib
import json
import threading
from collections import OrderedDict

import numpy as np

from . import _cabi
from .core import (EMIR_NAXIS1, EMIR_NAXIS2, EMIR_NBARS,
                   EMIR_NPIXPERSLIT_RECTIFIED)
from .set_wv_parameters import set_wv_parameters

_DTYPES = {"float32": (_cabi.F32, np.float32), "float64": (_cabi.F64, np.float64)}

# numerically relevant keys of RectWaveCoeff.contents[i]
_GEOMETRY_KEYS = ("bb_nc1_orig", "bb_nc2_orig", "bb_ns1_orig", "bb_ns2_orig",
                  "min_row_rectified", "max_row_rectified", "wpoly_coeff")


def order_fmap(ncoef, nmax_order=4):
    """Polynomial order from the number of coefficients (slitlet2d.py:211-212).

    Same error as numina's ``order_fmap`` above its cap of 4; ``nmax_order=5``
    admits BASELINE config 3 (an extension: the formula is the same).
    """
    order = 1
    while (order + 1) * (order + 2) // 2 != ncoef:
        order += 1
        if order > nmax_order:
            raise ValueError("order > " + str(nmax_order) + " not implemented")
    return order


def normalise_dtype(dtype):
    name = np.dtype(dtype).name
    if name not in _DTYPES:
        raise ValueError(f"dtype must be float32 or float64, not {name}")
    return name


def _valid_islitlets(rectwv_coeff):
    valid = list(range(1, EMIR_NBARS + 1))
    for idel in rectwv_coeff.missing_slitlets:
        valid.remove(idel)          # ValueError on unknown/duplicate, like the reference
    return valid


def _fingerprint(rectwv_coeff, resampling, dtype, max_order, inverse, grid, tiling, device):
    payload = [rectwv_coeff.global_integer_offset_x_pix,
               rectwv_coeff.global_integer_offset_y_pix,
               sorted(rectwv_coeff.missing_slitlets), resampling, dtype, max_order,
               inverse, grid, tiling, device]
    prefix = "tti" if inverse else "ttd"
    for entry in rectwv_coeff.contents:
        if prefix + "_aij" in entry:
            payload.append([entry[k] for k in _GEOMETRY_KEYS]
                           + [list(entry[prefix + "_aij"]), list(entry[prefix + "_bij"])])
        else:
            payload.append(None)
    blob = json.dumps(payload, default=float).encode("utf-8")
    return hashlib.md5(blob).hexdigest()


class RectWavePlan:
    """Device-resident tables for one (RectWaveCoeff, resampling, dtype)."""

    def __init__(self, rectwv_coeff, resampling=2, dtype="float32", max_order=4,
                 grid=None, inverse=False, tile_k=0, stages=0, device=None):
        self._lib = _cabi.load()
        self._handle = ctypes.c_void_p()
        self.dtype = normalise_dtype(dtype)
        self.resampling = int(resampling)
        if grid is None:
            wvp = set_wv_parameters(rectwv_coeff.tags["filter"], rectwv_coeff.tags["grism"])
            grid = (wvp["crpix1_enlarged"], wvp["crval1_enlarged"],
                    wvp["cdelt1_enlarged"], wvp["naxis1_enlarged"])
        self.crpix1, self.crval1, self.cdelt1, self.naxis1_out = grid
        self.valid_islitlets = _valid_islitlets(rectwv_coeff)

        desc = _cabi.PlanDesc()
        desc.naxis1 = EMIR_NAXIS1
        desc.naxis2 = EMIR_NAXIS2
        desc.nslitlets = EMIR_NBARS
        desc.rows_per_slitlet = EMIR_NPIXPERSLIT_RECTIFIED
        offx = rectwv_coeff.global_integer_offset_x_pix
        offy = rectwv_coeff.global_integer_offset_y_pix
        if type(offx) is not int or type(offy) is not int:
            raise ValueError("Invalid non-integer offsets")
        desc.offx, desc.offy = offx, offy
        desc.resampling = self.resampling
        desc.dtype = _DTYPES[self.dtype][0]
        desc.max_order = int(max_order)
        desc.naxis1_out = int(self.naxis1_out)
        desc.crpix1 = float(self.crpix1)
        desc.crval1 = float(self.crval1)
        desc.cdelt1 = float(self.cdelt1)
        desc.tile_k = int(tile_k)
        desc.stages = int(stages)
        prefix = "tti" if inverse else "ttd"
        self.bbox = {}
        for islitlet in range(1, EMIR_NBARS + 1):
            sl = desc.slitlet[islitlet - 1]
            if islitlet not in self.valid_islitlets:
                sl.valid = 0
                continue
            entry = rectwv_coeff.contents[islitlet - 1]
            aij = entry[prefix + "_aij"]
            bij = entry[prefix + "_bij"]
            wpoly = entry["wpoly_coeff"]
            order_fmap(len(aij), int(max_order))
            if len(bij) != len(aij):
                raise ValueError("Unexpected different lengths of aij and bij")
            if len(wpoly) > _cabi.MAX_WPOLY:
                raise ValueError("too many wavelength polynomial coefficients")
            sl.valid = 1
            sl.bb_nc1_orig = int(entry["bb_nc1_orig"])
            sl.bb_nc2_orig = int(entry["bb_nc2_orig"])
            sl.bb_ns1_orig = int(entry["bb_ns1_orig"])
            sl.bb_ns2_orig = int(entry["bb_ns2_orig"])
            sl.min_row_rectified = int(entry["min_row_rectified"])
            sl.max_row_rectified = int(entry["max_row_rectified"])
            sl.ncoef = len(aij)
            sl.nwpoly = len(wpoly)
            for k, (a, b) in enumerate(zip(aij, bij)):
                sl.aij[k] = float(a)
                sl.bij[k] = float(b)
            for k, c in enumerate(wpoly):
                sl.wpoly[k] = float(c)
            self.bbox[islitlet] = (sl.bb_ns2_orig - sl.bb_ns1_orig + 1,
                                   sl.bb_nc2_orig - sl.bb_nc1_orig + 1)
        # the plan lives on `device` (default: the thread's current one); the thread's current
        # device is the same after the call as before
        previous = _cabi.check(self._lib.emirwv_get_device())
        self.device = previous if device is None else int(device)
        if self.device != previous:
            _cabi.check(self._lib.emirwv_set_device(self.device))
        try:
            _cabi.check(self._lib.emirwv_plan_create(ctypes.byref(desc),
                                                     ctypes.byref(self._handle)))
        finally:
            if self.device != previous:
                self._lib.emirwv_set_device(previous)
        self.naxis2_out = EMIR_NBARS * EMIR_NPIXPERSLIT_RECTIFIED

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            self._lib.emirwv_plan_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def np_dtype(self):
        return _DTYPES[self.dtype][1]

    @property
    def out_shape(self):
        return (self.naxis2_out, int(self.naxis1_out))

    def info(self):
        info = _cabi.PlanInfo()
        _cabi.check(self._lib.emirwv_plan_get_info(self._handle, ctypes.byref(info)))
        return info.asdict()

    def set_tuning(self, frames_per_cta=0):
        _cabi.check(self._lib.emirwv_plan_set_tuning(self._handle, int(frames_per_cta)))

    # -- execution -----------------------------------------------------------
    def coverage(self):
        """(55, 2) int32: output columns [lo, hi) of each slitlet that can hold flux."""
        cover = np.zeros((EMIR_NBARS, 2), dtype=np.int32)
        _cabi.check(self._lib.emirwv_plan_get_coverage(
            self._handle, cover.ctypes.data_as(ctypes.c_void_p)))
        return cover

    def run_host(self, frames, out=None, jminmax=None, out_zeroed=False):
        """Host arrays in, host arrays out (copies pipelined inside the library).

        frames: (n, 2048, 2048) C-contiguous array of the plan dtype.
        out_zeroed: ``out`` already holds zeros outside the wavelength coverage of this
        plan (zero-initialised, or filled by an earlier call with the same plan): only the
        covered columns are copied back (EMIRWV_HOST_OUT_ZEROED).
        Returns (out (n, 2090, naxis1_out), jminmax (n, 55, 2) int32, raw 0-based).
        """
        frames = np.ascontiguousarray(frames, dtype=self.np_dtype)
        if frames.ndim == 2:
            frames = frames[None]
        self._check_frames(frames.shape)
        n = frames.shape[0]
        if out is None:
            out = np.empty((n,) + self.out_shape, dtype=self.np_dtype)
            out_zeroed = False
        if jminmax is None:
            jminmax = np.empty((n, EMIR_NBARS, 2), dtype=np.int32)
        assert out.flags.c_contiguous and out.dtype == self.np_dtype
        _cabi.check(self._lib.emirwv_run_host_ex(
            self._handle, frames.ctypes.data_as(ctypes.c_void_p),
            out.ctypes.data_as(ctypes.c_void_p), n,
            jminmax.ctypes.data_as(ctypes.c_void_p),
            _cabi.HOST_OUT_ZEROED if out_zeroed else 0))
        return out, jminmax

    # -- the recipes' flow in front of the path, and the ABBA product behind it -------------
    _RAW_DTYPES = {"float32": _cabi.F32, "float64": _cabi.F64, "uint16": _cabi.U16}

    def set_calibration(self, bpm=None, bias=None, dark=None, flat=None):
        """Calibration images of ``flow(newimg)`` (``recipes/spec/abba.py:371``,
        ``core/recipe.py:25-61``) for :meth:`run_host_raw` / ``abba_host(raw=True)``: (2048, 2048)
        arrays or None; float64 if any of bias / dark / flat is float64 (numpy's promotion),
        else float32; ``flat`` values <= 0 count as 1.0 (``flatfield.py:43``)."""
        cals = [None if c is None else np.asarray(c) for c in (bias, dark, flat)]
        want64 = any(c is not None and c.dtype == np.float64 for c in cals)
        cdt = np.float64 if want64 else np.float32
        cals = [None if c is None else np.ascontiguousarray(c, dtype=cdt) for c in cals]
        mask = None if bpm is None else np.ascontiguousarray(np.asarray(bpm) != 0, dtype=np.uint8)
        for c in cals + [mask]:
            if c is not None and c.shape != (EMIR_NAXIS2, EMIR_NAXIS1):
                raise ValueError("calibration images must be 2048 x 2048")
        vp = ctypes.c_void_p
        ptr = [None if c is None else c.ctypes.data_as(vp) for c in [mask] + cals]
        _cabi.check(self._lib.emirwv_plan_set_calibration(
            self._handle, *ptr, _cabi.F64 if want64 else _cabi.F32))

    def run_host_raw(self, raw, out=None, jminmax=None, out_zeroed=False):
        """RAW frames (uint16 / float32 / float64 host array) through the plan's calibration
        flow and the path: ``(out, jminmax)`` like :meth:`run_host`.  float32 plans only."""
        raw = np.ascontiguousarray(raw)
        if raw.ndim == 2:
            raw = raw[None]
        if raw.dtype.name not in self._RAW_DTYPES:
            raise ValueError("raw frames must be uint16, float32 or float64")
        self._check_frames(raw.shape)
        n = raw.shape[0]
        if out is None:
            out = np.empty((n,) + self.out_shape, dtype=self.np_dtype)
            out_zeroed = False
        if jminmax is None:
            jminmax = np.empty((n, EMIR_NBARS, 2), dtype=np.int32)
        assert out.flags.c_contiguous and out.dtype == self.np_dtype
        vp = ctypes.c_void_p
        _cabi.check(self._lib.emirwv_run_host_raw(
            self._handle, raw.ctypes.data_as(vp), self._RAW_DTYPES[raw.dtype.name],
            out.ctypes.data_as(vp), n, jminmax.ctypes.data_as(vp),
            _cabi.HOST_OUT_ZEROED if out_zeroed else 0))
        return out, jminmax

    def abba_host(self, frames, labels, offsets=None, method="mean", low=3.0, high=3.0,
                  raw=False, out=None):
        """The ABBA product of host frames in one call (``abba.py:351-381,545-606``): every
        frame rectified, shifted by its vertical offset, A and B images combined with
        ``method`` ("mean", "median", "sigmaclip"), float32(A) - float32(B).  Only the
        (2090, naxis1) float32 image comes back.  ``raw``: the frames are raw and go through
        the plan's calibration flow first."""
        frames = np.ascontiguousarray(frames)
        self._check_frames(frames.shape)
        n = frames.shape[0]
        if raw:
            if frames.dtype.name not in self._RAW_DTYPES:
                raise ValueError("raw frames must be uint16, float32 or float64")
            code = self._RAW_DTYPES[frames.dtype.name]
        else:
            frames = np.ascontiguousarray(frames, dtype=self.np_dtype)
            code = _DTYPES[self.dtype][0]
        lab = np.ascontiguousarray(labels, dtype=np.int8)
        if lab.shape != (n,):
            raise ValueError("one label per frame is needed")
        off = None
        if offsets is not None:
            off = np.ascontiguousarray(offsets, dtype=np.float64)
            if off.shape != (n,):
                raise ValueError("one offset per frame is needed")
        methods = {"mean": _cabi.MEAN, "median": _cabi.MEDIAN, "sigmaclip": _cabi.SIGMACLIP}
        if method not in methods:
            raise ValueError(f"Unexpected method: {method}")
        if out is None:
            out = np.empty(self.out_shape, dtype=np.float32)
        assert out.flags.c_contiguous and out.dtype == np.float32 and out.shape == self.out_shape
        vp = ctypes.c_void_p
        _cabi.check(self._lib.emirwv_abba_host(
            self._handle, frames.ctypes.data_as(vp), code, n, lab.ctypes.data_as(vp),
            off.ctypes.data_as(vp) if off is not None else None, methods[method], float(low),
            float(high), out.ctypes.data_as(vp), _cabi.HOST_RAW if raw else 0))
        return out

    def run_device(self, d_in, d_out, nframes, d_jminmax=0, stream=0, out_zeroed=False):
        """Raw device pointers (ints); asynchronous on ``stream``.  ``out_zeroed``: the
        output buffer already holds zeros outside the coverage (EMIRWV_RUN_OUT_ZEROED)."""
        _cabi.check(self._lib.emirwv_run_ex(
            self._handle, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out), int(nframes),
            ctypes.c_void_p(d_jminmax) if d_jminmax else None,
            ctypes.c_void_p(stream) if stream else None,
            _cabi.RUN_OUT_ZEROED if out_zeroed else 0))

    def run_torch(self, frames, out=None, jminmax=None, out_zeroed=False):
        """CUDA tensors in / out on torch's current stream (no host copies)."""
        import torch
        if frames.dim() == 2:
            frames = frames.unsqueeze(0)
        tdtype = torch.float32 if self.dtype == "float32" else torch.float64
        if frames.dtype != tdtype or not frames.is_cuda or not frames.is_contiguous():
            raise ValueError(f"frames must be a contiguous CUDA tensor of {tdtype}")
        self._check_frames(tuple(frames.shape))
        n = frames.shape[0]
        if out is None:
            out = torch.empty((n,) + self.out_shape, dtype=tdtype, device=frames.device)
            out_zeroed = False
        if jminmax is None:
            jminmax = torch.empty((n, EMIR_NBARS, 2), dtype=torch.int32, device=frames.device)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        self.run_device(frames.data_ptr(), out.data_ptr(), n, jminmax.data_ptr(), stream,
                        out_zeroed)
        return out, jminmax

    def rectify_slitlet_torch(self, frame, islitlet):
        """Kernel 1 alone: (39, bbw) rectified rows min_row..max_row+1 of a slitlet."""
        import torch
        bbw = self.bbox[islitlet][1]
        rect = torch.empty((_cabi.ROWS_SCANNED, bbw), dtype=frame.dtype, device=frame.device)
        stream = torch.cuda.current_stream(frame.device).cuda_stream
        _cabi.check(self._lib.emirwv_rectify_slitlet(
            self._handle, ctypes.c_void_p(frame.data_ptr()), int(islitlet),
            ctypes.c_void_p(rect.data_ptr()), ctypes.c_void_p(stream) if stream else None))
        return rect

    # -- parity hooks ----------------------------------------------------------
    def debug_geometry(self, islitlet, weights=True):
        bbw = self.bbox[islitlet][1]
        k = self.info()["footprint"]
        rows = np.empty((_cabi.ROWS_SCANNED, bbw), dtype=np.int32)
        cols = np.empty_like(rows)
        ok = np.empty(rows.shape, dtype=np.uint8)
        w = np.empty(rows.shape + (k, k), dtype=np.float64) if weights else None
        _cabi.check(self._lib.emirwv_debug_geometry(
            self._handle, int(islitlet), rows.ctypes.data_as(ctypes.c_void_p),
            cols.ctypes.data_as(ctypes.c_void_p), ok.ctypes.data_as(ctypes.c_void_p),
            w.ctypes.data_as(ctypes.c_void_p) if weights else None))
        return rows, cols, ok.astype(bool), w

    def debug_borders(self, islitlet):
        nb = int(self.naxis1_out) + 1
        idx = np.empty(nb, dtype=np.int32)
        t = np.empty(nb, dtype=np.float64)
        _cabi.check(self._lib.emirwv_debug_borders(
            self._handle, int(islitlet), idx.ctypes.data_as(ctypes.c_void_p),
            t.ctypes.data_as(ctypes.c_void_p)))
        return idx, t

    def _check_frames(self, shape):
        # Slitlet2D.extract_slitlet2d (slitlet2d.py:359-363)
        if shape[-1] != EMIR_NAXIS1:
            raise ValueError("Unexpected naxis1")
        if shape[-2] != EMIR_NAXIS2:
            raise ValueError("Unexpected naxis2")


# --------------------------------------------------------------------- plan cache
_CACHE = OrderedDict()
_CACHE_LOCK = threading.Lock()
_CACHE_SIZE = 2          # a plan holds up to a few hundred MB of device tables


_KEY_MEMO = {}           # quick look -> full fingerprint of the product it was taken from


def _quick_look(rectwv_coeff, options):
    """What changes when a product is edited in place, without serialising all of it: the
    object, its uuid, offsets, missing slitlets and one coefficient of every table."""
    prefix = "tti" if options[3] else "ttd"
    probe = []
    for entry in rectwv_coeff.contents:
        coef = entry.get(prefix + "_aij")
        if coef:
            wpoly = entry["wpoly_coeff"]
            probe += [coef[0], coef[-1], entry[prefix + "_bij"][0], wpoly[0], wpoly[-1]]
    probe = hash(tuple(probe))
    return (id(rectwv_coeff), getattr(rectwv_coeff, "uuid", None),
            rectwv_coeff.global_integer_offset_x_pix, rectwv_coeff.global_integer_offset_y_pix,
            tuple(rectwv_coeff.missing_slitlets), len(rectwv_coeff.contents), probe,
            rectwv_coeff.tags.get("grism", ""), rectwv_coeff.tags.get("filter", "")) + options


def get_plan(rectwv_coeff, resampling=2, dtype="float32", max_order=4, grid=None,
             inverse=False, tile_k=0, stages=0, device=None):
    """Cached ``RectWavePlan`` for this product and options."""
    dtype = normalise_dtype(dtype)
    if device is None:          # the key names the device the plan will really live on
        device = _cabi.check(_cabi.load().emirwv_get_device())
    options = (int(resampling), dtype, int(max_order), bool(inverse),
               grid if grid is None else tuple(grid), (tile_k, stages), device)
    # the full fingerprint serialises every coefficient (0.6 ms): the loop over exposures of a
    # recipe passes the same object again and again, so it is remembered per quick look
    look = _quick_look(rectwv_coeff, options)
    key = _KEY_MEMO.get(look)
    if key is None:
        key = _fingerprint(rectwv_coeff, *options)
        key += ":" + rectwv_coeff.tags.get("grism", "") + ":" + rectwv_coeff.tags.get("filter", "")
        if len(_KEY_MEMO) > 64:
            _KEY_MEMO.clear()
        _KEY_MEMO[look] = key
    with _CACHE_LOCK:
        plan = _CACHE.get(key)
        if plan is not None:
            _CACHE.move_to_end(key)
            return plan
    plan = RectWavePlan(rectwv_coeff, resampling, dtype, max_order, grid, inverse,
                        tile_k, stages, device)
    with _CACHE_LOCK:
        _CACHE[key] = plan
        while len(_CACHE) > _CACHE_SIZE:
            _CACHE.popitem(last=False)      # freed when the last user drops it
    return plan


def clear_plan_cache():
    with _CACHE_LOCK:
        _CACHE.clear()
    _KEY_MEMO.clear()

"""Pinned (page-locked) host arrays for the results of the single-frame drop-in.

``apply_rectwv_coeff`` returns a new float32 image per call (``apply_rectwv_coeff.py:128``:
``np.zeros``).  A device -> host copy into fresh pageable memory costs several times the copy
into page-locked memory (page faults + the driver's staging), and page-locking a buffer per
call costs as much again, so result arrays are carved from a small pool of buffers from
``emirwv_host_alloc``: an array handed out is an ordinary ``numpy.ndarray``; when the caller
drops it (and every view of it) its buffer goes back to the pool -- the usual loop over
exposures keeps reusing one or two buffers.
"""

import ctypes
import threading
import weakref

import numpy as np

from . import _cabi

_LOCK = threading.Lock()
_FREE = {}                   # nbytes -> [pointers]
_MAX_FREE_BYTES = 1 << 30    # what the pool keeps for reuse; beyond that buffers are released
_free_bytes = 0


def _give_back(ptr, nbytes):
    global _free_bytes
    try:
        with _LOCK:
            if _free_bytes + nbytes <= _MAX_FREE_BYTES:
                _FREE.setdefault(nbytes, []).append(ptr)
                _free_bytes += nbytes
                return
        _cabi.load().emirwv_host_free(ctypes.c_void_p(ptr))
    except Exception:        # interpreter shutdown
        pass


def empty(shape, dtype):
    """An uninitialised C-contiguous array in page-locked memory (pageable if that fails)."""
    global _free_bytes
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype=dtype)
    ptr = None
    with _LOCK:
        bucket = _FREE.get(nbytes)
        if bucket:
            ptr = bucket.pop()
            _free_bytes -= nbytes
    if ptr is None:
        ptr = _cabi.load().emirwv_host_alloc(nbytes)
        if not ptr:
            return np.empty(shape, dtype=dtype)
    raw = (ctypes.c_byte * nbytes).from_address(ptr)
    weakref.finalize(raw, _give_back, ptr, nbytes)
    return np.frombuffer(raw, dtype=dtype).reshape(shape)


def trim():
    """Release every buffer the pool holds for reuse."""
    global _free_bytes
    with _LOCK:
        ptrs = [p for bucket in _FREE.values() for p in bucket]
        _FREE.clear()
        _free_bytes = 0
    for ptr in ptrs:
        _cabi.load().emirwv_host_free(ctypes.c_void_p(ptr))

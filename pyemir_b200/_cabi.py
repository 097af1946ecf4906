"""ctypes binding of ``libemirwv.so`` (C ABI declared in ``include/emirwv.h``).

Thin by design: structures, prototypes, error translation.  The library must be
present — there is no fallback implementation; loading or calling it without the
CUDA build fails loudly.
"""

import ctypes
import os

from . import _build

MAX_SLITLETS = 55
MAX_COEF = 21
MAX_WPOLY = 16
ROWS_SCANNED = 39

F32, F64, U16 = 0, 1, 2
HOST_OUT_ZEROED = 1
HOST_RAW = 2
RUN_OUT_ZEROED = 1
MEAN, MEDIAN, SIGMACLIP = 0, 1, 2
NEAREST, AREA, BILINEAR = 1, 2, 3

E_VALUE, E_INDEX, E_CUDA, E_NOMEM, E_LIMIT = -1, -2, -3, -4, -5

# every symbol include/emirwv.h declares
EXPORTED_SYMBOLS = (
    "emirwv_version", "emirwv_last_error", "emirwv_device_count", "emirwv_set_device", "emirwv_get_device",
    "emirwv_plan_create", "emirwv_plan_destroy", "emirwv_plan_get_info",
    "emirwv_plan_set_tuning", "emirwv_run", "emirwv_run_ex", "emirwv_run_host",
    "emirwv_run_host_ex", "emirwv_plan_get_coverage",
    "emirwv_host_alloc", "emirwv_host_free", "emirwv_rectify_slitlet",
    "emirwv_rectify2d_host", "emirwv_debug_geometry", "emirwv_debug_borders",
    "emirwv_median_slitlets", "emirwv_median_slitlets_host",
    "emirwv_abba_partial", "emirwv_abba_finish", "emirwv_stack_median", "emirwv_abba_median",
    "emirwv_prereduce", "emirwv_prereduce_ex", "emirwv_stack_sigmaclip",
    "emirwv_abba_sigmaclip", "emirwv_shift_rows", "emirwv_plan_set_calibration",
    "emirwv_run_host_raw", "emirwv_abba_host", "emirwv_gather", "emirwv_abba_reduce",
    "emirwv_peer_create", "emirwv_peer_connect", "emirwv_peer_disconnect", "emirwv_peer_destroy", "emirwv_peer_timed_out",
    "emirwv_abba_mean_push",
)


class SlitletDesc(ctypes.Structure):
    _fields_ = [
        ("valid", ctypes.c_int32),
        ("bb_nc1_orig", ctypes.c_int32),
        ("bb_nc2_orig", ctypes.c_int32),
        ("bb_ns1_orig", ctypes.c_int32),
        ("bb_ns2_orig", ctypes.c_int32),
        ("min_row_rectified", ctypes.c_int32),
        ("max_row_rectified", ctypes.c_int32),
        ("ncoef", ctypes.c_int32),
        ("nwpoly", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("aij", ctypes.c_double * MAX_COEF),
        ("bij", ctypes.c_double * MAX_COEF),
        ("wpoly", ctypes.c_double * MAX_WPOLY),
    ]


class PlanDesc(ctypes.Structure):
    _fields_ = [
        ("naxis1", ctypes.c_int32),
        ("naxis2", ctypes.c_int32),
        ("nslitlets", ctypes.c_int32),
        ("rows_per_slitlet", ctypes.c_int32),
        ("offx", ctypes.c_int32),
        ("offy", ctypes.c_int32),
        ("resampling", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("max_order", ctypes.c_int32),
        ("naxis1_out", ctypes.c_int32),
        ("crpix1", ctypes.c_double),
        ("crval1", ctypes.c_double),
        ("cdelt1", ctypes.c_double),
        ("tile_k", ctypes.c_int32),
        ("stages", ctypes.c_int32),
        ("slitlet", SlitletDesc * MAX_SLITLETS),
    ]


class PlanInfo(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32),
        ("resampling", ctypes.c_int32),
        ("footprint", ctypes.c_int32),
        ("naxis1_out", ctypes.c_int32),
        ("naxis2_out", ctypes.c_int32),
        ("ntiles", ctypes.c_int32),
        ("ntiles_empty", ctypes.c_int32),
        ("tile_k", ctypes.c_int32),
        ("stages", ctypes.c_int32),
        ("window_rows", ctypes.c_int32),
        ("window_cols", ctypes.c_int32),
        ("max_span", ctypes.c_int32),
        ("frames_per_cta", ctypes.c_int32),
        ("threads", ctypes.c_int32),
        ("smem_bytes", ctypes.c_int32),
        ("ctas_per_sm", ctypes.c_int32),
        ("table_bytes", ctypes.c_int64),
        ("in_frame_bytes", ctypes.c_int64),
        ("out_frame_bytes", ctypes.c_int64),
    ]

    def asdict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class EmirwvError(RuntimeError):
    """CUDA / limit failure reported by libemirwv."""


_lib = None


def library_path():
    """The product library, or the experiment variant named by ``EMIRWV_LIB``
    (a file name inside the package directory or an absolute path)."""
    override = os.environ.get("EMIRWV_LIB")
    if override:
        return override if os.path.isabs(override) else os.path.join(
            os.path.dirname(_build.LIB_PATH), override)
    return _build.LIB_PATH


def load():
    """Load libemirwv.so and declare prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -m pyemir_b200._build` "
            "(nvcc, sm_100a).  pyemir_b200 has no CPU fallback.")
    lib = ctypes.CDLL(path)
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    lib.emirwv_version.restype = i32
    lib.emirwv_version.argtypes = []
    lib.emirwv_last_error.restype = ctypes.c_char_p
    lib.emirwv_last_error.argtypes = []
    lib.emirwv_device_count.restype = i32
    lib.emirwv_device_count.argtypes = []
    lib.emirwv_get_device.restype = i32
    lib.emirwv_get_device.argtypes = []
    lib.emirwv_set_device.restype = i32
    lib.emirwv_set_device.argtypes = [i32]
    lib.emirwv_plan_create.restype = i32
    lib.emirwv_plan_create.argtypes = [ctypes.POINTER(PlanDesc), ctypes.POINTER(vp)]
    lib.emirwv_plan_destroy.restype = None
    lib.emirwv_plan_destroy.argtypes = [vp]
    lib.emirwv_plan_get_info.restype = i32
    lib.emirwv_plan_get_info.argtypes = [vp, ctypes.POINTER(PlanInfo)]
    lib.emirwv_plan_set_tuning.restype = i32
    lib.emirwv_plan_set_tuning.argtypes = [vp, i32]
    lib.emirwv_run.restype = i32
    lib.emirwv_run.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.emirwv_run_ex.restype = i32
    lib.emirwv_run_ex.argtypes = [vp, vp, vp, i32, vp, vp, ctypes.c_uint32]
    lib.emirwv_run_host.restype = i32
    lib.emirwv_run_host.argtypes = [vp, vp, vp, i32, vp]
    lib.emirwv_run_host_ex.restype = i32
    lib.emirwv_run_host_ex.argtypes = [vp, vp, vp, i32, vp, ctypes.c_uint32]
    lib.emirwv_plan_get_coverage.restype = i32
    lib.emirwv_plan_get_coverage.argtypes = [vp, vp]
    lib.emirwv_host_alloc.restype = vp
    lib.emirwv_host_alloc.argtypes = [ctypes.c_int64]
    lib.emirwv_host_free.restype = None
    lib.emirwv_host_free.argtypes = [vp]
    lib.emirwv_rectify_slitlet.restype = i32
    lib.emirwv_rectify_slitlet.argtypes = [vp, vp, i32, vp, vp]
    lib.emirwv_rectify2d_host.restype = i32
    lib.emirwv_rectify2d_host.argtypes = [vp, i32, i32, vp, vp, i32, i32, i32, vp]
    lib.emirwv_median_slitlets.restype = i32
    lib.emirwv_median_slitlets.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.emirwv_median_slitlets_host.restype = i32
    lib.emirwv_median_slitlets_host.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32]
    lib.emirwv_abba_partial.restype = i32
    lib.emirwv_abba_partial.argtypes = [vp, vp, i32, ctypes.c_int64, i32, vp, vp, i32, vp]
    lib.emirwv_abba_finish.restype = i32
    lib.emirwv_abba_finish.argtypes = [vp, vp, i32, i32, ctypes.c_int64, vp, vp]
    lib.emirwv_stack_median.restype = i32
    lib.emirwv_stack_median.argtypes = [vp, i32, vp, i32, ctypes.c_int64, i32, vp, vp]
    lib.emirwv_abba_median.restype = i32
    lib.emirwv_abba_median.argtypes = [vp, i32, vp, i32, vp, i32, ctypes.c_int64, i32, vp, vp]
    lib.emirwv_prereduce.restype = i32
    lib.emirwv_prereduce.argtypes = [vp, i32, vp, i32, ctypes.c_int64, vp, vp, vp, i32, vp]
    lib.emirwv_prereduce_ex.restype = i32
    lib.emirwv_prereduce_ex.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]
    f64 = ctypes.c_double
    lib.emirwv_stack_sigmaclip.restype = i32
    lib.emirwv_stack_sigmaclip.argtypes = [vp, i32, vp, i32, ctypes.c_int64, i32, f64, f64, vp, vp]
    lib.emirwv_abba_sigmaclip.restype = i32
    lib.emirwv_abba_sigmaclip.argtypes = [vp, i32, vp, i32, vp, i32, ctypes.c_int64, i32, f64, f64,
                                          vp, vp]
    lib.emirwv_shift_rows.restype = i32
    lib.emirwv_shift_rows.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp]
    lib.emirwv_plan_set_calibration.restype = i32
    lib.emirwv_plan_set_calibration.argtypes = [vp, vp, vp, vp, vp, i32]
    lib.emirwv_run_host_raw.restype = i32
    lib.emirwv_run_host_raw.argtypes = [vp, vp, i32, vp, i32, vp, ctypes.c_uint32]
    lib.emirwv_abba_host.restype = i32
    lib.emirwv_abba_host.argtypes = [vp, vp, i32, i32, vp, vp, i32, f64, f64, vp, ctypes.c_uint32]
    lib.emirwv_gather.restype = i32
    lib.emirwv_gather.argtypes = [vp, vp, i32, i32, ctypes.c_int64, vp, vp, i32, vp]
    lib.emirwv_abba_reduce.restype = i32
    lib.emirwv_abba_reduce.argtypes = [vp, vp, ctypes.c_int64, vp, i32, vp]
    lib.emirwv_peer_create.restype = i32
    lib.emirwv_peer_create.argtypes = [ctypes.c_int64, i32, i32, ctypes.POINTER(vp), vp]
    lib.emirwv_peer_connect.restype = i32
    lib.emirwv_peer_connect.argtypes = [vp, vp]
    lib.emirwv_peer_disconnect.restype = i32
    lib.emirwv_peer_disconnect.argtypes = [vp]
    lib.emirwv_peer_destroy.restype = None
    lib.emirwv_peer_destroy.argtypes = [vp]
    lib.emirwv_peer_timed_out.restype = i32
    lib.emirwv_peer_timed_out.argtypes = [vp]
    lib.emirwv_abba_mean_push.restype = i32
    lib.emirwv_abba_mean_push.argtypes = [vp, vp, vp, i32, i32, vp, vp, i32, i32, i32,
                                          ctypes.POINTER(vp), vp]
    lib.emirwv_debug_geometry.restype = i32
    lib.emirwv_debug_geometry.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.emirwv_debug_borders.restype = i32
    lib.emirwv_debug_borders.argtypes = [vp, i32, vp, vp]
    _lib = lib
    return lib


def check(code):
    """Translate a negative return code into the exception the reference raises."""
    if code >= 0:
        return code
    message = load().emirwv_last_error().decode("utf-8", "replace")
    if code == E_VALUE:
        raise ValueError(message)
    if code == E_INDEX:
        raise IndexError(message)
    if code == E_NOMEM:
        raise MemoryError(message)
    raise EmirwvError(message)

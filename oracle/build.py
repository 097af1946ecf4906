"""Build recipe for the oracle's C helper (gcc on oracle/clippix.c).

ORACLE — test infrastructure only.
"""

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "clippix.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_LIB = os.path.join(_OUT_DIR, "liboracle_clippix.so")


def build(force=False):
    """Compile clippix.c into oracle/_build/liboracle_clippix.so (idempotent)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_LIB)
            and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC)):
        return _LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


_lib = None


def load_clippix():
    """Return the ctypes handle of the helper, building it when gcc is there."""
    global _lib
    if _lib is not None:
        return _lib
    path = _LIB
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_SRC):
        path = build()
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lp = ctypes.POINTER(ctypes.c_long)
    lib.oracle_resample.restype = None
    lib.oracle_resample.argtypes = [dp, ctypes.c_long, ctypes.c_long, dp, dp,
                                    ctypes.c_long, lp, lp, dp, ctypes.c_long,
                                    ctypes.c_long]
    lib.oracle_quad_weights.restype = None
    lib.oracle_quad_weights.argtypes = [dp, dp, ctypes.c_long, ctypes.c_long,
                                        ctypes.c_long, ctypes.c_long, dp]
    lib.oracle_poly_box_area.restype = ctypes.c_double
    lib.oracle_poly_box_area.argtypes = [dp, dp, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_double, ctypes.c_double,
                                         ctypes.c_double]
    _lib = lib
    return lib


if __name__ == "__main__":
    print(build(force=True))

"""The C-ABI library loads on a CPU-only box and exports what include/emirwv.h declares."""

import ctypes
import os
import re

import pytest

from pyemir_b200 import _build, _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "emirwv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emirwv_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_cabi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    path = _build.build_library()            # no-op when the .so is current
    lib = ctypes.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert _cabi.load().emirwv_version() == 200


def test_struct_sizes_match_the_header():
    # emirwv_slitlet_desc: 10 int32 + (21 + 21 + 16) doubles; plan_desc: 10 int32,
    # 3 doubles, 2 int32, 55 slitlets; plan_info: 16 int32 + 3 int64
    assert ctypes.sizeof(_cabi.SlitletDesc) == 10 * 4 + 58 * 8
    assert ctypes.sizeof(_cabi.PlanDesc) == 40 + 24 + 8 + 55 * ctypes.sizeof(_cabi.SlitletDesc)
    assert ctypes.sizeof(_cabi.PlanInfo) == 16 * 4 + 3 * 8


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _cabi.load()
    assert lib.emirwv_device_count() == _cabi.E_CUDA
    with pytest.raises(_cabi.EmirwvError):
        _cabi.check(lib.emirwv_set_device(0))
    from pyemir_b200 import synthetic
    from pyemir_b200.plan import RectWavePlan
    with pytest.raises(_cabi.EmirwvError):
        RectWavePlan(synthetic.make_config(1))

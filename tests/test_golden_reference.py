"""Golden vectors produced by the reference's OWN code (tests/golden/make_golden.py).

The fixtures were generated in the build container by executing the reference's
apply_rectwv_coeff.py, slitlet2d.py, set_wv_parameters.py and core/__init__.py from
/root/reference (numina's arithmetic came from the oracle restatement, see the generator's
docstring).  They pin everything pyemir itself does on the path:

* CPU: the oracle's restated orchestration and the host mirrors reproduce them exactly;
* GPU: the CUDA drop-in reproduces the header cards exactly and the flux to the float32 cast.
"""

import hashlib
import importlib.util
import json
import os

import numpy as np
import pytest

from oracle import pipeline
from pyemir_b200 import core
from pyemir_b200.set_wv_parameters import set_wv_parameters

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load_generator():
    spec = importlib.util.spec_from_file_location("make_golden",
                                                  os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # imports oracle + pyemir_b200 only; never the reference
    return mod


GEN = _load_generator()
with open(os.path.join(GOLDEN, "apply_rectwv_coeff.json")) as _fh:
    META = json.load(_fh)
with open(os.path.join(GOLDEN, "set_wv_parameters.json")) as _fh:
    WVP = json.load(_fh)
SUB = np.load(os.path.join(GOLDEN, "apply_rectwv_coeff_subsample.npz"))
CASES = {c[0]: c for c in GEN.CASES}


def _cards(hdul):
    return [list(c) for c in GEN.cards_of(hdul[0].header)]


def _golden_cards(name):
    return [list(c) for c in META["cases"][name]["cards"]]


# ---------------------------------------------------------------- host mirrors
def test_core_constants_match_reference():
    for key, value in WVP["constants"].items():
        assert getattr(core, key) == value, key


@pytest.mark.parametrize("combo", sorted(WVP["table"]))
def test_set_wv_parameters_matches_reference(combo):
    grism, filt = combo.split("+")
    want = WVP["table"][combo]
    if "ValueError" in want:
        with pytest.raises(ValueError) as exc:
            set_wv_parameters(filt, grism)
        assert str(exc.value) == want["ValueError"]
        return
    got = set_wv_parameters(filt, grism)
    assert sorted(got) == sorted(want)
    for key, value in want.items():
        if isinstance(value, dict) and "Polynomial" in value:
            assert [float(c) for c in got[key].coef] == value["Polynomial"], key
        elif isinstance(value, list):
            assert np.asarray(got[key]).tolist() == value, key
        else:
            assert got[key] == value, key


# ---------------------------------------------------------------- oracle orchestration
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_pipeline_reproduces_reference_orchestration(name):
    case = CASES[name]
    want = META["cases"][name]
    hdul, coeff = GEN.build_case(*case)
    out = pipeline.apply_rectwv_coeff(hdul, coeff, args_resampling=case[3])
    data = out[0].data
    assert list(data.shape) == want["shape"] and str(data.dtype) == want["dtype"]
    assert np.array_equal(data[::META["row_step"], ::META["col_step"]], SUB[name])
    assert hashlib.sha256(np.ascontiguousarray(data).tobytes()).hexdigest() == want["sha256"]
    assert _cards(out) == _golden_cards(name)


def test_oracle_pipeline_errors_match_reference():
    errors = META["errors"]
    hdul, coeff = GEN.build_case(*GEN.CASES[0])
    hdul[0].header["filter"] = "H"
    with pytest.raises(ValueError) as exc:
        pipeline.apply_rectwv_coeff(hdul, coeff)
    assert str(exc.value) == errors["filter_mismatch"]
    hdul, coeff = GEN.build_case(*GEN.CASES[0])
    hdul[0].header["grism"] = "K"
    with pytest.raises(ValueError) as exc:
        pipeline.apply_rectwv_coeff(hdul, coeff)
    assert str(exc.value) == errors["grism_mismatch"]
    hdul, coeff = GEN.build_case(*GEN.CASES[0])
    hdul[0].header["xdtu"] = float(hdul[0].header["xdtu"]) + 0.5
    with pytest.raises(ValueError) as exc:
        pipeline.apply_rectwv_coeff(hdul, coeff, args_ignore_dtu_configuration=False)
    assert str(exc.value) == errors["dtu_mismatch"]
    hdul, coeff = GEN.build_case(*GEN.CASES[0])
    with pytest.raises(ValueError) as exc:
        pipeline.apply_rectwv_coeff(hdul, coeff, args_resampling=3)
    assert str(exc.value) == errors["bad_resampling"]


# ---------------------------------------------------------------- the CUDA drop-in
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_dropin_reproduces_reference_vectors(name):
    from pyemir_b200.apply_rectwv_coeff import apply_rectwv_coeff
    case = CASES[name]
    want = META["cases"][name]
    hdul, coeff = GEN.build_case(*case)
    # float64 kernels + the reference's float32 cast: equal to one float32 ulp
    h64 = GEN.synthetic.make_hdulist(hdul[0].data.astype(np.float64), coeff)
    out = apply_rectwv_coeff(h64, coeff, args_resampling=case[3])
    data = out[0].data
    assert list(data.shape) == want["shape"] and str(data.dtype) == want["dtype"]
    # IMNSLT/IMXSLT/JMNSLT/JMXSLT, WCS and HISTORY cards: bit exact, same order
    assert _cards(out) == _golden_cards(name)
    sub = data[::META["row_step"], ::META["col_step"]]
    ref = SUB[name]
    assert np.array_equal(sub == 0, ref == 0)                 # masks: exact
    assert np.all(np.abs(sub - ref) <= np.spacing(np.abs(ref)) + 1e-12 * np.abs(ref).max())
    assert abs(float(data.astype(np.float64).sum()) - want["sum"]) <= 1e-6 * abs(want["sum"])
    # float32 kernels: flux tolerance of the north star (1e-5 relative)
    out32 = apply_rectwv_coeff(hdul, coeff, args_resampling=case[3])
    assert _cards(out32) == _golden_cards(name)
    sub32 = out32[0].data[::META["row_step"], ::META["col_step"]]
    scale = np.abs(ref).max()
    assert np.all(np.abs(sub32 - ref) <= 1e-5 * np.abs(ref) + 2e-7 * scale)


@pytest.mark.gpu
def test_cuda_dropin_errors_match_reference():
    from pyemir_b200.apply_rectwv_coeff import apply_rectwv_coeff
    errors = META["errors"]
    hdul, coeff = GEN.build_case(*GEN.CASES[0])
    hdul[0].header["filter"] = "H"
    with pytest.raises(ValueError) as exc:
        apply_rectwv_coeff(hdul, coeff)
    assert str(exc.value) == errors["filter_mismatch"]
    hdul, coeff = GEN.build_case(*GEN.CASES[0])
    hdul[0].header["xdtu"] = float(hdul[0].header["xdtu"]) + 0.5
    with pytest.raises(ValueError) as exc:
        apply_rectwv_coeff(hdul, coeff, args_ignore_dtu_configuration=False)
    assert str(exc.value) == errors["dtu_mismatch"]


# ---------------------------------------------------------------- median_slitlets_rectified
MEDIAN_KEYS = sorted(k for k in META["median"] if k.startswith("mode"))


@pytest.fixture(scope="module")
def median_input():
    # the oracle's image is bit-identical to the reference's (sha256 test above), so the
    # generator's input can be rebuilt without the reference
    return GEN.median_input(pipeline)


def _check_median(out, key):
    want = META["median"][key]
    data = out[0].data
    assert list(data.shape) == want["shape"] and data.dtype == np.float32
    assert hashlib.sha256(np.ascontiguousarray(data).tobytes()).hexdigest() == want["sha256"]


@pytest.mark.parametrize("key", MEDIAN_KEYS)
def test_oracle_median_slitlets_reproduces_reference(median_input, key):
    want = META["median"][key]
    out = pipeline.median_slitlets_rectified(median_input, mode=want["mode"],
                                             list_useful_slitlets=want["useful"])
    _check_median(out, key)


def test_oracle_median_slitlets_errors(median_input):
    from pyemir_b200.hdu import copy_img
    bad = copy_img(median_input)
    bad[0].data = bad[0].data[:100]
    with pytest.raises(ValueError) as exc:
        pipeline.median_slitlets_rectified(bad)
    assert str(exc.value) == META["median"]["error_naxis2"]
    bad = copy_img(median_input)
    bad[0].header["instrume"] = "OSIRIS"
    with pytest.raises(ValueError) as exc:
        pipeline.median_slitlets_rectified(bad)
    assert str(exc.value) == META["median"]["error_instrume"]


@pytest.mark.gpu
@pytest.mark.parametrize("key", MEDIAN_KEYS)
def test_cuda_median_slitlets_reproduces_reference(median_input, key):
    from pyemir_b200.median_slitlets_rectified import median_slitlets_rectified
    want = META["median"][key]
    before = median_input[0].data.copy()
    out = median_slitlets_rectified(median_input, mode=want["mode"],
                                    list_useful_slitlets=want["useful"])
    assert np.array_equal(median_input[0].data, before, equal_nan=True)    # input untouched
    _check_median(out, key)                                                # bit exact


@pytest.mark.gpu
def test_cuda_median_slitlets_errors(median_input):
    from pyemir_b200.hdu import copy_img
    from pyemir_b200.median_slitlets_rectified import median_slitlets_rectified
    bad = copy_img(median_input)
    bad[0].data = bad[0].data[:100]
    with pytest.raises(ValueError) as exc:
        median_slitlets_rectified(bad)
    assert str(exc.value) == META["median"]["error_naxis2"]
    bad = copy_img(median_input)
    bad[0].header["instrume"] = "OSIRIS"
    with pytest.raises(ValueError) as exc:
        median_slitlets_rectified(bad)
    assert str(exc.value) == META["median"]["error_instrume"]


# ---------------------------------------------------------------- ABBA bookkeeping
@pytest.mark.parametrize("key", sorted(k for k in META["abba"] if k != "error"))
def test_get_isky_matches_reference(key):
    from pyemir_b200.abba import build_full_set, get_isky
    basic_pattern, repeat = key.split("_")
    want = META["abba"][key]
    assert [get_isky(i, basic_pattern, int(repeat)) for i in range(len(want))] == want
    # the sky image of an A is a B and vice versa
    full_set = build_full_set(basic_pattern, int(repeat), 3)
    assert len(full_set) == len(want)
    assert all(full_set[i] != full_set[j] for i, j in enumerate(want))


def test_get_isky_error_matches_reference():
    from pyemir_b200.abba import build_full_set, get_isky
    with pytest.raises(ValueError) as exc:
        get_isky(0, "ABAB", 1)
    assert str(exc.value) == META["abba"]["error"]
    with pytest.raises(ValueError):
        build_full_set("ABAB", 1, 1)
    assert build_full_set("ABBA", 2, 2) == "AABBBBAA" * 2

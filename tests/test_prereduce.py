"""Pre-reduction in front of the path: flat field pinned by vectors from
the reference's own FlatFieldCorrector, bias / dark restated."""

import hashlib
import json
import os

import numpy as np
import pytest

from oracle import pipeline as oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "flatfield.json")) as _fd:
    GOLDEN = json.load(_fd)

import sys
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_flatfield import make_inputs, sha_canonical  # noqa: E402  (no reference read)

RESULTS = np.load(os.path.join(HERE, "golden", "flatfield_results.npz"))


def _sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


@pytest.mark.parametrize("name", sorted(GOLDEN["cases"]))
def test_oracle_flatfield_matches_reference_vectors(name):
    case = GOLDEN["cases"][name]
    data, flat = make_inputs(name, case["img_dtype"], case["flat_dtype"])
    got = oracle.flatfield_correct(data, flat)
    assert got.dtype.name == case["result_dtype"]
    assert np.array_equal(got, RESULTS[name], equal_nan=True)
    assert sha_canonical(got) == case["sha256"] == sha_canonical(RESULTS[name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN["cases"]))
def test_flatfield_corrector_mirror_bit_exact(name):
    """The CUDA stage behind the reference's class interface reproduces the reference's
    output bit for bit, fixes the flat in place like it and writes the same cards."""
    from pyemir_b200 import hdu
    from pyemir_b200.prereduce import FlatFieldCorrector
    case = GOLDEN["cases"][name]
    data, flat = make_inputs(name, case["img_dtype"], case["flat_dtype"])
    node = FlatFieldCorrector(flat, calibid="uuid:golden-flat")
    assert _sha(flat) == case["flat_fixed_sha256"]
    assert (np.isnan(node.flat_stats) and np.isnan(case["flat_stats"])) or \
        float(node.flat_stats) == case["flat_stats"]
    img = hdu.HDUList([hdu.PrimaryHDU(data.copy(), hdu.Header())])
    res = node.run(img)
    arr = res["primary"].data
    assert arr.dtype.name == case["result_dtype"]
    # bit exact; NaNs compare as NaNs (CUDA returns its canonical NaN, numpy keeps the payload)
    assert np.array_equal(arr, RESULTS[name], equal_nan=True), \
        int((~((arr == RESULTS[name]) | (np.isnan(arr) & np.isnan(RESULTS[name])))).sum())
    assert np.array_equal(np.signbit(arr[arr == 0]), np.signbit(RESULTS[name][RESULTS[name] == 0]))
    assert sha_canonical(arr) == case["sha256"]
    cards = [[c[0], c[1]] for c in res["primary"].header.cards
             if "correction time" not in str(c[1])]
    assert cards == case["cards"]
    assert any("correction time" in str(c[1]) for c in res["primary"].header.cards)


@pytest.mark.gpu
@pytest.mark.parametrize("img_dtype,cal_dtype", [("float32", "float32"), ("uint16", "float32"),
                                                 ("float32", "float64"), ("float64", "float32"),
                                                 ("uint16", "float64")])
@pytest.mark.parametrize("shape", [(5, 64, 200), (6, 21, 33)])
def test_prereduce_stack_bit_exact(img_dtype, cal_dtype, shape):
    """(5, 64, 200): 16-byte loads, a batch of four frames and a tail; (6, 21, 33): odd size,
    scalar path."""
    import torch
    from pyemir_b200.prereduce import prereduce_torch
    rng = np.random.default_rng(3)
    if img_dtype == "uint16":
        frames = rng.integers(0, 65535, size=shape).astype(np.uint16)
    else:
        frames = rng.normal(2000.0, 500.0, size=shape).astype(img_dtype)
    bias = rng.normal(100.0, 3.0, size=shape[1:]).astype(cal_dtype)
    dark = rng.normal(5.0, 1.0, size=shape[1:]).astype(cal_dtype)
    flat = rng.normal(1.0, 0.3, size=shape[1:]).astype(cal_dtype)
    flat[3, :20] = 0.0
    flat[4, :20] = -1.0
    d_frames = torch.from_numpy(frames).cuda()
    for use in [(1, 1, 1), (1, 0, 0), (0, 1, 1), (0, 0, 1), (0, 0, 0)]:
        kw = {k: v for k, v, u in zip(("bias", "dark", "flat"), (bias, dark, flat), use) if u}
        got = prereduce_torch(d_frames, **kw).cpu().numpy()
        want = oracle.prereduce(frames, **kw)
        assert got.dtype == np.float32 and np.array_equal(got, want, equal_nan=True), use
    if img_dtype == "float32":                                   # in place
        work = d_frames.clone()
        out = prereduce_torch(work, bias=bias, dark=dark, flat=flat, out=work)
        assert out.data_ptr() == work.data_ptr()
        assert np.array_equal(out.cpu().numpy(), oracle.prereduce(frames, bias, dark, flat))
    with pytest.raises(ValueError):
        prereduce_torch(d_frames, flat=flat[:10])


@pytest.mark.gpu
def test_prereduce_then_rectify(coeff_j, frame_j):
    """flow(newimg) -> apply_rectwv_coeff, device resident, against the oracle chain."""
    import torch
    from oracle.pipeline import apply_rectwv_coeff as oracle_apply
    from pyemir_b200 import synthetic
    from pyemir_b200.plan import get_plan
    from pyemir_b200.prereduce import prereduce_torch
    rng = np.random.default_rng(8)
    flat = rng.normal(1.0, 0.05, size=frame_j.shape).astype(np.float32)
    bias = rng.normal(20.0, 1.0, size=frame_j.shape).astype(np.float32)
    raw = (frame_j * flat + bias).astype(np.float32)
    corrected = prereduce_torch(torch.from_numpy(raw).cuda(), bias=bias, flat=flat)
    plan = get_plan(coeff_j, 2, "float32")
    out, _ = plan.run_torch(corrected)
    want_in = oracle.prereduce(raw, bias=bias, flat=flat)
    assert np.array_equal(corrected.cpu().numpy(), want_in)
    hdul = synthetic.make_hdulist(want_in, coeff_j)
    _, ref64 = oracle_apply(hdul, coeff_j, only_needed_rows=True, return_float64=True)
    got = out[0].cpu().numpy()
    scale = np.abs(ref64).max()
    assert np.all(np.abs(got - ref64) <= 1e-5 * np.abs(ref64) + 2e-7 * scale)
    assert np.array_equal(got == 0, ref64 == 0)

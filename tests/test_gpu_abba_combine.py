"""The rest of the ABBA product on the GPU (recipes/spec/abba.py:545-610), through the C ABI:
sigma-clip combination, the per-frame vertical offsets of a device-resident stack in one
launch, the whole `combine_abba`, and the bad-pixel stage of the pre-reduction
(core/correctors.py:23-42).  Everything bit exact against the oracle restatements
(numina arithmetic: parity unpinned, see oracle/pipeline.py).
"""

import numpy as np
import pytest

from oracle import numina_restated as nr
from oracle import pipeline as oracle

pytestmark = pytest.mark.gpu


def _stack(rng, n, shape, dtype):
    x = rng.normal(500.0, 20.0, size=(n,) + shape).astype(dtype)
    # cosmic rays and dead pixels
    hits = rng.integers(0, [n, shape[0], shape[1]], size=(300, 3))
    x[hits[:, 0], hits[:, 1], hits[:, 2]] += 3000.0
    hits = rng.integers(0, [n, shape[0], shape[1]], size=(100, 3))
    x[hits[:, 0], hits[:, 1], hits[:, 2]] = 0.0
    return x


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("repeat,nseq", [(1, 1), (1, 3), (2, 3), (3, 5), (8, 2), (16, 2)])
def test_abba_sigmaclip_matches_oracle(dtype, repeat, nseq):
    """2 .. 64 images per nodding position (every register-array size of the kernel)."""
    import torch
    from pyemir_b200 import abba
    rng = np.random.default_rng(10 * repeat + nseq)
    full_set = abba.build_full_set("ABBA", repeat, nseq)
    n = len(full_set)
    stack = _stack(rng, n, (33, 515), dtype)
    stack[:, 5, :40] = 7.0                       # std = 0
    stack[0, 6, 7] = np.nan
    stack[n - 1, 6, 8] = np.inf
    frames = torch.from_numpy(stack).cuda()
    index_a = [i for i, c in enumerate(full_set) if c == "A"]
    for low, high in ((3.0, 3.0), (1.5, 2.5)):
        one = abba.stack_sigmaclip(frames, index_a, low, high)
        torch.cuda.synchronize()
        want = oracle.sigmaclip_stack(stack[index_a], low, high)
        assert one.dtype == torch.float64
        assert np.array_equal(one.cpu().numpy(), want, equal_nan=True)
        got = abba.combine_abba(frames, full_set, None, "sigmaclip", {"low": low, "high": high})
        torch.cuda.synchronize()
        with np.errstate(invalid="ignore"):
            want = oracle.combine_abba(stack, full_set, None, "sigmaclip", low, high)
        assert got.dtype == torch.float32
        assert np.array_equal(got.cpu().numpy(), want, equal_nan=True)
    # the clipping does something: the plain mean differs where a cosmic ray sits
    mean = abba.combine_abba(frames, full_set, None, "mean").cpu().numpy()
    if n >= 12:
        assert (mean != want).sum() > 50


def test_sigmaclip_errors():
    import torch
    from pyemir_b200 import abba
    from pyemir_b200._cabi import EmirwvError
    frames = torch.zeros((70, 4, 8), dtype=torch.float32, device="cuda")
    with pytest.raises(EmirwvError):
        abba.stack_sigmaclip(frames, list(range(65)))
    with pytest.raises(ValueError):
        abba.stack_sigmaclip(frames, [1, 2], low=-1.0)
    with pytest.raises(ValueError):
        abba.combine_abba(frames[:4].contiguous(), "ABBA", None, "sum")
    with pytest.raises(ValueError):
        abba.combine_abba(frames[:4].contiguous(), "ABB")


@pytest.mark.parametrize("in_dtype,out_dtype", [(np.float32, np.float32), (np.float64, np.float64),
                                                (np.float64, np.float32), (np.float32, np.float64)])
def test_shift_frames_bit_exact(in_dtype, out_dtype):
    """One launch for a stack with a different offset per frame (zero, integer, fractional,
    beyond the image) against the oracle's shift_image2d frame by frame."""
    import torch
    from pyemir_b200 import abba
    rng = np.random.default_rng(8)
    offsets = [0.0, -1.3, 0.3712, 2.0, -0.5, 17.0001, -3.9999, 1e-4, 80.0]
    stack = rng.normal(1000.0, 300.0, size=(len(offsets), 61, 259)).astype(in_dtype)
    tdt = {np.float32: torch.float32, np.float64: torch.float64}
    got = abba.shift_frames(torch.from_numpy(stack).cuda(), offsets, out_dtype=tdt[out_dtype])
    torch.cuda.synchronize()
    got = got.cpu().numpy()
    assert got.dtype == out_dtype
    for f, off in enumerate(offsets):
        want = stack[f] if off == 0 else nr.shift_image2d(stack[f], yoffset=off)
        assert np.array_equal(got[f], want.astype(out_dtype)), (f, off)
    assert not got[8].any()                      # shifted out of the image
    assert np.array_equal(got[3][2:], stack[3][:-2].astype(out_dtype))


def test_shift_frames_full_size_product_corner():
    """A full-size product (2090 x 3400): the far rows and columns, where the products of the
    shoelace sum round differently."""
    import torch
    from pyemir_b200 import abba
    rng = np.random.default_rng(9)
    img = np.zeros((1, 2090, 3400), dtype=np.float32)
    img[0, 1990:, 3290:] = rng.normal(1000.0, 300.0, size=(100, 110))
    got = abba.shift_frames(torch.from_numpy(img).cuda(), [0.4321], out_dtype=torch.float64)
    want = nr.rectify2d(img[0], [-0.0, 1.0, 0.0], [-0.4321, 0.0, 1.0], 2, rows=range(1985, 2090))
    assert np.array_equal(got.cpu().numpy()[0, 1985:], want[1985:])


@pytest.mark.parametrize("method", ["mean", "median", "sigmaclip"])
def test_combine_abba_with_offsets_matches_oracle(method):
    """abba.py:545-610 in one call: offsets, combination, float32(A) - float32(B)."""
    import torch
    from pyemir_b200 import abba
    rng = np.random.default_rng(21)
    full_set = abba.build_full_set("ABBA", 1, 3)
    n = len(full_set)
    stack = _stack(rng, n, (47, 321), np.float32)
    offsets = [0.0, 0.7312, -0.7312, 0.0, 1.25, 0.0, 0.0, -2.0, 0.0005, 0.0, 0.0, 3.3333]
    got = abba.combine_abba(torch.from_numpy(stack).cuda(), full_set, offsets, method)
    torch.cuda.synchronize()
    want = oracle.combine_abba(stack, full_set, offsets, method)
    assert got.dtype == torch.float32
    assert np.array_equal(got.cpu().numpy(), want, equal_nan=True)


@pytest.mark.parametrize("img_dtype,cal_dtype", [("float32", "float32"), ("uint16", "float32"),
                                                 ("float64", "float64"), ("uint16", "float64")])
@pytest.mark.parametrize("shape", [(3, 64, 200), (2, 21, 33)])
def test_prereduce_with_bad_pixel_mask_bit_exact(img_dtype, cal_dtype, shape):
    import torch
    from pyemir_b200.prereduce import prereduce_torch
    rng = np.random.default_rng(4)
    if img_dtype == "uint16":
        frames = rng.integers(0, 65535, size=shape).astype(np.uint16)
    else:
        frames = rng.normal(2000.0, 500.0, size=shape).astype(img_dtype)
        frames[0, 10, 11] = np.nan
    mask = (rng.random(shape[1:]) < 0.03).astype(np.uint8)
    mask[0, 0] = mask[0, 1] = mask[-1, -1] = 1
    mask[10, 10] = 7
    mask[12:19, 20:27] = 1                       # the centre of the block has no good neighbour
    bias = rng.normal(100.0, 3.0, size=shape[1:]).astype(cal_dtype)
    dark = rng.normal(5.0, 1.0, size=shape[1:]).astype(cal_dtype)
    flat = rng.normal(1.0, 0.1, size=shape[1:]).astype(cal_dtype)
    d = torch.from_numpy(frames).cuda()
    for kw in ({}, {"bias": bias, "dark": dark, "flat": flat}, {"flat": flat}):
        got = prereduce_torch(d, bpm=mask, **kw)
        torch.cuda.synchronize()
        want = oracle.prereduce_bpm(frames, mask, **kw)
        assert np.array_equal(got.cpu().numpy(), want, equal_nan=True)
    # without a mask the entry is the old one
    got = prereduce_torch(d, bias=bias)
    assert np.array_equal(got.cpu().numpy(), oracle.prereduce(frames, bias=bias), equal_nan=True)
    if img_dtype == "float32":
        with pytest.raises(ValueError):
            prereduce_torch(d, bpm=mask, out=d)


def test_bad_pixel_corrector_node_full_frame():
    """The node on a full 2048 x 2048 raw frame with 1 % of bad pixels."""
    from pyemir_b200.hdu import HDUList, PrimaryHDU
    from pyemir_b200.prereduce import BadPixelCorrector
    rng = np.random.default_rng(12)
    raw = rng.integers(100, 40000, size=(2048, 2048)).astype(np.uint16)
    mask = (rng.random((2048, 2048)) < 0.01).astype(np.int32)
    img = HDUList([PrimaryHDU(raw.copy())])
    out = BadPixelCorrector(mask, calibid="ID-BPM")(img)
    want = oracle.process_bpm_median(raw, mask).astype(np.float32)
    assert out[0].data.dtype == np.float32
    assert np.array_equal(out[0].data, want)
    assert out[0].header["NUM-BPM"] == "ID-BPM"
    good = mask == 0
    assert np.array_equal(out[0].data[good], raw[good].astype(np.float32))


# ------------------------------------------------------------------ host-buffer entries
def _raw_case(coeff_j, frame_j, n):
    rng = np.random.default_rng(31)
    scale = rng.uniform(0.5, 1.5, size=n)
    raw = np.clip(frame_j[None] * scale[:, None, None] * 8.0 + 1000.0, 0, 65535).astype(np.uint16)
    bias = rng.normal(1000.0, 2.0, size=frame_j.shape).astype(np.float32)
    dark = rng.normal(2.0, 0.5, size=frame_j.shape).astype(np.float32)
    flat = rng.normal(8.0, 0.2, size=frame_j.shape).astype(np.float32)
    bpm = (rng.random(frame_j.shape) < 0.004).astype(np.uint8)
    return raw, bias, dark, flat, bpm


def test_run_host_raw_equals_flow_then_path(coeff_j, frame_j):
    """emirwv_run_host_raw (uint16 frames + the plan's calibration) = prereduce kernel + path,
    bit for bit, dense and sparse copy-back, pageable and page-locked output."""
    import torch
    from pyemir_b200 import _pinned
    from pyemir_b200.plan import RectWavePlan
    from pyemir_b200.prereduce import prereduce_torch
    n = 11                                       # three pipeline slots, a short last chunk
    raw, bias, dark, flat, bpm = _raw_case(coeff_j, frame_j, n)
    plan = RectWavePlan(coeff_j, 2, "float32")
    plan.set_calibration(bpm=bpm, bias=bias, dark=dark, flat=flat)
    red = prereduce_torch(torch.from_numpy(raw).cuda(), bias, dark, flat, bpm=bpm)
    want, want_jm = plan.run_torch(red)
    want, want_jm = want.cpu().numpy(), want_jm.cpu().numpy()
    got, jm = plan.run_host_raw(raw)
    assert np.array_equal(got, want) and np.array_equal(jm, want_jm)
    # sparse copy-back: page-locked buffer (one kernel per chunk writes the covered columns
    # straight into it) and pageable buffer (strided copies)
    for out in (_pinned.empty(want.shape, np.float32), np.empty(want.shape, np.float32)):
        out[...] = 0
        got, jm = plan.run_host_raw(raw, out=out, out_zeroed=True)
        assert np.array_equal(got, want) and np.array_equal(jm, want_jm)
    # float32 "raw" frames, no bad-pixel mask
    plan.set_calibration(flat=flat)
    rawf = raw[:3].astype(np.float32)
    got, _ = plan.run_host_raw(rawf)
    red = prereduce_torch(torch.from_numpy(rawf).cuda(), flat=flat)
    assert np.array_equal(got, plan.run_torch(red)[0].cpu().numpy())
    with pytest.raises(ValueError):
        RectWavePlan(coeff_j, 2, "float64").run_host_raw(raw[:1])
    plan.close()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_run_host_batch_into_pinned_buffer_zeros_written_on_the_host(coeff_j, frame_j, dtype,
                                                                     monkeypatch):
    """A batch into a page-locked buffer: the covered columns travel, host threads write the
    zeros; the buffer starts out full of NaNs and must come back identical to the plain copy
    (pageable buffer) and to the all-bytes-over-PCIe variant."""
    from pyemir_b200 import _pinned
    from pyemir_b200.plan import RectWavePlan
    n = 9
    rng = np.random.default_rng(5)
    frames = (frame_j[None] * rng.uniform(0.5, 1.5, size=(n, 1, 1))).astype(dtype)
    plan = RectWavePlan(coeff_j, 2, dtype)
    want, want_jm = plan.run_host(frames)                    # pageable: plain copies
    out = _pinned.empty(want.shape, want.dtype)
    for threads in ("3", "1", "0"):
        monkeypatch.setenv("EMIRWV_HOST_ZERO_THREADS", threads)
        out[...] = np.nan
        got, jm = plan.run_host(frames, out=out)
        assert np.array_equal(got, want) and np.array_equal(jm, want_jm), threads
    plan.close()


@pytest.mark.parametrize("method", ["mean", "median", "sigmaclip"])
@pytest.mark.parametrize("shifted", [False, True])
def test_abba_host_equals_device_pipeline(coeff_j, frame_j, method, shifted):
    """emirwv_abba_host: frames in, one combined image out = path + shift + combination run
    kernel by kernel on device tensors, bit for bit (the sums of the mean run in frame order
    although three pipeline slots work concurrently)."""
    import torch
    from pyemir_b200 import abba
    from pyemir_b200.plan import RectWavePlan
    from pyemir_b200.prereduce import prereduce_torch
    n = 12
    raw, bias, dark, flat, bpm = _raw_case(coeff_j, frame_j, n)
    full_set = abba.build_full_set("ABBA", 1, 3)
    labels = abba.labels_of(full_set)
    offsets = [0.0, 0.7312, -0.7312, 0.0, 1.25, 0.0, 0.0, -2.0, 0.0005, 0.0, 0.0, 3.3333] \
        if shifted else None
    plan = RectWavePlan(coeff_j, 2, "float32")
    plan.set_calibration(bpm=bpm, bias=bias, dark=dark, flat=flat)
    red = prereduce_torch(torch.from_numpy(raw).cuda(), bias, dark, flat, bpm=bpm)
    prod, _ = plan.run_torch(red)
    want = abba.combine_abba(prod, full_set, offsets, method).cpu().numpy()
    got = plan.abba_host(raw, labels, offsets, method, raw=True)
    assert got.dtype == np.float32 and got.shape == plan.out_shape
    assert np.array_equal(got, want, equal_nan=True)
    # reduced float32 frames in (no flow), an unused frame in the sequence
    lab2 = list(labels)
    lab2[5] = 0
    got = plan.abba_host(red.cpu().numpy(), lab2, offsets, method)
    fs2 = [i for i in range(n) if i != 5]
    want = abba.combine_abba(prod[fs2].contiguous(), "".join(full_set[i] for i in fs2),
                             None if offsets is None else [offsets[i] for i in fs2], method)
    assert np.array_equal(got, want.cpu().numpy(), equal_nan=True)
    with pytest.raises(ValueError):
        plan.abba_host(raw, [1] * n, None, method, raw=True)       # no B image
    plan.close()


def test_abba_host_mean_against_oracle_small(coeff_j, frame_j):
    """The whole chain against the CPU oracle for three slitlets: rectification (float32
    tolerance), shift + mean combination on the oracle's own products (bit exact)."""
    import copy
    from oracle.pipeline import apply_rectwv_coeff as oracle_apply
    from pyemir_b200 import abba, synthetic
    from pyemir_b200.plan import RectWavePlan
    coeff = copy.deepcopy(coeff_j)
    coeff.missing_slitlets = [i for i in range(1, 56) if i not in (3, 27, 50)]
    frames = np.stack([frame_j, frame_j * 0.5, frame_j * 0.25, frame_j * 2.0]).astype(np.float32)
    offsets = [0.0, 0.4, -0.4, 0.0]
    plan = RectWavePlan(coeff, 2, "float32")
    got = plan.abba_host(frames, abba.labels_of("ABBA"), offsets, "mean")
    prods = [oracle_apply(synthetic.make_hdulist(f, coeff), coeff, only_needed_rows=True)[0].data
             for f in frames]
    want = oracle.combine_abba(np.stack(prods), "ABBA", offsets, "mean")
    scale = np.abs(want).max()
    assert np.abs(got - want).max() < 2e-5 * scale
    assert np.array_equal(got == 0, want == 0) or np.abs(got - want).max() < 2e-5 * scale
    plan.close()

"""Stand-ins for compute-sanitizer (closed on this pool, profiles/r02_sanitizer_unavailable.txt):
guard bands around every buffer the kernels write, poison around the frames they read, repeats
and concurrent streams of the same launch compared bit for bit."""

import numpy as np
import pytest

from pyemir_b200 import synthetic

pytestmark = pytest.mark.gpu

GUARD = 1 << 16          # elements on either side


def _guarded(shape, dtype, fill):
    import torch
    n = int(np.prod(shape))
    buf = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device="cuda")
    return buf, buf[GUARD:GUARD + n].view(shape)


def _bands_intact(buf, fill):
    import torch
    lo, hi = buf[:GUARD], buf[-GUARD:]
    if isinstance(fill, float) and fill != fill:
        return bool(torch.isnan(lo).all() and torch.isnan(hi).all())
    return bool((lo == fill).all() and (hi == fill).all())


@pytest.mark.parametrize("number,dtype,mode", [(2, "float32", 2), (2, "float64", 2), (2, "float32", 1),
                                               (4, "float32", 2), ("4b", "float64", 2)])
def test_hot_kernel_stays_inside_its_buffers(number, dtype, mode):
    """Frames surrounded by NaN poison (a tap with weight that strays outside the frames would
    poison the product), products and JMN/JMX buffers surrounded by guard bands; 7 frames: a
    full frame group and a short one; first and last slitlets touch the detector edges."""
    import torch
    from pyemir_b200.plan import RectWavePlan
    coeff = synthetic.make_config(number)
    plan = RectWavePlan(coeff, mode, dtype)
    tdt = torch.float32 if dtype == "float32" else torch.float64
    frames = torch.from_numpy(synthetic.make_frames(coeff, 7, seed=5).astype(plan.np_dtype)).cuda()
    ref_out, ref_jm = plan.run_torch(frames)
    torch.cuda.synchronize()
    in_buf, d_in = _guarded(frames.shape, tdt, float("nan"))
    d_in.copy_(frames)
    out_buf, d_out = _guarded(ref_out.shape, tdt, float("nan"))
    jm_buf, d_jm = _guarded(ref_jm.shape, torch.int32, -77)
    plan.run_torch(d_in, d_out, d_jm)
    torch.cuda.synchronize()
    assert torch.equal(d_out, ref_out) and torch.equal(d_jm, ref_jm)
    assert not torch.isnan(d_out).any()
    assert _bands_intact(out_buf, float("nan")) and _bands_intact(jm_buf, -77)
    assert _bands_intact(in_buf, float("nan"))
    plan.close()


def test_repeats_and_concurrent_streams_are_bit_identical():
    """Twenty launches of the same batch, alternating between two streams that run
    concurrently into separate buffers: every product equals the first one."""
    import torch
    from pyemir_b200.plan import RectWavePlan
    coeff = synthetic.make_config(2)
    plan = RectWavePlan(coeff, 2, "float32")
    frames = torch.from_numpy(synthetic.make_frames(coeff, 16, seed=6)).cuda()
    ref_out, ref_jm = plan.run_torch(frames)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [(torch.empty_like(ref_out), torch.empty_like(ref_jm)) for _ in streams]
    for rep in range(10):
        for st, (o, j) in zip(streams, outs):
            o.fill_(float("nan"))
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                plan.run_torch(frames, o, j)
        torch.cuda.synchronize()
        for o, j in outs:
            assert torch.equal(o, ref_out) and torch.equal(j, ref_jm), rep
    plan.close()


def test_side_kernels_stay_inside_their_buffers():
    """Guard bands around the outputs of the kernels either side of the path."""
    import torch
    from pyemir_b200 import abba
    from pyemir_b200.prereduce import prereduce_torch
    rng = np.random.default_rng(3)
    n, h, w = 12, 45, 333                         # odd sizes: scalar paths and tails
    stack = torch.from_numpy(rng.normal(500, 20, size=(n, h, w)).astype(np.float32)).cuda()
    full_set = abba.build_full_set("ABBA", 1, 3)
    offsets = [0.0, 0.7, -0.7, 0.0, 1.25, 0.0, 0.0, -44.0, 0.0005, 0.0, 0.0, 60.0]
    sh_buf, sh = _guarded(stack.shape, torch.float32, float("nan"))
    abba.shift_frames(stack, offsets, out=sh)
    torch.cuda.synchronize()
    assert _bands_intact(sh_buf, float("nan")) and not torch.isnan(sh).any()
    for method in ("mean", "median", "sigmaclip"):
        assert not torch.isnan(abba.combine_abba(stack, full_set, offsets, method)).any()
    raw = torch.from_numpy(rng.integers(0, 60000, size=(5, 63, 201)).astype(np.uint16)).cuda()
    mask = (rng.random((63, 201)) < 0.05).astype(np.uint8)
    mask[0, :] = 1
    mask[:, -1] = 1
    pr_buf, pr = _guarded(raw.shape, torch.float32, float("nan"))
    prereduce_torch(raw, flat=np.full((63, 201), 2.0, np.float32), bpm=mask, out=pr)
    torch.cuda.synchronize()
    assert _bands_intact(pr_buf, float("nan")) and not torch.isnan(pr).any()

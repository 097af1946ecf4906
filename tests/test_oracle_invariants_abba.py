"""numina-independent properties of the restatements behind the ABBA product and the bad-pixel
stage (oracle/pipeline.py: sigmaclip_stack, combine_abba, process_bpm_median; numina_restated:
shift_image2d).  They are restatements of third-party code that cannot be imported here (PARITY
UNPINNED), so analytic properties are what vouches for them."""

import numpy as np
import pytest

from oracle import numina_restated as nr
from oracle import pipeline as oracle


def test_sigmaclip_without_outliers_is_the_mean(rng):
    stack = rng.normal(100.0, 1.0, size=(9, 7, 11))
    # 9 samples: the largest deviation is at most (n-1)/sqrt(n) = 2.67 std: nothing is clipped
    assert np.array_equal(oracle.sigmaclip_stack(stack, 3.0, 3.0),
                          stack.sum(axis=0) / 9.0)


def test_sigmaclip_drops_a_cosmic_ray_and_nothing_else(rng):
    stack = rng.normal(100.0, 1.0, size=(20, 5, 5))
    clean = stack.copy()
    stack[7, 2, 3] += 5000.0
    got = oracle.sigmaclip_stack(stack, 3.0, 3.0)
    keep = [i for i in range(20) if i != 7]
    assert got[2, 3] == pytest.approx(clean[keep, 2, 3].mean(), abs=1e-12)
    untouched = np.ones((5, 5), bool)
    untouched[2, 3] = False
    assert np.array_equal(got[untouched], oracle.sigmaclip_stack(clean, 3.0, 3.0)[untouched])


def test_sigmaclip_is_shift_and_scale_equivariant_and_idempotent(rng):
    stack = rng.normal(0.0, 1.0, size=(12, 40))
    stack[3, :10] += 50.0
    base = oracle.sigmaclip_stack(stack, 2.5, 2.5)
    # exact powers of two keep every rounding in place
    assert np.array_equal(oracle.sigmaclip_stack(stack * 4.0, 2.5, 2.5), base * 4.0)
    shifted = oracle.sigmaclip_stack(stack + 1024.0, 2.5, 2.5)
    assert np.allclose(shifted, base + 1024.0, rtol=0, atol=1e-9)
    # a stack that was clipped once already: clipping its survivors again changes nothing
    again = oracle.sigmaclip_stack(np.concatenate([stack[:3], stack[4:]]), 2.5, 2.5)
    assert np.allclose(again[:10], base[:10], rtol=0, atol=1e-12)


def test_sigmaclip_edge_cases():
    one = np.array([[3.0, -2.0]])
    assert np.array_equal(oracle.sigmaclip_stack(one), one[0])              # one sample
    same = np.full((5, 3), 7.0)
    assert np.array_equal(oracle.sigmaclip_stack(same), same[0])            # std = 0
    with np.errstate(invalid="ignore"):
        nan = oracle.sigmaclip_stack(np.array([[1.0], [np.nan], [2.0]]))
    assert np.isnan(nan[0])
    # limits below one standard deviation can throw out everything: the last mean stays
    two = np.array([[0.0], [10.0]])
    assert oracle.sigmaclip_stack(two, 0.1, 0.1)[0] == 5.0


def test_shift_image2d_properties(rng):
    img = rng.normal(100.0, 10.0, size=(40, 23))
    # integer shifts move pixels exactly and compose
    a = nr.shift_image2d(nr.shift_image2d(img, yoffset=3), yoffset=-1)
    b = nr.shift_image2d(img, yoffset=2)
    assert np.array_equal(a[3:-1], b[3:-1])      # (the last row left the image on the way)
    assert np.array_equal(b[2:], img[:-2]) and not b[:2].any()
    # a fractional shift conserves the flux of every column away from the edges and mixes
    # exactly two rows with weights that sum to one
    c = nr.shift_image2d(img, yoffset=0.3)
    assert np.allclose(c.sum(axis=0) + 0.3 * img[-1], img.sum(axis=0), rtol=1e-12)
    assert np.allclose(c[5], 0.3 * img[4] + 0.7 * img[5], rtol=1e-12)
    # opposite shifts are not an identity (two interpolations) but conserve flux inside
    d = nr.shift_image2d(c, yoffset=-0.3)
    assert abs(d[3:-3].sum() - img[3:-3].sum()) < 1e-2 * abs(img[3:-3].sum())


@pytest.mark.parametrize("method", ["mean", "median", "sigmaclip"])
def test_combine_abba_properties(rng, method):
    # (multiples of 1/64: every sum and difference below is exact in float32)
    base = (np.round(rng.normal(100.0, 5.0, size=(30, 17)) * 64.0) / 64.0).astype(np.float32)
    # identical A and B images: the difference vanishes whatever the method
    frames = np.stack([base] * 8)
    assert not oracle.combine_abba(frames, "ABBAABBA", None, method).any()
    # A = base + signal, B = base: the signal comes back exactly (float32 values, exact sums)
    signal = np.zeros_like(base)
    signal[10:14, 3:9] = 64.0
    frames = np.stack([base + signal, base, base, base + signal] * 2)
    assert np.array_equal(oracle.combine_abba(frames, "ABBAABBA", None, method), signal)
    # zero offsets are no offsets; an integer offset moves the signal by whole rows
    zero = oracle.combine_abba(frames, "ABBAABBA", [0.0] * 8, method)
    assert np.array_equal(zero, signal)
    moved = oracle.combine_abba(frames, "ABBAABBA", [2.0] * 8, method)
    assert np.array_equal(moved[:-2], signal[2:])           # yoffset = -offset: rows move up
    with pytest.raises(ValueError):
        oracle.combine_abba(frames, "ABBAABBX", None, method)
    with pytest.raises(ValueError):
        oracle.combine_abba(frames, "ABBAABBA", None, "sum")


def test_process_bpm_median_properties(rng):
    img = rng.normal(1000.0, 30.0, size=(25, 31))
    none = np.zeros(img.shape, np.uint8)
    assert np.array_equal(oracle.process_bpm_median(img, none), img)        # nothing masked
    mask = none.copy()
    mask[12, 15] = 1
    out = oracle.process_bpm_median(img, mask)
    win = np.delete(img[10:15, 13:18].ravel(), 12)                          # 24 good neighbours
    assert out[12, 15] == np.median(win)
    assert np.array_equal(out[mask == 0], img[mask == 0])                   # good pixels kept
    # a constant image is a fixed point; corners use the clipped window
    flat = np.full((9, 9), 5.0)
    m = np.zeros((9, 9), np.uint8)
    m[0, 0] = m[8, 8] = m[4, 4] = 1
    assert np.array_equal(oracle.process_bpm_median(flat, m), flat)
    # a fully masked window gives the fill value, and masked neighbours do not contribute
    m = np.zeros((9, 9), np.uint8)
    m[2:7, 2:7] = 1
    out = oracle.process_bpm_median(flat * 0 + np.arange(81.0).reshape(9, 9), m, fill=-1.0)
    assert out[4, 4] == -1.0
    ramp = np.arange(81.0).reshape(9, 9)
    good = [ramp[r, c] for r in range(0, 5) for c in range(0, 5) if m[r, c] == 0]
    assert out[2, 2] == np.median(good)

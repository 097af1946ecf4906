"""GPU parity, part 2 (``-m gpu``): the kernel instantiations and limits that the BASELINE
products never reach (footprints K = 2 and K = 4, the EMIRWV_E_LIMIT exits), the full 55
slitlets of BASELINE configs 3, 4 and 4b in both dtypes and both resampling modes, and the
overlap areas of the plan against an integration that shares no code with the clippers
(tests/area_scanline.py).  Reference: slitlet2d.py:413-423, apply_rectwv_coeff.py:152-159.
"""

import numpy as np
import pytest

from oracle import numina_restated as nr
from oracle.pipeline import apply_rectwv_coeff as oracle_apply
from pyemir_b200 import synthetic
from pyemir_b200._cabi import EmirwvError
from pyemir_b200.apply_rectwv_coeff import rectify_frames
from pyemir_b200.plan import RectWavePlan

import area_scanline as scan
from parity_util import (assert_flux_close, assert_flux_close_f32, only_slitlets, scaled_map)

pytestmark = pytest.mark.gpu


def _check_product(frame, coeff, mode, max_order=4, what=""):
    """float64 and float32 CUDA products of one frame against the float64 oracle."""
    hdul = synthetic.make_hdulist(frame, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, args_resampling=mode, only_needed_rows=True,
                              return_float64=True, nmax_order=max_order)
    assert np.count_nonzero(ref64) > 0
    out64, kw64 = rectify_frames(frame.astype(np.float64), coeff, mode, dtype="float64",
                                 max_order=max_order)
    assert_flux_close(out64[0], ref64, 1e-12, 1e-12, what + " float64")
    assert np.array_equal(out64[0] == 0, ref64 == 0)
    out32, kw32 = rectify_frames(frame, coeff, mode, dtype="float32", max_order=max_order)
    assert_flux_close_f32(out32[0], ref64, what + " float32")
    assert np.array_equal(out32[0] == 0, ref64 == 0)
    hdr = ref[0].header
    for kw in (kw64, kw32):
        for islitlet, (imn, imx, jmn, jmx) in enumerate(kw[0], start=1):
            assert (imn, imx, jmn, jmx) == tuple(
                hdr[f"{k}{islitlet:02d}"] for k in ("IMNSLT", "IMXSLT", "JMNSLT", "JMXSLT")), islitlet


# ------------------------------------------------------------------ footprints K = 4 and K = 2
@pytest.mark.parametrize("su,sv,keep", [(2.1, 1.0, (5, 30, 51)), (1.0, 2.1, (28,))])
def test_footprint_4_matches_oracle(coeff_j, frame_j, su, sv, keep):
    """A map stretched by more than 2 along one axis: an output pixel covers four source
    columns (su) or four source rows (sv: the fourth ring-slot word of the records) --
    k_rectwv<*, 4, *>, which no BASELINE product instantiates."""
    coeff = scaled_map(only_slitlets(coeff_j, keep), keep, su=su, sv=sv)
    for dtype in ("float64", "float32"):
        plan = RectWavePlan(coeff, 2, dtype)
        assert plan.info()["footprint"] == 4, plan.info()
        plan.close()
    _check_product(frame_j, coeff, 2, what=f"K=4 su={su} sv={sv}")


def test_footprint_2_area_mode_matches_oracle(coeff_j, frame_j):
    """A map shrunk to one half: every footprint fits 2 x 2 source pixels (K = 2 in mode 2,
    records without the compact form)."""
    keep = (2, 29, 54)
    coeff = scaled_map(only_slitlets(coeff_j, keep), keep, su=0.5, sv=0.5)
    plan = RectWavePlan(coeff, 2, "float32")
    assert plan.info()["footprint"] == 2
    plan.close()
    _check_product(frame_j, coeff, 2, what="K=2")


def test_footprint_4_batch_tail_bit_identical(coeff_j):
    """K = 4 on a 7-frame batch: frame grouping inside a CTA does not change a bit."""
    keep = (5, 30)
    coeff = scaled_map(only_slitlets(coeff_j, keep), keep, su=2.1)
    frames = synthetic.make_frames(coeff_j, 7, seed=12)
    plan = RectWavePlan(coeff, 2, "float32")
    out, jmm = plan.run_host(frames)
    for i in (0, 3, 6):
        one, jmm1 = plan.run_host(frames[i:i + 1])
        assert np.array_equal(one[0], out[i]) and np.array_equal(jmm1[0], jmm[i])
    plan.close()


# ------------------------------------------------------------------ EMIRWV_E_LIMIT exits
def test_limit_footprint_too_large(coeff_j):
    coeff = scaled_map(only_slitlets(coeff_j, (30,)), (30,), su=4.5)
    with pytest.raises(EmirwvError, match="distortion too strong: an output pixel covers"):
        RectWavePlan(coeff, 2, "float32")
    # nearest neighbour has a footprint of one pixel whatever the stretch
    plan = RectWavePlan(coeff, 1, "float32")
    assert plan.info()["footprint"] == 1
    plan.close()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_limit_window_does_not_fit(coeff_j, dtype):
    coeff = scaled_map(only_slitlets(coeff_j, (30,)), (30,), shear=0.8)
    with pytest.raises(EmirwvError, match="distortion too strong: the source window"):
        RectWavePlan(coeff, 2, dtype)


# ------------------------------------------------------------------ full frames, configs 3 / 4 / 4b
@pytest.mark.parametrize("number,max_order", [(3, 5), (4, 4), ("4b", 4)])
@pytest.mark.parametrize("mode", [1, 2])
def test_all_55_slitlets_both_dtypes(number, max_order, mode):
    """Every slitlet of the product (the round-1 tests took 3-4 of them), float64 to 1e-12,
    float32 to 1e-5 (+ local floor), coverage mask and JMN/JMX bit-exact."""
    coeff = synthetic.make_config(number)
    frame = synthetic.make_frames(coeff, 1, seed=700 + mode)[0]
    _check_product(frame, coeff, mode, max_order, what=f"config {number} mode {mode}")


# ------------------------------------------------------------------ overlap areas, independently
@pytest.mark.parametrize("number,max_order,islitlet", [(1, 4, 33), (3, 5, 7), (3, 5, 49)])
def test_plan_weights_match_scanline_integration(number, max_order, islitlet):
    """emirwv_debug_geometry weights (what the kernels multiply with) against the chord
    integration of tests/area_scanline.py for a sample of pixels, and the sum of a pixel's
    weights against the shoelace area of its mapped quadrilateral for all interior pixels.
    Tolerance: the rounding of a shoelace sum in absolute bounding-box coordinates."""
    coeff = synthetic.make_config(number)
    plan = RectWavePlan(coeff, 2, "float64", max_order=max_order)
    entry = coeff.contents[islitlet - 1]
    order = nr.order_fmap(len(entry["ttd_aij"]), max_order)
    g_row, g_col, _, g_w = plan.debug_geometry(islitlet)
    K = g_w.shape[-1]
    bbh, bbw = plan.bbox[islitlet]
    rng = np.random.default_rng(5 * islitlet)
    y0 = entry["min_row_rectified"]
    eps = np.finfo(float).eps
    worst = 0.0
    for _ in range(300):
        y, j = int(rng.integers(0, 39)), int(rng.integers(0, bbw))
        xs = np.array([j - .5, j - .5, j + .5, j + .5])
        ys = np.array([y0 + y - .5, y0 + y + .5, y0 + y + .5, y0 + y - .5])
        qx, qy = nr.fmap(order, entry["ttd_aij"], entry["ttd_bij"], xs, ys)
        want = scan.quad_weights(qx, qy, int(g_row[y, j]), int(g_col[y, j]), K, K)
        # pixels outside the bounding box carry no weight (rectify2d clips to the image)
        for a in range(K):
            for b in range(K):
                r, c = g_row[y, j] + a, g_col[y, j] + b
                if not (0 <= r < bbh and 0 <= c < bbw):
                    want[a, b] = 0.0
        tol = 8 * eps * max(np.abs(qx).max(), 1.0) * max(np.abs(qy).max(), 1.0)
        worst = max(worst, float(np.abs(g_w[y, j] - want).max() / tol))
        assert np.allclose(g_w[y, j], want, rtol=0, atol=tol), (y, j, g_w[y, j], want)
    assert worst < 1.0
    # sum of weights = area of the mapped pixel, every interior pixel of the slitlet
    jj, yy = np.meshgrid(np.arange(20, bbw - 20, dtype=float), np.arange(3, 36, dtype=float) + y0)
    cx, cy = [], []
    for dj, di in ((-.5, -.5), (-.5, .5), (.5, .5), (.5, -.5)):
        u, v = nr.fmap(order, entry["ttd_aij"], entry["ttd_bij"], (jj + dj).ravel(), (yy + di).ravel())
        cx.append(u)
        cy.append(v)
    cx, cy = np.array(cx), np.array(cy)
    inside = (cy.min(0) > 1.0) & (cy.max(0) < bbh - 2.0) & (cx.min(0) > 1.0) & (cx.max(0) < bbw - 2.0)
    mx, my = cx.mean(0), cy.mean(0)
    px, py = cx - mx, cy - my                                   # shoelace about the centre
    area = 0.5 * np.abs(sum(px[k] * py[(k + 1) % 4] - px[(k + 1) % 4] * py[k] for k in range(4)))
    total = g_w[3:36, 20:bbw - 20].sum(axis=(-1, -2)).ravel()
    tol = 32 * eps * np.abs(cx).max() * np.abs(cy).max()
    assert inside.sum() > 10000
    assert np.all(np.abs(total - area)[inside] <= tol), float(np.abs(total - area)[inside].max())
    plan.close()

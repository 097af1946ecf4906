"""CUDA path vs oracle, through the C ABI (needs a B200: ``-m gpu``).

Tolerances (BASELINE.json north_star): bit-exact for index maps, interval indices,
slitlet masks / zero patterns and header integers; 1e-12 relative for float64
flux; 1e-5 relative for float32 flux.  "Relative" is applied per pixel with an
absolute floor (tests/parity_util.py): for float64 1e-12 x the frame's flux scale
(pixels at the edge of the wavelength coverage are differences of nearly equal
numbers in the oracle itself), for float32 1e-6 x the LOCAL flux (the brightest
pixel among the output pixel's neighbours); every float32 comparison records how
many pixels pass the bare 1e-5 relative bound (gpurun_out/parity_stats.jsonl).
"""

import copy

import numpy as np
import pytest

from oracle import numina_restated as nr
from oracle.pipeline import apply_rectwv_coeff as oracle_apply
from pyemir_b200 import synthetic
from pyemir_b200.apply_rectwv_coeff import (apply_rectwv_coeff, apply_rectwv_coeff_batch,
                                            rectify_frames)
from pyemir_b200.plan import RectWavePlan, clear_plan_cache, get_plan
from pyemir_b200.set_wv_parameters import set_wv_parameters

import proto_tables as proto
from parity_util import (assert_flux_close, assert_flux_close_f32, grid_of, header_cards,
                         only_slitlets)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle_j(coeff_j, frame_j):
    """Oracle output of BASELINE config 1 (one J+J frame), modes 1 and 2, float64."""
    hdul = synthetic.make_hdulist(frame_j, coeff_j, with_mecs=True)
    res = {}
    for mode in (1, 2):
        res[mode] = oracle_apply(hdul, coeff_j, args_resampling=mode,
                                 only_needed_rows=True, return_float64=True)
    return hdul, res


# ------------------------------------------------------------------ plan geometry
def test_nearest_index_maps_bit_exact_all_slitlets(coeff_j):
    plan = get_plan(coeff_j, 1, "float64")
    assert plan.info()["footprint"] == 1
    for islitlet in range(1, 56):
        entry = coeff_j.contents[islitlet - 1]
        bbh, bbw = plan.bbox[islitlet]
        rows = list(range(entry["min_row_rectified"], entry["max_row_rectified"] + 2))
        _, iv, ju, ok = nr.rectify2d(np.zeros((bbh, bbw)), entry["ttd_aij"], entry["ttd_bij"],
                                     1, rows=rows, return_maps=True)
        g_row, g_col, g_ok, _ = plan.debug_geometry(islitlet, weights=False)
        assert np.array_equal(g_row, iv), islitlet
        assert np.array_equal(g_col, ju), islitlet
        assert np.array_equal(g_ok, ok), islitlet


def test_nearest_index_maps_bit_exact_order5():
    coeff = synthetic.make_config(3)
    plan = RectWavePlan(coeff, 1, "float64", max_order=5)
    for islitlet in (1, 20, 55):
        entry = coeff.contents[islitlet - 1]
        bbh, bbw = plan.bbox[islitlet]
        rows = list(range(entry["min_row_rectified"], entry["max_row_rectified"] + 2))
        _, iv, ju, ok = nr.rectify2d(np.zeros((bbh, bbw)), entry["ttd_aij"], entry["ttd_bij"],
                                     1, nmax_order=5, rows=rows, return_maps=True)
        g_row, g_col, g_ok, _ = plan.debug_geometry(islitlet, weights=False)
        assert np.array_equal(g_row, iv) and np.array_equal(g_col, ju)
        assert np.array_equal(g_ok, ok)
    plan.close()


@pytest.mark.parametrize("mode", [2, 3])
def test_weight_tables_match_independent_statement(coeff_j, mode):
    plan = get_plan(coeff_j, mode, "float64")
    for islitlet in (1, 33):
        entry = coeff_j.contents[islitlet - 1]
        br, bc, w, K = proto.rectify_tables(entry, mode)
        g_row, g_col, _, g_w = plan.debug_geometry(islitlet)
        assert plan.info()["footprint"] >= K
        Kp = g_w.shape[-1]
        ns1, nc1 = entry["bb_ns1_orig"], entry["bb_nc1_orig"]
        assert np.array_equal(g_row + ns1 - 1, br) and np.array_equal(g_col + nc1 - 1, bc)
        # same clipping order and individually rounded operations as the oracle:
        # the float64 weight tables are bit-identical
        assert np.array_equal(g_w[..., :K, :K], w), np.abs(g_w[..., :K, :K] - w).max()
        if Kp > K:
            assert np.all(g_w[..., K:, :] == 0) and np.all(g_w[..., :, K:] == 0)
        # mode 2: weights of an interior pixel add up to the mapped pixel's area (≈ Jacobian)
        if mode == 2 and islitlet == 33:      # slitlet 1 is cut by the detector edge
            total = g_w.sum(axis=(-1, -2))[5:30, 100:1900]
            assert np.all((total > 0.8) & (total < 1.2))


def test_border_tables_bit_exact(coeff_j):
    plan = get_plan(coeff_j, 2, "float64")
    for islitlet in (1, 28, 55):
        tb = proto.rebin_tables(coeff_j.contents[islitlet - 1], **grid_of(coeff_j))
        idx, t = plan.debug_borders(islitlet)
        assert np.array_equal(idx, tb["idx"])
        assert np.array_equal(t, tb["t"])


def test_kernel1_alone_matches_oracle_rectify(coeff_j, frame_j):
    import torch
    plan = get_plan(coeff_j, 2, "float64")
    frame = torch.from_numpy(frame_j.astype(np.float64)).cuda()
    for islitlet in (2, 40):
        entry = coeff_j.contents[islitlet - 1]
        ns1, ns2 = entry["bb_ns1_orig"], entry["bb_ns2_orig"]
        rows = list(range(entry["min_row_rectified"], entry["max_row_rectified"] + 2))
        want = nr.rectify2d(frame_j[ns1 - 1:ns2].astype(float), entry["ttd_aij"],
                            entry["ttd_bij"], 2, rows=rows)[rows]
        got = plan.rectify_slitlet_torch(frame, islitlet).cpu().numpy()
        assert np.allclose(got, want, rtol=1e-13, atol=1e-13 * np.abs(want).max())


# ------------------------------------------------------------------ full path parity
@pytest.mark.parametrize("mode", [1, 2])
def test_float64_parity_full_frame(coeff_j, frame_j, oracle_j, mode):
    _, res = oracle_j
    ref_hdul, ref64 = res[mode]
    out, keywords = rectify_frames(frame_j.astype(np.float64), coeff_j, mode, dtype="float64")
    got = out[0]
    assert got.dtype == np.float64 and got.shape == (2090, 3400)
    # float64: 1e-12 relative; floor 1e-12 x frame scale (the oracle differences a
    # running sum ~1e4 x larger than a pixel, i.e. carries ~1e-12 relative noise itself)
    assert_flux_close(got, ref64, 1e-12, 1e-12, f"mode {mode}")
    assert np.array_equal(got == 0, ref64 == 0)           # coverage mask, bit-exact
    hdr = ref_hdul[0].header
    for islitlet, (imn, imx, jmn, jmx) in enumerate(keywords[0], start=1):
        assert (imn, imx, jmn, jmx) == tuple(
            hdr[f"{k}{islitlet:02d}"] for k in ("IMNSLT", "IMXSLT", "JMNSLT", "JMXSLT"))


@pytest.mark.parametrize("mode", [1, 2])
def test_float32_parity_full_frame(coeff_j, frame_j, oracle_j, mode):
    _, res = oracle_j
    ref_hdul, ref64 = res[mode]
    out, keywords = rectify_frames(frame_j, coeff_j, mode, dtype="float32")
    got = out[0]
    assert got.dtype == np.float32
    # float32: 1e-5 relative, floor 1e-6 x local flux (parity_util.assert_flux_close_f32)
    assert_flux_close_f32(got, ref64, f"mode {mode}")
    assert np.array_equal(got == 0, ref64 == 0)
    hdr = ref_hdul[0].header
    for islitlet, (imn, imx, jmn, jmx) in enumerate(keywords[0], start=1):
        assert (jmn, jmx) == (hdr[f"JMNSLT{islitlet:02d}"], hdr[f"JMXSLT{islitlet:02d}"])


def test_dropin_hdulist_matches_oracle(coeff_j, frame_j, oracle_j):
    hdul, res = oracle_j
    ref_hdul, _ = res[2]
    before = copy.deepcopy(hdul[0].header.cards)
    got = apply_rectwv_coeff(hdul, coeff_j)                   # default kwargs, like the recipes
    assert hdul[0].header.cards == before and hdul[0].data is frame_j   # input untouched
    assert got[0].data.dtype == np.float32 and got[0].data.shape == (2090, 3400)
    assert "MECS" in got and got["MECS"].header["XDTU"] == hdul["MECS"].header["XDTU"]
    assert header_cards(got) == header_cards(ref_hdul)       # keywords, values, comments, order
    assert_flux_close_f32(got[0].data, ref_hdul[0].data.astype(np.float64))
    for key in ("CD1_1", "PCD1_1", "PCRPIX1"):
        assert key not in got[0].header
    assert got[0].header["CRVAL1"] == 11200.0 and got[0].header["CTYPE1"] == "WAVE"


def test_dropin_float64_input_uses_float64_kernels(coeff_j, frame_j, oracle_j):
    hdul, res = oracle_j
    ref_hdul, _ = res[2]
    h64 = synthetic.make_hdulist(frame_j.astype(np.float64), coeff_j)
    got = apply_rectwv_coeff(h64, coeff_j)
    # float64 arithmetic then the same float32 cast as the reference: equal to 1 ulp
    a, b = got[0].data, ref_hdul[0].data
    assert a.dtype == np.float32
    assert np.all(np.abs(a - b) <= np.spacing(np.abs(b)) + 1e-12 * np.abs(b).max())


def test_error_behaviour(coeff_j, frame_j):
    hdul = synthetic.make_hdulist(frame_j, coeff_j)
    bad = copy.deepcopy(hdul)
    bad[0].header["FILTER"] = "H"
    with pytest.raises(ValueError, match="Filter name does not match!"):
        apply_rectwv_coeff(bad, coeff_j)
    bad = copy.deepcopy(hdul)
    bad[0].header["GRISM"] = "K"
    with pytest.raises(ValueError, match="Grism name does not match!"):
        apply_rectwv_coeff(bad, coeff_j)
    bad = copy.deepcopy(hdul)
    bad[0].header["XDTU"] = 1.0
    apply_rectwv_coeff(bad, coeff_j)                                   # warning only
    with pytest.raises(ValueError, match="DTU configurations do not match!"):
        apply_rectwv_coeff(bad, coeff_j, args_ignore_dtu_configuration=False)
    with pytest.raises(ValueError, match="Unexpected resampling value=7"):
        apply_rectwv_coeff(hdul, coeff_j, args_resampling=7)
    # 3 is the bilinear extension: refused like in the reference unless asked for by name
    with pytest.raises(ValueError, match="Unexpected resampling value=3"):
        apply_rectwv_coeff(hdul, coeff_j, args_resampling=3)
    assert apply_rectwv_coeff(hdul, coeff_j, args_resampling=3,
                              allow_extensions=True)[0].data.shape == (2090, 3400)
    small = synthetic.make_hdulist(frame_j[:100], coeff_j)
    with pytest.raises(ValueError, match="Unexpected naxis2"):
        apply_rectwv_coeff(small, coeff_j)
    c5 = synthetic.make_config(3)
    h5 = synthetic.make_hdulist(frame_j, c5)
    with pytest.raises(ValueError, match="order > 4 not implemented"):
        apply_rectwv_coeff(h5, c5)
    cbad = copy.deepcopy(coeff_j)
    cbad.global_integer_offset_x_pix = 1.0
    with pytest.raises(ValueError, match="non-integer offsets"):
        apply_rectwv_coeff(hdul, cbad)
    cbad = copy.deepcopy(coeff_j)
    cbad.contents[9]["max_row_rectified"] = (cbad.contents[9]["bb_ns2_orig"]
                                             - cbad.contents[9]["bb_ns1_orig"])
    cbad.contents[9]["min_row_rectified"] = cbad.contents[9]["max_row_rectified"] - 37
    with pytest.raises(IndexError):
        apply_rectwv_coeff(hdul, cbad)


# ------------------------------------------------------------------ edge cases
def test_missing_slitlets_and_offsets(frame_j):
    coeff = synthetic.make_config(1, missing_slitlets=[1, 55, 17], offx=3, offy=-2)
    hdul = synthetic.make_hdulist(frame_j, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, only_needed_rows=True, return_float64=True)
    got = apply_rectwv_coeff(hdul, coeff, dtype="float64")
    assert header_cards(got) == header_cards(ref)
    for islitlet in (1, 17, 55):
        assert np.all(got[0].data[(islitlet - 1) * 38: islitlet * 38] == 0)
        assert got[0].header[f"IMNSLT{islitlet:02d}"] == 0
    out, _ = rectify_frames(frame_j.astype(np.float64), coeff, 2, dtype="float64")
    assert_flux_close(out[0], ref64, 1e-12, 1e-12)
    assert np.array_equal(out[0] == 0, ref64 == 0)


def test_order5_distortion(frame_j):
    coeff = only_slitlets(synthetic.make_config(3), [3, 30, 53])
    hdul = synthetic.make_hdulist(frame_j, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, only_needed_rows=True, return_float64=True,
                              nmax_order=5)
    out, kw = rectify_frames(frame_j.astype(np.float64), coeff, 2, dtype="float64", max_order=5)
    assert_flux_close(out[0], ref64, 1e-12, 1e-12)
    assert np.array_equal(out[0] == 0, ref64 == 0)
    out32, _ = rectify_frames(frame_j, coeff, 2, dtype="float32", max_order=5)
    assert_flux_close_f32(out32[0], ref64)


@pytest.mark.parametrize("number", [4, "4b"])
def test_lr_grisms_float64(frame_j, number):
    coeff = only_slitlets(synthetic.make_config(number), [4, 29, 55])
    hdul = synthetic.make_hdulist(frame_j, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, only_needed_rows=True, return_float64=True)
    out, kw = rectify_frames(frame_j.astype(np.float64), coeff, 2, dtype="float64")
    assert out.shape[1:] == ref64.shape and ref64.shape[1] in (1270, 1435)
    assert_flux_close(out[0], ref64, 1e-12, 1e-12)
    assert np.array_equal(out[0] == 0, ref64 == 0)
    got = apply_rectwv_coeff(hdul, coeff, dtype="float64")
    assert header_cards(got) == header_cards(ref)


def test_bilinear_extension(coeff_j, frame_j):
    coeff = only_slitlets(coeff_j, [12, 44])
    hdul = synthetic.make_hdulist(frame_j, coeff)
    _, ref64 = oracle_apply(hdul, coeff, args_resampling=3, only_needed_rows=True,
                            allow_extensions=True, return_float64=True)
    out, _ = rectify_frames(frame_j.astype(np.float64), coeff, 3, dtype="float64")
    assert_flux_close(out[0], ref64, 1e-12, 1e-12)


def test_integer_input(coeff_j, frame_j):
    coeff = only_slitlets(coeff_j, [8])
    data = np.clip(frame_j, 0, 60000).astype(np.uint16)
    hdul = synthetic.make_hdulist(data, coeff)
    ref = oracle_apply(hdul, coeff, only_needed_rows=True)
    got = apply_rectwv_coeff(hdul, coeff)
    assert header_cards(got) == header_cards(ref)
    assert_flux_close_f32(got[0].data, ref[0].data.astype(np.float64))


# ------------------------------------------------------------------ size-independent properties
@pytest.fixture(scope="module")
def batch64(coeff_j, scene_j):
    """BASELINE config-2 sized batch: 64 frames (float32)."""
    return synthetic.make_frames(coeff_j, 64, seed=77, scene=scene_j)


def test_batch_properties_64_frames(coeff_j, batch64):
    import torch
    plan = get_plan(coeff_j, 2, "float32")
    frames = torch.from_numpy(batch64).cuda()
    out, jmm = plan.run_torch(frames)
    torch.cuda.synchronize()
    # determinism and independence of the frame grouping inside a CTA
    for g in (1, 2, 4):
        plan.set_tuning(g)
        out2, jmm2 = plan.run_torch(frames[:7])
        torch.cuda.synchronize()
        assert torch.equal(out2, out[:7]) and torch.equal(jmm2, jmm[:7]), g
    plan.set_tuning(0)
    # homogeneity: every stage is positively homogeneous, scaling by 2 is exact in binary
    out3, _ = plan.run_torch(frames[:4] * 2.0)
    assert torch.equal(out3, out[:4] * 2.0)
    # zero frames -> exact zeros and no useful channel
    outz, jmmz = plan.run_torch(torch.zeros_like(frames[:2]))
    assert not outz.any()
    assert bool((jmmz[..., 1] == -1).all()) and bool((jmmz[..., 0] == 2**31 - 1).all())
    # flux conservation of the rebin: sum over wavelength of a slitlet row equals the
    # rectified row's flux inside the output grid (float64 sums of float32 data)
    islitlet = 28
    entry = coeff_j.contents[islitlet - 1]
    tb = proto.rebin_tables(entry, **grid_of(coeff_j))
    rect = plan.rectify_slitlet_torch(frames[5], islitlet).double().cpu().numpy()
    rows = out[5, (islitlet - 1) * 38: islitlet * 38].double().cpu().numpy()
    starts_before = tb["idx"][0] == 0 and tb["t"][0] == 0
    ends_after = tb["idx"][-1] == tb["nchan"] - 1 and np.isclose(tb["tau"][-1], 1.0)
    assert starts_before and ends_after          # J slitlet 28 lies inside the grid
    assert np.allclose(rows.sum(axis=1), rect[:38].sum(axis=1), rtol=2e-6)
    # host-buffer entry point agrees bit for bit with the device entry point
    host_out, host_jmm = plan.run_host(batch64[:11])
    assert np.array_equal(host_out, out[:11].cpu().numpy())
    assert np.array_equal(host_jmm, jmm[:11].cpu().numpy())


def test_batch_api_equals_single_calls(coeff_j, batch64):
    hduls = [synthetic.make_hdulist(batch64[i], coeff_j) for i in range(3)]
    batch = apply_rectwv_coeff_batch(hduls, coeff_j)
    for hdul, got in zip(hduls, batch):
        single = apply_rectwv_coeff(hdul, coeff_j)
        assert np.array_equal(single[0].data, got[0].data)
        assert header_cards(single) == header_cards(got)


def test_plan_cache_reuse(coeff_j):
    clear_plan_cache()
    a = get_plan(coeff_j, 2, "float32")
    b = get_plan(copy.deepcopy(coeff_j), 2, "float32")
    assert a is b
    c = copy.deepcopy(coeff_j)
    c.global_integer_offset_x_pix += 1
    assert get_plan(c, 2, "float32") is not a


# ------------------------------------------------------------------ per-slitlet API
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_rectify2d_matches_oracle_all_rows(coeff_j, frame_j, mode):
    from pyemir_b200.slitlet2d import rectify2d
    entry = coeff_j.contents[19]
    crop = frame_j[entry["bb_ns1_orig"] - 1:entry["bb_ns2_orig"]].astype(float)
    want = nr.rectify2d(crop, entry["ttd_aij"], entry["ttd_bij"], mode)
    got = rectify2d(crop, entry["ttd_aij"], entry["ttd_bij"], mode)
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.array_equal(got, want)          # same operation order: bit-identical float64


def test_slitlet2d_mirror_direct_and_inverse(coeff_j, frame_j):
    from oracle.pipeline import OracleSlitlet2D
    from pyemir_b200.slitlet2d import Slitlet2D
    for islitlet in (3, 54):
        slt = Slitlet2D(islitlet, coeff_j, debugplot=0)
        ref = OracleSlitlet2D(islitlet, coeff_j)
        assert (slt.iminslt, slt.imaxslt, slt.ttd_order) == (ref.iminslt, ref.imaxslt, 2)
        crop = slt.extract_slitlet2d(frame_j)
        assert np.array_equal(crop, ref.extract_slitlet2d(frame_j))
        for resampling in (1, 2):
            rect = slt.rectify(crop, resampling)
            assert np.array_equal(rect, ref.rectify(crop, resampling))
            back = slt.rectify(rect, resampling, inverse=True)
            assert np.array_equal(back, ref.rectify(rect, resampling, inverse=True))
        with pytest.raises(ValueError, match="Unexpected resampling value=3"):
            slt.rectify(crop, 3)
        with pytest.raises(ValueError, match="Unexpected slitlet2d_rect naxis2"):
            slt.rectify(crop[:-1], 2)
    with pytest.raises(ValueError, match="Unexpected naxis1"):
        Slitlet2D(3, coeff_j).extract_slitlet2d(frame_j[:, :100])
    broken = copy.deepcopy(coeff_j)
    del broken.contents[2]["corr_yrect_a"]
    with pytest.raises(KeyError):
        Slitlet2D(3, broken)


# ------------------------------------------------------------------ remaining BASELINE configs
def test_config2_hh_float32_full_frame():
    """BASELINE config 2 geometry (grism H + filter H), one full frame, float32."""
    coeff = synthetic.make_config(2)
    frame = synthetic.make_frames(coeff, 1, seed=1002)[0]
    hdul = synthetic.make_hdulist(frame, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, only_needed_rows=True, return_float64=True)
    got = apply_rectwv_coeff(hdul, coeff)
    assert header_cards(got) == header_cards(ref)
    assert_flux_close_f32(got[0].data, ref64)
    assert np.array_equal(got[0].data == 0, ref64 == 0)


def test_config3_order5_batch_tail_and_nearest():
    """Order-5 map (K + Ksp), modes 1 and 2, a 7-frame batch (odd tail of a G=4 group)."""
    coeff = only_slitlets(synthetic.make_config(3), [2, 27, 28, 55])
    frames = synthetic.make_frames(coeff, 7, seed=1003)
    for mode in (1, 2):
        out, kw = rectify_frames(frames, coeff, mode, dtype="float32", max_order=5)
        for i in (0, 6):
            hdul = synthetic.make_hdulist(frames[i], coeff)
            ref, ref64 = oracle_apply(hdul, coeff, args_resampling=mode, only_needed_rows=True,
                                      return_float64=True, nmax_order=5)
            assert_flux_close_f32(out[i], ref64, f"mode {mode} frame {i}")
            assert np.array_equal(out[i] == 0, ref64 == 0)
            for islitlet, (_, _, jmn, jmx) in enumerate(kw[i], start=1):
                assert (jmn, jmx) == (ref[0].header[f"JMNSLT{islitlet:02d}"],
                                      ref[0].header[f"JMXSLT{islitlet:02d}"])


def test_lr_float32_unaligned_rows():
    """LR + HK: naxis1 = 1435 (rows not 16-byte aligned: scalar zero-fill path)."""
    coeff = only_slitlets(synthetic.make_config("4b"), [10, 40])
    frame = synthetic.make_frames(coeff, 1, seed=1004)[0]
    hdul = synthetic.make_hdulist(frame, coeff)
    _, ref64 = oracle_apply(hdul, coeff, only_needed_rows=True, return_float64=True)
    out, _ = rectify_frames(frame, coeff, 2, dtype="float32")
    assert out.shape == (1, 2090, 1435)
    assert_flux_close_f32(out[0], ref64)
    assert np.array_equal(out[0] == 0, ref64 == 0)


# ------------------------------------------------------------------ kernel variants, host entry
def test_item_distribution_variants_bit_identical(monkeypatch):
    """Which CTA takes which item (per-launch counter vs fixed stride), the order of the tiles
    and the moment the copy warp requests an item's first rows do not change a single bit of
    the product or of the JMN/JMX scan."""
    coeff = only_slitlets(synthetic.make_config(2), list(range(1, 56, 3)))
    fr = synthetic.make_frames(coeff, 9, seed=21)
    res = []
    for env in ({}, {"EMIRWV_DYNAMIC": "0"}, {"EMIRWV_TILE_ORDER": "0", "EMIRWV_EARLY": "0"},
                {"EMIRWV_DYNAMIC": "0", "EMIRWV_TILE_ORDER": "0", "EMIRWV_EARLY": "0"}):
        for key in ("EMIRWV_DYNAMIC", "EMIRWV_TILE_ORDER", "EMIRWV_EARLY"):
            monkeypatch.delenv(key, raising=False)
        for key, value in env.items():
            monkeypatch.setenv(key, value)
        plan = RectWavePlan(coeff, 2, "float32")
        out, jmm = plan.run_host(fr)
        out2, jmm2 = plan.run_host(fr)                       # counters are per launch
        assert np.array_equal(out, out2) and np.array_equal(jmm, jmm2)
        res.append((out, jmm))
        plan.close()
    for out, jmm in res[1:]:
        assert np.array_equal(out, res[0][0]) and np.array_equal(jmm, res[0][1])
    assert np.count_nonzero(res[0][0]) > 0


@pytest.mark.parametrize("number", [1, 4])
def test_float64_two_stage_pipeline_bit_identical(monkeypatch, number):
    """The depth of the copy pipeline (rows in flight per CTA: EMIRWV_NSTAGE_F64, or the
    `stages` field of the plan description) does not change a bit of the result."""
    coeff = only_slitlets(synthetic.make_config(number), [1, 2, 28, 54, 55])
    fr = synthetic.make_frames(coeff, 5, seed=31, dtype=np.float64)
    res = []
    for stages in ("4", "2"):
        monkeypatch.setenv("EMIRWV_NSTAGE_F64", stages)
        plan = RectWavePlan(coeff, 2, "float64")
        res.append(plan.run_host(fr))
        assert plan.info()["stages"] == int(stages)
        plan.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.count_nonzero(res[0][0]) > 0


@pytest.mark.parametrize("number,dtype", [(2, "float32"), ("4b", "float64")])
def test_run_host_out_zeroed_equals_full_copy(number, dtype):
    """EMIRWV_HOST_OUT_ZEROED copies back only the covered columns of every slitlet; on a
    zeroed buffer the result is the full product, and everything outside
    emirwv_plan_get_coverage is exactly zero in the full copy."""
    coeff = synthetic.make_config(number)
    coeff.missing_slitlets = [3, 55]
    fr = synthetic.make_frames(coeff, 6, seed=5, dtype=np.dtype(dtype).type)
    plan = RectWavePlan(coeff, 2, dtype)
    full, jmm_full = plan.run_host(fr)
    cover = plan.coverage()
    assert cover.shape == (55, 2) and tuple(cover[2]) == (0, 0) and tuple(cover[54]) == (0, 0)
    inside = np.zeros(full.shape[1:], dtype=bool)
    for s, (lo, hi) in enumerate(cover):
        assert 0 <= lo <= hi <= full.shape[2]
        inside[s * 38:(s + 1) * 38, lo:hi] = True
    assert np.all(full[:, ~inside] == 0)
    assert 0.3 < inside.mean() < 0.9
    out = np.zeros_like(full)
    sparse, jmm = plan.run_host(fr, out=out, out_zeroed=True)
    assert sparse is out
    assert np.array_equal(sparse, full) and np.array_equal(jmm, jmm_full)
    # a reused buffer: second call with other frames over the first result
    fr2 = fr[::-1].copy()
    plan.run_host(fr2, out=out, out_zeroed=True)
    assert np.array_equal(out, full[::-1])
    plan.close()


# ---------------------------------------------------------------- median_slitlets_rectified
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [0, 1])
def test_median_slitlets_matches_numpy(dtype, mode):
    """k_median_slitlets against numpy.median (what the reference calls,
    median_slitlets_rectified.py:89): random stacks with ties, signed zeros, infinities and
    NaNs; ragged width (not a multiple of the block)."""
    from pyemir_b200.median_slitlets_rectified import median_frames
    rng = np.random.default_rng(7)
    n, naxis1 = 3, 1271
    stack = rng.normal(100.0, 30.0, size=(n, 2090, naxis1)).astype(dtype)
    stack[0, :, 10:60] = np.round(stack[0, :, 10:60] / 25.0) * 25.0      # many ties
    stack[1, 38 * 3:38 * 4, :] = 0.0
    stack[1, 38 * 3 + 5, 100:200] = -0.0
    stack[2, 38 * 7 + 11, 300] = np.nan
    stack[2, 38 * 8 + 1, 301] = np.inf
    stack[2, 38 * 8 + 2, 301] = -np.inf
    got = median_frames(stack, mode=mode)
    want = np.median(stack.reshape(n, 55, 38, naxis1), axis=2)
    if mode == 0:
        want = np.repeat(want, 38, axis=1)
    assert got.dtype == dtype and got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)                      # bit exact


@pytest.mark.gpu
def test_median_slitlets_torch_device_resident(coeff_j, frame_j):
    """Product of the hot kernel -> median kernel, both device resident."""
    import torch
    from pyemir_b200.median_slitlets_rectified import median_frames_torch
    plan = get_plan(coeff_j, 2, "float32")
    frames = torch.from_numpy(np.stack([frame_j, frame_j[::-1].copy()])).cuda()
    out, _ = plan.run_torch(frames)
    med = median_frames_torch(out, mode=1)
    torch.cuda.synchronize()
    want = np.median(out.cpu().numpy().reshape(2, 55, 38, -1), axis=2)
    assert np.array_equal(med.cpu().numpy(), want, equal_nan=True)
    with pytest.raises(ValueError):
        median_frames_torch(out[:, :100].contiguous(), mode=1)


# ---------------------------------------------------------------- ABBA mean combination
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_abba_mean_matches_oracle(dtype):
    """emirwv_abba_partial / _finish against the restated combination (abba.py:579-606):
    float64 sums in frame order, float32 means, float32 difference: bit exact."""
    import torch
    from oracle.pipeline import combine_abba_mean as oracle_abba
    from pyemir_b200 import abba
    rng = np.random.default_rng(11)
    full_set = abba.build_full_set("ABBA", 2, 2)                     # 16 frames
    stack = (rng.normal(500.0, 200.0, size=(len(full_set), 77, 1031))).astype(dtype)
    labels = abba.labels_of(full_set)
    frames = torch.from_numpy(stack).cuda()
    got = abba.combine_abba_mean(frames, labels, full_set.count("A"), full_set.count("B"))
    torch.cuda.synchronize()
    want = oracle_abba(stack, full_set)
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
    # the same in two chunks (accumulating partial sums), with an unused frame in between
    sums = abba.partial_sums(frames[:5].contiguous(), labels[:5])
    sums = abba.partial_sums(frames[4:].contiguous(), [0] + labels[5:], sums=sums)
    got2 = abba.finish(sums, full_set.count("A"), full_set.count("B"))
    assert np.array_equal(got2.cpu().numpy(), want)
    with pytest.raises(ValueError):
        abba.finish(sums, 0, 8)
    with pytest.raises(ValueError):
        abba.partial_sums(frames, labels[:3])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("repeat,nseq,width", [(1, 1, 517), (1, 1, 516), (1, 3, 516), (2, 3, 517),
                                               (2, 4, 516), (3, 5, 517), (8, 2, 516), (16, 2, 517)])
def test_abba_median_matches_oracle(dtype, repeat, nseq, width):
    """emirwv_stack_median (+ _finish) against numpy.median per nodding position: 2 .. 64
    images per position (every register-array size of the kernel), ties, signed zeros,
    infinities and NaNs; bit exact."""
    import torch
    from oracle.pipeline import combine_abba_median as oracle_abba
    from pyemir_b200 import abba
    rng = np.random.default_rng(100 * repeat + nseq)
    full_set = abba.build_full_set("ABBA", repeat, nseq)
    n = len(full_set)
    # width 516: 16-byte loads, four pixels per thread; 517: scalar path
    stack = rng.normal(500.0, 200.0, size=(n, 40, width)).astype(dtype)
    stack[:, 3, :] = np.round(stack[:, 3, :] / 100.0)                 # many ties
    stack[::2, 4, :50] = 0.0
    stack[1::2, 4, :50] = -0.0
    stack[0, 5, 7] = np.nan
    stack[n - 1, 5, 8] = np.nan
    stack[1, 6, 9] = np.inf
    stack[2, 6, 10] = -np.inf
    stack[:, 7, 11] = np.inf
    labels = abba.labels_of(full_set)
    frames = torch.from_numpy(stack).cuda()
    index_a = [i for i, c in enumerate(full_set) if c == "A"]
    med_a = abba.stack_median(frames, index_a)
    torch.cuda.synchronize()
    want_a = np.median(stack[index_a].astype(np.float64), axis=0)
    assert med_a.dtype == torch.float64
    assert np.array_equal(med_a.cpu().numpy(), want_a, equal_nan=True)
    got = abba.combine_abba_median(frames, labels, [n])
    torch.cuda.synchronize()
    with np.errstate(invalid="ignore"):
        want = oracle_abba(stack, full_set)
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want, equal_nan=True)
    # unequal numbers of A and B images (one A image dropped)
    index_b = [i for i, c in enumerate(full_set) if c == "B"]
    if len(index_a) > 1:
        got2 = abba.abba_median_difference(frames, index_a[1:], index_b).cpu().numpy()
        with np.errstate(invalid="ignore"):
            want2 = (np.median(stack[index_a[1:]].astype(np.float64), axis=0).astype(np.float32)
                     - np.median(stack[index_b].astype(np.float64), axis=0).astype(np.float32))
        assert np.array_equal(got2, want2, equal_nan=True)


@pytest.mark.gpu
def test_abba_median_errors():
    import torch
    from pyemir_b200 import abba
    from pyemir_b200._cabi import EmirwvError
    frames = torch.zeros((70, 4, 8), dtype=torch.float32, device="cuda")
    with pytest.raises(EmirwvError):
        abba.stack_median(frames, list(range(65)))                    # more than 64 images
    with pytest.raises(ValueError):
        abba.stack_median(frames, [70])                               # outside the stack
    with pytest.raises(ValueError):
        abba.stack_median(frames, [])
    with pytest.raises(ValueError):
        abba.combine_abba_median(frames[:4].contiguous(), [1, 1, 1, 1], [4])
    one = abba.stack_median(frames, [3])
    assert float(one.abs().max()) == 0.0


@pytest.mark.gpu
def test_abba_mean_of_rectified_frames(coeff_j, frame_j):
    """Hot kernel -> ABBA combination, device resident: A - B of identical exposures is zero
    where both contribute, and the combination of (A, B) = (2x, x) returns x."""
    import torch
    from pyemir_b200 import abba
    plan = get_plan(coeff_j, 2, "float32")
    base = torch.from_numpy(frame_j).cuda()
    frames = torch.stack([base * 2.0, base, base, base * 2.0])
    out, _ = plan.run_torch(frames)
    diff = abba.combine_abba_mean(out, abba.labels_of("ABBA"), 2, 2)
    torch.cuda.synchronize()
    assert torch.equal(diff, out[1])          # mean(2x, 2x) - mean(x, x) = x exactly


@pytest.mark.gpu
def test_shift_image2d_matches_oracle_and_integer_shift():
    from pyemir_b200.abba import shift_image2d
    rng = np.random.default_rng(3)
    img = rng.normal(100.0, 10.0, size=(40, 300))
    got = shift_image2d(img, yoffset=-1.3)                            # abba.py:556 usage
    want = nr.shift_image2d(img, yoffset=-1.3)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    # an integer offset moves the pixels exactly; what enters from outside is zero
    got = shift_image2d(img, xoffset=3, yoffset=2)
    assert np.array_equal(got[2:, 3:], img[:-2, :-3])
    assert not got[:2].any() and not got[:, :3].any()
    # flux is conserved away from the borders
    got = shift_image2d(img, xoffset=0.25, yoffset=-0.5)
    assert abs(got[2:-2, 2:-2].sum() - img[2:-2, 2:-2].sum()) < 0.02 * img[2:-2, 2:-2].sum()

"""The plan-table formulation used by the CUDA kernels agrees with the oracle.

CPU only.  tests/proto_tables.py states the tables (frame-independent rectification
weights, rebin border tables) and the telescoped rebin; here it is compared with
the oracle's literal cumulative-sum + Steffen + difference pipeline.
"""

import copy

import numpy as np
import pytest

from oracle.pipeline import apply_rectwv_coeff as oracle_apply
from pyemir_b200 import synthetic
from pyemir_b200.set_wv_parameters import set_wv_parameters

import proto_tables as proto


def only_slitlets(coeff, keep):
    c = copy.deepcopy(coeff)
    c.missing_slitlets = [i for i in range(1, 56) if i not in keep]
    return c


def grid_of(coeff):
    p = set_wv_parameters(coeff.tags["filter"], coeff.tags["grism"])
    return dict(naxis1=p["naxis1_enlarged"], cdelt1=p["cdelt1_enlarged"],
                crval1=p["crval1_enlarged"], crpix1=p["crpix1_enlarged"])


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_prototype_matches_oracle_float64(coeff_j, frame_j, mode):
    keep = [1, 28, 55]
    coeff = only_slitlets(coeff_j, keep)
    hdul = synthetic.make_hdulist(frame_j, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, args_resampling=mode, only_needed_rows=True,
                              allow_extensions=True, return_float64=True)
    scale = np.abs(ref64).max()
    for isl in keep:
        entry = coeff.contents[isl - 1]
        rows, rect, K = proto.run_slitlet(frame_j, entry, grid_of(coeff), mode)
        got = rows[:38]
        want = ref64[(isl - 1) * 38: isl * 38]
        # the oracle differences a running sum of ~1e5..1e6 -> ~1e-10 absolute noise
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * scale), \
            (isl, np.abs(got - want).max())
        assert np.array_equal(got == 0, want == 0)
        hdr = ref[0].header
        assert proto.jminmax(rows) == (hdr["JMNSLT%02d" % isl], hdr["JMXSLT%02d" % isl])
        assert K == {1: 1, 2: K, 3: 2}[mode] and K <= 3


def test_prototype_float32_within_1e5(coeff_j, frame_j):
    keep = [14]
    coeff = only_slitlets(coeff_j, keep)
    hdul = synthetic.make_hdulist(frame_j, coeff)
    ref, ref64 = oracle_apply(hdul, coeff, args_resampling=2, only_needed_rows=True,
                              return_float64=True)
    entry = coeff.contents[13]
    rows, _, _ = proto.run_slitlet(frame_j, entry, grid_of(coeff), 2, dtype=np.float32)
    assert rows.dtype == np.float32
    want = ref64[13 * 38: 14 * 38]
    scale = np.abs(want).max()
    err = np.abs(rows[:38] - want)
    assert np.all(err <= 1e-5 * np.abs(want) + 1e-6 * scale), err.max()
    assert np.array_equal(rows[:38] == 0, want == 0)


def test_prototype_with_global_offsets(frame_j):
    coeff = synthetic.make_config(1, offx=3, offy=-2)
    keep = [1, 30]
    coeff = only_slitlets(coeff, keep)
    hdul = synthetic.make_hdulist(frame_j, coeff)
    _, ref64 = oracle_apply(hdul, coeff, args_resampling=2, only_needed_rows=True,
                            return_float64=True)
    scale = np.abs(ref64).max()
    for isl in keep:
        rows, _, _ = proto.run_slitlet(frame_j, coeff.contents[isl - 1], grid_of(coeff),
                                       2, offx=3, offy=-2)
        want = ref64[(isl - 1) * 38: isl * 38]
        assert np.allclose(rows[:38], want, rtol=1e-12, atol=1e-12 * scale)

"""Host-side plumbing that needs no GPU: CPU lists of the NUMA placement, the pinned pool's
fallback, the plan-key memo, the gating of the bilinear extension."""

import copy

import numpy as np
import pytest

from pyemir_b200 import _pinned, distributed, synthetic
from pyemir_b200.plan import _fingerprint, _quick_look


def test_parse_cpulist():
    assert distributed._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert distributed._parse_cpulist("") == set()
    assert distributed._parse_cpulist("5") == {5}


def test_bind_is_harmless_without_a_gpu(monkeypatch):
    import os
    before = os.sched_getaffinity(0)
    info = distributed.bind_to_gpu_numa(0)
    assert info["numa_node"] is None and os.sched_getaffinity(0) == before
    monkeypatch.setenv("EMIRWV_NUMA_BIND", "0")
    assert distributed.bind_to_gpu_numa(0)["skipped"] == "EMIRWV_NUMA_BIND=0"


def test_pinned_pool_falls_back_to_pageable_memory():
    a = _pinned.empty((3, 5), np.float32)         # no CUDA device here: cudaHostAlloc fails
    assert a.shape == (3, 5) and a.dtype == np.float32 and a.flags.c_contiguous
    a[...] = 1.5
    assert float(a.sum()) == 22.5
    assert _pinned.empty((0, 4), np.float64).size == 0
    _pinned.trim()


def test_quick_look_changes_with_everything_the_fingerprint_covers():
    coeff = synthetic.make_config(1)
    opts = (2, "float32", 4, False, None, (0, 0), 0)
    base_look, base_key = _quick_look(coeff, opts), _fingerprint(coeff, *opts)
    assert _quick_look(coeff, opts) == base_look
    edits = []
    c = copy.deepcopy(coeff); c.global_integer_offset_x_pix = 3; edits.append(c)
    c = copy.deepcopy(coeff); c.missing_slitlets = [4]; edits.append(c)
    c = copy.deepcopy(coeff); c.contents[10]["ttd_aij"][0] += 1e-9; edits.append(c)
    c = copy.deepcopy(coeff); c.contents[10]["wpoly_coeff"][-1] *= 1.0001; edits.append(c)
    for c in edits:
        # (a different object: the id differs anyway; compare what follows the id)
        assert _quick_look(c, opts)[1:] != base_look[1:]
        assert _fingerprint(c, *opts) != base_key
    assert _quick_look(coeff, (1,) + opts[1:]) != base_look
    # an edit in place is seen although the object is the same
    before = _quick_look(coeff, opts)
    coeff.contents[3]["ttd_bij"][0] += 0.5
    assert _quick_look(coeff, opts) != before


def test_bilinear_extension_is_refused_like_in_the_reference():
    """args_resampling=3 raises the reference's error before any device work."""
    from pyemir_b200.apply_rectwv_coeff import apply_rectwv_coeff, apply_rectwv_coeff_batch
    coeff = synthetic.make_config(1)
    frame = np.zeros((2048, 2048), dtype=np.float32)
    hdul = synthetic.make_hdulist(frame, coeff)
    with pytest.raises(ValueError, match="Unexpected resampling value=3"):
        apply_rectwv_coeff(hdul, coeff, args_resampling=3)
    with pytest.raises(ValueError, match="Unexpected resampling value=3"):
        apply_rectwv_coeff_batch([hdul], coeff, args_resampling=3)
    # the input is not modified and shares no header with the failed call
    assert hdul[0].data is frame

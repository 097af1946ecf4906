"""Host mirrors of get_islitlet / useful_mos_xpixels against vectors from the reference's own
functions (tests/golden/make_golden_useful_xpixels.py executes them): the consumer of the
JMNSLTnn / JMXSLTnn keywords that apply_rectwv_coeff writes (recipes/spec/abba.py:423-440)."""

import json
import os
import sys

import numpy as np
import pytest

from pyemir_b200.get_islitlet import get_islitlet
from pyemir_b200.useful_mos_xpixels import useful_mos_xpixels

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_useful_xpixels as gen  # noqa: E402  (inputs only; no reference read)

with open(os.path.join(HERE, "golden", "useful_mos_xpixels.json")) as _fd:
    GOLDEN = json.load(_fd)


def test_get_islitlet_matches_reference():
    for key, want in GOLDEN["get_islitlet"].items():
        ipixel = float(key) if "." in key else int(key)
        if isinstance(want, dict):
            with pytest.raises(ValueError) as err:
                get_islitlet(ipixel)
            assert str(err.value) == want["ValueError"]
        else:
            assert get_islitlet(ipixel) == want, key
    # every row of the rectified image: slitlet i owns rows (i-1)*38+1 .. i*38
    rows = np.arange(1, 2091)
    assert [get_islitlet(int(r)) for r in rows] == list((rows - 1) // 38 + 1)


@pytest.mark.parametrize("case", gen.CASES, ids=[c[0] for c in gen.CASES])
def test_useful_mos_xpixels_matches_reference(case):
    name, kind, vpix, nrem, regions = case
    want = GOLDEN["cases"][name]
    data = np.zeros((gen.NAXIS2, gen.NAXIS1), dtype=np.float32)
    kwargs = dict(vpix_region=vpix, npix_removed_near_ohlines=nrem,
                  list_valid_wvregions=regions, debugplot=0, oh_catalogue=gen.oh_catalogue())
    errors = [k for k in want if k.endswith("Error")]
    if errors:
        exc = {"ValueError": ValueError, "IndexError": IndexError}[errors[0]]
        with pytest.raises(exc) as err:
            useful_mos_xpixels(data, gen.make_header(kind), **kwargs)
        assert str(err.value) == want[errors[0]]
        return
    mask = useful_mos_xpixels(data, gen.make_header(kind), **kwargs)
    assert mask.dtype == np.bool_ and list(mask.shape) == want["shape"]
    assert int(mask.sum()) == want["count"]
    assert gen.runs_of(mask) == want["runs"]


def test_oh_catalogue_is_required_when_emirdrp_is_absent():
    data = np.zeros((gen.NAXIS2, gen.NAXIS1), dtype=np.float32)
    with pytest.raises(ValueError, match="OH line catalogue"):
        useful_mos_xpixels(data, gen.make_header(), (10, 20), npix_removed_near_ohlines=2)


def test_consumes_the_header_the_dropin_writes():
    """Works on the Header object apply_rectwv_coeff returns (case-insensitive keywords)."""
    from pyemir_b200.hdu import Header
    hdr = Header()
    for key, value in gen.make_header().items():
        hdr[key.upper()] = value
    mask = useful_mos_xpixels(None, hdr, (5, 30))
    assert gen.runs_of(mask) == GOLDEN["cases"]["one_slitlet"]["runs"]

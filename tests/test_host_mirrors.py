"""Host-side mirrors of the reference interface (no GPU, no oracle needed)."""

import numpy as np
import pytest

from pyemir_b200.dtu import DtuConf
from pyemir_b200.hdu import HDUList, Header, ImageHDU, PrimaryHDU, copy_img
from pyemir_b200.products import RectWaveCoeff
from pyemir_b200.set_wv_parameters import set_wv_parameters
from pyemir_b200 import synthetic


# ---- DtuConf: known answers from the reference's tests/instrument/components/test_dtu.py
def test_dtu_values_lowercase_dict():
    conf = DtuConf.from_header(dict(xdtu=1.0, ydtu=1.0, zdtu=1.0))   # test_dtu.py:17-19
    assert conf.coor == [1.0, 1.0, 1.0]


def test_dtu_values_raises():
    with pytest.raises(KeyError):                                      # test_dtu.py:22-24
        DtuConf.from_header(dict(kwargs=1.0, kwargs1=1.0))


def test_dtu_coor_r_known_answer():
    conf = DtuConf.from_header(synthetic.DTU_HEADER_EXAMPLE)           # test_dtu.py:54-59
    assert np.allclose(conf.coor, [-205.679000854492, -24.4878005981445, -463.765991210938])
    assert np.allclose(conf.coor_r, [-17.206285211909773, -0.7186057740102463,
                                     0.055023193358977096])


def test_dtu_formatter_known_answer():
    expected = (                                                       # test_dtu.py:62-77
        "<DtuConf instance>\n"
        "- XDTU..: -205.679\n- YDTU..:  -24.488\n- ZDTU..: -463.766\n"
        "- XDTU_0: -205.651\n- YDTU_0:  -24.479\n- ZDTU_0: -463.821\n"
        "- XDTU_F:    0.923\n- YDTU_F:    0.972\n- ZDTU_F:    1.000\n")
    assert DtuConf.from_header(synthetic.DTU_HEADER_EXAMPLE).describe() == expected


def test_dtu_hash_known_answer():
    conf = DtuConf.from_header(synthetic.DTU_HEADER_EXAMPLE)           # test_dtu.py:117-120
    assert conf.outdict()["dtuhash"] == "271c6b3476b06c527196b8a96d72c9d6"


def test_dtu_eq_case_insensitive_and_3_decimals():
    upper = DtuConf.from_header(Header(synthetic.DTU_HEADER_EXAMPLE))
    lower = DtuConf.from_header({k.lower(): v for k, v in synthetic.DTU_HEADER_EXAMPLE.items()})
    assert upper == lower                                              # test_dtu.py:27-32
    moved = dict(synthetic.DTU_HEADER_EXAMPLE)
    moved["XDTU"] += 0.0004
    assert DtuConf.from_header(moved) == upper
    moved["XDTU"] += 0.01
    assert DtuConf.from_header(moved) != upper


def test_dtu_from_img_prefers_mecs():
    primary = Header({"XDTU": 1.0, "YDTU": 2.0, "ZDTU": 3.0})
    mecs = Header({"XDTU": 4.0, "YDTU": 5.0, "ZDTU": 6.0})
    img = HDUList([PrimaryHDU(np.zeros((2, 2)), primary)])
    assert DtuConf.from_img(img).coor == [1.0, 2.0, 3.0]
    img.append(ImageHDU(None, mecs, name="MECS"))
    assert DtuConf.from_img(img).coor == [4.0, 5.0, 6.0]


# ---- wavelength grid table (set_wv_parameters.py:64-148)
@pytest.mark.parametrize("grism,filt,crval1,cdelt1,naxis1", [
    ("J", "J", 11200.0, 0.77, 3400), ("H", "H", 14500.0, 1.22, 3400),
    ("K", "Ksp", 19100.0, 1.73, 3400), ("LR", "YJ", 8900.0, 3.56, 1270),
    ("LR", "HK", 14500.0, 6.83, 1435)])
def test_set_wv_parameters_table(grism, filt, crval1, cdelt1, naxis1):
    p = set_wv_parameters(filt, grism)
    assert p["crpix1_enlarged"] == 1.0
    assert (p["crval1_enlarged"], p["cdelt1_enlarged"], p["naxis1_enlarged"]) == \
        (crval1, cdelt1, naxis1)
    assert p["poly_crval1_linear"](0.0) > 0


def test_set_wv_parameters_errors():
    with pytest.raises(ValueError):
        set_wv_parameters("Z", "J")
    with pytest.raises(ValueError):
        set_wv_parameters("J", "Q")
    with pytest.raises(ValueError, match="invalid filter_name and grism_name"):
        set_wv_parameters("H", "J")


# ---- Header / HDUList stand-ins
def test_header_semantics():
    h = Header()
    h["filter"] = "J"
    assert h["FILTER"] == "J" and "Filter" in h
    h["crpix1"] = (1.0, "reference pixel")
    assert h.comments["CRPIX1"] == "reference pixel"
    h["crpix1"] = 2.0                       # value only: comment kept
    assert h["crpix1"] == 2.0 and h.comments["crpix1"] == "reference pixel"
    h["history"] = "one"
    h["history"] = "two"
    assert h["history"] == ["one", "two"]
    h.remove("crpix1")
    assert "crpix1" not in h
    with pytest.raises(KeyError):
        h.remove("crpix1")
    assert h.get("nothing", 5) == 5
    with pytest.raises(KeyError):
        h["nothing"]


def test_copy_img_is_deep():
    img = HDUList([PrimaryHDU(np.ones((2, 2)), Header({"A": 1})),
                   ImageHDU(None, Header({"B": 2}), name="MECS")])
    dup = copy_img(img)
    dup[0].data[0, 0] = 7.0
    dup[0].header["A"] = 3
    assert img[0].data[0, 0] == 1.0 and img[0].header["A"] == 1
    assert "MECS" in dup and dup["MECS"].header["B"] == 2


# ---- RectWaveCoeff state round trip
def test_rectwv_coeff_json_roundtrip(tmp_path, coeff_j):
    path = tmp_path / "coeff.json"
    coeff_j.writeto(str(path))
    back = RectWaveCoeff._datatype_load(str(path))
    assert back.uuid == coeff_j.uuid and back.tags == coeff_j.tags
    assert back.missing_slitlets == coeff_j.missing_slitlets
    assert back.contents[27]["ttd_aij"] == coeff_j.contents[27]["ttd_aij"]
    assert type(back.global_integer_offset_x_pix) is int
    state = back.__getstate__()
    for key in ("instrument", "uuid", "tags", "type", "meta_info", "quality_control",
                "global_integer_offset_x_pix", "global_integer_offset_y_pix",
                "total_slitlets", "missing_slitlets", "contents"):
        assert key in state


def test_synthetic_product_is_well_formed(coeff_j):
    assert len(coeff_j.contents) == 55
    for entry in coeff_j.contents:
        assert entry["max_row_rectified"] - entry["min_row_rectified"] + 1 == 38
        bbh = entry["bb_ns2_orig"] - entry["bb_ns1_orig"] + 1
        assert entry["max_row_rectified"] + 1 < bbh
        assert len(entry["ttd_aij"]) == 6 and len(entry["wpoly_coeff"]) >= 4


# ------------------------------------------------------------------ ABBA / pre-reduction host logic
def test_slab_range_partitions_rows():
    from pyemir_b200.abba import slab_range
    for height in (0, 1, 5, 38, 2090):
        for world in (1, 2, 3, 8):
            spans = [slab_range(height, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == height
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
    assert slab_range(2090, 7, 8) == (1834, 2090)


def test_combine_abba_median_argument_checks():
    """Bookkeeping errors are raised before any device work (no GPU needed)."""
    import torch
    from pyemir_b200 import abba
    frames = torch.zeros((4, 3, 5))

    def never(*args):
        raise AssertionError("the combination must not start")

    with pytest.raises(ValueError, match="both A and B"):
        abba.combine_abba_median(frames, [1, 1, 1, 1], [4], median_difference=never)
    with pytest.raises(ValueError, match="one label per frame"):
        abba.combine_abba_median(frames, [1, -1, -1, 1, 1], [4], median_difference=never)
    # single process: the stand-in sees the whole stack and the A / B index lists
    seen = {}

    def fake(stack, index_a, index_b):
        seen["shape"], seen["a"], seen["b"] = tuple(stack.shape), index_a, index_b
        return torch.ones(stack.shape[1:], dtype=torch.float32)

    out = abba.combine_abba_median(frames, abba.labels_of("ABBA"), [4], median_difference=fake)
    assert seen == {"shape": (4, 3, 5), "a": [0, 3], "b": [1, 2]} and out.shape == (3, 5)


def test_device_entry_points_reject_host_tensors():
    """The CUDA entry points take device tensors only: no silent CPU path."""
    import torch
    from pyemir_b200 import abba
    from pyemir_b200.prereduce import prereduce_torch
    cpu = torch.zeros((2, 4, 8))
    with pytest.raises(ValueError, match="CUDA tensor"):
        abba.stack_median(cpu, [0, 1])
    with pytest.raises(ValueError, match="CUDA tensor"):
        abba.abba_median_difference(cpu, [0], [1])
    with pytest.raises(ValueError, match="CUDA tensor"):
        abba.partial_sums(cpu, [1, -1])
    with pytest.raises(ValueError, match="CUDA tensor"):
        prereduce_torch(cpu, flat=np.ones((4, 8), dtype=np.float32))


def test_flatfield_corrector_constructor_mirrors_reference():
    """`processing/flatfield.py:33-45`: the flat is fixed in place, its mean kept."""
    from pyemir_b200.prereduce import FlatFieldCorrector
    flat = np.array([[1.0, 0.0, -2.0], [0.5, 2.0, 1.5]], dtype=np.float32)
    node = FlatFieldCorrector(flat, calibid="uuid:x")
    assert node.flatdata is flat
    assert np.array_equal(flat, np.array([[1.0, 1.0, 1.0], [0.5, 2.0, 1.5]], dtype=np.float32))
    assert node.flat_stats == flat.mean() and node.calibid == "uuid:x" and node.dtype == "float32"
    assert node.update_variance is False


def test_prereduce_input_widening_follows_numpy_promotion():
    from pyemir_b200.prereduce import as_supported_frames
    for dtype, want in (("float32", "float32"), ("float64", "float64"), ("uint16", "uint16"),
                        ("int16", "float32"), ("uint8", "float32"), ("float16", "float32"),
                        ("int32", "float64"), ("uint32", "float64"), ("int64", "float64")):
        arr = as_supported_frames(np.arange(12, dtype=dtype).reshape(3, 4)[:, ::2])
        assert arr.dtype.name == want and arr.flags.c_contiguous
        # the same type numpy computes in when such an image meets a float32 flat
        if dtype != "uint16":
            assert (np.zeros(1, dtype=dtype) / np.ones(1, dtype=np.float32)).dtype == arr.dtype

"""Golden vectors of the flat-field corrector, produced by the REFERENCE's own class.

Run in the build container only (it reads /root/reference; the fixture it writes travels):

    python tests/golden/make_golden_flatfield.py

`emirdrp/processing/flatfield.py` runs unmodified.  Its base class `numina.processing.Corrector`
is absent (numina is not installed): a shim supplies what FlatFieldCorrector uses from it --
the constructor's attributes (`datamodel`, `calibid`, `dtype`) and `get_imgid` -- and a
datamodel whose `get_data(img)` returns the primary array.  astropy's HDUList is replaced by
pyemir_b200.hdu.  The arithmetic (`flatdata[flatdata <= 0] = 1.0`, `data / flatdata`,
`.astype(dtype)`) is plain numpy inside the reference file, so these vectors PIN the stage.
"""

import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/emirdrp"
sys.path.insert(0, ROOT)

from pyemir_b200 import hdu  # noqa: E402

SHAPE = (96, 128)
RESULTS = {}
# (name, image dtype, flat dtype)
CASES = [("f32_f32", "float32", "float32"), ("f32_f64", "float32", "float64"),
         ("u16_f32", "uint16", "float32"), ("f64_f32", "float64", "float32")]


def make_inputs(name, img_dtype, flat_dtype):
    """Seeded image and flat of one case (shared with the tests)."""
    rng = np.random.default_rng(sum(map(ord, name)))
    if img_dtype == "uint16":
        data = rng.integers(0, 65535, size=SHAPE).astype(np.uint16)
    else:
        data = rng.normal(1000.0, 300.0, size=SHAPE).astype(img_dtype)
        data[5, 7] = np.nan
        data[5, 8] = np.inf
    flat = rng.normal(1.0, 0.2, size=SHAPE).astype(flat_dtype)
    flat[0, :10] = 0.0                  # "To avoid NaN": become 1.0
    flat[1, :10] = -0.5
    if name == "f64_f32":
        flat[2, 3] = np.nan             # stays NaN (nan <= 0 is False)
        flat[2, 4] = np.inf
    return data, flat


def sha_canonical(arr):
    """sha256 of the array with every NaN replaced by one bit pattern (numpy keeps the payload
    of an input NaN, CUDA arithmetic returns its canonical NaN: both are "NaN")."""
    arr = np.ascontiguousarray(arr).copy()
    if arr.dtype.kind == "f":
        arr[np.isnan(arr)] = np.nan
    return hashlib.sha256(arr.tobytes()).hexdigest()


class _Corrector:
    """What FlatFieldCorrector needs from numina.processing.Corrector."""

    def __init__(self, datamodel=None, calibid="calibid-unknown", dtype="float32"):
        self.datamodel = datamodel
        self.calibid = calibid
        self.dtype = dtype

    def get_imgid(self, img):
        return "imgid-golden"


class _DataModel:
    def get_data(self, img):
        return img["primary"].data


def load_reference_class():
    numina = types.ModuleType("numina")
    numina.__path__ = []
    proc = types.ModuleType("numina.processing")
    proc.Corrector = _Corrector
    numina.processing = proc
    sys.modules["numina"] = numina
    sys.modules["numina.processing"] = proc
    spec = importlib.util.spec_from_file_location(
        "ref_flatfield", os.path.join(REF, "processing", "flatfield.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.FlatFieldCorrector


def main():
    cls = load_reference_class()
    out = {"shape": list(SHAPE), "cases": {}}
    for name, img_dtype, flat_dtype in CASES:
        data, flat = make_inputs(name, img_dtype, flat_dtype)
        flat_in = flat.copy()
        node = cls(flat_in, datamodel=_DataModel(), calibid="uuid:golden-flat")
        img = hdu.HDUList([hdu.PrimaryHDU(data.copy(), hdu.Header())])
        res = node.run(img)
        arr = res["primary"].data
        cards = [[c[0], c[1]] for c in res["primary"].header.cards]
        RESULTS[name] = arr
        out["cases"][name] = {
            "img_dtype": img_dtype, "flat_dtype": flat_dtype,
            "result_dtype": arr.dtype.name,
            "sha256": sha_canonical(arr),
            "flat_fixed_sha256": hashlib.sha256(flat_in.tobytes()).hexdigest(),
            "flat_stats": float(node.flat_stats),
            "sample": [float(v) if np.isfinite(v) else repr(float(v))
                       for v in arr[::17, ::23].ravel()[:12]],
            "cards": [c for c in cards if "correction time" not in str(c[1])],
        }
    np.savez_compressed(os.path.join(HERE, "flatfield_results.npz"), **RESULTS)
    with open(os.path.join(HERE, "flatfield.json"), "w") as fd:
        json.dump(out, fd, indent=1)
    print("wrote", os.path.join(HERE, "flatfield.json"))


if __name__ == "__main__":
    main()

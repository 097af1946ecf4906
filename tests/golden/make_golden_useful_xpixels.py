"""Golden vectors of useful_mos_xpixels / get_islitlet, produced by the REFERENCE's own code.

Run in the build container only (it reads /root/reference; the fixture it writes travels):

    python tests/golden/make_golden_useful_xpixels.py

`emirdrp/processing/wavecal/useful_mos_xpixels.py` and `get_islitlet.py` run unmodified; the
two numina display modules they import (never reached with debugplot=0) are stubs, and
`pkgutil.get_data` is answered with the small synthetic OH catalogue of `oh_catalogue()` below
instead of the file shipped with emirdrp, so that the tests need nothing from the reference.
"""

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

NAXIS1, NAXIS2 = 3400, 2090
WCS = dict(naxis1=NAXIS1, naxis2=NAXIS2, crpix1=1.0, crval1=11200.0, cdelt1=0.77)


def make_header(kind="normal"):
    """Header dict with the keywords the function reads (lower case like the reference's
    accesses; a plain dict is what the function needs)."""
    hdr = dict(WCS)
    for islitlet in range(1, 56):
        jmn = 300 + 11 * islitlet
        jmx = jmn + 2100
        if kind == "empty_first" and islitlet == 1:
            jmn, jmx = 0, 0
        hdr["jmnslt{:02d}".format(islitlet)] = jmn
        hdr["jmxslt{:02d}".format(islitlet)] = jmx
    return hdr


def oh_catalogue():
    """Two columns of wavelengths inside and outside the grid (format of Oliva et al. 2013:
    the function uses columns 0 and 1)."""
    rng = np.random.default_rng(2013)
    first = np.sort(rng.uniform(11000.0, 14000.0, size=40))
    return np.column_stack([first, first + rng.uniform(0.5, 3.0, size=40), np.ones(40)])


# (name, header kind, vpix_region, npix_removed_near_ohlines, list_valid_wvregions)
CASES = [
    ("one_slitlet", "normal", (5, 30), 0, None),
    ("rounding", "normal", (38.4, 38.6), 0, None),
    ("three_slitlets", "normal", (30, 100), 0, None),
    ("last_rows", "normal", (2053, 2090), 0, None),
    ("wvregions", "normal", (40, 70), 0, [(11500.0, 11800.5), (12500.0, 12600.0), (5000, 11210)]),
    ("wvregions_strings", "normal", (40, 70), 0, [("12000", "12100.5")]),
    ("oh_removed", "normal", (400, 430), 3, None),
    ("oh_and_regions", "normal", (400, 430), 2, [(11400.0, 13500.0)]),
    ("empty_first", "empty_first", (1, 38), 0, None),
    ("err_order", "normal", (30, 10), 0, None),
    ("err_range_low", "normal", (0, 10), 0, None),
    ("err_range_high", "normal", (10, 2091), 0, None),
    ("err_wvregion_order", "normal", (10, 20), 0, [(12000.0, 11000.0)]),
    ("err_no_region", "normal", (10, 20), 0, [(20000.0, 21000.0)]),
    ("err_all_removed", "normal", (10, 20), 4000, None),
]


def load_reference():
    def pkg(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    for name in ("numina", "numina.array", "numina.array.display", "emirdrp",
                 "emirdrp.processing", "emirdrp.processing.wavecal"):
        pkg(name)
    x = types.ModuleType("numina.array.display.ximshow")
    x.ximshow = lambda *a, **k: None
    sys.modules[x.__name__] = x
    y = types.ModuleType("numina.array.display.pause_debugplot")
    y.pause_debugplot = lambda *a, **k: None
    sys.modules[y.__name__] = y

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    load("emirdrp.core", os.path.join(REF, "core", "__init__.py"))
    gi = load("emirdrp.processing.wavecal.get_islitlet",
              os.path.join(REF, "processing", "wavecal", "get_islitlet.py"))
    um = load("emirdrp.processing.wavecal.useful_mos_xpixels",
              os.path.join(REF, "processing", "wavecal", "useful_mos_xpixels.py"))
    text = "\n".join(" ".join("%.6f" % v for v in row) for row in oh_catalogue())
    um.pkgutil = types.SimpleNamespace(get_data=lambda package, name: text.encode("utf8"))
    return gi.get_islitlet, um.useful_mos_xpixels


def runs_of(mask):
    """1-based inclusive runs of True: compact and readable in the fixture."""
    idx = np.flatnonzero(mask)
    if idx.size == 0:
        return []
    cuts = np.flatnonzero(np.diff(idx) > 1)
    starts = np.concatenate(([idx[0]], idx[cuts + 1]))
    stops = np.concatenate((idx[cuts], [idx[-1]]))
    return [[int(a) + 1, int(b) + 1] for a, b in zip(starts, stops)]


def main():
    get_islitlet, useful = load_reference()
    out = {"get_islitlet": {}, "cases": {}}
    for ipixel in (1, 2, 37, 38, 39, 76, 77, 1000, 2052, 2053, 2089, 2090, 38.5, 0, -3, 2091):
        try:
            out["get_islitlet"][repr(ipixel)] = get_islitlet(ipixel)
        except ValueError as exc:
            out["get_islitlet"][repr(ipixel)] = {"ValueError": str(exc)}
    data = np.zeros((NAXIS2, NAXIS1), dtype=np.float32)
    for name, kind, vpix, nrem, regions in CASES:
        try:
            mask = useful(data, make_header(kind), vpix_region=vpix,
                          npix_removed_near_ohlines=nrem, list_valid_wvregions=regions,
                          debugplot=0)
            out["cases"][name] = {"dtype": mask.dtype.name, "shape": list(mask.shape),
                                  "count": int(mask.sum()), "runs": runs_of(mask)}
        except (ValueError, IndexError) as exc:
            out["cases"][name] = {type(exc).__name__: str(exc)}
    with open(os.path.join(HERE, "useful_mos_xpixels.json"), "w") as fd:
        json.dump(out, fd, indent=1)
    print("wrote", os.path.join(HERE, "useful_mos_xpixels.json"))


if __name__ == "__main__":
    main()

"""Prototype (numpy, CPU) of a PIXEL-CENTRIC border phase for k_rectwv -- next-round design aid.

Today a lane owns one output border and fetches the knots (flux, slopes) of the source pixel the
border falls in from shared memory.  Pixel-centric: the lane that owns source pixel i evaluates
the output borders that fall INSIDE pixel i from its own registers; only the partial flux P of
the first border of the next non-empty pixel (and the flux of one empty pixel in between) moves
between lanes.  This script (1) states that evaluation in numpy and checks it against the
border-centric statement in tests/proto_tables.py bit for bit, and (2) prints, for the BASELINE
configs, how many borders fall into a pixel and how often the simple two-slot scheme applies.

usage: python scratch/pixel_centric_proto.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import proto_tables as proto  # noqa: E402
from pyemir_b200 import synthetic  # noqa: E402
from pyemir_b200.set_wv_parameters import set_wv_parameters  # noqa: E402


def live_range(tb, naxis1):
    """Borders [qa, qb]: qa = last border extrapolated below the spectrum (or 0), qb = first
    border extrapolated above it (or naxis1).  Outputs outside [qa, qb) are exact zeros."""
    idx, t = tb["idx"], tb["t"]
    h = 1.0 / tb["invh"]
    below = (idx == 0) & (t == 0.0)
    above = (idx == tb["nchan"] - 1) & (t == h[-1])
    qa = int(np.nonzero(below)[0].max()) if below.any() else 0
    qb = int(np.nonzero(above)[0].min()) if above.any() else naxis1
    return qa, qb


def rebin_row_pixel_centric(rect, tb, naxis1, max_slots=2):
    """Same output as proto.rebin_row, organised by source pixel.

    Returns (row, fallback) -- fallback is True when some pixel holds more than max_slots
    borders or two empty pixels follow each other (the border-centric path must take over)."""
    dtype = rect.dtype
    s, yp = proto.steffen_knot_slopes(rect, tb)
    idx = tb["idx"]
    t = tb["t"].astype(dtype)
    tau = tb["tau"].astype(dtype)
    nchan = tb["nchan"]
    qa, qb = live_range(tb, naxis1)
    out = np.zeros(naxis1, dtype=dtype)
    # borders of every pixel inside the live range
    first = np.full(nchan, -1)
    count = np.zeros(nchan, dtype=int)
    for q in range(qa, qb + 1):
        i = idx[q]
        if count[i] == 0:
            first[i] = q
        count[i] += 1
    if count.max() > max_slots:
        return None, True

    def partial(i, q):                      # P of border q inside pixel i: own registers only
        yp0, yp1, si = yp[i], yp[i + 1], s[i]
        A = yp0 + yp1 - 2 * si
        B = si - yp0 - A
        return t[q] * (yp0 + tau[q] * (B + tau[q] * A))

    pfirst = np.zeros(nchan, dtype=dtype)
    for i in range(nchan):
        if count[i]:
            pfirst[i] = partial(i, first[i])
    for i in range(nchan):
        for slot in range(count[i]):
            q = first[i] + slot
            if q >= qb:
                continue                    # the last border only closes the previous output
            p = partial(i, q)
            if slot + 1 < count[i]:         # next border in the same pixel
                out[q] = dtype.type(0) + (partial(i, q + 1) - p)
            else:
                nxt = i + 1
                whole = rect[i]
                if count[nxt] == 0:         # one empty pixel in between
                    whole = whole + rect[nxt]
                    nxt += 1
                    if count[nxt] == 0:
                        return None, True
                out[q] = whole + (pfirst[nxt] - p)
    return out, False


def main():
    rng = np.random.default_rng(0)
    for number in (2, 3, 4, "4b"):
        coeff = synthetic.make_config(number)
        p = set_wv_parameters(coeff.tags["filter"], coeff.tags["grism"])
        grid = dict(naxis1=p["naxis1_enlarged"], cdelt1=p["cdelt1_enlarged"],
                    crval1=p["crval1_enlarged"], crpix1=p["crpix1_enlarged"])
        hist = np.zeros(5, dtype=np.int64)
        fallbacks = checked = 0
        for islitlet in (1, 14, 28, 41, 55):
            entry = coeff.contents[islitlet - 1]
            tb = proto.rebin_tables(entry, **grid)
            qa, qb = live_range(tb, grid["naxis1"])
            cnt = np.bincount(tb["idx"][qa:qb + 1], minlength=tb["nchan"])
            inside = cnt[tb["idx"][qa]:tb["idx"][qb] + 1]
            hist += np.bincount(np.minimum(inside, 4), minlength=5)
            for dtype in (np.float64, np.float32):
                rect = rng.normal(100.0, 30.0, size=tb["nchan"]).astype(dtype)
                want = proto.rebin_row(rect, tb)
                got, fb = rebin_row_pixel_centric(rect, tb, grid["naxis1"])
                checked += 1
                if fb:
                    fallbacks += 1
                else:
                    assert np.array_equal(got, want), (number, islitlet, dtype)
        frac = hist / hist.sum()
        print(f"config {number}: borders per source pixel inside the coverage "
              f"0: {frac[0]:.3f}  1: {frac[1]:.3f}  2: {frac[2]:.3f}  3: {frac[3]:.3f}  4+: {frac[4]:.3f};"
              f"  rows checked {checked}, two-slot scheme not applicable {fallbacks}")


if __name__ == "__main__":
    main()

"""Numpy prototype of the plan tables used by the CUDA kernels (TEST HELPER ONLY).

The CUDA library turns a ``RectWaveCoeff`` into frame-independent tables once (the
"plan") and then streams frames through them.  This file states the
same tables and the same per-row evaluation in a few lines of numpy so that

* the formulation itself (telescoped cumulative-flux difference, border and
  coverage conventions, ghost row for JMNSLT/JMXSLT) is checked against the oracle
  on the CPU, before any GPU time is spent, and
* GPU tests can diff the library's exported tables against an independent
  statement of what they must contain.

It is not an implementation of the product path and nothing under
``pyemir_b200/`` imports it.
"""

import ctypes

import numpy as np
from numpy.polynomial.polynomial import polyval

from oracle import numina_restated as nr
from oracle.build import load_clippix

NROWS = 39          # 38 pasted rows + the extra row scanned for JMNSLT/JMXSLT


def slitlet_geometry(entry, offx=0, offy=0):
    return dict(
        ns1=entry["bb_ns1_orig"], nc1=entry["bb_nc1_orig"],
        bbh=entry["bb_ns2_orig"] - entry["bb_ns1_orig"] + 1,
        bbw=entry["bb_nc2_orig"] - entry["bb_nc1_orig"] + 1,
        min_row=entry["min_row_rectified"], offx=offx, offy=offy)


def rectify_tables(entry, mode, offx=0, offy=0, nmax_order=5, naxis=2048):
    """Return (base_row, base_col, weights[K,K], K) per (row y, column j).

    base_* are FRAME coordinates of tap (0,0); weights are 0 for taps outside the
    slitlet bounding box or outside the frame.
    """
    g = slitlet_geometry(entry, offx, offy)
    order = nr.order_fmap(len(entry["ttd_aij"]), nmax_order)
    aij = [float(c) for c in entry["ttd_aij"]]
    bij = [float(c) for c in entry["ttd_bij"]]
    rows = g["min_row"] + np.arange(NROWS)
    jj, ii = np.meshgrid(np.arange(g["bbw"], dtype=float), rows.astype(float))
    if mode == 1:
        u, v = nr.fmap(order, aij, bij, jj.ravel(), ii.ravel())
        bc = np.rint(u).astype(int).reshape(jj.shape)
        br = np.rint(v).astype(int).reshape(jj.shape)
        K = 1
        w = np.ones(jj.shape + (1, 1))
    elif mode == 3:
        u, v = nr.fmap(order, aij, bij, jj.ravel(), ii.ravel())
        u = u.reshape(jj.shape)
        v = v.reshape(jj.shape)
        bc = np.floor(u).astype(int)
        br = np.floor(v).astype(int)
        fu = u - bc
        fv = v - br
        K = 2
        w = np.empty(jj.shape + (2, 2))
        w[..., 0, 0] = (1 - fv) * (1 - fu)
        w[..., 0, 1] = (1 - fv) * fu
        w[..., 1, 0] = fv * (1 - fu)
        w[..., 1, 1] = fv * fu
    else:
        dj = np.array([-0.5, -0.5, 0.5, 0.5])
        di = np.array([-0.5, 0.5, 0.5, -0.5])
        xx = (jj[..., None] + dj).ravel()
        yy = (ii[..., None] + di).ravel()
        u, v = nr.fmap(order, aij, bij, xx, yy)
        u = u.reshape(jj.shape + (4,))
        v = v.reshape(jj.shape + (4,))
        bc = np.floor(u.min(-1) + 0.5).astype(int)
        br = np.floor(v.min(-1) + 0.5).astype(int)
        hc = np.maximum(np.ceil(u.max(-1) - 0.5).astype(int), bc)
        hr = np.maximum(np.ceil(v.max(-1) - 0.5).astype(int), br)
        K = int(max((hc - bc).max(), (hr - br).max())) + 1
        w = np.zeros(jj.shape + (K, K))
        lib = load_clippix()
        dp = ctypes.POINTER(ctypes.c_double)
        buf = np.zeros(K * K)
        for y in range(jj.shape[0]):
            for j in range(jj.shape[1]):
                qx = np.ascontiguousarray(u[y, j])
                qy = np.ascontiguousarray(v[y, j])
                lib.oracle_quad_weights(qx.ctypes.data_as(dp), qy.ctypes.data_as(dp),
                                        int(br[y, j]), int(bc[y, j]), K, K,
                                        buf.ctypes.data_as(dp))
                w[y, j] = buf.reshape(K, K)
    # zero the taps outside the bounding box or outside the (shifted) frame
    for a in range(K):
        for b in range(K):
            r = br + a
            c = bc + b
            inside = (r >= 0) & (r < g["bbh"]) & (c >= 0) & (c < g["bbw"])
            fr = r + g["ns1"] - 1 - offy
            fc = c + g["nc1"] - 1 - offx
            inside &= (fr >= 0) & (fr < naxis) & (fc >= 0) & (fc < naxis)
            w[..., a, b] = np.where(inside, w[..., a, b], 0.0)
    base_row = br + g["ns1"] - 1 - offy
    base_col = bc + g["nc1"] - 1 - offx
    return base_row, base_col, w, K


def rebin_tables(entry, naxis1, cdelt1, crval1, crpix1=1.0):
    """Per border k: (idx, t, tau); per knot interval i: invh; per knot i: rr."""
    nchan = entry["bb_nc2_orig"] - entry["bb_nc1_orig"] + 1
    old_x_borders = np.arange(-0.5, nchan)
    old_x_borders += crpix1
    xw = polyval(old_x_borders, entry["wpoly_coeff"])
    new_wl = crval1 + cdelt1 * np.arange(naxis1)
    nb = nr.map_borders(new_wl)
    h = xw[1:] - xw[:-1]
    assert np.all(h > 0)
    below = nb < xw[0]
    above = nb > xw[-1]
    idx = np.searchsorted(xw, nb).clip(1, nchan) - 1
    t = nb - xw[idx]
    idx = np.where(below, 0, idx)
    t = np.where(below, 0.0, t)
    idx = np.where(above, nchan - 1, idx)
    t = np.where(above, h[nchan - 1], t)
    invh = 1.0 / h
    tau = t * invh[idx]
    rr = np.zeros(nchan + 1)
    rr[1:nchan] = h[1:] / (h[:-1] + h[1:])
    return dict(idx=idx.astype(np.int32), t=t, tau=tau, invh=invh, rr=rr,
                nchan=nchan)


def apply_rectify(frame, base_row, base_col, w, dtype=np.float64):
    """rect[y, j] = sum_ab w[y,j,a,b] * frame[base_row+a, base_col+b] (0 weight => skip)."""
    K = w.shape[-1]
    n = frame.shape[0]
    rect = np.zeros(base_row.shape, dtype=dtype)
    wt = w.astype(dtype)
    fr = frame.astype(dtype)
    for a in range(K):
        for b in range(K):
            r = np.clip(base_row + a, 0, n - 1)
            c = np.clip(base_col + b, 0, n - 1)
            rect += wt[..., a, b] * fr[r, c]
    return rect


def steffen_knot_slopes(rect, tb):
    """s (nchan) and yp (nchan+1) of one row; pure elementwise neighbour math."""
    dtype = rect.dtype
    invh = tb["invh"].astype(dtype)
    rr = tb["rr"].astype(dtype)
    nchan = tb["nchan"]
    s = rect * invh
    yp = np.zeros(nchan + 1, dtype=dtype)
    sl = s[:-1]
    sr = s[1:]
    p = sl * rr[1:nchan] + sr * (1 - rr[1:nchan])
    yp[1:nchan] = (np.sign(sl) + np.sign(sr)) * np.minimum(
        np.minimum(np.abs(sl), np.abs(sr)), dtype.type(0.5) * np.abs(p))
    return s, yp


def rebin_row(rect, tb):
    """Telescoped cumulative-flux difference: no prefix sum, no cancellation."""
    dtype = rect.dtype
    s, yp = steffen_knot_slopes(rect, tb)
    idx = tb["idx"]
    t = tb["t"].astype(dtype)
    tau = tb["tau"].astype(dtype)
    yp0 = yp[idx]
    yp1 = yp[idx + 1]
    si = s[idx]
    A = yp0 + yp1 - 2 * si
    B = si - yp0 - A
    P = t * (yp0 + tau * (B + tau * A))
    out = P[1:] - P[:-1]
    span = idx[1:] - idx[:-1]
    whole = np.zeros_like(out)
    for step in range(int(span.max()) if span.size else 0):
        take = span > step
        whole[take] += rect[idx[:-1][take] + step]
    return whole + out


def run_slitlet(frame, entry, grid, mode, offx=0, offy=0, dtype=np.float64):
    """Rows (39, naxis1) of one slitlet: 38 output rows + the ghost row."""
    br, bc, w, K = rectify_tables(entry, mode, offx, offy)
    tb = rebin_tables(entry, **grid)
    rect = apply_rectify(frame, br, bc, w, dtype)
    rows = np.stack([rebin_row(rect[y], tb) for y in range(NROWS)])
    return rows, rect, K


def jminmax(rows):
    """1-based first/last non-zero channel over the 39 rows, (0, 0) when none."""
    nz = np.nonzero(np.any(rows != 0, axis=0))[0]
    if nz.size == 0:
        return 0, 0
    return int(nz[0]) + 1, int(nz[-1]) + 1

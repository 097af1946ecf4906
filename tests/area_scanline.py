"""Area of (convex quadrilateral ∩ axis-aligned box) by exact scan-line integration.

An independent statement of the quantity numina's ``_resample`` computes for
``rectify2d(..., resampling=2)`` (slitlet2d.py:413-423): it shares no code and no idea
with the Sutherland-Hodgman clippers of ``oracle/clippix.c`` and
``pyemir_b200/csrc/geom.cuh``.  No polygon is ever clipped here:

    area = ∫_{x0}^{x1} max(0, min(hi(x), y1) - max(lo(x), y0)) dx

where [lo(x), hi(x)] is the chord a vertical line at x cuts out of the convex polygon.
The integrand is piecewise linear in x; its breakpoints are the vertices' abscissae and
the abscissae where an edge crosses y = y0 or y = y1, so the midpoint rule between
consecutive breakpoints is exact (up to rounding, ~1e-15).
"""

import numpy as np


def _chord(px, py, x):
    """(lo, hi) of the polygon's intersection with the vertical line at x, or None."""
    lo, hi = np.inf, -np.inf
    n = len(px)
    for i in range(n):
        xa, ya, xb, yb = px[i], py[i], px[(i + 1) % n], py[(i + 1) % n]
        if xa == xb:
            continue                      # vertical edge: the neighbours give the chord
        if (xa - x) * (xb - x) > 0.0:
            continue                      # the edge lies on one side of the line
        y = ya + (yb - ya) * ((x - xa) / (xb - xa))
        lo = min(lo, y)
        hi = max(hi, y)
    return (lo, hi) if hi >= lo else None


def quad_box_area(px, py, x0, x1, y0, y1):
    """Area of the convex polygon (px, py) inside [x0, x1] x [y0, y1]."""
    px = [float(v) for v in px]
    py = [float(v) for v in py]
    cuts = {float(x0), float(x1)}
    n = len(px)
    for i in range(n):
        xa, ya, xb, yb = px[i], py[i], px[(i + 1) % n], py[(i + 1) % n]
        if x0 < xa < x1:
            cuts.add(xa)
        for level in (y0, y1):
            if ya != yb and (ya - level) * (yb - level) < 0.0:
                xc = xa + (xb - xa) * ((level - ya) / (yb - ya))
                if x0 < xc < x1:
                    cuts.add(xc)
    cuts = sorted(cuts)
    area = 0.0
    for a, b in zip(cuts[:-1], cuts[1:]):
        chord = _chord(px, py, 0.5 * (a + b))
        if chord is None:
            continue
        height = min(chord[1], y1) - max(chord[0], y0)
        if height > 0.0:
            area += height * (b - a)
    return area


def shoelace(px, py):
    """|area| of a simple polygon."""
    px = np.asarray(px, dtype=float)
    py = np.asarray(py, dtype=float)
    return 0.5 * abs(float(np.dot(px, np.roll(py, -1)) - np.dot(py, np.roll(px, -1))))


def quad_weights(px, py, i0, j0, ny, nx):
    """Overlap areas with the (ny, nx) block of unit pixels whose first one is (i0, j0);
    pixel (i, j) covers [j-.5, j+.5] x [i-.5, i+.5] (numina's convention)."""
    w = np.zeros((ny, nx))
    for a in range(ny):
        for b in range(nx):
            w[a, b] = quad_box_area(px, py, j0 + b - 0.5, j0 + b + 0.5,
                                    i0 + a - 0.5, i0 + a + 0.5)
    return w

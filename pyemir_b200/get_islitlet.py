"""Slitlet number of a row of the rectified image (``processing/wavecal/get_islitlet.py:16-47``)."""

from .core import EMIR_NBARS, EMIR_NPIXPERSLIT_RECTIFIED


def get_islitlet(ipixel):
    """Return islitlet (1 .. EMIR_NBARS) from a 1-based Y-pixel coordinate of the rectified image.

    Same results and error messages as the reference: rows 1..38 belong to slitlet 1, 39..76
    to slitlet 2, and so on (``iminslt = (i - 1) * 38 + 1``, ``slitlet2d.py:198-199``).
    """
    nrows = EMIR_NPIXPERSLIT_RECTIFIED * EMIR_NBARS
    if ipixel < 1:
        raise ValueError("ipixel={} cannot be < 1".format(ipixel))
    if ipixel > nrows:
        raise ValueError("ipixel={} cannot be > {}".format(ipixel, nrows))
    # int(x / 38) + 1, one less on the last row of a slitlet == ceil(x / 38) for integers;
    # the reference's float arithmetic is kept for non-integer arguments
    islitlet = int(ipixel / EMIR_NPIXPERSLIT_RECTIFIED) + 1
    if ipixel % EMIR_NPIXPERSLIT_RECTIFIED == 0:
        islitlet -= 1
    return islitlet

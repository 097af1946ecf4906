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
    # the reference counts int(x / 38) + 1 and takes one off on the last row of a slitlet:
    # that is the ceiling of x / 38, written here as a floor division (exact for integers,
    # same result for the fractional arguments the recipes can pass)
    return int(-(-ipixel // EMIR_NPIXPERSLIT_RECTIFIED))

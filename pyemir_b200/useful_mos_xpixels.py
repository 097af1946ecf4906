"""Useful X-axis pixels of a rectified, wavelength-calibrated MOS image (host mirror).

``emirdrp/processing/wavecal/useful_mos_xpixels.py:22-144``: the consumer of the
``JMNSLTnn`` / ``JMXSLTnn`` keywords that ``apply_rectwv_coeff`` writes, called by the ABBA
recipe right behind the path (``recipes/spec/abba.py:423-440``).  Pure bookkeeping on the
header: no pixel arithmetic, so it stays on the host.  Pinned by golden vectors produced with
the reference's own function (``tests/golden/make_golden_useful_xpixels.py``).
"""

import numpy as np

from .get_islitlet import get_islitlet


def _oh_catalogue(oh_catalogue):
    """Two-column table of OH line wavelengths (the reference ships Oliva et al. 2013)."""
    if oh_catalogue is None:
        import pkgutil
        from io import StringIO
        try:
            raw = pkgutil.get_data("emirdrp.instrument.configs", "Oliva_etal_2013.dat")
        except (ImportError, OSError):
            raw = None
        if raw is None:
            raise ValueError("an OH line catalogue is needed (oh_catalogue=...): "
                             "emirdrp.instrument.configs/Oliva_etal_2013.dat is not installed")
        return np.genfromtxt(StringIO(raw.decode("utf8")))
    if isinstance(oh_catalogue, (str, bytes)):
        return np.genfromtxt(oh_catalogue)
    return np.asarray(oh_catalogue, dtype=float)


def _pixel_of(wave, crval1, cdelt1, crpix1):
    # int() truncates towards zero, like the reference's int((w - crval1) / cdelt1 + crpix1 + 0.5)
    return int((wave - crval1) / cdelt1 + crpix1 + 0.5)


def useful_mos_xpixels(reduced_mos_data, base_header, vpix_region, npix_removed_near_ohlines=0,
                       list_valid_wvregions=None, debugplot=0, *, oh_catalogue=None):
    """Boolean mask (naxis1,) of the useful pixels in the wavelength direction.

    Same arguments, result and errors as the reference (``debugplot`` is accepted and ignored;
    ``reduced_mos_data`` is only displayed there).  ``oh_catalogue`` (extension): the OH line
    table as an array or a file name instead of the file shipped with emirdrp.

    The reference widens the column range over the slitlets of ``vpix_region`` with
    ``JMNSLTnn`` for BOTH limits (``useful_mos_xpixels.py:53-57`` reads ``jmnslt`` twice); that
    behaviour is reproduced, not corrected.
    """
    naxis1 = base_header["naxis1"]
    naxis2 = base_header["naxis2"]
    crpix1 = base_header["crpix1"]
    crval1 = base_header["crval1"]
    cdelt1 = base_header["cdelt1"]

    nsmin = int(vpix_region[0] + 0.5)
    nsmax = int(vpix_region[1] + 0.5)
    if nsmin > nsmax:
        raise ValueError("vpix_region values in wrong order")
    if nsmin < 1 or nsmax > naxis2:
        raise ValueError("vpix_region outside valid range")

    islitlet_min = get_islitlet(nsmin)
    islitlet_max = get_islitlet(nsmax)
    ncmin = base_header["jmnslt{:02d}".format(islitlet_min)]
    ncmax = base_header["jmxslt{:02d}".format(islitlet_min)]
    for islitlet in range(islitlet_min + 1, islitlet_max + 1):
        first = base_header["jmnslt{:02d}".format(islitlet)]
        ncmin = min(ncmin, first)
        ncmax = max(ncmax, first)            # sic: the reference uses JMNSLT here as well

    pixels = np.arange(1, naxis1 + 1)
    if list_valid_wvregions is None:
        # the reference indexes xisok[ipix - 1] for ipix in ncmin..ncmax: a limit of 0 (no
        # flux found, JMNSLT = 0) wraps to the last pixel, a limit beyond naxis1 raises
        if ncmax > naxis1 or (ncmin <= ncmax and ncmin < 1 - naxis1):
            raise IndexError("index {} is out of bounds for axis 0 with size {}".format(
                (ncmax if ncmax > naxis1 else ncmin) - 1, naxis1))
        xisok_wvreg = (pixels >= ncmin) & (pixels <= ncmax)
        if ncmin <= ncmax and ncmin < 1:
            wrapped = np.arange(ncmin, min(ncmax, 0) + 1) - 1          # negative indices
            xisok_wvreg[wrapped] = True
    else:
        xisok_wvreg = np.zeros(naxis1, dtype=bool)
        for wvregion in list_valid_wvregions:
            wvmin = float(wvregion[0])
            wvmax = float(wvregion[1])
            if wvmin > wvmax:
                raise ValueError("wvregion values in wrong order: {}, {}".format(wvmin, wvmax))
            minpix = _pixel_of(wvmin, crval1, cdelt1, crpix1)
            maxpix = _pixel_of(wvmax, crval1, cdelt1, crpix1)
            xisok_wvreg |= (pixels >= minpix) & (pixels <= maxpix)
        if np.sum(xisok_wvreg) < 1:
            raise ValueError("no valid wavelength ranges provided")

    xisok_oh = np.ones(naxis1, dtype=bool)
    nrem = int(npix_removed_near_ohlines)
    if nrem > 0:
        catlines = _oh_catalogue(oh_catalogue)
        for waveline in np.concatenate((catlines[:, 1], catlines[:, 0])):
            expected_pixel = _pixel_of(waveline, crval1, cdelt1, crpix1)
            xisok_oh &= ~((pixels >= expected_pixel - nrem) & (pixels <= expected_pixel + nrem))

    xisok = np.logical_and(xisok_wvreg, xisok_oh)
    if np.sum(xisok) < 1:
        raise ValueError("no valid wavelength range available after removing OH lines")
    return xisok

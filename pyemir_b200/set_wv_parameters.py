"""Output wavelength grid for each grism + filter combination.

Mirrors ``emirdrp/processing/wavecal/set_wv_parameters.py:19-154`` (same
function name, argument order, returned keys and error behaviour).  The table
is data, kept as one dict per combination instead of an if/elif chain.
"""

import numpy as np

from .core import EMIR_NAXIS1_ENLARGED, EMIR_VALID_FILTERS, EMIR_VALID_GRISMS

# (grism, filter) -> parameters; numbers from set_wv_parameters.py:64-148
_WV_TABLE = {
    ("J", "J"): dict(
        islitlet_min=2, islitlet_max=54, nbrightlines=[18],
        crval1_linear=[1.25122231e04, -4.83412417e00, 5.31657015e-04],
        cdelt1_linear=[7.73411692e-01, -3.28055653e-05, -1.20813896e-08],
        crval1_enlarged=11200.0, cdelt1_enlarged=0.77,
        naxis1_enlarged=EMIR_NAXIS1_ENLARGED,
        wvmin_expected=None, wvmax_expected=None,
        wvmin_useful=None, wvmax_useful=None),
    ("H", "H"): dict(
        islitlet_min=2, islitlet_max=54, nbrightlines=[12],
        crval1_linear=[1.65536274e04, -7.63517173e00, 7.74790265e-04],
        cdelt1_linear=[1.21327515e00, 1.42140078e-05, -1.27489119e-07],
        crval1_enlarged=14500.0, cdelt1_enlarged=1.22,
        naxis1_enlarged=EMIR_NAXIS1_ENLARGED,
        wvmin_expected=None, wvmax_expected=None,
        wvmin_useful=None, wvmax_useful=None),
    ("K", "Ksp"): dict(
        islitlet_min=2, islitlet_max=54, nbrightlines=[12],
        crval1_linear=[2.21044741e04, -1.08737529e01, 9.05081653e-04],
        cdelt1_linear=[1.72696857e00, 2.35009351e-05, -1.02164228e-07],
        crval1_enlarged=19100.0, cdelt1_enlarged=1.73,
        naxis1_enlarged=EMIR_NAXIS1_ENLARGED,
        wvmin_expected=None, wvmax_expected=None,
        wvmin_useful=None, wvmax_useful=None),
    ("LR", "YJ"): dict(
        islitlet_min=4, islitlet_max=55, nbrightlines=[20],
        crval1_linear=[1.04272465e04, -2.33176855e01, 6.55101267e-03],
        cdelt1_linear=[3.49037727e00, 1.26008332e-03, -4.66149678e-06],
        crval1_enlarged=8900.0, cdelt1_enlarged=3.56,
        naxis1_enlarged=1270,
        wvmin_expected=4000, wvmax_expected=15000,
        wvmin_useful=8900, wvmax_useful=13400),
    ("LR", "HK"): dict(
        islitlet_min=4, islitlet_max=55, nbrightlines=[25],
        crval1_linear=[2.00704978e04, -4.07702886e01, -5.95247468e-03],
        cdelt1_linear=[6.54247758e00, 2.09061196e-03, -2.48206609e-06],
        crval1_enlarged=14500.0, cdelt1_enlarged=6.83,
        naxis1_enlarged=1435,
        wvmin_expected=8000, wvmax_expected=30000,
        wvmin_useful=14500, wvmax_useful=24300),
}


def set_wv_parameters(filter_name, grism_name):
    """Return the wavelength-calibration parameters of the rectified image.

    Same keys as the reference: ``crpix1_enlarged`` (always 1.0,
    set_wv_parameters.py:63), ``crval1_enlarged``, ``cdelt1_enlarged``,
    ``naxis1_enlarged``, ``islitlet_min/max``, ``nbrightlines``,
    ``poly_crval1_linear``, ``poly_cdelt1_linear``, ``wv{min,max}_{expected,useful}``.
    Raises ``ValueError`` for unknown names or combinations
    (set_wv_parameters.py:55-58,149-152).
    """
    if filter_name not in EMIR_VALID_FILTERS:
        raise ValueError("Unexpected filter_name:", filter_name)
    if grism_name not in EMIR_VALID_GRISMS:
        raise ValueError("Unexpected grism_name:", grism_name)
    try:
        row = _WV_TABLE[(grism_name, filter_name)]
    except KeyError:
        raise ValueError("invalid filter_name and grism_name combination") from None

    wv_parameters = {"crpix1_enlarged": 1.0}
    for key, value in row.items():
        if key in ("crval1_linear", "cdelt1_linear"):
            wv_parameters["poly_" + key] = np.polynomial.Polynomial(value)
        else:
            wv_parameters[key] = list(value) if isinstance(value, list) else value
    return wv_parameters

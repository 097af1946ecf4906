"""B200-native apply_rectwv_coeff: spectral rectification + wavelength calibration.

Drop-in for ``emirdrp.processing.wavecal.apply_rectwv_coeff`` (reference:
``processing/wavecal/apply_rectwv_coeff.py:35-287``) on hand-written sm_100a CUDA
kernels behind a C-ABI shared library.  See DESIGN.md.
"""

from .core import (EMIR_NAXIS1, EMIR_NAXIS1_ENLARGED, EMIR_NAXIS2, EMIR_NBARS,  # noqa: F401
                   EMIR_NPIXPERSLIT_RECTIFIED)
from .products import RectWaveCoeff  # noqa: F401
from .set_wv_parameters import set_wv_parameters  # noqa: F401

__version__ = "0.1.0"

"""EMIR constants used on the apply_rectwv_coeff path.

Values follow the reference's ``emirdrp/core/__init__.py:39-48``.
"""

EMIR_NBARS = 55
EMIR_NAXIS1 = 2048
EMIR_NAXIS2 = 2048
EMIR_NAXIS1_ENLARGED = 3400
EMIR_NPIXPERSLIT_RECTIFIED = 38
EMIR_VALID_GRISMS = ["J", "H", "K", "LR"]
EMIR_VALID_FILTERS = ["J", "H", "Ksp", "YJ", "HK"]

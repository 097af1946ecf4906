"""ORACLE — CPU restatement of the reference's apply_rectwv_coeff path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyemir_b200/`` imports, links or
executes this package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may use it, and only as the checker.

PARITY UNPINNED.  The arithmetic of the path lives in *numina*
(``pyproject.toml:32``: ``numina>=0.36.0``, no lock file, not vendored, not
installable here — no network).  The reference's own tests hold no golden vector
for rectification / wavelength resampling, and neither
numina nor astropy can be imported in this image, so the oracle could not be
pinned against outputs of the real reference.  What it follows:

* orchestration, header keywords, slitlet unpacking:
  ``emirdrp/processing/wavecal/apply_rectwv_coeff.py:35-287``,
  ``slitlet2d.py:145-218,340-431``, ``set_wv_parameters.py:19-154``;
* numina array kernels: the published algorithms restated in
  ``oracle/numina_restated.py`` and ``oracle/clippix.c``,
  anchored on the reference's call sites and on in-repo uses of the same
  functions (``synthetic_lines_rawdata.py:123-130``, ``slitlet2darc.py:1089-1113``);
* numina-independent invariants are tested for the oracle and
  for the CUDA path alike.
"""

from .build import load_clippix  # noqa: F401

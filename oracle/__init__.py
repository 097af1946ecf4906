"""ORACLE — CPU restatement of the reference's apply_rectwv_coeff path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pyemir_b200/`` imports, links or
executes this package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may use it, and only as the checker.

PARITY: pyemir's own code PINNED, numina arithmetic UNPINNED.  The arithmetic of
the path lives in *numina* (``pyproject.toml:32``: ``numina>=0.36.0``, no lock
file, not vendored, not installable here -- no network), and the reference's own
tests hold no golden vector for rectification / wavelength resampling
(SURVEY.md §4, §8(c)).  ``tests/golden/make_golden.py`` executes the reference's
own ``apply_rectwv_coeff.py``, ``slitlet2d.py``, ``set_wv_parameters.py`` and
``core/__init__.py`` with numina/astropy shimmed and commits the vectors;
``oracle/pipeline.py`` reproduces them bit for bit
(``tests/test_golden_reference.py``).  The numina kernels restated in
``oracle/numina_restated.py`` / ``oracle/clippix.c`` could NOT be pinned against
the real numina: PARITY UNPINNED for those.  What the oracle follows:

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

"""``RectWaveCoeff``: the calibration product consumed by apply_rectwv_coeff.

Host-side mirror of ``emirdrp/products.py:340-381``.  The reference class derives
from ``numina.types.structured.BaseStructuredCalibration`` (numina is not in this
image); the attributes and the JSON state layout that the hot path and the
recipes rely on are restated here: ``instrument``, ``uuid``, ``tags``,
``meta_info`` (with ``origin`` and ``dtu_configuration``), ``quality_control``
plus the four keys added at ``products.py:355-365`` and ``contents``.

``apply_rectwv_coeff`` is duck-typed on these attributes, so a real numina-backed
``RectWaveCoeff`` instance can be passed to it unchanged.
"""

import json
import uuid as _uuid

_STATE_KEYS = (
    "global_integer_offset_x_pix",
    "global_integer_offset_y_pix",
    "total_slitlets",
    "missing_slitlets",
)


class RectWaveCoeff:
    """Rectification and Wavelength Calibration Coefficients."""

    def __init__(self, instrument="unknown"):
        self.instrument = instrument
        self.uuid = str(_uuid.uuid1())
        self.tags = {"grism": "unknown", "filter": "unknown"}
        self.meta_info = {"origin": {}}
        self.quality_control = "UNKNOWN"
        self.global_integer_offset_x_pix = 0
        self.global_integer_offset_y_pix = 0
        self.total_slitlets = 0
        self.missing_slitlets = []
        self.contents = []

    def tag_names(self):
        return ["grism", "filter"]

    # -- state (same layout as the reference's JSON files) ---------------
    def __getstate__(self):
        state = {
            "instrument": self.instrument,
            "tags": self.tags,
            "uuid": self.uuid,
            "type": type(self).__name__,
            "meta_info": self.meta_info,
            "quality_control": str(getattr(self.quality_control, "name",
                                           self.quality_control)),
        }
        for key in _STATE_KEYS:
            state[key] = self.__dict__[key]
        state["contents"] = self.contents.copy()
        return state

    def __setstate__(self, state):
        self.instrument = state.get("instrument", "unknown")
        self.tags = state.get("tags", {})
        self.uuid = state.get("uuid", str(_uuid.uuid1()))
        self.meta_info = state.get("meta_info", {"origin": {}})
        self.quality_control = state.get("quality_control", "UNKNOWN")
        for key in _STATE_KEYS:
            self.__dict__[key] = state[key]
        self.contents = state["contents"].copy()

    # -- JSON ------------------------------------------------------------
    @classmethod
    def _datatype_load(cls, obj):
        """Load from a JSON file name (``apply_rectwv_coeff.py:361``)."""
        with open(obj, "r") as fd:
            state = json.load(fd)
        return cls.from_state(state)

    @classmethod
    def from_state(cls, state):
        result = cls.__new__(cls)
        result.__setstate__(state)
        return result

    def writeto(self, filename):
        with open(filename, "w") as fd:
            json.dump(self.__getstate__(), fd, indent=2, sort_keys=True)
        return filename

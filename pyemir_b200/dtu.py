"""DTU (detector translation unit) configuration check used by the hot path.

Host-side mirror of ``emirdrp/instrument/components/dtu.py:54-66,123-169`` and
``dtuaxis.py:78-118``: only what ``apply_rectwv_coeff.py:92-108`` needs —
``DtuConf.from_img``, ``DtuConf.from_header`` and equality through the
3-decimal string form of (coor, coor_f, coor_0) per axis.
"""

import hashlib


def _lookup(hdr, key, *default):
    """Case-insensitive ``hdr[key]`` for FITS headers and plain dicts.

    The reference wraps dicts in ``astropy.io.fits.Header``
    (dtuaxis.py:78-80), which folds keyword case.
    """
    if isinstance(hdr, dict):
        for k, v in hdr.items():
            if isinstance(k, str) and k.upper() == key.upper():
                return v
        if default:
            return default[0]
        raise KeyError(f"Keyword {key!r} not found.")
    if default:
        return hdr.get(key, default[0])
    return hdr[key]


class DTUAxis:
    """One DTU axis of movement (dtuaxis.py:16-118)."""

    def __init__(self, name):
        self.name = name
        self.coor = 0.0
        self.coor_f = 1.0
        self.coor_0 = 0.0

    def configure_with_header(self, hdr):
        self.coor = _lookup(hdr, f"{self.name}DTU")
        self.coor_f = _lookup(hdr, f"{self.name}DTU_F", 1.0)
        self.coor_0 = _lookup(hdr, f"{self.name}DTU_0", 0.0)

    @property
    def coor_r(self):
        return (self.coor / self.coor_f) - self.coor_0

    def stringify(self, ndig=3):
        m1 = [round(r, ndig) for r in (self.coor, self.coor_f, self.coor_0)]
        return f"{m1[0]:+010.3f}:{m1[1]:+010.3f}:{m1[2]:+010.3f}"

    def __hash__(self):
        return hash(self.stringify())

    def __eq__(self, other):
        if isinstance(other, DTUAxis):
            return (self.name.lower() == other.name.lower()
                    and self.stringify(3) == other.stringify(3))
        return NotImplemented

    def __ne__(self, other):
        return not self == other


class DtuConf:
    """DTU configuration of an image or of a calibration product."""

    def __init__(self, name="DTU"):
        self.name = name
        self.xaxis = DTUAxis("X")
        self.yaxis = DTUAxis("Y")
        self.zaxis = DTUAxis("Z")

    def configure_with_header(self, hdr):
        for axis in (self.xaxis, self.yaxis, self.zaxis):
            axis.configure_with_header(hdr)

    def configure_with_img(self, img):
        # dtu.py:60-66: the MECS extension wins when present
        if "MECS" in img:
            hdr = img["MECS"].header
        else:
            hdr = img[0].header
        self.configure_with_header(hdr)

    @classmethod
    def from_header(cls, hdr, name="DTU"):
        obj = cls(name)
        obj.configure_with_header(hdr)
        return obj

    @classmethod
    def from_img(cls, img, name="DTU"):
        obj = cls(name)
        obj.configure_with_img(img)
        return obj

    @property
    def coor(self):
        return [self.xaxis.coor, self.yaxis.coor, self.zaxis.coor]

    @property
    def coor_r(self):
        return [self.xaxis.coor_r, self.yaxis.coor_r, self.zaxis.coor_r]

    def stringify(self, ndig=3):
        return (f"x={self.xaxis.stringify(ndig)}:y={self.yaxis.stringify(ndig)}"
                f":z={self.zaxis.stringify(ndig)}")

    def describe(self, ndig=3):
        rows = [("XDTU..", self.xaxis.coor), ("YDTU..", self.yaxis.coor),
                ("ZDTU..", self.zaxis.coor), ("XDTU_0", self.xaxis.coor_0),
                ("YDTU_0", self.yaxis.coor_0), ("ZDTU_0", self.zaxis.coor_0),
                ("XDTU_F", self.xaxis.coor_f), ("YDTU_F", self.yaxis.coor_f),
                ("ZDTU_F", self.zaxis.coor_f)]
        body = "".join(f"- {k}: {v:8.{ndig}f}\n" for k, v in rows)
        return "<DtuConf instance>\n" + body

    __repr__ = describe

    def outdict(self, ndigits=3):
        out = {"dtuhash": hashlib.md5(self.stringify(3).encode("utf-8")).hexdigest()}
        for ax, axis in (("x", self.xaxis), ("y", self.yaxis), ("z", self.zaxis)):
            out[f"{ax}dtu"] = round(axis.coor, ndigits)
            out[f"{ax}dtu_0"] = round(axis.coor_0, ndigits)
            out[f"{ax}dtu_f"] = round(axis.coor_f, ndigits)
        return out

    def __eq__(self, other):
        if isinstance(other, DtuConf):
            return (self.xaxis == other.xaxis and self.yaxis == other.yaxis
                    and self.zaxis == other.zaxis)
        return NotImplemented

    def __ne__(self, other):
        return not self == other

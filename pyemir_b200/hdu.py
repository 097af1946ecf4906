"""Light stand-ins for ``astropy.io.fits`` Header / HDU / HDUList.

astropy is not part of this image, and the hot path only needs a handful of
header and HDU-list operations:

* ``header[key]`` (case-insensitive), ``header[key] = value`` and
  ``header[key] = (value, comment)``, ``key in header``, ``header.remove(key)``,
  ``header.get(key, default)``, ``header["history"] = text`` (appends a card);
* ``hdul[0]``, ``hdul["MECS"]``, ``"MECS" in hdul``, iteration, ``hdu.copy()``.

``apply_rectwv_coeff`` is duck-typed: it works on real astropy objects when they
are available and on these classes otherwise (tests, bench, smoke).  Semantics
follow astropy where the reference relies on them (keyword case folding, comment
kept when only the value is updated, HISTORY/COMMENT being multi-card keywords).
"""

import copy as _copy

import numpy as np

_COMMENTARY = ("HISTORY", "COMMENT", "")


class _Card:
    __slots__ = ("keyword", "value", "comment")

    def __init__(self, keyword, value, comment=""):
        self.keyword = keyword
        self.value = value
        self.comment = comment

    def __iter__(self):
        return iter((self.keyword, self.value, self.comment))

    def __repr__(self):
        return f"Card({self.keyword!r}, {self.value!r}, {self.comment!r})"


class _Comments:
    def __init__(self, header):
        self._header = header

    def __getitem__(self, key):
        return self._header._find(key).comment

    def __setitem__(self, key, comment):
        self._header._find(key).comment = comment


class Header:
    """Ordered, case-insensitive FITS-like header."""

    def __init__(self, cards=None):
        self._cards = []
        if cards is None:
            return
        if isinstance(cards, Header):
            self._cards = [_Card(*c) for c in cards._cards]
        elif isinstance(cards, dict):
            for key, value in cards.items():
                self[key] = value
        else:
            for item in cards:
                item = tuple(item)
                if len(item) == 2:
                    self[item[0]] = item[1]
                else:
                    self[item[0]] = (item[1], item[2])

    # -- helpers ---------------------------------------------------------
    @staticmethod
    def _norm(key):
        if not isinstance(key, str):
            raise KeyError(key)
        return key.strip().upper()

    def _find(self, key):
        name = self._norm(key)
        for card in self._cards:
            if card.keyword == name:
                return card
        raise KeyError(f"Keyword {name!r} not found.")

    # -- mapping protocol -----------------------------------------------
    def __contains__(self, key):
        try:
            name = self._norm(key)
        except KeyError:
            return False
        return any(card.keyword == name for card in self._cards)

    def __getitem__(self, key):
        name = self._norm(key)
        if name in _COMMENTARY:
            values = [c.value for c in self._cards if c.keyword == name]
            if not values:
                raise KeyError(f"Keyword {name!r} not found.")
            return values
        return self._find(name).value

    def __setitem__(self, key, value):
        name = self._norm(key)
        comment = None
        if isinstance(value, tuple):
            if len(value) == 2:
                value, comment = value
            elif len(value) == 1:
                (value,) = value
            else:
                raise ValueError("A Header item may be set with (value, comment)")
        if name in _COMMENTARY:
            self._cards.append(_Card(name, value, ""))
            return
        for card in self._cards:
            if card.keyword == name:
                card.value = value
                if comment is not None:
                    card.comment = comment
                return
        self._cards.append(_Card(name, value, comment or ""))

    def __delitem__(self, key):
        self.remove(key)

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._cards)

    def __eq__(self, other):
        if not isinstance(other, Header):
            return NotImplemented
        return [tuple(c) for c in self._cards] == [tuple(c) for c in other._cards]

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def remove(self, key, ignore_missing=False, remove_all=False):
        name = self._norm(key)
        for i, card in enumerate(self._cards):
            if card.keyword == name:
                del self._cards[i]
                if remove_all:
                    self._cards = [c for c in self._cards if c.keyword != name]
                return
        if not ignore_missing:
            raise KeyError(f"Keyword '{name}' not found.")

    def keys(self):
        return [c.keyword for c in self._cards]

    def values(self):
        return [c.value for c in self._cards]

    def items(self):
        return [(c.keyword, c.value) for c in self._cards]

    @property
    def cards(self):
        return [tuple(c) for c in self._cards]

    @property
    def comments(self):
        return _Comments(self)

    def copy(self):
        return Header(self)

    def __repr__(self):
        lines = [f"{k:<8s}= {v!r} / {c}" for k, v, c in self._cards]
        return "\n".join(lines)


class ImageHDU:
    """An image HDU: a header plus an ndarray (or ``None``)."""

    def __init__(self, data=None, header=None, name=None):
        self.data = data
        self.header = header if isinstance(header, Header) else Header(header)
        if name is not None:
            self.header["EXTNAME"] = name

    @property
    def name(self):
        return str(self.header.get("EXTNAME", "")).upper()

    def copy(self):
        data = None if self.data is None else np.array(self.data, copy=True)
        return type(self)(data=data, header=self.header.copy())


class PrimaryHDU(ImageHDU):
    @property
    def name(self):
        return "PRIMARY"


class HDUList(list):
    """A list of HDUs addressable by position or by EXTNAME."""

    def _index_of(self, key):
        if isinstance(key, str):
            name = key.strip().upper()
            for i, hdu in enumerate(self):
                if hdu.name == name:
                    return i
            raise KeyError(f"Extension {key!r} not found.")
        return key

    def __getitem__(self, key):
        if isinstance(key, slice):
            return HDUList(list.__getitem__(self, key))
        return list.__getitem__(self, self._index_of(key))

    def __setitem__(self, key, value):
        list.__setitem__(self, self._index_of(key), value)

    def __contains__(self, item):
        if isinstance(item, str):
            try:
                self._index_of(item)
            except KeyError:
                return False
            return True
        return list.__contains__(self, item)

    def copy(self):
        return HDUList([hdu.copy() for hdu in self])

    def __deepcopy__(self, memo):
        return HDUList([_copy.deepcopy(hdu, memo) for hdu in self])


def copy_img(img):
    """Deep copy of an HDUList, one ``hdu.copy()`` per extension.

    Restates ``numina.frame.utils.copy_img`` as used at
    ``apply_rectwv_coeff.py:69`` for both astropy and stand-in HDU lists.
    """
    return type(img)([hdu.copy() for hdu in img])

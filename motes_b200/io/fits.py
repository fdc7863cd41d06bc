"""Minimal FITS reader / writer with the slice of the ``astropy.io.fits`` / ``astropy.table`` API that MOTES uses.

SURVEY.md §8f rank 2: the reference reads its frames with ``astropy.io.fits.open`` and writes its products
with ``PrimaryHDU`` / ``ImageHDU`` / ``BinTableHDU`` / ``HDUList.writeto``
(/root/reference/src/motes/io/motesio.py:63-70, 172-254, 438-524; io/gmosio.py:38-42, 146-153).  astropy is
not part of this image, so this module restates the part of the FITS standard (v4.0: 2880-byte blocks,
80-character header cards, big-endian data, binary-table extensions) those calls rely on.  It is host-side
I/O only - no arithmetic of the extraction path lives here.  ``motes_b200.io.standin.install()`` registers it
under the astropy module names so the unmodified reference can import it where astropy is missing.

Supported: primary and IMAGE HDUs with BITPIX 8/16/32/64/-32/-64 (BSCALE/BZERO applied on read, the
unsigned-integer convention kept), BINTABLE HDUs with fixed-width columns (L B I J K E D A, repeat counts;
other column types and heaps are carried through byte for byte), CONTINUE long strings, HIERARCH keywords,
COMMENT / HISTORY cards.  Not supported: compressed images, random groups, ASCII tables, memory mapping.
"""

from __future__ import annotations

import builtins
import copy as _copy
import os
import re

import numpy as np

BLOCK = 2880
CARD = 80

_COMMENTARY = ("COMMENT", "HISTORY", "")
_STRUCTURAL = re.compile(r"^(SIMPLE|XTENSION|BITPIX|NAXIS\d*|EXTEND|PCOUNT|GCOUNT|TFIELDS|END)$")
_TABLE_KEYS = re.compile(r"^(TTYPE|TFORM|TUNIT|TDIM|TNULL|TSCAL|TZERO|TDISP)\d+$")
_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_TFORM_DTYPE = {"L": "i1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "E": ">f4", "D": ">f8"}


class VerifyError(ValueError):
    """The bytes on disk (or the object to be written) break the FITS standard."""


# ---------------------------------------------------------------------------------------------- header
class Card:
    __slots__ = ("keyword", "value", "comment")

    def __init__(self, keyword, value=None, comment=""):
        self.keyword = _norm_key(keyword)
        self.value = value
        self.comment = comment or ""

    def __repr__(self):
        return f"Card({self.keyword!r}, {self.value!r}, {self.comment!r})"

    def __iter__(self):
        return iter((self.keyword, self.value, self.comment))


def _norm_key(key: str) -> str:
    key = key.strip().upper()
    if key.startswith("HIERARCH "):
        key = key[9:].strip()
    return key


def _format_value(value) -> str:
    if value is None:
        return ""
    if isinstance(value, (bool, np.bool_)):
        return f"{'T' if value else 'F':>20}"
    if isinstance(value, (int, np.integer)):
        return f"{int(value):>20d}"
    if isinstance(value, (float, np.floating)):
        v = float(value)
        if not np.isfinite(v):
            raise VerifyError(f"FITS cannot hold the floating-point value {v!r}")
        text = repr(v).upper()
        if "." not in text and "E" not in text:
            text += ".0"
        if "E" in text and "." not in text.split("E")[0]:
            mant, exp = text.split("E")
            text = f"{mant}.0E{exp}"
        if len(text) > 20:
            text = f"{v:.15G}"
            if "." not in text and "E" not in text:
                text += ".0"
        return f"{text:>20}"
    if isinstance(value, str):
        return "'" + f"{value.replace(chr(39), chr(39) * 2):<8}" + "'"
    raise VerifyError(f"unsupported header value type {type(value).__name__}")


def _card_images(card: Card) -> list:
    """The 80-character record(s) of one card (several when a long string needs CONTINUE)."""
    key, value, comment = card.keyword, card.value, card.comment
    if key in _COMMENTARY:
        text = "" if value is None else str(value)
        chunks = [text[i:i + 72] for i in range(0, len(text), 72)] or [""]
        return [f"{key:<8}{c:<72}" for c in chunks]
    hier = len(key) > 8 or " " in key or not re.fullmatch(r"[A-Z0-9_-]*", key)
    head = f"HIERARCH {key} = " if hier else f"{key:<8}= "
    body = _format_value(value)
    if isinstance(value, str) and len(head) + len(body) > CARD:
        return _continued_string(head, value, comment)
    if isinstance(value, str):
        body = f"{body:<20}"
    if hier and not isinstance(value, str):
        body = body.strip()
    image = head + body
    if comment:
        image += " / " + comment
    if len(image) > CARD:
        image = image[:CARD]                       # comments are truncated, values never are
        if len(head + body) > CARD:
            raise VerifyError(f"card {key!r} does not fit into 80 characters")
    return [f"{image:<80}"]


def _continued_string(head, value, comment):
    """Long-string convention: pieces end with '&' and go on in CONTINUE cards."""
    escaped = value.replace("'", "''")
    images, first = [], True
    room0 = CARD - len(head) - 3                   # quotes and the '&'
    while escaped or first:
        room = room0 if first else CARD - 10 - 3
        piece = escaped[:room]
        if piece.endswith("'") and (len(piece) - len(piece.rstrip("'"))) % 2 == 1:
            piece = piece[:-1]                     # never split an escaped quote
        escaped = escaped[len(piece):]
        more = "&" if escaped else ""
        start = head if first else "CONTINUE  "
        images.append(f"{start}'{piece}{more}'")
        first = False
    if comment:
        tail = images[-1] + " / " + comment
        if len(tail) <= CARD:
            images[-1] = tail
        else:
            images[-1] = images[-1][:-1] + "&'"
            images.append(f"CONTINUE  '' / {comment}"[:CARD])
    return [f"{im:<80}" for im in images]


_NUMBER = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([EeDd][+-]?\d+)?$")


def _parse_value(text: str):
    """Value and comment of the part of a card after the '= ' indicator."""
    s = text.lstrip()
    if s.startswith("'"):
        i, out = 1, []
        while i < len(s):
            if s[i] == "'":
                if i + 1 < len(s) and s[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(s[i])
            i += 1
        value = "".join(out).rstrip()
        rest = s[i + 1:]
        comment = rest.split("/", 1)[1].strip() if "/" in rest else ""
        return value, comment
    if "/" in s:
        raw, comment = s.split("/", 1)
        comment = comment.strip()
    else:
        raw, comment = s, ""
    raw = raw.strip()
    if raw == "":
        return None, comment
    if raw == "T":
        return True, comment
    if raw == "F":
        return False, comment
    if re.fullmatch(r"[+-]?\d+", raw):
        return int(raw), comment
    if _NUMBER.match(raw):
        return float(raw.replace("D", "E").replace("d", "e")), comment
    return raw, comment                            # complex numbers etc.: kept as text


def _parse_card(image: str):
    key = image[:8].rstrip()
    if key == "HIERARCH" and "=" in image:
        left, right = image[9:].split("=", 1)
        value, comment = _parse_value(right)
        return Card(left, value, comment)
    if key in _COMMENTARY or image[8:10] != "= ":
        if key == "CONTINUE":
            value, comment = _parse_value(image[8:])
            return Card("CONTINUE", value, comment)
        return Card(key, image[8:].rstrip(), "") if key in _COMMENTARY else Card(key, None, image[8:].strip())
    value, comment = _parse_value(image[10:])
    return Card(key, value, comment)


class _Comments:
    def __init__(self, header):
        self._h = header

    def __getitem__(self, key):
        return self._h._find(key).comment

    def __setitem__(self, key, text):
        self._h._find(key).comment = text


class Header:
    """Ordered FITS header; ``h[key]``, ``h[key] = value`` or ``= (value, comment)``, ``key in h``, ``h.comments``."""

    def __init__(self, cards=()):
        self.cards = []
        for c in cards:
            self.append(c)

    # -- lookup
    def _index(self, key):
        if isinstance(key, int):
            return key
        k = _norm_key(key)
        for i, c in enumerate(self.cards):
            if c.keyword == k:
                return i
        return None

    def _find(self, key):
        i = self._index(key)
        if i is None:
            raise KeyError(f"Keyword {key!r} not found.")
        return self.cards[i]

    def __contains__(self, key):
        return self._index(key) is not None

    def __getitem__(self, key):
        k = _norm_key(key) if isinstance(key, str) else key
        if k in ("COMMENT", "HISTORY"):
            return [c.value for c in self.cards if c.keyword == k]
        return self._find(key).value

    def get(self, key, default=None):
        i = self._index(key)
        return default if i is None else self.cards[i].value

    def __setitem__(self, key, value):
        comment = None
        if isinstance(value, tuple):
            if len(value) == 1:
                value = value[0]
            else:
                value, comment = value[0], value[1]
        if isinstance(value, np.generic):
            value = value.item()
        k = _norm_key(key)
        if k in _COMMENTARY:
            self.cards.append(Card(k, value, ""))
            return
        i = self._index(k)
        if i is None:
            self.cards.append(Card(k, value, comment or ""))
        else:
            self.cards[i].value = value
            if comment is not None:
                self.cards[i].comment = comment

    def __delitem__(self, key):
        i = self._index(key)
        if i is None:
            raise KeyError(f"Keyword {key!r} not found.")
        del self.cards[i]

    def append(self, card):
        if not isinstance(card, Card):
            card = Card(*card)
        self.cards.append(card)

    def set(self, key, value=None, comment=None):
        self[key] = (value, comment) if comment is not None else value

    def remove(self, key, ignore_missing=False):
        if key in self:
            del self[key]
        elif not ignore_missing:
            raise KeyError(key)

    @property
    def comments(self):
        return _Comments(self)

    def keys(self):
        return [c.keyword for c in self.cards]

    def values(self):
        return [c.value for c in self.cards]

    def items(self):
        return [(c.keyword, c.value) for c in self.cards]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.cards)

    def copy(self):
        return _copy.deepcopy(self)

    def __repr__(self):
        return "\n".join(im.rstrip() for c in self.cards for im in _card_images(c))

    # -- bytes
    def tobytes(self) -> bytes:
        images = [im for c in self.cards for im in _card_images(c)]
        images.append(f"{'END':<80}")
        text = "".join(images)
        text += " " * (-len(text) % BLOCK)
        try:
            return text.encode("ascii")
        except UnicodeEncodeError as exc:
            raise VerifyError("FITS headers hold ASCII text only") from exc

    @classmethod
    def fromfile(cls, fh):
        """Read header blocks up to END; returns None at a clean end of file."""
        cards, done = [], False
        first = True
        while not done:
            block = fh.read(BLOCK)
            if first and len(block) == 0:
                return None
            if len(block) != BLOCK:
                raise VerifyError("truncated FITS header")
            first = False
            text = block.decode("ascii", errors="replace")
            for i in range(0, BLOCK, CARD):
                image = text[i:i + CARD]
                if image.startswith("END") and image[3:].strip() == "":
                    done = True
                    break
                if image.strip() == "":
                    continue
                card = _parse_card(image)
                if card.keyword == "CONTINUE" and cards and isinstance(cards[-1].value, str) and cards[-1].value.endswith("&"):
                    cards[-1].value = cards[-1].value[:-1] + (card.value or "")
                    if card.comment:
                        cards[-1].comment = (cards[-1].comment + " " + card.comment).strip()
                    continue
                cards.append(card)
        h = cls()
        h.cards = cards
        return h


def _strip_structural(header, table=False, scaling=False):
    keep = Header()
    for c in header.cards:
        if _STRUCTURAL.match(c.keyword) or (table and _TABLE_KEYS.match(c.keyword)):
            continue
        if scaling and c.keyword in ("BSCALE", "BZERO"):
            continue
        keep.cards.append(_copy.copy(c))
    return keep


# ---------------------------------------------------------------------------------------------- HDUs
class _BaseHDU:
    def __init__(self, data=None, header=None, name=None):
        self.header = header.copy() if header is not None else Header()
        self._raw = None                     # data unit as read from disk
        self._header_at_read = None          # header as read from disk (describes _raw)
        self._data = None
        self._loaded = True
        if data is not None:
            self.data = data
        if name is not None:
            self.header["EXTNAME"] = name

    @property
    def name(self):
        return str(self.header.get("EXTNAME", "PRIMARY" if isinstance(self, PrimaryHDU) else "")).strip().upper()

    @name.setter
    def name(self, value):
        self.header["EXTNAME"] = value

    @property
    def ver(self):
        return int(self.header.get("EXTVER", 1))

    @property
    def data(self):
        if not self._loaded:
            self._data = self._decode(self._raw)
            self._loaded = True
        return self._data

    @data.setter
    def data(self, value):
        self._data = self._coerce(value)
        self._loaded = True
        self._raw = None

    def _coerce(self, value):
        return value

    def _decode(self, raw):
        raise NotImplementedError

    def _payload(self):
        """(structural cards, data bytes) of this HDU as it will be written."""
        raise NotImplementedError

    def tobytes(self, primary: bool) -> bytes:
        cards, payload = self._payload()
        table = isinstance(self, BinTableHDU)
        head = Header()
        if primary:
            head.cards.append(Card("SIMPLE", True, "conforms to FITS standard"))
        else:
            head.cards.append(Card("XTENSION", "BINTABLE" if table else "IMAGE",
                                   "binary table extension" if table else "Image extension"))
        head.cards.extend(cards)
        body = _strip_structural(self.header, table=table and self._raw is None, scaling=not table)
        if primary:
            head.cards.append(Card("EXTEND", True, ""))
        head.cards.extend(body.cards)
        return head.tobytes() + payload + b"\0" * (-len(payload) % BLOCK)


class _ImageBase(_BaseHDU):
    def _coerce(self, value):
        arr = np.asarray(value)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        if arr.dtype.kind not in "iuf":
            raise VerifyError(f"cannot store dtype {arr.dtype} in an image HDU")
        return arr

    def _decode(self, raw):
        h = self._header_at_read
        naxis = int(h.get("NAXIS", 0))
        if naxis == 0:
            return None
        shape = tuple(int(h[f"NAXIS{i}"]) for i in range(naxis, 0, -1))
        arr = np.frombuffer(bytearray(raw), dtype=_BITPIX_DTYPE[int(h["BITPIX"])], count=int(np.prod(shape))).reshape(shape)
        bscale, bzero = h.get("BSCALE", 1), h.get("BZERO", 0)
        if bscale == 1 and bzero == 0:
            return arr
        for key in ("BSCALE", "BZERO"):            # the array handed out is the physical one (as astropy does)
            self.header.remove(key, ignore_missing=True)
        bits = int(h["BITPIX"])
        if bscale == 1 and bits in (16, 32, 64) and bzero == 2 ** (bits - 1):      # unsigned-integer convention
            return (arr.astype(f"i{bits // 8}").view(f"u{bits // 8}") ^ np.array(1 << (bits - 1), dtype=f"u{bits // 8}"))
        out = arr.astype(np.float32 if abs(bits) <= 16 else np.float64)
        return out * bscale + bzero

    def _payload(self):
        if self._raw is not None and not self._loaded:      # never touched: the bytes go back as they came
            h = self._header_at_read
            cards = [Card(k, h[k], h.comments[k]) for k in h.keys()
                     if re.fullmatch(r"BITPIX|NAXIS\d*|PCOUNT|GCOUNT|BSCALE|BZERO", k)]
            return cards, self._raw
        arr = self.data
        cards = []
        if arr is None:
            cards += [Card("BITPIX", 8, "array data type"), Card("NAXIS", 0, "number of array dimensions")]
            payload = b""
        else:
            extra = []
            if arr.dtype.kind == "u" and arr.dtype.itemsize > 1:
                bits = arr.dtype.itemsize * 8
                arr = (arr ^ np.array(1 << (bits - 1), dtype=arr.dtype)).view(f"i{bits // 8}")
                extra = [Card("BSCALE", 1), Card("BZERO", 2 ** (bits - 1))]
            elif arr.dtype == np.int8:
                arr = arr.astype(np.int16)
            elif arr.dtype == np.float16:
                arr = arr.astype(np.float32)
            bitpix = {"u": 8, "i": arr.dtype.itemsize * 8, "f": -arr.dtype.itemsize * 8}[arr.dtype.kind]
            cards += [Card("BITPIX", bitpix, "array data type"), Card("NAXIS", arr.ndim, "number of array dimensions")]
            cards += [Card(f"NAXIS{i + 1}", n) for i, n in enumerate(reversed(arr.shape))]
            cards += extra
            payload = np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder(">"), copy=False).tobytes()
        if not isinstance(self, PrimaryHDU):
            cards += [Card("PCOUNT", 0, "number of parameters"), Card("GCOUNT", 1, "number of groups")]
        return cards, payload

    @property
    def shape(self):
        d = self.data
        return () if d is None else d.shape


class PrimaryHDU(_ImageBase):
    pass


class ImageHDU(_ImageBase):
    pass


class Column:
    def __init__(self, name, format, array=None, unit=None):
        self.name, self.format, self.array, self.unit = name, format, array, unit


def _tform_of(dtype: np.dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype.kind in "SU":
        return f"{dtype.itemsize // (4 if dtype.kind == 'U' else 1)}A"
    table = {("b", 1): "L", ("u", 1): "B", ("i", 2): "I", ("i", 4): "J", ("i", 8): "K", ("f", 4): "E", ("f", 8): "D"}
    try:
        return table[(dtype.kind, dtype.itemsize)]
    except KeyError:
        raise VerifyError(f"no binary-table column format for dtype {dtype}") from None


def _parse_tform(tform: str):
    m = re.fullmatch(r"\s*(\d*)([LXBIJKAEDCMPQ])(.*)", tform.strip())
    if not m:
        raise VerifyError(f"unreadable TFORM {tform!r}")
    return (int(m.group(1)) if m.group(1) else 1), m.group(2)


class Table:
    """Column store with the constructor MOTES uses: ``Table(array2d, names=..., dtype=...)``
    (/root/reference/src/motes/io/motesio.py:191-205, 235-239)."""

    def __init__(self, data=None, names=None, dtype=None):
        self.columns = {}
        if data is None:
            return
        if isinstance(data, np.ndarray) and data.dtype.names:
            for n in data.dtype.names:
                self.columns[n] = np.array(data[n])
            return
        if isinstance(data, dict):
            for n, col in data.items():
                self.columns[n] = np.asarray(col)
            return
        arr = np.asarray(data)
        cols = [arr[:, i] for i in range(arr.shape[1])] if arr.ndim == 2 else [np.asarray(c) for c in data]
        names = list(names) if names is not None else [f"col{i}" for i in range(len(cols))]
        if len(names) != len(cols):
            raise ValueError("Arguments 'names' and the data columns must have the same length")
        dtypes = list(dtype) if dtype is not None else [None] * len(cols)
        for n, c, t in zip(names, cols, dtypes):
            self.columns[n] = np.array(c, dtype=t) if t is not None else np.array(c)

    @property
    def colnames(self):
        return list(self.columns)

    def __getitem__(self, name):
        return self.columns[name]

    def __setitem__(self, name, col):
        self.columns[name] = np.asarray(col)

    def __len__(self):
        return len(next(iter(self.columns.values()))) if self.columns else 0

    def as_array(self):
        rec = np.empty(len(self), dtype=[(n, c.dtype, c.shape[1:]) for n, c in self.columns.items()])
        for n, c in self.columns.items():
            rec[n] = c
        return rec


class BinTableHDU(_BaseHDU):
    def _coerce(self, value):
        if isinstance(value, Table):
            return value.as_array()
        arr = np.asarray(value)
        if not arr.dtype.names:
            raise VerifyError("a binary table needs a Table or a structured array")
        return arr

    @classmethod
    def from_columns(cls, columns, header=None, name=None):
        t = Table()
        for c in columns:
            t[c.name] = c.array
        return cls(t, header=header, name=name)

    @property
    def columns(self):
        return list(self.data.dtype.names)

    def _decode(self, raw):
        h = self._header_at_read
        rowbytes, nrows, nfields = int(h["NAXIS1"]), int(h["NAXIS2"]), int(h["TFIELDS"])
        fields, offset = [], 0
        for i in range(1, nfields + 1):
            repeat, code = _parse_tform(str(h[f"TFORM{i}"]))
            name = str(h.get(f"TTYPE{i}", f"col{i}")).strip()
            if code == "A":
                dt, shape, width = np.dtype(f"S{repeat}"), (), repeat
            elif code in _TFORM_DTYPE:
                dt = np.dtype(_TFORM_DTYPE[code])
                shape, width = ((repeat,) if repeat != 1 else ()), repeat * dt.itemsize
            elif code == "X":
                width = (repeat + 7) // 8
                dt, shape = np.dtype("u1"), (width,)
            elif code in "CM":
                dt = np.dtype(">c8" if code == "C" else ">c16")
                shape, width = ((repeat,) if repeat != 1 else ()), repeat * dt.itemsize
            else:                                  # P / Q descriptors: heap offsets, carried through
                dt = np.dtype(">i4" if code == "P" else ">i8")
                shape, width = (2 * repeat,), 2 * repeat * dt.itemsize
            fields.append((name, dt, shape, offset))
            offset += width
        if offset != rowbytes:
            raise VerifyError(f"TFORM widths add to {offset}, NAXIS1 is {rowbytes}")
        dtype = np.dtype({"names": [f[0] for f in fields], "formats": [(f[1], f[2]) if f[2] else f[1] for f in fields],
                          "offsets": [f[3] for f in fields], "itemsize": rowbytes})
        return np.frombuffer(bytearray(raw[:rowbytes * nrows]), dtype=dtype, count=nrows)

    def _payload(self):
        if self._raw is not None:                  # table read from disk: bytes and column cards go back unchanged
            h = self._header_at_read
            cards = [Card(k, h[k], h.comments[k]) for k in h.keys()
                     if re.fullmatch(r"BITPIX|NAXIS\d*|PCOUNT|GCOUNT|TFIELDS", k)]
            return cards, self._raw
        rec = self.data
        names = rec.dtype.names
        packed = np.empty(len(rec), dtype=[(n, rec.dtype[n].base.newbyteorder(">") if rec.dtype[n].base.kind not in "SU"
                                            else np.dtype(f"S{_parse_tform(_tform_of(rec.dtype[n].base))[0]}"),
                                            rec.dtype[n].shape) for n in names])
        for n in names:
            col = rec[n]
            if col.dtype.kind == "U":
                col = np.char.encode(col, "ascii")
            if col.dtype.kind == "b":
                col = np.where(col, ord("T"), ord("F")).astype("i1")
            packed[n] = col
        cards = [Card("BITPIX", 8, "array data type"), Card("NAXIS", 2, "number of array dimensions"),
                 Card("NAXIS1", packed.dtype.itemsize, "length of dimension 1"),
                 Card("NAXIS2", len(packed), "length of dimension 2"),
                 Card("PCOUNT", 0, "number of group parameters"), Card("GCOUNT", 1, "number of groups"),
                 Card("TFIELDS", len(names), "number of table fields")]
        for i, n in enumerate(names, 1):
            base, shape = rec.dtype[n].base, rec.dtype[n].shape
            repeat = int(np.prod(shape)) if shape else 1
            form = _tform_of(base)
            if base.kind not in "SU" and repeat != 1:
                form = f"{repeat}{form}"
            cards += [Card(f"TTYPE{i}", n), Card(f"TFORM{i}", form)]
        return cards, packed.tobytes()


# ---------------------------------------------------------------------------------------------- HDU list
class HDUList(list):
    def __init__(self, hdus=(), file=None):
        if isinstance(hdus, _BaseHDU):
            hdus = [hdus]
        super().__init__(hdus)
        self._file = file

    def _locate(self, key):
        if isinstance(key, (int, np.integer, slice)):
            return key
        ver = None
        if isinstance(key, tuple):
            key, ver = key
        want = str(key).strip().upper()
        for i, hdu in enumerate(self):
            if hdu.name == want and (ver is None or hdu.ver == ver):
                return i
        raise KeyError(f"Extension {key!r} not found.")

    def __getitem__(self, key):
        return super().__getitem__(self._locate(key))

    def __contains__(self, key):
        if isinstance(key, _BaseHDU):
            return any(h is key for h in self)
        try:
            self._locate(key)
            return True
        except KeyError:
            return False

    def index_of(self, key):
        return self._locate(key)

    def info(self):
        return [(i, h.name, h.ver, type(h).__name__, len(h.header)) for i, h in enumerate(self)]

    def writeto(self, name, overwrite=False):
        if not len(self):
            raise VerifyError("nothing to write")
        if os.path.exists(name) and not overwrite:
            raise OSError(f"File {name} already exists. If you mean to replace it then use the argument \"overwrite=True\".")
        first = self[0]
        if not isinstance(first, PrimaryHDU):
            raise VerifyError("the first HDU of a FITS file must be a PrimaryHDU")
        with builtins.open(name, "wb") as fh:
            for i, hdu in enumerate(self):
                if i > 0 and isinstance(hdu, PrimaryHDU):          # astropy turns a late PrimaryHDU into an image extension
                    ext = ImageHDU(hdu.data, header=hdu.header)
                    fh.write(ext.tobytes(primary=False))
                else:
                    fh.write(hdu.tobytes(primary=(i == 0)))

    def close(self):
        self._file = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __deepcopy__(self, memo):
        return HDUList([_copy.deepcopy(h, memo) for h in self])


def _data_bytes(header) -> int:
    naxis = int(header.get("NAXIS", 0))
    if naxis == 0:
        return 0
    count = 1
    for i in range(1, naxis + 1):
        count *= int(header[f"NAXIS{i}"])
    return abs(int(header["BITPIX"])) // 8 * int(header.get("GCOUNT", 1)) * (int(header.get("PCOUNT", 0)) + count)


def open(name, mode="readonly", **_ignored):      # noqa: A001  (astropy's name)
    """Read every HDU of ``name`` into memory (/root/reference/src/motes/io/motesio.py:63)."""
    if mode not in ("readonly", "r", "rb"):
        raise ValueError("this FITS shim opens files read-only")
    hdus = []
    with builtins.open(name, "rb") as fh:
        while True:
            header = Header.fromfile(fh)
            if header is None:
                break
            nbytes = _data_bytes(header)
            raw = fh.read(nbytes)
            if len(raw) != nbytes:
                raise VerifyError(f"{name}: truncated data unit")
            fh.seek(-nbytes % BLOCK, os.SEEK_CUR)
            if not hdus:
                if header.cards[0].keyword != "SIMPLE":
                    raise VerifyError(f"{name}: not a FITS file (no SIMPLE card)")
                hdu = PrimaryHDU()
            else:
                kind = str(header.get("XTENSION", "")).strip().upper()
                hdu = BinTableHDU() if kind in ("BINTABLE", "A3DTABLE") else ImageHDU() if kind == "IMAGE" else _OpaqueHDU(kind)
            hdu._header_at_read = header.copy()
            hdu.header = header            # structural cards stay visible (MOTES reads NAXIS1 from them, io/gmosio.py:66)
            hdu._raw, hdu._loaded = raw, False
            hdus.append(hdu)
    if not hdus:
        raise VerifyError(f"{name}: empty file")
    return HDUList(hdus, file=name)


class _OpaqueHDU(_BaseHDU):
    """Extension of a kind this shim does not interpret (ASCII table, ...): carried through byte for byte."""

    def __init__(self, kind=""):
        super().__init__()
        self.kind = kind

    def _decode(self, raw):
        return np.frombuffer(raw, dtype=np.uint8)

    def tobytes(self, primary):
        return self._header_at_read.tobytes() + self._raw + b"\0" * (-len(self._raw) % BLOCK)


def getheader(name, ext=0):
    return open(name)[ext].header


def getdata(name, ext=0):
    return open(name)[ext].data

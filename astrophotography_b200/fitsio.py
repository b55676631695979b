"""FITS I/O seam.

The reference keeps all file I/O in ``astropy.io.fits`` (every ``Ap*`` class has
its own ``_read_fits`` / ``_write_corrected_image``, e.g.
``core/ApCalibrate.py:260-328,348-404``); so does this package whenever astropy
is importable.  astropy is not installable in the authoring container or on the
GPU box (no network), so a small self-contained reader/writer for the subset
the hot path needs -- 2-D image HDUs, BITPIX 8/16/32/-32/-64, the unsigned
16-bit ``BZERO=32768`` convention (``uint=True``), image extensions -- stands in
when it is absent.  Either way the rest of the package only sees
``read_image`` / ``write_image`` / ``read_header`` and the small ``Header``
mapping below; arithmetic never happens here.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

try:
    from astropy.io import fits as _afits
    HAVE_ASTROPY = True
except Exception:                       # noqa: BLE001
    _afits = None
    HAVE_ASTROPY = False

BLOCK = 2880
_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_STRUCTURAL = ("SIMPLE", "XTENSION", "BITPIX", "NAXIS", "NAXIS1", "NAXIS2", "NAXIS3", "EXTEND",
               "PCOUNT", "GCOUNT", "BZERO", "BSCALE", "END")


class _Comments:
    def __init__(self, hdr):
        self._hdr = hdr

    def __getitem__(self, key):
        return self._hdr._cards[key.upper()][1]


class Header:
    """Ordered keyword -> (value, comment) mapping with HISTORY lines; the
    subset of ``astropy.io.fits.Header`` behaviour the Ap* classes rely on."""

    def __init__(self, cards=None):
        self._cards = OrderedDict()
        self.history = []
        self.comment_cards = []
        if cards:
            for k, v in (cards.items() if hasattr(cards, "items") else cards):
                self[k] = v

    def __contains__(self, key):
        return str(key).upper() in self._cards

    def __getitem__(self, key):
        return self._cards[str(key).upper()][0]

    def __setitem__(self, key, value):
        key = str(key).upper()
        if key == "HISTORY":
            self.history.append(str(value))
            return
        if key == "COMMENT":
            self.comment_cards.append(str(value))
            return
        if isinstance(value, tuple):
            val, com = (value + ("",))[:2]
        else:
            val, com = value, self._cards.get(key, (None, ""))[1]
        self._cards[key] = (val, com or "")

    def __delitem__(self, key):
        del self._cards[str(key).upper()]

    def __iter__(self):
        return iter(self._cards)

    def __len__(self):
        return len(self._cards)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def keys(self):
        return self._cards.keys()

    def items(self):
        return [(k, v[0]) for k, v in self._cards.items()]

    @property
    def comments(self):
        return _Comments(self)

    def copy(self):
        h = Header()
        h._cards = OrderedDict(self._cards)
        h.history = list(self.history)
        h.comment_cards = list(self.comment_cards)
        return h


# ---------------------------------------------------------------------------
# minimal reader
# ---------------------------------------------------------------------------
def _parse_value(raw):
    raw = raw.strip()
    if not raw:
        return None
    if raw.startswith("'"):
        end = 1
        out = []
        while end < len(raw):
            if raw[end] == "'":
                if end + 1 < len(raw) and raw[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(raw[end])
            end += 1
        return "".join(out).rstrip()
    if raw in ("T", "F"):
        return raw == "T"
    try:
        return int(raw)
    except ValueError:
        pass
    try:
        return float(raw.replace("D", "E").replace("d", "e"))
    except ValueError:
        return raw


def _split_value_comment(body):
    """Split ``value / comment`` honouring quoted strings."""
    inq = False
    for i, ch in enumerate(body):
        if ch == "'":
            inq = not inq
        elif ch == "/" and not inq:
            return body[:i], body[i + 1:].strip()
    return body, ""


def _read_header_block(f):
    hdr = Header()
    last_str_key = None
    while True:
        block = f.read(BLOCK)
        if len(block) < BLOCK:
            raise OSError("truncated FITS header")
        done = False
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if key == "HISTORY":
                hdr.history.append(card[8:].rstrip())
            elif key == "COMMENT":
                hdr.comment_cards.append(card[8:].rstrip())
            elif key == "CONTINUE" and last_str_key is not None:
                # long-string convention: the previous value ended in '&'
                val, com = _split_value_comment(card[8:])
                prev, pcom = hdr._cards[last_str_key]
                more = _parse_value(val)
                joined = (prev[:-1] if prev.endswith("&") else prev) + (more if isinstance(more, str) else "")
                hdr._cards[last_str_key] = (joined, pcom or com)
                if not joined.endswith("&"):
                    last_str_key = None
            elif key and card[8:10] == "= ":
                val, com = _split_value_comment(card[10:])
                pv = _parse_value(val)
                hdr._cards[key] = (pv, com)
                last_str_key = key if isinstance(pv, str) and pv.endswith("&") else None
        if done:
            for k, (v, c) in list(hdr._cards.items()):
                if isinstance(v, str) and v.endswith("&"):
                    hdr._cards[k] = (v[:-1], c)
            return hdr


def _data_nbytes(hdr):
    naxis = int(hdr.get("NAXIS", 0))
    if naxis == 0:
        return 0, ()
    shape = tuple(int(hdr[f"NAXIS{i}"]) for i in range(naxis, 0, -1))
    n = abs(int(hdr["BITPIX"])) // 8
    for s in shape:
        n *= s
    return n, shape


def _mini_read(path, ext, header_only=False):
    with open(path, "rb") as f:
        for hdu in range(ext + 1):
            hdr = _read_header_block(f)
            nbytes, shape = _data_nbytes(hdr)
            padded = (nbytes + BLOCK - 1) // BLOCK * BLOCK
            if hdu < ext:
                f.seek(padded, os.SEEK_CUR)
                continue
            if header_only or nbytes == 0:
                return None, hdr
            raw = np.frombuffer(f.read(nbytes), dtype=_BITPIX_DTYPE[int(hdr["BITPIX"])]).reshape(shape)
    bzero = hdr.get("BZERO", 0) or 0
    bscale = hdr.get("BSCALE", 1) or 1
    bitpix = int(hdr["BITPIX"])
    if bitpix == 16 and bscale == 1 and bzero == 32768:          # uint=True convention
        data = (raw.astype(np.int32) + 32768).astype(np.uint16)
    elif bscale == 1 and bzero == 0:
        data = raw.astype(raw.dtype.newbyteorder("="))
    else:                                                        # generic scaling, as astropy does
        data = (raw.astype(np.float32 if bitpix in (8, 16) else np.float64) * bscale + bzero)
    return data, hdr


# ---------------------------------------------------------------------------
# minimal writer
# ---------------------------------------------------------------------------
def _fmt_value(val):
    if isinstance(val, (bool, np.bool_)):
        return f"{'T' if val else 'F':>20}"
    if isinstance(val, (int, np.integer)):
        return f"{int(val):>20d}"
    if isinstance(val, (float, np.floating)):
        if not np.isfinite(val):
            # FITS has no NaN / inf tokens: written as a string (what astropy's 'silentfix' would leave readable)
            return f"'{str(float(val)).upper():<8}'"
        s = repr(float(val)).upper()
        if "E" not in s and "." not in s:
            s += ".0"
        return f"{s:>20}"
    s = str(val).replace("'", "''")
    return f"'{s:<8}'"


def _string_cards(key, val, com=""):
    """A string value that does not fit one card: OGIP long-string convention
    (value pieces ending in '&' + CONTINUE cards), as astropy writes them."""
    s = str(val).replace("'", "''")
    pieces = []
    while len(s) > 67:
        cut = 67
        if s[cut - 1] == "'" and (len(s[:cut]) - len(s[:cut].rstrip("'"))) % 2 == 1:
            cut -= 1                    # do not split an escaped quote pair
        pieces.append(s[:cut])
        s = s[cut:]
    pieces.append(s)
    cards = []
    for i, piece in enumerate(pieces):
        last = i == len(pieces) - 1
        text = f"'{piece}{'' if last else '&'}'"
        body = (f"{key:<8}= " if i == 0 else "CONTINUE  ") + text
        if last and com and len(body) + 3 + len(com) <= 80:
            body += f" / {com}"
        cards.append(body.ljust(80))
    return cards


def _cards_for(key, val, com=""):
    """One or more 80-character cards for ``key = val / com``."""
    if val is None:
        return [f"{key:<8}=".ljust(80)[:80]]
    fv = _fmt_value(val)
    if fv.startswith("'") and len(fv) > 70:
        return _string_cards(key, val, com)
    body = f"{key:<8}= {fv}"
    if com:
        room = 80 - len(body) - 3
        if room > 0:
            body += f" / {com[:room]}"
    return [body.ljust(80)]


def _card(key, val, com=""):
    return _cards_for(key, val, com)[0]


def _encode_header(shape, bitpix, hdr, primary, extname=None, bzero=None):
    cards = []
    cards.append(_card("SIMPLE", True, "conforms to FITS standard") if primary
                 else _card("XTENSION", "IMAGE", "Image extension"))
    cards.append(_card("BITPIX", bitpix, "array data type"))
    cards.append(_card("NAXIS", len(shape), "number of array dimensions"))
    for i, s in enumerate(reversed(shape)):
        cards.append(_card(f"NAXIS{i + 1}", s))
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards.append(_card("PCOUNT", 0))
        cards.append(_card("GCOUNT", 1))
        if extname:
            cards.append(_card("EXTNAME", extname, "extension name"))
    if bzero is not None:
        cards.append(_card("BSCALE", 1))
        cards.append(_card("BZERO", bzero))
    if hdr is not None:
        for key in hdr.keys():
            if key in _STRUCTURAL or (not primary and key == "EXTNAME"):
                continue
            cards.extend(_cards_for(key, hdr[key], hdr.comments[key]))
        for line in getattr(hdr, "comment_cards", []):
            cards.append(("COMMENT " + str(line))[:80].ljust(80))
        for line in getattr(hdr, "history", []):
            text = str(line)
            for i in range(0, max(len(text), 1), 72):       # long HISTORY text wraps over several cards
                cards.append(("HISTORY " + text[i:i + 72]).ljust(80))
    cards.append("END".ljust(80))
    head = "".join(cards).encode("ascii", "replace")
    return head + b" " * (-len(head) % BLOCK)


def _encode_hdu(data, hdr, primary, extname=None):
    data = None if data is None else np.asarray(data)
    bzero = None
    if data is None:
        bitpix, shape, payload = 8, (), b""
    else:
        if data.dtype == np.bool_:
            data = data.astype(np.uint8)
        if data.dtype == np.uint16:
            bitpix, bzero = 16, 32768
            payload = (data.astype(np.int32) - 32768).astype(">i2").tobytes()
        else:
            code = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}.get(data.dtype.str[1:])
            if code is None:
                raise TypeError(f"cannot write dtype {data.dtype} to FITS")
            bitpix = code
            payload = data.astype(data.dtype.newbyteorder(">")).tobytes()
        shape = data.shape
    payload += b"\0" * (-len(payload) % BLOCK)
    return _encode_header(shape, bitpix, hdr, primary, extname, bzero) + payload


def write_data_unit(path, payload, bitpix, shape, header=None, overwrite=True):
    """Write a primary image HDU whose data unit ``payload`` (a bytes-like object) is ALREADY in FITS byte
    order -- e.g. the big-endian float32 plane a GPU kernel produced: the header plus one ``write`` of the
    page-locked buffer, no host-side byte swap.  ``header``: the stand-in ``Header`` or an astropy header."""
    path = str(path)
    if os.path.exists(path) and not overwrite:
        raise OSError(f"{path} exists")
    buf = memoryview(payload).cast("B")
    want = abs(int(bitpix)) // 8 * int(np.prod(shape))
    if buf.nbytes != want:
        raise ValueError(f"write_data_unit: {buf.nbytes} bytes for a {shape} BITPIX={bitpix} image")
    hdr = header
    if hdr is not None and not isinstance(hdr, Header):     # an astropy header: cards through the stand-in
        h2 = Header()
        for key in hdr.keys():
            if key and key not in ("HISTORY", "COMMENT"):
                h2[key] = (hdr[key], hdr.comments[key])
        for line in header_history(hdr):
            h2["HISTORY"] = line
        hdr = h2
    with open(path, "wb") as f:
        f.write(_encode_header(tuple(shape), int(bitpix), hdr, True))
        f.write(buf)
        f.write(b"\0" * (-buf.nbytes % BLOCK))


# ---------------------------------------------------------------------------
# public seam
# ---------------------------------------------------------------------------
def read_image(path, ext=0):
    """Return ``(data, header)`` of image HDU ``ext`` (``uint=True`` semantics:
    BITPIX 16 with BZERO 32768 comes back as uint16)."""
    if HAVE_ASTROPY:
        with _afits.open(path, uint=True, do_not_scale_image_data=False) as hl:
            data = np.asarray(hl[ext].data)
            # FITS is big-endian and astropy keeps it that way ('>f4', '>i2'): torch.from_numpy refuses
            # non-native byte order, so normalise once here
            if data.dtype.byteorder not in ("=", "|"):
                data = data.astype(data.dtype.newbyteorder("="))
            return data, hl[ext].header.copy()
    return _mini_read(str(path), ext)


def header_history(hdr):
    """The HISTORY lines of a header as a list of str (astropy: ``hdr['HISTORY']``; stand-in: ``.history``)."""
    if hasattr(hdr, "history"):
        return list(hdr.history)
    return [str(h) for h in hdr["HISTORY"]] if "HISTORY" in hdr else []


def read_header(path, ext=0):
    if HAVE_ASTROPY:
        return _afits.getheader(path, ext)
    return _mini_read(str(path), ext, header_only=True)[1]


class ImageLayout:
    """Where the pixels of a 2-D image HDU sit in its file, so that a row band can be read straight into
    page-locked memory (one ``readinto``, no parsing, no intermediate array)."""

    def __init__(self, path, ext=0):
        self.path = str(path)
        with open(self.path, "rb") as f:
            for hdu in range(ext + 1):
                hdr = _read_header_block(f)
                nbytes, shape = _data_nbytes(hdr)
                if hdu < ext:
                    f.seek((nbytes + BLOCK - 1) // BLOCK * BLOCK, os.SEEK_CUR)
            self.offset = f.tell()
        self.header = hdr
        self.shape = shape
        self.bitpix = int(hdr["BITPIX"])
        self.bzero = hdr.get("BZERO", 0) or 0
        self.bscale = hdr.get("BSCALE", 1) or 1
        self.itemsize = abs(self.bitpix) // 8

    @property
    def raw_u16(self):
        """BITPIX=16 with BZERO=32768: the data unit can go to ``apgpu_stack_reduce_u16`` /
        ``apgpu_calibrate_u16`` as it is on disk (``u16_format='fits'``)."""
        return len(self.shape) == 2 and self.bitpix == 16 and self.bscale == 1 and self.bzero == 32768

    def same_pixels_as(self, other):
        return (self.shape, self.bitpix, self.bzero, self.bscale) == (other.shape, other.bitpix, other.bzero, other.bscale)

    def read_rows_raw(self, r0, r1, out):
        """Rows ``[r0, r1)`` exactly as stored (big-endian) into ``out``, a C-contiguous array of
        ``(r1 - r0) * NAXIS1 * itemsize`` bytes of any dtype."""
        w = self.shape[1]
        buf = memoryview(out).cast("B")
        want = (r1 - r0) * w * self.itemsize
        if buf.nbytes != want:
            raise ValueError(f"read_rows_raw: buffer of {buf.nbytes} bytes for {want} bytes of pixels")
        with open(self.path, "rb", buffering=0) as f:
            f.seek(self.offset + r0 * w * self.itemsize)
            got = 0
            while got < want:
                k = f.readinto(buf[got:])
                if not k:
                    raise OSError(f"truncated FITS data unit in {self.path}")
                got += k
        return out


def new_header(cards=None):
    if HAVE_ASTROPY:
        h = _afits.Header()
        for k, v in (cards or {}).items():
            h[k] = v
        return h
    return Header(cards)


def write_image(path, data, header=None, extensions=(), overwrite=True):
    """Write a primary image HDU plus optional ``(extname, data)`` image extensions."""
    path = str(path)
    if os.path.exists(path) and not overwrite:
        raise OSError(f"{path} exists")
    if HAVE_ASTROPY:
        hdus = [_afits.PrimaryHDU(data=data, header=header)]
        for name, d in extensions:
            hdus.append(_afits.ImageHDU(data=d, name=name))
        _afits.HDUList(hdus).writeto(path, output_verify="ignore", overwrite=True)
        return
    blob = _encode_hdu(data, header, True)
    for name, d in extensions:
        blob += _encode_hdu(d, None, False, extname=name)
    with open(path, "wb") as f:
        f.write(blob)

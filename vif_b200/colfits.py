"""Minimal reader/writer for vif's column-oriented FITS tables -- the on-disk format of the
qxmatch2 front end (reference: io/fits/table.hpp:1047-1086,1280-1310,
io/fits/base.hpp:148-152,283-287): one BINTABLE extension with ONE row; each column is a vector
cell `TFORMn = '<count><type>'`, multi-dimensional columns carry `TDIMn = '(fastest,...,slowest)'`.
cfitsio / astropy are not available in this environment, so this is written directly against the
FITS standard (2880-byte blocks, 80-character cards, big-endian data) and checked against the
reference's own fixture files (`test/data/sources.fits`, `table.fits`, read in tests).
Host-side I/O only: FITS/ASCII I/O stays on the host (north star).
"""
import re

import numpy as np

BLOCK = 2880
_TYPES = {"D": ">f8", "E": ">f4", "K": ">i8", "J": ">i4", "I": ">i2", "B": "u1", "L": "S1", "A": "S1"}
_CODES = {np.dtype("float64"): "D", np.dtype("float32"): "E", np.dtype("int64"): "K", np.dtype("uint64"): "K",
          np.dtype("int32"): "J", np.dtype("uint32"): "J", np.dtype("int16"): "I", np.dtype("uint8"): "B"}


def _parse_header(raw, off):
    cards = {}
    order = []
    while True:
        block = raw[off:off + BLOCK]
        if len(block) < BLOCK:
            raise ValueError("truncated FITS header")
        off += BLOCK
        for i in range(36):
            card = block[i * 80:(i + 1) * 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return cards, order, off
            if card[8:10] == "= ":
                val = card[10:]
                m = re.match(r"\s*'((?:[^']|'')*)'", val)
                if m:
                    v = m.group(1).replace("''", "'").rstrip()
                else:
                    v = val.split("/")[0].strip()
                cards[key] = v
                order.append(key)


def _data_size(cards):
    nax = int(cards.get("NAXIS", "0"))
    if nax == 0:
        return 0
    n = abs(int(cards["BITPIX"])) // 8
    for k in range(1, nax + 1):
        n *= int(cards["NAXIS%d" % k])
    return n * int(cards.get("GCOUNT", "1")) + int(cards.get("PCOUNT", "0"))


def read_columns(path, names=None):
    """Reads the first binary-table extension. Returns {UPPERCASE name: ndarray} (native byte
    order; dims from TDIM, slowest axis first as in vif's vec). `names` restricts / checks columns
    (case-insensitive, as the reference's read_table)."""
    raw = open(path, "rb").read()
    off = 0
    while off < len(raw):
        cards, _, data_off = _parse_header(raw, off)
        size = _data_size(cards)
        if cards.get("XTENSION", "") == "BINTABLE":
            if int(cards["NAXIS2"]) != 1:
                raise ValueError("%s: row-oriented tables are not supported (NAXIS2 = %s)" % (path, cards["NAXIS2"]))
            out, pos = {}, data_off
            want = None if names is None else {n.upper() for n in names}
            for k in range(1, int(cards["TFIELDS"]) + 1):
                m = re.match(r"(\d*)([A-Z])", cards["TFORM%d" % k])
                cnt, code = int(m.group(1) or 1), m.group(2)
                if code not in _TYPES:
                    raise ValueError("unsupported TFORM %r" % cards["TFORM%d" % k])
                dt = np.dtype(_TYPES[code])
                name = cards.get("TTYPE%d" % k, "COL%d" % k).upper()
                nbytes = cnt * dt.itemsize
                if want is None or name in want:
                    a = np.frombuffer(raw, dtype=dt, count=cnt, offset=pos)
                    if dt.kind in "fiu" and dt.itemsize > 1:
                        a = a.astype(dt.newbyteorder("="))
                    if code == "K" and cards.get("TZERO%d" % k, "") in ("9223372036854775808", "9.223372036854775808E18"):
                        a = (a.astype(np.int64) ^ np.int64(-2**63)).view(np.uint64)
                    tdim = cards.get("TDIM%d" % k)
                    if tdim:
                        dims = [int(v) for v in tdim.strip("() ").split(",")]
                        a = a.reshape(dims[::-1])
                    out[name] = a
                pos += nbytes
            if want is not None:
                missing = want - set(out)
                if missing:
                    raise KeyError("%s: missing column(s) %s" % (path, sorted(missing)))
            return out
        off = data_off + (size + BLOCK - 1) // BLOCK * BLOCK
    raise ValueError("%s: no binary table extension" % path)


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, int):
        v = "%20d" % value
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
        v = "%-20s" % v
    s = "%-8s= %s" % (key, v)
    if comment:
        s += " / " + comment
    return ("%-80s" % s)[:80]


def _pad(b, fill=b" "):
    r = (-len(b)) % BLOCK
    return b + fill * r


def write_columns(path, columns, extname="RESULT_BINARY"):
    """Writes {name: ndarray} as one column-oriented table (1 row, vector cells, TDIM for >1-d).
    uint64 is stored as 'K' with the same bit pattern (vif maps uint_t to 'K' as well, so
    npos reads back as -1 in signed readers -- io/fits/base.hpp:148-152)."""
    prim = [_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", 8), _card("NAXIS", 0),
            _card("EXTEND", True), "%-80s" % "END"]
    cols, width = [], 0
    for name, a in columns.items():
        a = np.ascontiguousarray(a)
        if a.size == 0:
            continue                      # cfitsio (and the reference) skip empty columns
        code = _CODES.get(a.dtype)
        if code is None:
            raise TypeError("unsupported dtype %s for column %s" % (a.dtype, name))
        be = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else a)
        be = be.astype(be.dtype.newbyteorder(">"))
        cols.append((name.upper(), code, a.shape, be))
        width += be.nbytes
    hdr = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8), _card("NAXIS", 2),
           _card("NAXIS1", width, "width of table in bytes"), _card("NAXIS2", 1, "number of rows in table"),
           _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(cols)), _card("EXTNAME", extname)]
    for k, (name, code, shape, be) in enumerate(cols, 1):
        hdr.append(_card("TTYPE%d" % k, name, "label for field"))
        hdr.append(_card("TFORM%d" % k, "%d%s" % (be.size, code), "format of field"))
        if len(shape) > 1:
            hdr.append(_card("TDIM%d" % k, "(" + ",".join(str(v) for v in shape[::-1]) + ")"))
    hdr.append("%-80s" % "END")
    with open(path, "wb") as f:
        f.write(_pad("".join(prim).encode("ascii")))
        f.write(_pad("".join(hdr).encode("ascii")))
        data = b"".join(be.tobytes() for _, _, _, be in cols)
        f.write(_pad(data, b"\0"))

"""qxmatch2 front end (reference: tools/qxmatch2/qxmatch2.cpp:74-157) on the B200 library.

    python -m vif_b200.qxmatch2 cats=[file1.fits,file2.fits] output=out.fits [nth=1] [thread=1]
        [verbose] [quiet] [pos=...] [radec1=[ra,dec]] [radec2=[ra,dec]] [brute] [no_mirror]

Same `key=value` / flag arguments, same input columns (`<pos.>RA`, `<pos.>DEC` of a one-row
column-oriented table), same output table (`ID`, `D`, `RID`, `RD`, qxmatch.hpp:19-22). One file in
`cats` = self match (qxmatch2.cpp:133-148). FITS I/O stays on the host (vif_b200/colfits.py); the
match itself is the GPU path -- there is no CPU fallback.
"""
import sys

import numpy as np

from . import colfits
from .qxmatch import qxmatch

USAGE = "usage: python -m vif_b200.qxmatch2 cats=[file1,file2] output=ofile [options=...]  (see module docstring)"


def _list(v):
    v = v.strip()
    if v.startswith("[") and v.endswith("]"):
        v = v[1:-1]
        return [s.strip() for s in v.split(",")] if v.strip() else []      # "[,pos]" -> ["", "pos"]
    return [v] if v else []


def parse_args(argv):
    a = dict(cats=[], output="", nth=1, thread=1, verbose=False, quiet=False, pos=[], radec1=["ra", "dec"],
             radec2=["ra", "dec"], brute=False, no_mirror=False)
    for arg in argv:
        key, eq, val = arg.lstrip("-").partition("=")
        if key not in a:
            print("warning: unrecognized program argument '%s'" % arg)      # the reference only warns
            continue
        cur = a[key]
        if isinstance(cur, bool):
            a[key] = True if not eq else val not in ("0", "false")
        elif isinstance(cur, list):
            a[key] = _list(val)
        elif isinstance(cur, int):
            a[key] = int(val)
        else:
            a[key] = val
    if a["quiet"]:
        a["verbose"] = False
    return a


def _read_cat(path, prefix, radec):
    names = [(prefix + radec[0]).upper(), (prefix + radec[1]).upper()]
    cols = colfits.read_columns(path, names)
    ra = np.asarray(cols[names[0]], dtype=np.float64).reshape(-1)
    dec = np.asarray(cols[names[1]], dtype=np.float64).reshape(-1)
    return ra, dec


def main(argv=None):
    a = parse_args(sys.argv[1:] if argv is None else argv)
    if not a["output"] or len(a["cats"]) not in (1, 2):
        if not a["quiet"]:
            print(USAGE)
        return 0
    pos = a["pos"]
    kw = dict(nth=a["nth"], thread=a["thread"], verbose=a["verbose"], no_mirror=a["no_mirror"],
              brute_force=a["brute"])
    if len(a["cats"]) == 2:
        pos = pos * 2 if len(pos) == 1 else (pos or ["", ""])
        pos = [p + "." if p else "" for p in pos]
        ra1, dec1 = _read_cat(a["cats"][0], pos[0], a["radec1"])
        ra2, dec2 = _read_cat(a["cats"][1], pos[1], a["radec2"])
        if ra1.size == 0 or dec1.size == 0:
            print("error: qxmatch: first catalog is empty")
        if ra2.size == 0 or dec2.size == 0:
            print("error: qxmatch: second catalog is empty")
        if a["verbose"]:
            print("qxmatch: crossmatching %d sources from '%s' with %d sources from '%s'..."
                  % (ra1.size, a["cats"][0], ra2.size, a["cats"][1]))
        res = qxmatch(ra1, dec1, ra2, dec2, **kw)
    else:
        pre = (pos[0] + ".") if pos and pos[0] else ""
        ra, dec = _read_cat(a["cats"][0], pre, a["radec1"])
        if ra.size == 0 or dec.size == 0:
            print("error: qxmatch: catalog is empty")
        if a["verbose"]:
            print("qxmatch: self matching %d sources from '%s'..." % (ra.size, a["cats"][0]))
        res = qxmatch(ra, dec, **kw)
    colfits.write_columns(a["output"], {"ID": res.id, "D": res.d, "RID": res.rid, "RD": res.rd})
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""Column-oriented FITS tables (vif_b200/colfits.py) and the qxmatch2 front end
(vif_b200/qxmatch2.py, reference tools/qxmatch2/qxmatch2.cpp:74-157)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from vif_b200 import colfits, catalogs as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FIXTURE = "/root/reference/test/data/sources.fits"


def test_round_trip_types_and_dims(tmp_path):
    cols = {"id": (np.arange(24, dtype=np.uint64).reshape(2, 12)), "d": np.linspace(0, 1, 24).reshape(2, 12),
            "rid": np.array([2**64 - 1, 7, 9], dtype=np.uint64), "rd": np.array([np.nan, 1.5, np.inf]),
            "f": np.arange(5, dtype=np.float32), "j": np.arange(5, dtype=np.int32)}
    p = str(tmp_path / "t.fits")
    colfits.write_columns(p, cols)
    raw = open(p, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30].startswith(b"SIMPLE  =")
    got = colfits.read_columns(p)
    assert list(got) == ["ID", "D", "RID", "RD", "F", "J"]
    assert got["ID"].shape == (2, 12) and np.array_equal(got["ID"].view(np.uint64), cols["id"])
    assert np.array_equal(got["D"], cols["d"])
    assert np.array_equal(got["RID"].view(np.uint64), cols["rid"])          # npos survives as a bit pattern
    assert np.array_equal(got["RD"], cols["rd"], equal_nan=True)
    assert got["F"].dtype == np.float32 and got["J"].dtype == np.int32
    # header layout of the reference: one row, vector cells, TDIM fastest-axis-first
    hdr = raw[2880:2880 * 2].decode("ascii")
    assert "NAXIS2  =                    1" in hdr and "TFORM1  = '24K" in hdr and "TDIM1   = '(12,2)" in hdr
    sub = colfits.read_columns(p, ["rd", "ID"])
    assert set(sub) == {"RD", "ID"}
    with pytest.raises(KeyError):
        colfits.read_columns(p, ["nope"])


@pytest.mark.skipif(not os.path.exists(REF_FIXTURE), reason="reference fixture not present on this box")
def test_reads_the_reference_fixture():
    """test/data/sources.fits is a real one-row table with RA/DEC vector columns written by
    MWRFITS -- the format qxmatch2 consumes (qxmatch2.cpp:13-19)."""
    c = colfits.read_columns(REF_FIXTURE, ["ra", "dec"])
    assert c["RA"].shape == (2491,) and c["DEC"].shape == (2491,)
    assert abs(c["RA"][0] - 53.17343521) < 1e-7 and abs(c["DEC"][0] + 27.63717461) < 1e-7
    assert np.all(np.isfinite(c["RA"])) and 52 < c["RA"].min() and c["RA"].max() < 54.5


def test_argument_parsing():
    from vif_b200.qxmatch2 import parse_args
    a = parse_args(["cats=[a.fits,b.fits]", "output=o.fits", "nth=3", "-verbose", "brute", "pos=[x]", "radec2=[alpha,delta]"])
    assert a["cats"] == ["a.fits", "b.fits"] and a["output"] == "o.fits" and a["nth"] == 3
    assert a["verbose"] and a["brute"] and not a["no_mirror"] and a["pos"] == ["x"] and a["radec2"] == ["alpha", "delta"]
    assert parse_args(["quiet", "verbose"])["verbose"] is False


def _cli(args):
    return subprocess.run([sys.executable, "-m", "vif_b200.qxmatch2"] + args, cwd=ROOT, capture_output=True, text=True)


def test_cli_help_without_output():
    p = _cli([])
    assert p.returncode == 0 and "usage" in p.stdout


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path, oracle):
    ra1, dec1 = G.c1(61, 8000)
    ra2, dec2 = G.c1(62, 7000)
    f1, f2, fo = (str(tmp_path / n) for n in ("a.fits", "b.fits", "o.fits"))
    colfits.write_columns(f1, {"ra": ra1, "dec": dec1})
    colfits.write_columns(f2, {"pos.alpha": ra2.astype(np.float32), "pos.delta": dec2.astype(np.float32)})
    p = _cli(["cats=[%s,%s]" % (f1, f2), "output=" + fo, "nth=2", "verbose", "pos=[,pos]", "radec2=[alpha,delta]"])
    assert p.returncode == 0, p.stdout + p.stderr
    assert "crossmatching 8000 sources" in p.stdout
    out = colfits.read_columns(fo)
    exp = oracle.oracle_qxmatch(ra1, dec1, ra2.astype(np.float32).astype(np.float64),
                                dec2.astype(np.float32).astype(np.float64), nth=2, thread=4)
    assert np.array_equal(out["ID"].view(np.uint64), exp["id"]) and np.array_equal(out["RID"].view(np.uint64), exp["rid"])
    assert np.allclose(out["D"], exp["d"], rtol=1e-9, atol=0) and np.allclose(out["RD"], exp["rd"], rtol=1e-9, atol=0)
    # one catalog = self match; rid/rd are written as npos / NaN columns
    p = _cli(["cats=[%s]" % f1, "output=" + fo, "nth=3", "quiet"])
    assert p.returncode == 0 and p.stdout == ""
    out = colfits.read_columns(fo)
    exp = oracle.oracle_qxmatch(ra1, dec1, ra1, dec1, nth=3, self=True, thread=4)
    assert np.array_equal(out["ID"].view(np.uint64), exp["id"]) and np.all(out["RID"] == -1)

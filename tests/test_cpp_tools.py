"""The reference's own C++ callers of the path, compiled UNMODIFIED against the drop-in (tests/cpp/Makefile,
built in the authoring container where the reference's headers are; the binaries travel to the GPU box):

  tests/_build/qxmatch2            /root/reference/tools/qxmatch2/qxmatch2.cpp:74-157, the front end the north
                                   star names: column-FITS RA/DEC in, ID/D/RID/RD out
  tests/_build/catalog_merge_test  astro/catalog_merge.hpp:122-245 (catalog_pool::add_catalog): qxmatch on
                                   views with nth = 2, thread = 4, cache file, mutual-best merge
  tests/_build/colfits_tool        round trip through include/vif_b200/io/colfits.hpp alone (no GPU)

cfitsio is absent, so vif::fits is served by the header-only shim; the Python twin (vif_b200/colfits.py)
writes the inputs and reads the outputs, which checks the two implementations against each other."""
import os
import subprocess

import numpy as np
import pytest

from vif_b200 import catalogs as G
from vif_b200 import colfits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def exe(name):
    p = os.path.join(BUILD, name)
    if not os.path.exists(p):
        pytest.skip("%s not built (needs the reference headers: python -c 'import __graft_entry__ as g; g.build()')" % name)
    return p


def test_colfits_shim_round_trip(tmp_path):
    """Python writes, the C++ shim reads (with type conversion float32 -> double) and rewrites, Python reads."""
    ra, dec = G.c1(7, 1000)
    ident = np.arange(2 * 1000, dtype=np.uint64).reshape(2, 1000)
    ident[1, 5] = np.uint64(0xFFFFFFFFFFFFFFFF)                     # npos survives the K column
    fin, fout = str(tmp_path / "in.fits"), str(tmp_path / "out.fits")
    colfits.write_columns(fin, {"ra": ra, "dec": dec.astype(np.float32), "id": ident, "n": np.array([42], dtype=np.uint64)})
    p = subprocess.run([exe("colfits_tool"), fin, fout, "more"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "1000 1000 2x1000 42" in p.stdout
    out = colfits.read_columns(fout)
    assert np.array_equal(out["RA"], ra) and np.array_equal(out["DEC"], dec.astype(np.float32).astype(np.float64))
    assert out["ID"].shape == (2, 1000) and np.array_equal(out["ID"].view(np.uint64), ident) and int(out["N"][0]) == 42


def test_colfits_shim_reads_the_references_fixture(tmp_path):
    """test/data/sources.fits of the reference (written by cfitsio) through the shim."""
    src = "/root/reference/test/data/sources.fits"
    if not os.path.exists(src):
        pytest.skip("reference fixtures are only present in the authoring container")
    fout = str(tmp_path / "out.fits")
    p = subprocess.run([exe("colfits_tool"), src, fout], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    a, b = colfits.read_columns(src, ["RA", "DEC"]), colfits.read_columns(fout)
    assert np.array_equal(np.asarray(a["RA"], dtype=np.float64), b["RA"])
    assert np.array_equal(np.asarray(a["DEC"], dtype=np.float64), b["DEC"])


def test_missing_column_is_print_and_exit(tmp_path):
    fin = str(tmp_path / "in.fits")
    colfits.write_columns(fin, {"ra": np.zeros(3)})
    p = subprocess.run([exe("colfits_tool"), fin, str(tmp_path / "o.fits")], capture_output=True, text=True)
    assert p.returncode == 1 and "cannot find column 'DEC'" in p.stdout + p.stderr


@pytest.mark.gpu
def test_reference_qxmatch2_tool_on_the_gpu(tmp_path, oracle):
    ra1, dec1 = G.c1(61, 8000)
    ra2, dec2 = G.c1(62, 7000)
    f1, f2, fo = (str(tmp_path / n) for n in ("a.fits", "b.fits", "o.fits"))
    colfits.write_columns(f1, {"ra": ra1, "dec": dec1})
    colfits.write_columns(f2, {"pos.alpha": ra2.astype(np.float32), "pos.delta": dec2.astype(np.float32)})
    p = subprocess.run([exe("qxmatch2"), "cats=[%s,%s]" % (f1, f2), "output=" + fo, "nth=2", "verbose", "pos=[,pos]",
                        "radec2=[alpha,delta]", "thread=4"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "crossmatching 8000 sources" in p.stdout
    out = colfits.read_columns(fo)
    exp = oracle.oracle_qxmatch(ra1, dec1, ra2.astype(np.float32).astype(np.float64),
                                dec2.astype(np.float32).astype(np.float64), nth=2, thread=4)
    assert np.array_equal(out["ID"].view(np.uint64), exp["id"]) and np.array_equal(out["RID"].view(np.uint64), exp["rid"])
    assert np.allclose(out["D"], exp["d"], rtol=1e-9, atol=0) and np.allclose(out["RD"], exp["rd"], rtol=1e-9, atol=0)
    # one catalog = self match (qxmatch2.cpp:131-149); brute force flag
    p = subprocess.run([exe("qxmatch2"), "cats=[%s]" % f1, "output=" + fo, "nth=3", "quiet"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout == "", p.stdout + p.stderr
    out = colfits.read_columns(fo)
    exp = oracle.oracle_qxmatch(ra1, dec1, ra1, dec1, nth=3, self=True, thread=4)
    assert np.array_equal(out["ID"].view(np.uint64), exp["id"]) and np.all(out["RID"] == -1)
    p = subprocess.run([exe("qxmatch2"), "cats=[%s,%s]" % (f1, f1), "output=" + fo, "brute", "no_mirror"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    out = colfits.read_columns(fo)
    exp = oracle.oracle_qxmatch(ra1, dec1, ra1, dec1, brute_force=True, no_mirror=True, thread=4)
    assert np.array_equal(out["ID"].view(np.uint64), exp["id"]) and "RID" not in out
    # the tool's own help path needs no GPU and no files
    assert "usage: qxmatch2" in subprocess.run([exe("qxmatch2")], capture_output=True, text=True).stdout


@pytest.mark.gpu
def test_reference_catalog_merge_on_the_gpu(tmp_path, oracle):
    """catalog_pool::add_catalog through the drop-in; expectation restated from catalog_merge.hpp:191-232 with
    the oracle's qxmatch(nth = 2): the duplicate-id quirk of the binned nth = 2 search is part of it."""
    n1, n2 = 6000, 5000
    ra1, dec1 = G.c1(71, n1)
    rb, db = G.c1(72, n2)
    ra2, dec2 = rb.copy(), db.copy()
    ra2[:3000] = ra1[:3000] + 2e-5                 # 3000 counterparts 0.07" away, 2000 new sources
    dec2[:3000] = dec1[:3000] - 1e-5
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        for a in (ra1, dec1, ra2, dec2):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    p = subprocess.run([exe("catalog_merge_test"), str(fin), str(fout), str(n1), str(n2), str(tmp_path / "cache" / "xm")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "cross-matching second" in p.stdout and "reading data from" in p.stdout      # pass 2 used the cache file
    raw = np.fromfile(fout, dtype=np.uint64)
    ngal, nidm = int(raw[0]), int(raw[1])
    idm = raw[2:2 + nidm]; o = 2 + nidm
    origin = raw[o:o + ngal]; o += ngal
    pra = raw[o:o + ngal].view(np.float64); o += ngal
    pdec = raw[o:o + ngal].view(np.float64); o += ngal
    d0 = raw[o:o + n2].view(np.float64)
    xm = oracle.oracle_qxmatch(ra2, dec2, ra1, dec1, nth=2, thread=4)          # cra[sel] = catalog 2 vs the pool (= catalog 1)
    e_idm, e_ra, e_dec, e_org = [], list(ra1), list(dec1), [1] * n1
    for i in range(n2):
        j = int(xm["id"][0, i])
        if int(xm["rid"][j]) == i:
            e_idm.append(j)
        else:
            e_idm.append(len(e_ra)); e_ra.append(ra2[i]); e_dec.append(dec2[i]); e_org.append(2)
    assert ngal == len(e_ra) and np.array_equal(idm, np.array(e_idm, dtype=np.uint64))
    assert np.array_equal(origin, np.array(e_org, dtype=np.uint64))
    assert np.array_equal(pra, np.array(e_ra)) and np.array_equal(pdec, np.array(e_dec))
    assert np.allclose(d0, xm["d"][0], rtol=1e-9, atol=0)
    assert int(np.sum(np.array(e_idm) < n1)) >= 2900          # the displaced copies (plus chance mutual-best pairs)
    cache = [f for f in os.listdir(tmp_path / "cache") if f.endswith(".fits")]
    assert len(cache) == 1
    c = colfits.read_columns(str(tmp_path / "cache" / cache[0]))
    assert set(c) >= {"SEL", "NGAL", "ID", "D", "RID", "RD"} and c["ID"].shape == (2, n2)
    assert np.array_equal(c["ID"].view(np.uint64), xm["id"])

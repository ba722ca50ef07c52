"""The C++ drop-in header (include/vif_b200/astro/qxmatch.hpp) used the way the reference's callers
use qxmatch: vif vectors in, qxmatch_res out (tests/cpp/dropin_test.cpp, compiled in the authoring
container against the reference's core headers; the binary travels to the GPU box)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_same_result
from vif_b200 import catalogs as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_build", "dropin_test")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="dropin_test not built (needs the reference headers)")


def run_exe(tmp_path, ra1, dec1, ra2, dec2, nth, flags):
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        for a in (ra1, dec1, ra2, dec2):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    p = subprocess.run([EXE, str(fin), str(fout), str(ra1.size), str(ra2.size), str(nth), flags],
                       capture_output=True, text=True)
    return p, fout


def read_out(fout):
    raw = np.fromfile(fout, dtype=np.uint64)
    k, n1, nrid, nrd, nbest, nlost = (int(v) for v in raw[:6])
    o = 6
    ident = raw[o:o + k * n1].reshape(k, n1); o += k * n1
    d = raw[o:o + k * n1].view(np.float64).reshape(k, n1); o += k * n1
    rid = raw[o:o + nrid]; o += nrid
    rd = raw[o:o + nrd].view(np.float64); o += nrd
    best = raw[o:o + nbest]
    return {"id": ident, "d": d, "rid": rid, "rd": rd}, best, nlost


@needs_exe
@pytest.mark.gpu
@pytest.mark.parametrize("nth,flags,kw", [(1, "-", {}), (2, "v", dict(nth=2)), (5, "s", dict(nth=5, self=True)),
                                          (1, "b", dict(brute_force=True)), (3, "m", dict(nth=3, no_mirror=True))],
                         ids=["catalogs", "views_float", "self", "brute", "no_mirror"])
def test_dropin_matches_oracle(tmp_path, oracle, nth, flags, kw):
    n1, n2 = (3000, 2500) if "b" in flags else (20000, 15000)
    ra1, dec1 = G.c1(51, n1)
    ra2, dec2 = G.c1(52, n2)
    if "v" in flags:                      # the test program narrows catalog 2 to float
        ra2 = ra2.astype(np.float32).astype(np.float64)
        dec2 = dec2.astype(np.float32).astype(np.float64)
    p, fout = run_exe(tmp_path, ra1, dec1, ra2, dec2, nth, flags)
    assert p.returncode == 0, p.stdout + p.stderr
    got, best, nlost = read_out(fout)
    if "s" in flags:
        exp = oracle.oracle_qxmatch(ra1, dec1, ra1, dec1, thread=4, **kw)
    else:
        exp = oracle.oracle_qxmatch(ra1, dec1, ra2, dec2, thread=4, **kw)
    assert_same_result(got, exp, 1e-9, flags)
    if not kw.get("no_mirror") and not kw.get("self"):
        keep = exp["rid"][exp["id"][0]] == np.arange(n1, dtype=np.uint64)
        assert np.array_equal(best, np.nonzero(keep)[0].astype(np.uint64))
        assert nlost == n1 - keep.sum()


@needs_exe
@pytest.mark.gpu
def test_dropin_error_is_print_and_exit(tmp_path):
    """vif_check semantics: message on stdout, exit(EXIT_FAILURE) (core/error.hpp:330-335)."""
    ra = np.array([0.1, np.nan, 0.3])
    p, _ = run_exe(tmp_path, ra, ra, np.array([0.1, 0.2]), np.array([0.1, 0.2]), 1, "-")
    assert p.returncode == 1
    assert "first RA and Dec coordinates contain invalid values" in p.stdout + p.stderr


@needs_exe
def test_dropin_fails_loudly_without_gpu(tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("checks the no-GPU behaviour")
    except ImportError:
        pass
    a, b = G.c1(1, 100)
    p, _ = run_exe(tmp_path, a, b, a[:50], b[:50], 1, "-")
    assert p.returncode == 1 and "CUDA" in p.stdout + p.stderr

"""The C-ABI library loads, exports every symbol include/*.h declares, and refuses to compute
without a GPU (no CPU fallback). No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from vif_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vif_b200_qxmatch.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 9
    for name in decl:
        assert hasattr(lib, name), name
    assert sorted(_lib.SYMBOLS) == decl
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    for name in decl:
        assert re.search(r"\bT %s\b" % name, out), name


def test_header_compiles_as_c():
    """The boundary is plain C."""
    code = '#include "vif_b200_qxmatch.h"\nint main(void){ xm_params p; (void)p; return sizeof(xm_stats) > 0 ? 0 : 1; }\n'
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=code.encode(), capture_output=True)
    assert p.returncode == 0, p.stderr.decode()


def test_struct_layout_matches_header():
    code = ('#include <stdio.h>\n#include <stddef.h>\n#include "vif_b200_qxmatch.h"\n'
            'int main(void){ printf("%zu %zu %zu %zu %zu", sizeof(xm_params), sizeof(xm_grid), sizeof(xm_stats),'
            ' offsetof(xm_params, device), offsetof(xm_params, bbox1)); return 0; }\n')
    exe = os.path.join(ROOT, "tests", "_build", "layout")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=code.encode(), check=True)
    sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_lib.XmParams), C.sizeof(_lib.XmGrid), C.sizeof(_lib.XmStats),
                     _lib.XmParams.device.offset, _lib.XmParams.bbox1.offset]


def test_version():
    assert b"sm_100a" in _lib.load().xm_version()


def _gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_gpu_present(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_gpu():
    from vif_b200 import qxmatch, QxmatchError
    with pytest.raises(QxmatchError) as e:
        qxmatch(np.zeros(4), np.zeros(4), np.ones(4), np.ones(4))
    assert e.value.code == _lib.XM_ERR_CUDA


def test_dimension_mismatch_message():
    from vif_b200 import qxmatch, QxmatchError
    with pytest.raises(QxmatchError, match="first RA and Dec dimensions do not match"):
        qxmatch(np.zeros(4), np.zeros(5), np.ones(4), np.ones(4))
    with pytest.raises(QxmatchError, match="second RA and Dec dimensions do not match"):
        qxmatch(np.zeros(4), np.zeros(4), np.ones(3), np.ones(4))


def test_no_oracle_in_product():
    """The product never imports, links or executes anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "vif_b200")):
        if os.sep + "lib" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(base, f), errors="replace").read()
                if re.search(r"\boracle\b", txt, flags=re.I) and f not in ("__init__.py",):
                    for line in txt.splitlines():
                        if re.search(r"import\s+oracle|from\s+oracle|oracle/|liboracle|qxmatch_oracle", line):
                            bad.append((f, line.strip()))
    for hdr in ("vif_b200_qxmatch.h",):
        assert "oracle" not in open(os.path.join(ROOT, "include", hdr)).read().lower()
    assert not bad, bad
    out = subprocess.check_output(["ldd", _lib.LIB_PATH]).decode()
    assert "oracle" not in out and "qxmatch_ref" not in out

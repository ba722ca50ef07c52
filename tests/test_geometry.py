"""Host-side scalar geometry and ring enumeration of the product library (no GPU needed: these
entry points are host-only) against the golden values dumped from the reference and against the
oracle port."""
import ctypes as C
import os

import numpy as np
import pytest

from vif_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
GEO = np.load(os.path.join(HERE, "golden", "geometry.npz"))


def product_grid(bb1, bb2, n2, nth, linear):
    lib = _lib.load()
    g = _lib.XmGrid()
    b1 = (C.c_double * 4)(*bb1)
    b2 = (C.c_double * 4)(*bb2)
    assert lib.xm_grid_from_bbox(b1, b2, n2, nth, int(linear), C.byref(g)) == 0
    return g.as_dict()


def test_area_matches_reference_golden():
    for box, area in zip(GEO["boxes"], GEO["areas"]):
        g = product_grid(box, box, 1000, 1, False)
        assert g["area"] == area, (box, g["area"], area)


@pytest.mark.parametrize("linear", [False, True])
def test_grid_matches_oracle_port(oracle, linear):
    rng = np.random.default_rng(5)
    for box in GEO["boxes"]:
        if linear and (box[0] == box[1] or box[2] == box[3]):
            continue
        for n2, nth in ((1, 1), (3, 5), (1000, 1), (10_000_000, 1), (100_000_000, 5), (12345, 2)):
            shrink = rng.random(4) * 0.1
            bb1 = [box[0] - shrink[0], box[1] + shrink[1], box[2] + shrink[2] * 0, box[3] + shrink[3]]
            a = product_grid(bb1, box, n2, nth, linear)
            b = oracle.oracle_grid(bb1, box, n2, nth, linear)
            for k in ("rra0", "rra1", "rdec0", "rdec1", "cell_size", "dra", "ddec", "nc2", "area"):
                assert a[k] == b[k], (k, box, n2, nth, a[k], b[k])
            if b["nra"] < 2**62 and b["ndec"] < 2**62:
                assert a["nra"] == b["nra"] and a["ndec"] == b["ndec"]


def test_headline_grid_values():
    """C2 = 890 x 890 cells of 40.51 arcsec."""
    g = product_grid([0, 10, -5, 5], [0, 10, -5, 5], 10_000_000, 1, False)
    assert (g["nra"], g["ndec"]) == (890, 890)
    assert abs(g["cell_size"] - 40.51) < 0.01


@pytest.mark.parametrize("dedupe", [0, 1])
def test_ring_enumeration_matches_reference_tables(dedupe):
    lib = _lib.load()
    checked = 0
    for nx, ny in GEO["grids"]:
        for k in range(0, min(max(nx, ny), 14)):
            bx = GEO[f"ring_{nx}_{ny}_{k}_bx"]
            by = GEO[f"ring_{nx}_{ny}_{k}_by"]
            cap = bx.size + 64
            gx = (C.c_int32 * cap)()
            gy = (C.c_int32 * cap)()
            n = lib.xm_ring_entries(k, int(nx), int(ny), dedupe, gx, gy, cap)
            got = list(zip(gx[:n], gy[:n]))
            exp = list(zip(bx.tolist(), by.tolist()))
            if dedupe:
                seen, uniq = set(), []
                for e in exp:
                    if e not in seen:
                        seen.add(e)
                        uniq.append(e)
                exp = uniq
            assert got == exp, (nx, ny, k)
            checked += 1
    assert checked > 50


def test_ring_counts():
    lib = _lib.load()
    n = [lib.xm_ring_entries(k, 91, 91, 0, None, None, 0) for k in range(4)]
    assert n == [1, 32, 24, 36]

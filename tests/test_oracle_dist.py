"""The oracle's restatements of angdist (vector forms), angdist_less and qdist against the compiled
reference (astro.hpp:43-103, qxmatch.hpp:684-717): bit-exact."""
import numpy as np
import pytest

from vif_b200 import catalogs as G


@pytest.fixture(scope="module")
def pts():
    ra1, dec1 = G.fullsky(91, 4000)
    ra2, dec2 = G.fullsky(92, 4000)
    ra1[:5] = ra2[:5]; dec1[:5] = dec2[:5]          # zero distances
    return ra1, dec1, ra2, dec2


def test_angdist_vector_forms(oracle, have_reference, pts):
    if not have_reference:
        pytest.skip("reference sources not mounted")
    ra1, dec1, ra2, dec2 = pts
    a = oracle.oracle_angdist_vec(ra1, dec1, ra2, dec2)
    b = oracle.ref_angdist_vec(ra1, dec1, ra2, dec2)
    assert np.array_equal(a, b) and a[0] == 0.0 and a.max() > 3600.0*90
    a = oracle.oracle_angdist_vec(ra1, dec1, ra2[:1], dec2[:1])          # scalar second position
    b = oracle.ref_angdist_vec(ra1, dec1, ra2[:1], dec2[:1])
    assert np.array_equal(a, b)
    assert a[7] == oracle.oracle_angdist(ra1[7], dec1[7], ra2[0], dec2[0])


def test_angdist_less(oracle, have_reference, pts):
    if not have_reference:
        pytest.skip("reference sources not mounted")
    ra1, dec1, ra2, dec2 = pts
    for radius in (1.0, 3600.0*30, 3600.0*100):
        a = oracle.oracle_angdist_less(ra1, dec1, ra2[0], dec2[0], radius)
        b = oracle.ref_angdist_less(ra1, dec1, ra2[0], dec2[0], radius)
        assert np.array_equal(a, b)
    assert 0 < a.sum() < a.size


def test_qdist(oracle, have_reference):
    if not have_reference:
        pytest.skip("reference sources not mounted")
    ra, dec = G.c1(93, 300)
    a = oracle.oracle_qdist(ra, dec)
    b = oracle.ref_qdist(ra, dec)
    assert np.array_equal(a, b)
    assert np.all(np.triu(a) == 0.0) and np.all(a[np.tril_indices(300, -1)] > 0.0)
    assert a[5, 2] == oracle.oracle_angdist_vec(ra[2:3], dec[2:3], ra[5:6], dec[5:6])[0] or \
        abs(a[5, 2] - oracle.oracle_angdist(ra[2], dec[2], ra[5], dec[5])) < 1e-9*a[5, 2]

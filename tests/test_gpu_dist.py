"""angdist / angdist_less / qdist kernels against the oracle's restatements (which tests/test_oracle_dist.py
pins bit-exactly to the compiled reference). Distances: <= 1e-9 relative (north-star tolerance; the
device cos / asin and the large-argument sine are CUDA's libm), exact zeros preserved."""
import numpy as np
import pytest
import torch

from vif_b200 import angdist, angdist_less, qdist, QxmatchError, catalogs as G

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _close(got, exp):
    assert got.shape == exp.shape
    assert np.array_equal(got == 0.0, exp == 0.0)
    nz = exp != 0.0
    assert np.all(np.abs(got[nz] - exp[nz]) <= RTOL*np.abs(exp[nz]))


def test_angdist_vector_and_scalar_forms(oracle):
    ra1, dec1 = G.fullsky(91, 200_003)
    ra2, dec2 = G.fullsky(92, 200_003)
    ra1[:5] = ra2[:5]; dec1[:5] = dec2[:5]
    ra2[5:1000] = ra1[5:1000] + 1e-4; dec2[5:1000] = np.clip(dec1[5:1000] - 2e-4, -90, 90)     # arcsec-scale pairs
    _close(angdist(ra1, dec1, ra2, dec2), oracle.oracle_angdist_vec(ra1, dec1, ra2, dec2))
    _close(angdist(ra1, dec1, float(ra2[7]), float(dec2[7])), oracle.oracle_angdist_vec(ra1, dec1, ra2[7:8], dec2[7:8]))
    t = angdist(torch.from_numpy(ra1).cuda(), torch.from_numpy(dec1).cuda(), torch.from_numpy(ra2).cuda(),
                torch.from_numpy(dec2).cuda())
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), angdist(ra1, dec1, ra2, dec2))
    # close pairs: the sine branch is bit-identical to glibc, only cos / asin can differ by an ulp or two
    near = angdist(ra1[5:1000], dec1[5:1000], ra2[5:1000], dec2[5:1000])
    exp = oracle.oracle_angdist_vec(ra1[5:1000], dec1[5:1000], ra2[5:1000], dec2[5:1000])
    assert np.all(np.abs(near - exp) <= 1e-13*exp)
    with pytest.raises(QxmatchError, match="position sets dimensions do not match"):
        angdist(ra1, dec1, ra2[:10], dec2[:10])


def test_angdist_less(oracle):
    ra1, dec1 = G.fullsky(93, 300_001)
    for ra2, dec2, radius in ((10.0, -20.0, 3600.0*25), (359.9, 89.0, 3600.0*3), (180.0, 0.0, 1.0)):
        got = angdist_less(ra1, dec1, ra2, dec2, radius)
        exp = oracle.oracle_angdist_less(ra1, dec1, ra2, dec2, radius)
        bad = np.flatnonzero(got != exp)
        # a boolean may differ only for a point within rounding of the radius
        d = oracle.oracle_angdist_vec(ra1[bad], dec1[bad], np.array([ra2]), np.array([dec2]))
        assert np.all(np.abs(d - radius) <= 1e-9*radius), (bad.size, d)
        assert got.dtype == bool and got.shape == ra1.shape


def test_qdist(oracle):
    ra, dec = G.c1(94, 777)
    ra[10] = ra[3]; dec[10] = dec[3]                        # a zero below the diagonal
    got = qdist(ra, dec)
    exp = oracle.oracle_qdist(ra, dec)
    assert got.shape == (777, 777) and got[10, 3] == 0.0
    assert np.all(np.triu(got) == 0.0)
    il = np.tril_indices(777, -1)
    assert np.all(np.abs(got[il] - exp[il]) <= RTOL*np.abs(exp[il]))
    fs = G.fullsky(95, 1500)
    g2, e2 = qdist(*fs), oracle.oracle_qdist(*fs)
    il = np.tril_indices(1500, -1)
    assert np.all(np.triu(g2) == 0.0) and np.all(np.abs(g2[il] - e2[il]) <= RTOL*np.abs(e2[il]))
    assert qdist(np.zeros(0), np.zeros(0)).shape == (0, 0)

import numpy as np
import torch

from vif_b200 import catalogs as G


def test_numpy_and_torch_generators_agree_bitwise():
    for seed, stream, start in ((1, 0, 0), (2, 1, 12345), (7, 15, 2**33)):
        a = G.uniform(seed, stream, 4096, start)
        b = G.uniform_torch(seed, stream, 4096, "cpu", start).numpy()
        assert np.array_equal(a, b)
        assert 0.0 <= a.min() and a.max() < 1.0


def test_rows_regenerate_independently():
    full = G.uniform(3, 2, 1000)
    part = G.uniform(3, 2, 100, start=500)
    assert np.array_equal(full[500:600], part)


def test_config_boxes():
    ra, dec = G.c2(1, 5000)
    assert 0 <= ra.min() and ra.max() < 10 and -5 <= dec.min() and dec.max() < 5
    ra, dec = G.c3_clustered(3, 4)
    assert ra.size == 10**4 and 0 < ra.min() and ra.max() < 10 and -5 < dec.min() and dec.max() < 5
    ra, dec = G.fullsky(1, 5000)
    assert 0 <= ra.min() and ra.max() < 360 and -90 <= dec.min() and dec.max() <= 90
    rt, dt = G.c2(1, 5000, xp="torch", device="cpu")
    assert np.array_equal(rt.numpy(), G.c2(1, 5000)[0])

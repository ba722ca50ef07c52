"""The library's restatement of glibc 2.39's sin/cos (csrc/xm_libm_glibc.cuh) against the host libm itself:
the reference's proxies carry the host libm's last bits (qxmatch.hpp:174-204), so bit-exact ids need the
device to reproduce them. Host side here (no GPU); the device side in the gpu test below."""
import ctypes as C
import ctypes.util

import numpy as np
import pytest

from vif_b200 import _lib

LIBM = C.CDLL(ctypes.util.find_library("m"))
LIBM.sin.restype = LIBM.cos.restype = C.c_double
LIBM.sin.argtypes = LIBM.cos.argtypes = [C.c_double]


def _args(n, seed):
    rng = np.random.default_rng(seed)
    x = (rng.random(n) * 2 - 1) * 2.42
    x[::3] *= rng.random(x[::3].size) ** 3            # denser towards 0: the binned path's half differences
    x[:8] = [0.0, 1e-300, 2.0 ** -27, -2.0 ** -26, 0.126, -0.855468, 0.855469, np.pi / 2]
    return x


def test_host_libm_is_recognised_and_restated_bit_for_bit():
    lib = _lib.load()
    model = lib.xm_libm_model()
    assert model in (1, 2), "host libm is not glibc 2.39's sin/cos: the device would fall back to the CUDA libm"
    x = _args(200_000, 5)
    for which, fn in ((0, LIBM.sin), (1, LIBM.cos)):
        bad = sum(1 for v in x if lib.xm_libm_eval_host(which, model, float(v)) != fn(float(v)))
        assert bad == 0, "%d of %d %s values differ from the host libm" % (bad, x.size, "sin" if which == 0 else "cos")
    assert np.isnan(lib.xm_libm_eval_host(0, model, 2.5)) and np.isnan(lib.xm_libm_eval_host(1, model, -3.0))


def test_the_two_models_differ_somewhere():
    """Otherwise the probe could not tell them apart."""
    lib = _lib.load()
    x = _args(20_000, 6)
    assert any(lib.xm_libm_eval_host(1, 1, float(v)) != lib.xm_libm_eval_host(1, 2, float(v)) for v in x)


@pytest.mark.gpu
def test_device_sin_cos_equal_the_host_libm():
    import torch
    lib = _lib.load()
    model = lib.xm_libm_model()
    assert model in (1, 2)
    x = _args(2_000_000, 7)
    xd = torch.from_numpy(x).cuda()
    out = torch.empty_like(xd)
    err = C.create_string_buffer(256)
    vsin, vcos = np.vectorize(LIBM.sin), np.vectorize(LIBM.cos)
    for which, fn in ((0, vsin), (1, vcos)):
        torch.cuda.synchronize()
        assert lib.xm_libm_eval_dev(which, model, xd.data_ptr(), x.size, out.data_ptr(), -1, None, err, 256) == 0, err.value
        got = out.cpu().numpy()
        exp = fn(x)
        assert np.array_equal(got, exp), "%d device values differ from the host libm" % int(np.sum(got != exp))

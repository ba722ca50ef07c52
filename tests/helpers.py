import numpy as np


def assert_same_result(got, exp, rtol=1e-9, what=""):
    """ids bit-exact; distances within `rtol` relative (north star: 1e-9), NaN/inf in the same slots."""
    for k in ("id", "rid"):
        a = np.asarray(got[k] if isinstance(got, dict) else getattr(got, k))
        b = np.asarray(exp[k])
        assert a.shape == b.shape, f"{what} {k}: shape {a.shape} vs {b.shape}"
        nd = int(np.sum(a != b))
        assert nd == 0, f"{what} {k}: {nd} of {a.size} ids differ (first at {np.argwhere(a != b)[:3].tolist()})"
    for k in ("d", "rd"):
        a = np.asarray(got[k] if isinstance(got, dict) else getattr(got, k))
        b = np.asarray(exp[k])
        assert a.shape == b.shape, f"{what} {k}: shape {a.shape} vs {b.shape}"
        assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what} {k}: NaN slots differ"
        assert np.array_equal(np.isinf(a), np.isinf(b)), f"{what} {k}: inf slots differ"
        fin = np.isfinite(b)
        if fin.any():
            rel = np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300))
            assert rel <= rtol, f"{what} {k}: max relative error {rel:.3e} > {rtol}"

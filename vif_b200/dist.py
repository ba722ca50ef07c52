"""Python mirror of the reference's distance helpers of the cross-match family, on the GPU:
`angdist` (astro/astro.hpp:33-81, degrees in, arcsec out; vector or scalar second position),
`angdist_less` (astro/astro.hpp:83-103) and `qdist` (astro/qxmatch.hpp:684-717, n x n matrix with
ret[j, i] for j > i). CUDA float64 tensors are used in place and a CUDA tensor is returned; numpy
arrays / scalars go through torch for the copies (plumbing) and numpy arrays come back.
No CPU implementation: without the CUDA library and a device every call raises."""
import ctypes as C

import numpy as np

from . import _lib
from .qxmatch import QxmatchError, _is_torch_cuda


def _dev(a, device):
    """-> (contiguous CUDA float64 tensor, came_from_host)"""
    import torch
    if _is_torch_cuda(a):
        if a.dtype != torch.float64:
            raise TypeError("device inputs must be float64")
        return a.contiguous().reshape(-1), False
    if not torch.cuda.is_available():
        raise QxmatchError(_lib.XM_ERR_CUDA, "no CUDA device: vif_b200 has no CPU implementation")
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1))
    return t.to(device if device is not None else "cuda"), True


def _check(rc, err):
    if rc != _lib.XM_OK:
        raise QxmatchError(rc, err.value.decode("utf-8", "replace") or ("error code %d" % rc))


def _same_dims(a, b, msg):
    if tuple(np.shape(a)) != tuple(np.shape(b)):
        raise QxmatchError(_lib.XM_ERR_INVALID_ARG, "%s (%s vs %s)" % (msg, list(np.shape(a)), list(np.shape(b))))


def angdist(ra1, dec1, ra2, dec2, device=None):
    """Angular distance [arcsec] between positions in degrees; ra2/dec2 may be scalars."""
    import torch
    lib = _lib.load()
    _same_dims(ra1, dec1, "first RA and Dec dimensions do not match")       # messages of astro.hpp:47-52
    _same_dims(ra2, dec2, "second RA and Dec dimensions do not match")
    scalar2 = np.ndim(ra2) == 0
    if not scalar2:
        _same_dims(ra1, ra2, "position sets dimensions do not match")
    shape = tuple(np.shape(ra1))
    a1, host = _dev(ra1, device)
    b1, _ = _dev(dec1, a1.device)
    a2, _ = _dev(np.reshape(ra2, (1,)) if scalar2 else ra2, a1.device)
    b2, _ = _dev(np.reshape(dec2, (1,)) if scalar2 else dec2, a1.device)
    out = torch.empty(a1.numel(), dtype=torch.float64, device=a1.device)
    err = C.create_string_buffer(512)
    torch.cuda.current_stream(a1.device).synchronize()
    _check(lib.xm_angdist_dev(a1.data_ptr(), b1.data_ptr(), a2.data_ptr(), b2.data_ptr(), a1.numel(), 1 if scalar2 else 0,
                              out.data_ptr(), a1.device.index or 0, None, err, 512), err)
    out = out.reshape(shape)
    return out.cpu().numpy() if host else out


def angdist_less(ra1, dec1, ra2, dec2, radius, device=None):
    """True where (ra1, dec1) [deg] lies within `radius` arcsec of the position (ra2, dec2) [deg]."""
    import torch
    lib = _lib.load()
    _same_dims(ra1, dec1, "RA and Dec dimensions do not match")            # astro.hpp:86-87
    shape = tuple(np.shape(ra1))
    a1, host = _dev(ra1, device)
    b1, _ = _dev(dec1, a1.device)
    out = torch.empty(a1.numel(), dtype=torch.uint8, device=a1.device)
    err = C.create_string_buffer(512)
    torch.cuda.current_stream(a1.device).synchronize()
    _check(lib.xm_angdist_less_dev(a1.data_ptr(), b1.data_ptr(), a1.numel(), float(ra2), float(dec2), float(radius),
                                   out.data_ptr(), a1.device.index or 0, None, err, 512), err)
    out = out.reshape(shape).bool()
    return out.cpu().numpy() if host else out


def qdist(ra, dec, device=None):
    """All-pairs distance matrix [arcsec]: ret[j, i] for j > i, zero elsewhere (as the reference)."""
    import torch
    lib = _lib.load()
    _same_dims(ra, dec, "first RA and Dec dimensions do not match")         # qxmatch.hpp:688-689
    a, host = _dev(ra, device)
    b, _ = _dev(dec, a.device)
    n = a.numel()
    out = torch.empty((n, n), dtype=torch.float64, device=a.device)
    err = C.create_string_buffer(512)
    torch.cuda.current_stream(a.device).synchronize()
    _check(lib.xm_qdist_dev(a.data_ptr(), b.data_ptr(), n, out.data_ptr(), a.device.index or 0, None, err, 512), err)
    return out.cpu().numpy() if host else out

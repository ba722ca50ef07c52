"""Python mirror of the reference interface `vif::astro::qxmatch` (qxmatch.hpp:133-634).

    res = qxmatch(ra1, dec1, ra2, dec2, nth=1, thread=1, verbose=False, self=False,
                  brute_force=False, no_mirror=False, linear=False)
    res.id  (nth, n1) uint64    res.d  (nth, n1) float64 [arcsec]
    res.rid (n2,)     uint64    res.rd (n2,)     float64 [arcsec]      (empty when no_mirror)

Same keyword names, meaning and defaults as `qxmatch_params` (qxmatch.hpp:31-39); `qxmatch(ra, dec)`
with one catalog is the self-match overload (qxmatch.hpp:629-634). Unmatched slots hold NPOS / NaN
exactly as in the reference.

Inputs may be numpy arrays / sequences (host path: xm_qxmatch_f64, H2D and D2H copies inside the
call) or CUDA torch tensors of dtype float64 (device path: xm_qxmatch_dev, results are returned as
CUDA tensors, ids as int64 bit patterns with NPOS == -1).

Error behaviour: the reference prints a message and calls exit(EXIT_FAILURE) (vif_check). The
Python mirror raises QxmatchError with the same message instead; the C++ drop-in header keeps the
print + exit behaviour.

All computation happens in the CUDA library; there is no CPU path here.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib

NPOS = np.uint64(0xFFFFFFFFFFFFFFFF)


class QxmatchError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


@dataclass
class QxmatchParams:
    """Field-for-field mirror of qxmatch_params (qxmatch.hpp:31-39) + GPU extras."""
    thread: int = 1
    nth: int = 1
    verbose: bool = False
    self: bool = False
    brute_force: bool = False
    no_mirror: bool = False
    linear: bool = False
    # extras
    device: int = -1
    bbox1: tuple = None        # (ra_min, ra_max, dec_min, dec_max) of the WHOLE catalog 1 when only
    exact_only: bool = False   # a row shard of it is passed (multi-GPU)
    exact_grid: bool = False   # brute_force result through the exact spherical grid search
    fwd_range: tuple = None    # (begin, count) of catalog-1 rows to match   } partitioned runs: both catalogs
    rev_range: tuple = None    # (begin, count) of catalog-2 rows to reverse  } whole, device path only
    bbox2: tuple = None        # with bbox1: catalog 2's bounding box -> no host round trip at the head of the call
    cat2_ready_event: int = 0  # cudaEvent_t handle: nothing reads catalog 2 before it (broadcast still in flight)

    def to_c(self):
        p = _lib.XmParams()
        p.thread = int(self.thread)
        p.nth = max(int(self.nth), 0)
        p.verbose, p.self, p.brute_force = int(self.verbose), int(self.self), int(self.brute_force)
        p.no_mirror, p.linear, p.exact_only = int(self.no_mirror), int(self.linear), int(self.exact_only)
        p.exact_grid = int(self.exact_grid)
        p.device = int(self.device)
        if self.fwd_range is not None or self.rev_range is not None:
            p.has_ranges = 1
            p.fwd_begin, p.fwd_count = (int(v) for v in (self.fwd_range or (0, 0)))
            p.rev_begin, p.rev_count = (int(v) for v in (self.rev_range or (0, 0)))
        if self.bbox1 is not None:
            p.has_bbox1 = 1
            for k in range(4):
                p.bbox1[k] = float(self.bbox1[k])
        if self.bbox2 is not None:
            p.has_bbox2 = 1
            for k in range(4):
                p.bbox2[k] = float(self.bbox2[k])
        if self.cat2_ready_event:
            p.cat2_ready_event = int(self.cat2_ready_event)
        return p


@dataclass
class QxmatchRes:
    """Mirror of qxmatch_res (qxmatch.hpp:8-17)."""
    id: object
    d: object
    rid: object
    rd: object
    stats: dict = field(default_factory=dict)


def _is_torch_cuda(x):
    try:
        import torch
    except ImportError:       # pragma: no cover
        return False
    return isinstance(x, torch.Tensor) and x.is_cuda


def _raise(rc, errbuf):
    raise QxmatchError(rc, errbuf.value.decode("utf-8", "replace") or ("error code %d" % rc))


def last_stats(device=-1):
    st = _lib.XmStats()
    rc = _lib.load().xm_last_stats(int(device), C.byref(st))
    if rc != _lib.XM_OK:
        raise QxmatchError(rc, "xm_last_stats failed")
    return st.as_dict()


def release_memory(device=-1):
    """Hands the scratch memory cached between calls back to the driver (xm_release_memory)."""
    rc = _lib.load().xm_release_memory(int(device))
    if rc != _lib.XM_OK:
        raise QxmatchError(rc, "xm_release_memory failed")


def qxmatch(ra1, dec1, ra2=None, dec2=None, params=None, out=None, **kw):
    """See module docstring. `params` may be a QxmatchParams; keywords override its fields.
    `out`: preallocated result buffers. Host path: dict of numpy arrays id/d/rid/rd, e.g. views of
    pinned memory so that the D2H copies run at full PCIe rate. Device path: dict with CUDA tensors
    for any of rid/rd (int64 / float64, n2 elements), e.g. views of a peer-mapped exchange buffer."""
    p = QxmatchParams(**{**(params.__dict__ if params is not None else {}), **kw})
    if ra2 is None and dec2 is None:
        # qxmatch.hpp:629-634: one catalog -> self match against itself
        p.self = True
        ra2, dec2 = ra1, dec1
    elif ra2 is None or dec2 is None:
        raise TypeError("qxmatch: ra2 and dec2 must be given together")
    lib = _lib.load()
    if _is_torch_cuda(ra1):
        return _qxmatch_torch(lib, ra1, dec1, ra2, dec2, p, out)
    return _qxmatch_host(lib, ra1, dec1, ra2, dec2, p, out)


def _check_dims(s1, s2, which):
    if tuple(s1) != tuple(s2):
        # message of qxmatch.hpp:140-143
        raise QxmatchError(_lib.XM_ERR_INVALID_ARG,
                           "%s RA and Dec dimensions do not match (%s vs %s)" % (which, list(s1), list(s2)))


def _qxmatch_host(lib, ra1, dec1, ra2, dec2, p, out=None):
    same = ra2 is ra1 and dec2 is dec1
    a1 = np.ascontiguousarray(ra1, dtype=np.float64)
    b1 = np.ascontiguousarray(dec1, dtype=np.float64)
    _check_dims(a1.shape, b1.shape, "first")
    if same:
        a2, b2 = a1, b1
    else:
        a2 = np.ascontiguousarray(ra2, dtype=np.float64)
        b2 = np.ascontiguousarray(dec2, dtype=np.float64)
    _check_dims(a2.shape, b2.shape, "second")
    a1, b1, a2, b2 = a1.reshape(-1), b1.reshape(-1), a2.reshape(-1), b2.reshape(-1)
    n1, n2 = a1.size, a2.size
    nth = max(int(p.nth), 1)
    nrev = 0 if p.no_mirror else n2
    if out is not None:
        ident, d, rid, rd = out["id"], out["d"], out["rid"], out["rd"]
        for arr, shape, dt in ((ident, (nth, n1), np.uint64), (d, (nth, n1), np.float64),
                               (rid, (nrev,), np.uint64), (rd, (nrev,), np.float64)):
            if arr.shape != shape or arr.dtype != dt or not arr.flags.c_contiguous:
                raise TypeError("out[...] arrays must be C-contiguous with the result shapes/dtypes")
    else:
        ident = np.empty((nth, n1), dtype=np.uint64)
        d = np.empty((nth, n1), dtype=np.float64)
        rid = np.empty(nrev, dtype=np.uint64)
        rd = np.empty(nrev, dtype=np.float64)
    cp = p.to_c()
    err = C.create_string_buffer(1024)
    rc = lib.xm_qxmatch_f64(a1.ctypes.data, b1.ctypes.data, n1, a2.ctypes.data, b2.ctypes.data, n2,
                            C.byref(cp), ident.ctypes.data, d.ctypes.data,
                            rid.ctypes.data if rid.size else None, rd.ctypes.data if rd.size else None,
                            err, 1024)
    if rc != _lib.XM_OK:
        _raise(rc, err)
    return QxmatchRes(ident, d, rid, rd, last_stats(p.device))


def _qxmatch_torch(lib, ra1, dec1, ra2, dec2, p, out=None):
    import torch
    for t in (ra1, dec1, ra2, dec2):
        if not (_is_torch_cuda(t) and t.dtype == torch.float64 and t.is_contiguous()):
            raise TypeError("device path needs contiguous CUDA float64 tensors")
    _check_dims(ra1.shape, dec1.shape, "first")
    _check_dims(ra2.shape, dec2.shape, "second")
    dev = ra1.device
    n1, n2 = ra1.numel(), ra2.numel()
    nth = max(int(p.nth), 1)
    ranged = p.fwd_range is not None or p.rev_range is not None
    n1o = int((p.fwd_range or (0, 0))[1]) if ranged else n1          # rows the outputs cover
    n2o = int((p.rev_range or (0, 0))[1]) if ranged else n2
    ident = torch.empty((nth, n1o), dtype=torch.int64, device=dev)
    d = torch.empty((nth, n1o), dtype=torch.float64, device=dev)
    nrev = 0 if p.no_mirror else n2o
    rid = (out or {}).get("rid")
    rd = (out or {}).get("rd")
    for t, dt in ((rid, torch.int64), (rd, torch.float64)):
        if t is not None and not (_is_torch_cuda(t) and t.dtype == dt and t.is_contiguous() and t.numel() == nrev
                                  and t.device == dev):
            raise TypeError("out['rid'] / out['rd'] must be contiguous CUDA int64 / float64 tensors of n2 elements")
    if rid is None:
        rid = torch.empty(nrev, dtype=torch.int64, device=dev)
    if rd is None:
        rd = torch.empty(nrev, dtype=torch.float64, device=dev)
    p.device = dev.index if dev.index is not None else torch.cuda.current_device()
    cp = p.to_c()
    err = C.create_string_buffer(1024)
    # the library runs on its own stream: make torch's pending work on these tensors visible first
    torch.cuda.current_stream(dev).synchronize()
    rc = lib.xm_qxmatch_dev(ra1.data_ptr(), dec1.data_ptr(), n1, ra2.data_ptr(), dec2.data_ptr(), n2,
                            C.byref(cp), ident.data_ptr(), d.data_ptr(),
                            rid.data_ptr() if rid.numel() else None,
                            rd.data_ptr() if rd.numel() else None, None, err, 1024)
    if rc != _lib.XM_OK:
        _raise(rc, err)
    return QxmatchRes(ident, d, rid, rd, last_stats(p.device))


@dataclass
class IdPair:
    """Mirror of id_pair (qxmatch.hpp:636-639)."""
    id1: object
    id2: object
    lost: object


def xmatch_clean_best(res):
    """Mutual-best filter (qxmatch.hpp:641-663): keeps row i when rid[id[i]] == i.
    CUDA tensors (device path results) are filtered on the device by xm_clean_best_dev while they
    are still resident; host arrays with numpy (the reference's helper is host code too)."""
    ident, rid = res.id, res.rid
    if _is_torch_cuda(ident):
        import torch
        lib = _lib.load()
        n1, n2 = ident.shape[1], rid.numel()
        dev = ident.device
        id0 = ident[0].contiguous()
        out = [torch.empty(n1, dtype=torch.int64, device=dev) for _ in range(3)]
        counts = (C.c_uint64 * 2)()
        err = C.create_string_buffer(512)
        torch.cuda.current_stream(dev).synchronize()
        rc = lib.xm_clean_best_dev(id0.data_ptr(), rid.data_ptr() if n2 else None, n1, n2, out[0].data_ptr(),
                                   out[1].data_ptr(), out[2].data_ptr(), counts, dev.index or 0, None, err, 512)
        if rc != _lib.XM_OK:
            _raise(rc, err)
        return IdPair(out[0][:counts[0]], out[1][:counts[0]], out[2][:counts[1]])
    id0 = np.asarray(ident)[0]
    rid = np.asarray(rid)
    i = np.arange(id0.size, dtype=np.uint64)
    ok = id0 < np.uint64(rid.size)
    keep = np.zeros(id0.size, dtype=bool)
    keep[ok] = rid[id0[ok]] == i[ok]
    return IdPair(i[keep], id0[keep], i[~keep])

"""ctypes bindings of oracle/_build/libqxmatch_oracle.so (the C port, qxmatch_oracle.c) and of
oracle/_ref/libqxmatch_ref.so (the unmodified reference, ref_wrapper.cpp). TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(_HERE, "_build", "libqxmatch_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libqxmatch_ref.so")
NPOS = np.uint64(0xFFFFFFFFFFFFFFFF)

_u64 = C.c_uint64
_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint64)
_ip = C.POINTER(C.c_int64)


class OracleGrid(C.Structure):
    _fields_ = [("rra0", C.c_double), ("rra1", C.c_double), ("rdec0", C.c_double),
                ("rdec1", C.c_double), ("cell_size", C.c_double), ("dra", C.c_double),
                ("ddec", C.c_double), ("nra", _u64), ("ndec", _u64), ("nc2", _u64),
                ("area", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class OracleInfo(C.Structure):
    _fields_ = [("n_eval_fwd", _u64), ("n_eval_rev", _u64),
                ("max_ring_fwd", _u64), ("max_ring_rev", _u64)]


def build(force=False):
    """Compile the port (always) and the reference (only where /root/reference exists)."""
    if force or not os.path.exists(_PORT) or \
            os.path.getmtime(_PORT) < os.path.getmtime(os.path.join(_HERE, "qxmatch_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle"])
    if os.path.isdir("/root/reference/include/vif") and (
            force or not os.path.exists(_REF) or
            os.path.getmtime(_REF) < os.path.getmtime(os.path.join(_HERE, "ref_wrapper.cpp"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_port = None
_ref = None


def _load_port():
    global _port
    if _port is None:
        build()
        lib = C.CDLL(_PORT)
        lib.oracle_qxmatch.restype = C.c_int
        lib.oracle_qxmatch.argtypes = [_dp, _dp, _u64, _dp, _dp, _u64, _u64, _u64, C.c_int, C.c_int,
                                       C.c_int, C.c_int, _up, _dp, _up, _dp,
                                       C.POINTER(OracleGrid), C.POINTER(OracleInfo)]
        lib.oracle_grid.restype = None
        lib.oracle_grid.argtypes = [_dp, _dp, _u64, _u64, C.c_int, C.POINTER(OracleGrid)]
        lib.oracle_ring.restype = _u64
        lib.oracle_ring.argtypes = [_u64, _u64, C.c_double, _u64, _ip, _ip, _u64, _dp]
        lib.oracle_field_area_bbox.restype = C.c_double
        lib.oracle_field_area_bbox.argtypes = [C.c_double] * 4
        lib.oracle_angdist.restype = C.c_double
        lib.oracle_angdist.argtypes = [C.c_double] * 4
        _bind_dist(lib, "oracle")
        _port = lib
    return _port


def have_ref():
    return os.path.exists(_REF) or os.path.isdir("/root/reference/include/vif")


def _load_ref():
    global _ref
    if _ref is None:
        build()
        lib = C.CDLL(_REF)
        lib.ref_qxmatch.restype = _u64
        lib.ref_qxmatch.argtypes = [_dp, _dp, _u64, _dp, _dp, _u64, _u64, _u64, C.c_int, C.c_int,
                                    C.c_int, C.c_int, _up, _dp, _up, _dp]
        lib.ref_field_area_bbox.restype = C.c_double
        lib.ref_field_area_bbox.argtypes = [C.c_double] * 4
        lib.ref_angdist.restype = C.c_double
        lib.ref_angdist.argtypes = [C.c_double] * 4
        _bind_dist(lib, "ref")
        lib.ref_ring.restype = _u64
        lib.ref_ring.argtypes = [_u64, _u64, C.c_double, _u64, _ip, _ip, _u64, _dp]
        lib.ref_hardware_threads.restype = _u64
        _ref = lib
    return _ref


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a, t):
    return a.ctypes.data_as(t)


def _run(fn, ra1, dec1, ra2, dec2, nth, thread, self, brute_force, no_mirror, linear, extra=()):
    ra1, dec1, ra2, dec2 = _f64(ra1), _f64(dec1), _f64(ra2), _f64(dec2)
    n1, n2 = ra1.size, ra2.size
    k = max(int(nth), 1)
    ident = np.empty((k, n1), dtype=np.uint64)
    d = np.empty((k, n1), dtype=np.float64)
    rid = np.empty(0 if no_mirror else n2, dtype=np.uint64)
    rd = np.empty(0 if no_mirror else n2, dtype=np.float64)
    rc = fn(_p(ra1, _dp), _p(dec1, _dp), n1, _p(ra2, _dp), _p(dec2, _dp), n2, int(nth), int(thread),
            int(self), int(brute_force), int(no_mirror), int(linear),
            _p(ident, _up), _p(d, _dp), _p(rid, _up), _p(rd, _dp), *extra)
    return rc, {"id": ident, "d": d, "rid": rid, "rd": rd}


def oracle_qxmatch(ra1, dec1, ra2, dec2, nth=1, thread=1, self=False, brute_force=False,
                   no_mirror=False, linear=False, with_info=False):
    """The C port (qxmatch_oracle.c). Returns dict(id (nth,n1) u64, d, rid, rd[, grid, info])."""
    lib = _load_port()
    g, info = OracleGrid(), OracleInfo()
    rc, res = _run(lib.oracle_qxmatch, ra1, dec1, ra2, dec2, nth, thread, self, brute_force,
                   no_mirror, linear, extra=(C.byref(g), C.byref(info)))
    if rc != 0:
        raise ValueError("RA and Dec coordinates contain invalid values (infinite or NaN)")
    if with_info:
        res["grid"] = g.as_dict()
        res["info"] = {k: getattr(info, k) for k, _ in info._fields_}
    return res


def ref_qxmatch(ra1, dec1, ra2, dec2, nth=1, thread=1, self=False, brute_force=False,
                no_mirror=False, linear=False):
    """The unmodified reference (vif::astro::qxmatch, qxmatch.hpp:133)."""
    lib = _load_ref()
    _, res = _run(lib.ref_qxmatch, ra1, dec1, ra2, dec2, nth, thread, self, brute_force,
                  no_mirror, linear)
    return res


def oracle_grid(bb1, bb2, n2, nth=1, linear=False):
    lib = _load_port()
    g = OracleGrid()
    b1, b2 = _f64(bb1), _f64(bb2)
    lib.oracle_grid(_p(b1, _dp), _p(b2, _dp), int(n2), max(int(nth), 1), int(linear), C.byref(g))
    return g.as_dict()


def _ring(fn, nx, ny, cs, k):
    cap = 4 * (k + 3) * (k + 3) + 8
    bx = np.zeros(cap, dtype=np.int64)
    by = np.zeros(cap, dtype=np.int64)
    md = C.c_double()
    n = fn(int(nx), int(ny), float(cs), int(k), _p(bx, _ip), _p(by, _ip), cap, C.byref(md))
    return bx[:n].copy(), by[:n].copy(), md.value


def oracle_ring(nx, ny, cs, k):
    return _ring(_load_port().oracle_ring, nx, ny, cs, k)


def ref_ring(nx, ny, cs, k):
    return _ring(_load_ref().ref_ring, nx, ny, cs, k)


def oracle_field_area_bbox(r0, r1, d0, d1):
    return _load_port().oracle_field_area_bbox(r0, r1, d0, d1)


def ref_field_area_bbox(r0, r1, d0, d1):
    return _load_ref().ref_field_area_bbox(r0, r1, d0, d1)


def _bind_dist(lib, prefix):
    f = getattr(lib, prefix + "_angdist_vec")
    f.restype = None
    f.argtypes = [_dp, _dp, _u64, _dp, _dp, _u64, _dp]
    f = getattr(lib, prefix + "_angdist_less")
    f.restype = None
    f.argtypes = [_dp, _dp, _u64, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_uint8)]
    f = getattr(lib, prefix + "_qdist")
    f.restype = None
    f.argtypes = [_dp, _dp, _u64, _dp]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def _dist_calls(lib, prefix):
    def angdist_vec(ra1, dec1, ra2, dec2):
        a1, b1, a2, b2 = _f64(ra1), _f64(dec1), _f64(ra2), _f64(dec2)
        out = np.empty(a1.size)
        getattr(lib, prefix + "_angdist_vec")(a1.ctypes.data_as(_dp), b1.ctypes.data_as(_dp), a1.size,
                                              a2.ctypes.data_as(_dp), b2.ctypes.data_as(_dp), a2.size,
                                              out.ctypes.data_as(_dp))
        return out

    def angdist_less(ra1, dec1, ra2, dec2, radius):
        a1, b1 = _f64(ra1), _f64(dec1)
        out = np.empty(a1.size, dtype=np.uint8)
        getattr(lib, prefix + "_angdist_less")(a1.ctypes.data_as(_dp), b1.ctypes.data_as(_dp), a1.size, float(ra2),
                                               float(dec2), float(radius), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def qdist(ra, dec):
        a, b = _f64(ra), _f64(dec)
        out = np.empty((a.size, a.size))
        getattr(lib, prefix + "_qdist")(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), a.size, out.ctypes.data_as(_dp))
        return out
    return angdist_vec, angdist_less, qdist


def oracle_angdist_vec(ra1, dec1, ra2, dec2):
    """astro.hpp:43-81 (ra2/dec2 of one element = the scalar overload)."""
    return _dist_calls(_load_port(), "oracle")[0](ra1, dec1, ra2, dec2)


def oracle_angdist_less(ra1, dec1, ra2, dec2, radius):
    """astro.hpp:83-103."""
    return _dist_calls(_load_port(), "oracle")[1](ra1, dec1, ra2, dec2, radius)


def oracle_qdist(ra, dec):
    """qxmatch.hpp:684-717."""
    return _dist_calls(_load_port(), "oracle")[2](ra, dec)


def ref_angdist_vec(ra1, dec1, ra2, dec2):
    return _dist_calls(_load_ref(), "ref")[0](ra1, dec1, ra2, dec2)


def ref_angdist_less(ra1, dec1, ra2, dec2, radius):
    return _dist_calls(_load_ref(), "ref")[1](ra1, dec1, ra2, dec2, radius)


def ref_qdist(ra, dec):
    return _dist_calls(_load_ref(), "ref")[2](ra, dec)


def oracle_angdist(a, b, c, d):
    return _load_port().oracle_angdist(a, b, c, d)


def ref_angdist(a, b, c, d):
    return _load_ref().ref_angdist(a, b, c, d)

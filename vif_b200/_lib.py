"""ctypes binding of include/vif_b200_qxmatch.h. Fails loudly when the library is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libvif_b200_qxmatch.so")

u64, i32, f64 = C.c_uint64, C.c_int32, C.c_double
dp, up, vp = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_void_p

XM_OK, XM_ERR_INVALID_ARG, XM_ERR_NONFINITE_1, XM_ERR_NONFINITE_2, XM_ERR_GRID_TOO_LARGE, \
    XM_ERR_CUDA, XM_ERR_NO_MEMORY = range(7)


class XmParams(C.Structure):
    _fields_ = [("struct_size", u64), ("thread", u64), ("nth", u64), ("verbose", C.c_uint8), ("self", C.c_uint8),
                ("brute_force", C.c_uint8), ("no_mirror", C.c_uint8), ("linear", C.c_uint8),
                ("has_bbox1", C.c_uint8), ("exact_only", C.c_uint8), ("exact_grid", C.c_uint8),
                ("device", i32), ("has_ranges", i32), ("bbox1", f64 * 4),
                ("fwd_begin", u64), ("fwd_count", u64), ("rev_begin", u64), ("rev_count", u64),
                ("has_bbox2", u64), ("bbox2", f64 * 4), ("cat2_ready_event", C.c_void_p)]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.struct_size = C.sizeof(XmParams)


class XmGrid(C.Structure):
    _fields_ = [("rra0", f64), ("rra1", f64), ("rdec0", f64), ("rdec1", f64), ("cell_size", f64),
                ("dra", f64), ("ddec", f64), ("nra", u64), ("ndec", u64), ("nc2", u64), ("area", f64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class XmStats(C.Structure):
    _fields_ = [("struct_size", u64), ("grid", XmGrid), ("ms_total", f64), ("ms_bbox", f64), ("ms_keys", f64),
                ("ms_sort", f64), ("ms_index", f64), ("ms_forward", f64), ("ms_reverse", f64),
                ("ms_search", f64), ("launches", u64), ("rows_refined_fwd", u64),
                ("rows_refined_rev", u64), ("rows_ambiguous", u64), ("evals_fwd", u64),
                ("evals_rev", u64), ("sin_model", u64), ("ms_host", f64), ("ms_filter_fwd", f64), ("ms_filter_rev", f64),
                ("cos_model", u64)]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.struct_size = C.sizeof(XmStats)

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("grid", "struct_size")}
        d["grid"] = self.grid.as_dict()
        return d


# every symbol include/vif_b200_qxmatch.h declares (tests check the export list against the header)
SYMBOLS = ["xm_bbox_neg_dev", "xm_libm_model", "xm_libm_eval_host", "xm_libm_eval_dev", "xm_band_partition_dev", "xm_band_finish_dev", "xm_band_home_dev", "xm_fp64_peak", "xm_version", "xm_init", "xm_shutdown", "xm_release_memory", "xm_rev_merge_p2p", "xm_angdist_dev", "xm_angdist_less_dev", "xm_qdist_dev", "xm_qxmatch_f64", "xm_qxmatch_dev",
           "xm_grid_from_bbox", "xm_bbox_dev", "xm_last_stats", "xm_sort_pairs_dev", "xm_ring_entries",
           "xm_rev_merge_pack", "xm_rev_merge_mask", "xm_rev_merge_unpack", "xm_clean_best_dev"]

_lib = None


def load():
    """Loads the shared library. No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "vif_b200: %s is missing -- build it with `python -m vif_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.xm_version.restype = C.c_char_p
    lib.xm_version.argtypes = []
    lib.xm_init.restype = i32
    lib.xm_init.argtypes = [i32, C.c_char_p, u64]
    lib.xm_shutdown.restype = None
    lib.xm_shutdown.argtypes = []
    lib.xm_release_memory.restype = C.c_int32
    lib.xm_release_memory.argtypes = [C.c_int32]
    sig = [vp, vp, u64, vp, vp, u64, C.POINTER(XmParams), vp, vp, vp, vp]
    lib.xm_qxmatch_f64.restype = i32
    lib.xm_qxmatch_f64.argtypes = sig + [C.c_char_p, u64]
    lib.xm_qxmatch_dev.restype = i32
    lib.xm_qxmatch_dev.argtypes = sig + [vp, C.c_char_p, u64]
    lib.xm_grid_from_bbox.restype = i32
    lib.xm_grid_from_bbox.argtypes = [dp, dp, u64, u64, i32, C.POINTER(XmGrid)]
    lib.xm_bbox_dev.restype = i32
    lib.xm_bbox_dev.argtypes = [vp, vp, u64, i32, vp, dp, up, C.c_char_p, u64]
    lib.xm_bbox_neg_dev.restype = i32
    lib.xm_bbox_neg_dev.argtypes = [vp, vp, u64, i32, vp, vp, C.c_char_p, u64]
    lib.xm_band_partition_dev.restype = i32
    lib.xm_band_partition_dev.argtypes = [C.c_void_p, C.c_void_p, u64, u64, C.POINTER(f64), i32, f64, i32, C.c_void_p, u64,
                                          C.c_void_p, C.POINTER(u64), i32, C.c_void_p, C.c_char_p, u64]
    lib.xm_band_finish_dev.restype = i32
    lib.xm_band_finish_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, u64, f64, C.c_void_p, i32, C.c_void_p,
                                       C.c_char_p, u64]
    lib.xm_band_home_dev.restype = i32
    lib.xm_band_home_dev.argtypes = [C.c_void_p, C.c_void_p, u64, u64, C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_char_p, u64]
    lib.xm_libm_model.restype = i32
    lib.xm_libm_eval_host.restype = f64
    lib.xm_libm_eval_host.argtypes = [i32, i32, f64]
    lib.xm_libm_eval_dev.restype = i32
    lib.xm_libm_eval_dev.argtypes = [i32, i32, C.c_void_p, u64, C.c_void_p, i32, C.c_void_p, C.c_char_p, u64]
    lib.xm_fp64_peak.restype = i32
    lib.xm_fp64_peak.argtypes = [i32, i32, C.POINTER(f64), C.c_char_p, u64]
    lib.xm_last_stats.restype = i32
    lib.xm_last_stats.argtypes = [i32, C.POINTER(XmStats)]
    lib.xm_sort_pairs_dev.restype = i32
    lib.xm_sort_pairs_dev.argtypes = [vp, vp, u64, i32, i32, vp, C.c_char_p, u64]
    lib.xm_ring_entries.restype = u64
    lib.xm_ring_entries.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, i32, C.POINTER(i32),
                                    C.POINTER(i32), u64]
    lib.xm_clean_best_dev.restype = i32
    lib.xm_clean_best_dev.argtypes = [vp, vp, u64, u64, vp, vp, vp, up, i32, vp, C.c_char_p, u64]
    lib.xm_rev_merge_pack.restype = i32
    lib.xm_rev_merge_pack.argtypes = [vp, vp, u64, u64, vp, i32, vp, C.c_char_p, u64]
    lib.xm_rev_merge_mask.restype = i32
    lib.xm_rev_merge_mask.argtypes = [vp, vp, vp, u64, i32, vp, C.c_char_p, u64]
    lib.xm_rev_merge_unpack.restype = i32
    lib.xm_rev_merge_unpack.argtypes = [vp, vp, vp, u64, f64, i32, vp, C.c_char_p, u64]
    lib.xm_angdist_dev.restype = i32
    lib.xm_angdist_dev.argtypes = [vp, vp, vp, vp, u64, i32, vp, i32, vp, C.c_char_p, u64]
    lib.xm_angdist_less_dev.restype = i32
    lib.xm_angdist_less_dev.argtypes = [vp, vp, u64, f64, f64, f64, vp, i32, vp, C.c_char_p, u64]
    lib.xm_qdist_dev.restype = i32
    lib.xm_qdist_dev.argtypes = [vp, vp, u64, vp, i32, vp, C.c_char_p, u64]
    lib.xm_rev_merge_p2p.restype = i32
    lib.xm_rev_merge_p2p.argtypes = [C.POINTER(vp), C.POINTER(vp), u64, i32, i32, f64, i32, vp, C.c_char_p, u64]
    _lib = lib
    return lib

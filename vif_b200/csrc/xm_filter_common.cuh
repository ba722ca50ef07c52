// Pieces shared by the filter kernels (xm_search_fast.cu: nth = 1, xm_search_topk.cu: nth = 2..8).
#pragma once
#include "xm_common.cuh"
#include "xm_math_exact.cuh"

namespace xm {
namespace filt {

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

// 5x5 block offsets, nearest first: 4 sides, 4 corners, 16 outer cells
static __constant__ int8_t c_bx[24] = {0, 1, 0, -1,   1, -1, -1, 1,
                                0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1};
static __constant__ int8_t c_by[24] = {1, 0, -1, 0,   1, 1, -1, -1,
                                2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1, 2, 2};

__device__ __forceinline__ Rec load_rec(const Rec* p) {
    Rec r;
    unsigned long long w;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.c), "=l"(w) : "l"(p));
    r.idx = (uint32_t)w;
    r.key = (uint32_t)(w >> 32);
    return r;
}

// result of one source, padded to a DRAM sector (a scattered 8-byte store costs a read-modify-write
// of the sector); unpack_results_kernel splits the records into id[] / d[] with coalesced accesses
struct __align__(32) ResRec { uint64_t id; double d; uint64_t pad0, pad1; };
// record of a row the filter could not decide: the unpack pass leaves such rows alone, so that the
// exact kernel can write them while the unpack pass is still running (no catalog has 2^64 - 2 rows)
constexpr uint64_t kUndecidedId = 0xfffffffffffffffeull;

__device__ __forceinline__ void store_res(ResRec* p, uint64_t id, double d) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%3};" :: "l"(p), "l"(id), "l"(__double_as_longlong(d)), "l"(0ll) : "memory");
}

// Running state of one source. Q >= 0, so the high 32 bits of its IEEE pattern are a monotone
// 21-bit-mantissa image of it: the loop tracks the smallest and second smallest HIGH WORDS with
// native 32-bit integer min/max (FP64 min/max are multi-instruction selects on sm_100) and the slot
// of the smallest. Two metrics with different high words are strictly ordered; metrics sharing a
// high word (within 2^-20 of each other) are told apart by the exact kernel.
constexpr int32_t kInfHi = 0x7ff00000;

// smallest / largest double whose high word is h
__device__ __forceinline__ double hi_floor(int32_t h) { return __hiloint2double(h, 0); }
__device__ __forceinline__ double hi_ceil(int32_t h) {
    return h >= kInfHi ? __hiloint2double(kInfHi, 0) : __hiloint2double(h + 1, 0);
}

// Stop rule of qxmatch.hpp:336-340 on the small-angle forms. 0: stop for certain, 1: continue for
// certain, 2: inside the slack. q1 is only known to lie in [q1_lo, q1_hi).
__device__ __forceinline__ int stop_decision(double q1_lo, double q1_hi, double tr, const GridView& g,
                                             double extra_delta = 0.0) {
    const double lo = tr - (g.stop_delta + extra_delta), hi = tr + (g.stop_delta + extra_delta);
    if (lo > 0.0 && q1_hi*(1.0 + g.eps_hi) <= lo*lo*(1.0 - lo*lo*(1.0/12.0) - 1e-12)) return 0;
    if (hi <= 0.0) return q1_lo > 0.0 ? 1 : 2;
    if (q1_lo*(1.0 - g.eps_lo) > hi*hi) return 1;
    return 2;
}

}  // namespace filt
}  // namespace xm

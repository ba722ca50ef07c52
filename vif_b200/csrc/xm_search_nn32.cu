// Nearest-neighbour filter kernel, second generation (nth = 1, spherical, cell-binned path): the hot
// kernel of the headline configuration. Same contract as ring_nn_filter_kernel (xm_search_fast.cu):
// decide, for almost every source, which candidate the reference's ring search
// (qxmatch.hpp:293-380) returns, and hand every row that cannot be decided RIGOROUSLY to the exact
// kernel (xm_search_exact.cu). What changed is the arithmetic and the work shape:
//
//   * FP32 metric in a per-tile frame. A tile = up to 256 consecutive sources of ONE grid column.
//     Its 3-column neighbourhood (rows +-1) is converted once, while it is staged in shared memory,
//     to 16-byte records {U = u*G, V, G} with u = x - X0, V = y - Y0 (X0 = centre of the tile's column,
//     Y0 = centre of the staged rows; differences taken in FP64, then rounded) and G = sqrt(cos dec).
//     For a source (us, vs, cs = cos dec) with ks = 1/sqrt(cs):
//         a = U - us*G            (= dx*sqrt(c2))
//         b = V*ks - vs*ks        (= dy/sqrt(c1))
//         Q' = a*a + b*b          (= Q/c1,  Q = dy^2 + dx^2*c1*c2 the small-angle metric of the FP64 filter)
//     4 FP32 operations and one LDS.128 per candidate (the FP64 kernel: 6 half-rate operations, two loads).
//   * the low 10 bits of Q' are replaced by the candidate's position in the stage, so the running
//     state is two integers (smallest, second smallest) updated with three VIMNMX; the winner's slot
//     falls out of the smallest. The 2^-13 relative blur this adds is part of the error band.
//   * error band. With |a - a_true| <= ea and |b - b_true| <= eb (derived below) the true scaled distance
//     sqrt(Q/c1) lies in [s(1-rho) - e, s(1+rho) + e], s = sqrt(Q'), e = ea + eb; all decisions use these
//     intervals: winner certain <=> hi(s1)(1+eps) < lo(s2); a cell is skipped when its lower bound exceeds
//     hi(s1)^2; the reference's stop rule (qxmatch.hpp:336-340) is evaluated on [lo, hi] with a slack.
//     Rows that fail any test go to the exact kernel through a device-side list (~3 per 10^4 rows on
//     the benchmark catalogs).
//   * culling in the tile frame needs four gaps (to the edges of the source's own cell): 4 side tests,
//     4 corner tests, 1 test for the 16 outer cells of the 5x5 block (a row that needs one of those is
//     flagged: it takes an empty neighbourhood, ~1e-5 of the rows).
//   * the (source, neighbour cell) pairs that survive are pooled per CTA and dealt one per thread, as in
//     the FP64 kernel, but the pool holds only (owner, cell) codes: the taker resolves the slot range.
//
// Error budget (tile frame; wx, wy cell widths, Ymax = largest |y - Y0| in the stage <= 32.5 wy).
// sqrt and 1/x are the SFU approximations (relative error <= 2^-22 each):
//   u, us, V, vs are FP64 differences rounded to FP32: |err| <= 2^-24 |value| (|u| <= 1.5 wx, |us| <= 0.5 wx).
//   G = sqrt(float(c2)): relative 1.2*2^-22; U = u*G rounded: 2^-24; a = fma(-us, G, U): 2^-24.
//       |a - a_true| <= |dx| 1.45*2^-22 + 3.5 wx 2^-24 <= 3.8*2^-22 wx                      (|dx| <= 2 wx)
//   ks = 1/sqrt(float(c1)): relative 2.2*2^-22; vs*ks rounded; b = fma(V, ks, -vsk):
//       |b - b_true| <= ks (|dy| 9.8*2^-24 + 3 Ymax 2^-24) <= ks 2^-24 (19.6 wy + 3 Ymax)   (|dy| <= 2 wy)
//   Q' = fma(a, a, b*b): relative 2^-23 -> sqrt: 2^-24. The kernel uses twice these bounds:
//       ea = 2^-19 wx,  eb = ks 2^-23 (20 wy + 3 Ymax),  rho = 2^-22 (+ 2^-21 for the approximate sqrt).
#include <cstdlib>

#include "xm_common.cuh"
#include "xm_math_exact.cuh"
#include "xm_filter_common.cuh"

namespace xm {

namespace {

using namespace filt;

#ifndef XM_NN32_MINB
#define XM_NN32_MINB 4
#endif
constexpr int kMaxT = 256;               // sources per tile (CTA size); 128 and 64 for sparser catalogs
constexpr int kCap = 1024;               // staged candidate records (10 index bits)
constexpr int kRowsMax = 63;             // staged cell rows (incl. the two halo rows); < smallest CTA size
constexpr int kQ = 4;                    // neighbour cells a source may queue; rows that need more are flagged
constexpr int kItems = kMaxT*kQ;
constexpr int32_t kInfF = 0x7f800000;    // +inf as FP32 bits
constexpr int32_t kIdxMask = 0x3ff;

struct __align__(16) TilePlan32 {
    uint32_t p0, cnt;                    // sources [p0, p0 + cnt) of the cell-ordered catalog
    uint32_t run_start[3];               // the three column runs of the neighbourhood (slots of `cand`)
    uint32_t run_cnt[3];
    int32_t  ya;                         // first staged cell row
    uint32_t xs;                         // the tile's grid column
    uint32_t nrows;                      // staged rows; 0: the neighbourhood does not fit the stage
    float    cmin;                       // lower bound of cos(dec) over the staged rows
};
static_assert(sizeof(TilePlan32) == 48, "TilePlan32 layout");

// tiles per grid column: ceil(sources in the column / tile)
__global__ void __launch_bounds__(256)
col_tiles_kernel(GridView g, CatView src, uint32_t tile, uint32_t* __restrict__ ntile) {
    const uint32_t x = blockIdx.x*blockDim.x + threadIdx.x;
    if (x > g.nra) return;
    uint32_t v = 0;
    if (x < g.nra) {
        const uint32_t ns = src.cstart[(x + 1)*g.ndec] - src.cstart[x*g.ndec];
        v = (ns + tile - 1)/tile;
    }
    ntile[x] = v;                        // ntile[nra] = 0: the scan leaves the total there
}

__global__ void __launch_bounds__(256)
plan_tiles32_kernel(GridView g, CatView src, CatView cand, const uint32_t* __restrict__ toff, uint32_t tile,
                    uint32_t ntiles_max, TilePlan32* __restrict__ plan) {
    const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= ntiles_max) return;
    TilePlan32 pl;
    memset(&pl, 0, sizeof(pl));
    const uint32_t total = toff[g.nra];
    if (t < total) {
        uint32_t lo = 0, hi = g.nra;                     // toff[lo] <= t < toff[hi]
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (toff[mid] <= t) lo = mid; else hi = mid;
        }
        const uint32_t xs = lo;
        const uint32_t cb = src.cstart[xs*g.ndec], ce = src.cstart[(xs + 1)*g.ndec];
        const uint32_t p0 = cb + (t - toff[xs])*tile;
        const uint32_t cnt = min(tile, ce - p0);
        const int ndec = (int)g.ndec, nra = (int)g.nra;
        const int yf = (int)(src.rec[p0].key - xs*g.ndec), yl = (int)(src.rec[p0 + cnt - 1].key - xs*g.ndec);
        const int ya = max(yf - 1, 0), yb = min(yl + 1, ndec - 1);
        const int nrows = yb - ya + 1;
        pl.p0 = p0; pl.cnt = cnt; pl.ya = ya; pl.xs = xs;
        if (nrows <= kRowsMax) {
            uint32_t tot = 0;
            for (int c = 0; c < 3; ++c) {
                const int cx = (int)xs - 1 + c;
                if (cx < 0 || cx >= nra) continue;
                const uint32_t b0 = cand.cstart[(uint32_t)cx*g.ndec + (uint32_t)ya];
                const uint32_t b1 = cand.cstart[(uint32_t)cx*g.ndec + (uint32_t)(ya + nrows)];
                pl.run_start[c] = b0; pl.run_cnt[c] = b1 - b0;
                tot += b1 - b0;
            }
            const double cm = fmin(g.cmin_row[ya], g.cmin_row[yb]);    // cos is concave on [-pi/2, pi/2]
            if (tot <= (uint32_t)kCap && cm >= 0.0) {
                pl.nrows = (uint32_t)nrows;
                pl.cmin = __double2float_rd(cm);
            }
        }
    }
    plan[t] = pl;
}

struct Src32 { float us, vsk, ks; int ry; };

template <bool SELF>
__device__ __forceinline__ void eval32(const float4* __restrict__ s_c, uint32_t li, float us, float vsk, float ks,
                                       uint32_t self_li, int32_t& h1, int32_t& h2) {
    const float4 f = s_c[li];
    const float a = fmaf(-us, f.z, f.x);
    const float b = fmaf(f.y, ks, -vsk);
    const float Q = fmaf(a, a, b*b);
    int32_t h = (__float_as_int(Q) & ~kIdxMask) | (int32_t)li;
    if (SELF && li == self_li) h = kInfF;                // never pair a source with itself
    h2 = min(h2, max(h1, h));
    h1 = min(h1, h);
}

template <bool SELF>
__device__ __forceinline__ void scan32(const float4* __restrict__ s_c, uint32_t qs, uint32_t qe, float us, float vsk,
                                       float ks, uint32_t self_li, int32_t& h1, int32_t& h2) {
    uint32_t li = qs;
    for (; li + 2 <= qe; li += 2) {
        eval32<SELF>(s_c, li, us, vsk, ks, self_li, h1, h2);
        eval32<SELF>(s_c, li + 1, us, vsk, ks, self_li, h1, h2);
    }
    if (li < qe) eval32<SELF>(s_c, li, us, vsk, ks, self_li, h1, h2);
}

__device__ __forceinline__ float sqrt_apx(float x) {      // MUFU.SQRT: relative error <= 2^-22
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_apx(float x) {       // MUFU.RCP: relative error <= 2^-22
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// [lo, hi] of the true scaled distance sqrt(Q/c1) of a tracked value h (low bits = stage position)
__device__ __forceinline__ float dist_hi(int32_t h, float e) {
    return fmaf(sqrt_apx(__int_as_float(h | kIdxMask)), 1.0f + 0x1p-20f, e);
}
__device__ __forceinline__ float dist_lo(int32_t h, float e) {
    return fmaxf(fmaf(sqrt_apx(__int_as_float(h & ~kIdxMask)), 1.0f - 0x1p-20f, -e), 0.0f);
}

// Stop rule of qxmatch.hpp:336-340 on distances [rad]: the best distance is only known to lie in
// [d_lo, d_hi], the ring radius tr within +-delta. 0: stop for certain, 1: continue for certain, 2: undecided.
__device__ __forceinline__ int stop_decision32(float d_lo, float d_hi, float tr, float delta, float eps_lo,
                                               float eps_hi) {
    const float lo = tr - delta, hi = tr + delta;
    // proxy(best) <= sin^2(lo/2)  <=  4p <= Q(1+eps_hi) and 4 sin^2(lo/2) >= lo^2 (1 - lo^2/12)
    if (lo > 0.0f && d_hi*(1.0f + eps_hi + 0x1p-20f) <= lo*(1.0f - lo*lo*(1.0f/12.0f) - 0x1p-20f)) return 0;
    if (hi <= 0.0f) return d_lo > 0.0f ? 1 : 2;
    if (d_lo*(1.0f - eps_lo - 0x1p-20f) > hi) return 1;
    return 2;
}

// What remains to be done for a source after ring 0 (its own cell), given the smallest tracked value h1.
//   * mask: the neighbour cells (bit e as in the pooling step) that can hold a candidate nearer than, or within
//     the error band of, the best so far. Gaps to the own cell's edges in the tile frame, shrunk by the rounding
//     of the frame coordinates and of the binning (fzx, fzy).
//   * mask == 0 (and the 16 outer cells of the 5x5 block out of reach as well): nothing the reference visits
//     later, however deep it goes, can replace or tie the best so far (qxmatch.hpp:318 is a strict <), so the
//     row is decided without looking at the stop rule.
//   * mask != 0: the reference's stop rule after ring 0 (qxmatch.hpp:336-340, max_dist = cs/2) says whether it
//     looks at those cells at all; certain stop -> decided, certain continue -> scan them, else flag.
//     No stop rule is needed after ring 1: depth 1 of the reference is the whole 5x5 block and whatever it
//     visits beyond the masked cells is out of reach by the bounds above.
// Returns the mask; flag: the exact kernel must decide the row.
__device__ __forceinline__ uint32_t ring0_decide(int32_t h1, float us, float vs, float ks, float gs, float c1f,
                                                 float ccen, int ry, int nrows, float cminT, float wxf, float wyf,
                                                 float ymax, float eband, float eps, float eps_lo, float eps_hi,
                                                 float csr, float fzx, float fzy, float stop_delta, bool& flag) {
    const float hi1 = dist_hi(h1, eband);
    const float ylo = ((float)ry - 0.5f*(float)nrows)*wyf;               // lower edge of the own cell row
    const float mx = fzx + 0x1p-22f*wxf, my = fzy + 0x1p-22f*ymax;
    const float hwx = 0.5f*wxf;
    const float sc = sqrt_apx(cminT)*(1.0f - 0x1p-20f);
    const float ksl = ks*(1.0f - 0x1p-19f);
    const float xl = fmaxf((us + hwx) - mx, 0.0f)*sc, xr = fmaxf((hwx - us) - mx, 0.0f)*sc;
    const float yd = fmaxf((vs - ylo) - my, 0.0f)*ksl, yu = fmaxf(((ylo + wyf) - vs) - my, 0.0f)*ksl;
    const float T = (h1 >= kInfF) ? __int_as_float(kInfF) : hi1*hi1*(1.0f + 2.0f*eps + 0x1p-18f);
    const float xl2 = xl*xl, xr2 = xr*xr, yd2 = yd*yd, yu2 = yu*yu;
    uint32_t mask = (yu2 < T ? 1u : 0u) | (xr2 < T ? 2u : 0u) | (yd2 < T ? 4u : 0u) | (xl2 < T ? 8u : 0u) |
                    (xr2 + yu2 < T ? 16u : 0u) | (xl2 + yu2 < T ? 32u : 0u) | (xl2 + yd2 < T ? 64u : 0u) |
                    (xr2 + yd2 < T ? 128u : 0u);
    // rows outside the grid do not exist (qxmatch.hpp:309-311); columns outside it are empty runs
    if (ry + 1 >= nrows) mask &= ~(1u | 16u | 32u);
    if (ry - 1 < 0) mask &= ~(4u | 64u | 128u);
    // the 16 outer cells of the 5x5 block lie a full cell further
    const float wxs = wxf*(1.0f - 0x1p-20f)*sc, wys = wyf*(1.0f - 0x1p-20f)*ksl;
    const float o = fminf(fminf((xl + wxs)*(xl + wxs), (xr + wxs)*(xr + wxs)),
                          fminf((yd + wys)*(yd + wys), (yu + wys)*(yu + wys)));
    if (!(o >= T) || __popc(mask) > kQ) { flag = true; return 0u; }
    if (mask == 0u) return 0u;
    // distance to the cell centre (small-angle form of qxmatch.hpp:298-299)
    const float dyc = vs - (ylo + 0.5f*wyf);
    const float cd = sqrt_apx(fmaf(us*us*ccen, c1f, dyc*dyc));
    // slack of the ring radius: the FP64 filter's (small-angle forms) + the FP32 evaluation of cd
    const float sdelta = stop_delta + 1.2f*(0x1p-22f*(wxf + 2.0f*ymax) + 0x1p-19f*cd) + 0x1p-20f*csr;
    const float lo1 = dist_lo(h1, eband);
    const int dec = stop_decision32(lo1*gs*(1.0f - 0x1p-20f), hi1*gs*(1.0f + 0x1p-20f), 0.4f*csr - 1.1f*cd, sdelta,
                                    eps_lo, eps_hi);
    if (dec == 0) return 0u;                     // the reference stops after ring 0: the best so far is its answer
    if (dec == 2) { flag = true; return 0u; }
    return mask;
}

template <bool SELF>
__global__ void __launch_bounds__(kMaxT, XM_NN32_MINB)
ring_nn32_kernel(GridView g, CatView src, CatView cand, const TilePlan32* __restrict__ plan,
                 ResRec* __restrict__ res, uint32_t* __restrict__ refine_list,
                 unsigned long long* refine_count, int reverse, SearchCounters* ctr, uint32_t ntiles,
                 uint32_t ahead) {
    __shared__ float4   s_c[kCap];
    __shared__ uint4    s_next[3];
    __shared__ float4   s_src[kMaxT];
    __shared__ uint32_t s_cs[3][kRowsMax + 1];
    __shared__ float    s_ccen[kRowsMax + 1];
    __shared__ uint32_t s_item[kItems];
    __shared__ int2     s_part[kItems];
    __shared__ uint32_t s_wtot[kMaxT/32];

    const uint4* pp = reinterpret_cast<const uint4*>(plan + blockIdx.x);
    const uint4 pa = __ldg(pp), pb = __ldg(pp + 1), pc = __ldg(pp + 2);
    const uint32_t p0 = pa.x, cnt = pa.y;
    if (cnt == 0) return;                                // grid is an upper bound of the tile count
    // The tile that will run on this SM slot about one wave from now: its plan is fetched here (the load
    // is consumed at the end of the kernel) so that its records can be pulled into L2 ahead of time.
    const bool pf = threadIdx.x < 3 && blockIdx.x + ahead < ntiles;
    uint4 pn = make_uint4(0u, 0u, 0u, 0u);
    if (pf) pn = __ldg(reinterpret_cast<const uint4*>(plan + blockIdx.x + ahead) + threadIdx.x);
    const uint32_t rs0 = pa.z, rs1 = pa.w, rs2 = pb.x, rc0 = pb.y, rc1 = pb.z, rc2 = pb.w;
    const int ya = (int)pc.x;
    const uint32_t xs = pc.y;
    const int nrows = (int)pc.z;
    const float cminT = __uint_as_float(pc.w);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nthr = blockDim.x;
    const bool live = (uint32_t)tid < cnt;
    const uint32_t p = p0 + tid;
    Rec me;
    me.x = me.y = me.c = 0.0; me.idx = 0; me.key = 0;
    if (live) me = load_rec(src.rec + p);

    if (nrows == 0) {                                    // neighbourhood too large for the stage
        if (live) {
            const unsigned long long pos = atomicAdd(refine_count, 1ull);
            refine_list[pos] = p;
            store_res(res + me.idx, kUndecidedId, 0.0);
        }
        return;
    }

    // ---- 0. stage the neighbourhood as FP32 records in the tile frame ---------------------------------
    // All global loads of the prologue are issued before the first use of any of them.
    const double X0 = g.ox + ((double)xs + 0.5)*g.wx;
    const double Y0 = g.oy + ((double)ya + 0.5*(double)nrows)*g.wy;
    const uint32_t tot = rc0 + rc1 + rc2;
    uint32_t csv[3] = {0u, 0u, 0u};
    double ccv = 0.0;
    if (tid <= nrows) {
#pragma unroll
        for (int col = 0; col < 3; ++col) {
            const int cx = (int)xs - 1 + col;
            if (cx >= 0 && cx < (int)g.nra) csv[col] = __ldg(cand.cstart + (uint32_t)cx*g.ndec + (uint32_t)(ya + tid));
        }
        if (tid < nrows) ccv = __ldg(g.ccen_row + ya + tid);
    }
    for (uint32_t base = 0; base < tot; base += 4u*(uint32_t)nthr) {
        Rec r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = base + (uint32_t)(k*nthr + tid);
            r[k].x = r[k].y = 0.0; r[k].c = 1.0;
            if (t < tot) {
                const uint32_t gi = t < rc0 ? rs0 + t : (t < rc0 + rc1 ? rs1 + (t - rc0) : rs2 + (t - rc0 - rc1));
                r[k] = load_rec(cand.rec + gi);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = base + (uint32_t)(k*nthr + tid);
            if (t < tot) {
                const float u = __double2float_rn(r[k].x - X0), v = __double2float_rn(r[k].y - Y0);
                const float G = sqrt_apx(__double2float_rn(r[k].c));
                s_c[t] = make_float4(u*G, v, G, 0.0f);
            }
        }
    }
    if (tid <= nrows) {
        s_cs[0][tid] = csv[0]; s_cs[1][tid] = csv[1]; s_cs[2][tid] = csv[2];
        if (tid < nrows) s_ccen[tid] = __double2float_rn(ccv);
    }

    // stage position of slot q of column c: q + d[c]
    const uint32_t d0 = 0u - rs0, d1 = rc0 - rs1, d2 = rc0 + rc1 - rs2;

    // the source in the same frame
    const float wxf = __double2float_rn(g.wx), wyf = __double2float_rn(g.wy);
    const int ry = live ? (int)(me.key - xs*g.ndec) - ya : 0;
    const float us = __double2float_rn(me.x - X0), vs = __double2float_rn(me.y - Y0);
    const float c1f = live ? __double2float_rn(me.c) : 1.0f;
    const float gs = sqrt_apx(c1f);
    const float ks = rcp_apx(gs);
    const float vsk = vs*ks;
    const float ymax = 0.5f*(float)nrows*wyf*(1.0f + 0x1p-20f) + __double2float_ru(g.fzy);
    // error band of the scaled distance (header): twice the derived bounds
    const float eband = fmaf(ks, 0x1p-23f*(20.0f*wyf + 3.0f*ymax), 0x1p-19f*wxf)*(1.0f + 0x1p-10f);
    const float eps_lo = __double2float_ru(g.eps_lo), eps_hi = __double2float_ru(g.eps_hi);
    const float eps = eps_lo + eps_hi + 0x1p-20f;
    const float csr = __double2float_rn(g.cs_rad);
    if (threadIdx.x < 3) s_next[threadIdx.x] = pn;

    __syncthreads();

    // ---- 1. ring 0: own cell -------------------------------------------------------------------------------
    int32_t h1 = kInfF, h2 = kInfF;
    const uint32_t self_li = SELF ? p + d1 : 0xffffffffu;
    uint32_t evals = 0;
    bool flag = false;
    uint32_t mask = 0;
    if (live) {
        const uint32_t qs = s_cs[1][ry] + d1, qe = s_cs[1][ry + 1] + d1;
        scan32<SELF>(s_c, qs, qe, us, vsk, ks, self_li, h1, h2);
        evals += qe - qs;
        mask = ring0_decide(h1, us, vs, ks, gs, c1f, s_ccen[ry], ry, nrows, cminT, wxf, wyf, ymax, eband, eps, eps_lo,
                            eps_hi, csr, __double2float_ru(g.fzx), __double2float_ru(g.fzy),
                            __double2float_ru(g.stop_delta), flag);
    }

    // ---- 2. pool the surviving (source, cell) pairs ------------------------------------------------------
    s_src[tid] = make_float4(us, vsk, ks, __int_as_float(ry));
    const uint32_t nq = __popc(mask);
    uint32_t inc = nq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wtot[wid] = inc;
    __syncthreads();
    uint32_t first = inc - nq, total = 0;
#pragma unroll
    for (int k = 0; k < kMaxT/32; ++k) {
        const uint32_t w = k < nthr/32 ? s_wtot[k] : 0u;
        if (k < wid) first += w;
        total += w;
    }
    {
        uint32_t m = mask, k = first;
        while (m) {
            const uint32_t e = __ffs(m) - 1;
            m &= m - 1;
            s_item[k++] = (uint32_t)tid | (e << 8);
        }
    }
    __syncthreads();

    // ---- L2 prefetch of the records of tile (blockIdx + ahead): 128-byte lines, spread over the CTA ----------
    {
        const uint4 na = s_next[0], nb = s_next[1];
        if (na.y) {                                       // cnt of that tile
            const uint32_t l0 = (nb.y + 3u) >> 2, l1 = (nb.z + 3u) >> 2, l2 = (nb.w + 3u) >> 2, ls = (na.y + 3u) >> 2;
            for (uint32_t t = tid; t < l0 + l1 + l2 + ls; t += nthr) {
                const Rec* q;
                if (t < l0) q = cand.rec + na.z + 4u*t;
                else if (t < l0 + l1) q = cand.rec + na.w + 4u*(t - l0);
                else if (t < l0 + l1 + l2) q = cand.rec + nb.x + 4u*(t - l0 - l1);
                else q = src.rec + na.x + 4u*(t - l0 - l1 - l2);
                asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
            }
        }
    }

    // ---- 3. one pooled pair per thread ---------------------------------------------------------------------
    for (uint32_t it = tid; it < total; it += nthr) {
        const uint32_t item = s_item[it];
        const uint32_t owner = item & 0xffu, e = item >> 8;
        // e: 0 (0,+1) 1 (+1,0) 2 (0,-1) 3 (-1,0) 4 (+1,+1) 5 (-1,+1) 6 (-1,-1) 7 (+1,-1)
        const int bx = (int)((0x20020121u >> (4*e)) & 3u) - 1;
        const int by = (int)((0x00221012u >> (4*e)) & 3u) - 1;
        const float4 o = s_src[owner];
        const int r = __float_as_int(o.w) + by;
        const uint32_t dd = bx < 0 ? d0 : (bx > 0 ? d2 : d1);
        const uint32_t qs = s_cs[bx + 1][r] + dd, qe = s_cs[bx + 1][r + 1] + dd;
        int32_t a1 = kInfF, a2 = kInfF;
        scan32<false>(s_c, qs, qe, o.x, o.y, o.z, 0xffffffffu, a1, a2);
        evals += qe - qs;
        s_part[it] = make_int2(a1, a2);
    }
    __syncthreads();
    for (uint32_t k = 0; k < nq; ++k) {
        const int2 a = s_part[first + k];
        h2 = min(min(h2, a.y), max(h1, a.x));
        h1 = min(h1, a.x);
    }

    // ---- 4. decide -----------------------------------------------------------------------------------------
    if (live) {
        const float hi1 = dist_hi(h1, eband);
        if (h1 >= kInfF) flag = true;
        // the winner must beat the runner-up by more than the band
        if (!flag && h2 < kInfF && !(hi1*(1.0f + eps) < dist_lo(h2, eband))) flag = true;

        if (flag) {
            const unsigned long long pos = atomicAdd(refine_count, 1ull);
            refine_list[pos] = p;
            store_res(res + me.idx, kUndecidedId, 0.0);
        } else {
            const uint32_t li = (uint32_t)(h1 & kIdxMask);
            const uint32_t slot = li < rc0 ? rs0 + li : (li < rc0 + rc1 ? rs1 + (li - rc0) : rs2 + (li - rc0 - rc1));
            // exact proxy of the winning pair in reference-order arithmetic -> arcsec
            const Rec w = load_rec(cand.rec + slot);
            const double pr = proxy_sph_small(me.x, me.y, me.c, w.x, w.y, w.c, g.sin_model);
            store_res(res + me.idx, (uint64_t)w.idx, proxy_to_true_small(pr));
        }
    }
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, o);
    if (lane == 0 && evals) atomicAdd(reverse ? &ctr->evals_rev : &ctr->evals_fwd, (unsigned long long)evals);
}

// ---- persistent, warp-specialised form (tiles of kCons = 192 sources) ------------------------------------------
// ncu on the kernel above (profiles/r2_nn32_kernel_ncu.txt): a third of the stall samples sit on the two
// dependent global-memory latencies at the head of every CTA (plan -> records) and another fifth on the
// barriers behind them; occupancy is register-bound (5 CTAs of 192), so more warps cannot hide them.
// Here a CTA lives for the whole launch and walks tiles blockIdx, blockIdx + grid, ...:
//   * 2 PRODUCER warps fetch tile t+1 (plan, candidate runs, source records, CSR offsets, row table),
//     convert it to the FP32 tile frame and fill the other of two shared-memory stages;
//   * 6 CONSUMER warps (one thread per source) run rings 0 and 1 on the current stage;
//   * full / empty hand-over through named barriers (bar.arrive / bar.sync), consumer-only barriers
//     for the pooling steps. The consumers never wait for global memory except for the winner's
//     record in the epilogue (their own record is requested when the tile starts and used at the end).
#ifndef XM_NN32_CONS
#define XM_NN32_CONS 192
#endif
constexpr int kCons = XM_NN32_CONS, kProd = 256 - kCons, kPThreads = kCons + kProd;
#ifndef XM_NN32_PB
#define XM_NN32_PB 3
#endif
constexpr int kPB = XM_NN32_PB;        // candidate records a producer thread keeps in flight
constexpr int kSrcPer = (kCons + kProd - 1)/kProd;    // source records per producer thread
constexpr int kBarFull = 1, kBarEmpty = 3, kBarCons = 5;          // named barriers: full[2] = 1,2; empty[2] = 3,4

struct __align__(16) Stage32 {
    float4   c[kCap];                    // candidates {U, V, G, -}
    float4   a[kCons];                   // sources {us, vsk, ks, ry}
    float4   b[kCons];                   // sources {vs, c1f, gs, -}
    uint4    plan[3];
    uint32_t cs[3][kRowsMax + 1];
    float    ccen[kRowsMax + 1];
};

struct __align__(16) Persist32 {
    Stage32  st[2];
    uint32_t item[kCons*kQ];
    int2     part[kCons*kQ];
    uint32_t wtot[kCons/32];
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory"); }

#ifndef XM_NN32_PCTAS
#define XM_NN32_PCTAS 4
#endif
constexpr int kPCtas = XM_NN32_PCTAS;  // persistent CTAs per SM (register-bound: 4 x 256 threads x 64 registers)

template <bool SELF, bool PF>
__global__ void __launch_bounds__(kPThreads, kPCtas)
ring_nn32p_kernel(GridView g, CatView src, CatView cand, const TilePlan32* __restrict__ plan,
                  const uint32_t* __restrict__ ntiles_dev, ResRec* __restrict__ res,
                  uint32_t* __restrict__ refine_list, unsigned long long* refine_count, int reverse,
                  SearchCounters* ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Persist32& sm = *reinterpret_cast<Persist32*>(smem_raw);
    const uint32_t ntiles = *ntiles_dev;
    const int tid = threadIdx.x;
    const float wxf = __double2float_rn(g.wx), wyf = __double2float_rn(g.wy);

    if (tid >= kCons) {
        // ================================ producers ================================
        // The plan of the next tile is requested while the current one is staged; a tile's own loads go out in
        // batches of four 32-byte records per thread (deeper batches spill: measured slower).
        const int pt = tid - kCons;
        int buf = 0;
        const uint4* pp0 = reinterpret_cast<const uint4*>(plan + blockIdx.x);
        uint4 pa = make_uint4(0, 0, 0, 0), pb = pa, pc = pa;
        if (blockIdx.x < ntiles) { pa = __ldg(pp0); pb = __ldg(pp0 + 1); pc = __ldg(pp0 + 2); }
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, buf ^= 1) {
            Stage32& S = sm.st[buf];
            const uint4 qa = pa, qb = pb, qc = pc;
            if (t + gridDim.x < ntiles) {
                const uint4* pn = reinterpret_cast<const uint4*>(plan + t + gridDim.x);
                pa = __ldg(pn); pb = __ldg(pn + 1); pc = __ldg(pn + 2);
            }
            const uint32_t p0 = qa.x, cnt = qa.y;
            const uint32_t rs0 = qa.z, rs1 = qa.w, rs2 = qb.x, rc0 = qb.y, rc1 = qb.z, rc2 = qb.w;
            const int ya = (int)qc.x;
            const uint32_t xs = qc.y;
            const int nrows = (int)qc.z;
            bar_sync(kBarEmpty + buf, kPThreads);                    // the consumers are done with this stage
            if (pt < 3) S.plan[pt] = pt == 0 ? qa : (pt == 1 ? qb : qc);
            if (nrows > 0) {
                const double X0 = g.ox + ((double)xs + 0.5)*g.wx;
                const double Y0 = g.oy + ((double)ya + 0.5*(double)nrows)*g.wy;
                const uint32_t tot = rc0 + rc1 + rc2;
                uint32_t csv[3] = {0u, 0u, 0u};
                double ccv = 0.0;
                if (pt <= nrows) {
#pragma unroll
                    for (int col = 0; col < 3; ++col) {
                        const int cx = (int)xs - 1 + col;
                        if (cx >= 0 && cx < (int)g.nra) csv[col] = __ldg(cand.cstart + (uint32_t)cx*g.ndec + (uint32_t)(ya + pt));
                    }
                    if (pt < nrows) ccv = __ldg(g.ccen_row + ya + pt);
                }
                // the tile's sources: raw record -> frame parameters. All records of a thread are requested at once
                // (one round trip instead of kSrcPer: the producers' chain of dependent latencies is the critical
                // path of the kernel)
                {
                    Rec sr[kSrcPer];
#pragma unroll
                    for (int j = 0; j < kSrcPer; ++j) {
                        const uint32_t k = (uint32_t)(j*kProd + pt);
                        sr[j].x = sr[j].y = 0.0; sr[j].c = 1.0; sr[j].idx = 0; sr[j].key = 0;
                        if (k < cnt) sr[j] = load_rec(src.rec + p0 + k);
                    }
#pragma unroll
                    for (int j = 0; j < kSrcPer; ++j) {
                        const uint32_t k = (uint32_t)(j*kProd + pt);
                        if (k < cnt) {
                            const float us = __double2float_rn(sr[j].x - X0), vs = __double2float_rn(sr[j].y - Y0);
                            const float c1f = __double2float_rn(sr[j].c);
                            const float gs = sqrt_apx(c1f);
                            const float ks = rcp_apx(gs);
                            const int ry = (int)(sr[j].key - xs*g.ndec) - ya;
                            S.a[k] = make_float4(us, vs*ks, ks, __int_as_float(ry));
                            S.b[k] = make_float4(vs, c1f, gs, 0.0f);
                        }
                    }
                }
                for (uint32_t base = 0; base < tot; base += (uint32_t)(kPB*kProd)) {
                    Rec r[kPB];
#pragma unroll
                    for (int k = 0; k < kPB; ++k) {
                        const uint32_t q = base + (uint32_t)(k*kProd + pt);
                        r[k].x = r[k].y = 0.0; r[k].c = 1.0;
                        if (q < tot) {
                            const uint32_t gi = q < rc0 ? rs0 + q : (q < rc0 + rc1 ? rs1 + (q - rc0) : rs2 + (q - rc0 - rc1));
                            r[k] = load_rec(cand.rec + gi);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kPB; ++k) {
                        const uint32_t q = base + (uint32_t)(k*kProd + pt);
                        if (q < tot) {
                            const float u = __double2float_rn(r[k].x - X0), v = __double2float_rn(r[k].y - Y0);
                            const float G = sqrt_apx(__double2float_rn(r[k].c));
                            S.c[q] = make_float4(u*G, v, G, 0.0f);
                        }
                    }
                }
                if (pt <= nrows) {
                    S.cs[0][pt] = csv[0]; S.cs[1][pt] = csv[1]; S.cs[2][pt] = csv[2];
                    if (pt < nrows) S.ccen[pt] = __double2float_rn(ccv);
                }
            }
            bar_arrive(kBarFull + buf, kPThreads);
            if (PF && t + gridDim.x < ntiles && pa.y) {
                // The next tile's records are pulled into L2 now (its plan has landed while this tile was staged):
                // 128-byte lines (4 records) covering the three candidate runs and the source range. No registers
                // are held for it, and the loads of the next iteration find at least their later batches in L2.
                const uint32_t b0 = pa.z >> 2, l0 = ((pa.z + pb.y + 3u) >> 2) - b0;
                const uint32_t b1 = pa.w >> 2, l1 = ((pa.w + pb.z + 3u) >> 2) - b1;
                const uint32_t b2 = pb.x >> 2, l2 = ((pb.x + pb.w + 3u) >> 2) - b2;
                const uint32_t bs = pa.x >> 2, ls = ((pa.x + pa.y + 3u) >> 2) - bs;
                for (uint32_t l = pt; l < l0 + l1 + l2 + ls; l += kProd) {
                    const Rec* q;
                    if (l < l0) q = cand.rec + 4u*(b0 + l);
                    else if (l < l0 + l1) q = cand.rec + 4u*(b1 + (l - l0));
                    else if (l < l0 + l1 + l2) q = cand.rec + 4u*(b2 + (l - l0 - l1));
                    else q = src.rec + 4u*(bs + (l - l0 - l1 - l2));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
                }
            }
        }
        return;
    }

    // ================================ consumers ================================
    const int lane = tid & 31, wid = tid >> 5;
    bar_arrive(kBarEmpty + 0, kPThreads);
    bar_arrive(kBarEmpty + 1, kPThreads);
    const float eps_lo = __double2float_ru(g.eps_lo), eps_hi = __double2float_ru(g.eps_hi);
    const float eps = eps_lo + eps_hi + 0x1p-20f;
    const float csr = __double2float_rn(g.cs_rad);
    const float fzxf = __double2float_ru(g.fzx), fzyf = __double2float_ru(g.fzy), sdel0 = __double2float_ru(g.stop_delta);
    uint32_t evals = 0;
    int buf = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, buf ^= 1) {
        Stage32& S = sm.st[buf];
        bar_sync(kBarFull + buf, kPThreads);
        const uint4 pa = S.plan[0], pb = S.plan[1], pc = S.plan[2];
        const uint32_t p0 = pa.x, cnt = pa.y;
        const uint32_t rs0 = pa.z, rs1 = pa.w, rs2 = pb.x, rc0 = pb.y, rc1 = pb.z;
        const int nrows = (int)pc.z;
        const float cminT = __uint_as_float(pc.w);
        const bool live = (uint32_t)tid < cnt;
        const uint32_t p = p0 + tid;
        Rec me;
        me.x = me.y = me.c = 0.0; me.idx = 0; me.key = 0;
        if (live) me = load_rec(src.rec + p);                     // used in the epilogue only
        if (nrows == 0) {                                          // neighbourhood too large for the stage
            bar_arrive(kBarEmpty + buf, kPThreads);
            if (live) {
                const unsigned long long pos = atomicAdd(refine_count, 1ull);
                refine_list[pos] = p;
                store_res(res + me.idx, kUndecidedId, 0.0);
            }
            continue;
        }
        const uint32_t d0 = 0u - rs0, d1 = rc0 - rs1, d2 = rc0 + rc1 - rs2;
        const float4 A = S.a[live ? tid : 0], B = S.b[live ? tid : 0];
        const float us = A.x, vsk = A.y, ks = A.z, vs = B.x, c1f = B.y, gs = B.z;
        const int ry = __float_as_int(A.w);
        const float ymax = 0.5f*(float)nrows*wyf*(1.0f + 0x1p-20f) + __double2float_ru(g.fzy);
        const float eband = fmaf(ks, 0x1p-23f*(20.0f*wyf + 3.0f*ymax), 0x1p-19f*wxf)*(1.0f + 0x1p-10f);

        // ---- ring 0 ----
        int32_t h1 = kInfF, h2 = kInfF;
        const uint32_t self_li = SELF ? p + d1 : 0xffffffffu;
        bool flag = false;
        uint32_t mask = 0;
        if (live) {
            const uint32_t qs = S.cs[1][ry] + d1, qe = S.cs[1][ry + 1] + d1;
            scan32<SELF>(S.c, qs, qe, us, vsk, ks, self_li, h1, h2);
            evals += qe - qs;
            mask = ring0_decide(h1, us, vs, ks, gs, c1f, S.ccen[ry], ry, nrows, cminT, wxf, wyf, ymax, eband, eps, eps_lo,
                                eps_hi, csr, fzxf, fzyf, sdel0, flag);
        }

        // ---- pool the surviving (source, cell) pairs (consumer-only barriers) ----
        const uint32_t nq = __popc(mask);
        uint32_t inc = nq;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) sm.wtot[wid] = inc;
        bar_sync(kBarCons, kCons);
        uint32_t first = inc - nq, total = 0;
#pragma unroll
        for (int k = 0; k < kCons/32; ++k) {
            const uint32_t w = sm.wtot[k];
            if (k < wid) first += w;
            total += w;
        }
        {
            uint32_t m = mask, k = first;
            while (m) {
                const uint32_t e = __ffs(m) - 1;
                m &= m - 1;
                sm.item[k++] = (uint32_t)tid | (e << 8);
            }
        }
        bar_sync(kBarCons, kCons);
        for (uint32_t it = tid; it < total; it += kCons) {
            const uint32_t item = sm.item[it];
            const uint32_t owner = item & 0xffu, e = item >> 8;
            const int bx = (int)((0x20020121u >> (4*e)) & 3u) - 1;
            const int by = (int)((0x00221012u >> (4*e)) & 3u) - 1;
            const float4 o = S.a[owner];
            const int r = __float_as_int(o.w) + by;
            const uint32_t dd = bx < 0 ? d0 : (bx > 0 ? d2 : d1);
            const uint32_t qs = S.cs[bx + 1][r] + dd, qe = S.cs[bx + 1][r + 1] + dd;
            int32_t a1 = kInfF, a2 = kInfF;
            scan32<false>(S.c, qs, qe, o.x, o.y, o.z, 0xffffffffu, a1, a2);
            evals += qe - qs;
            sm.part[it] = make_int2(a1, a2);
        }
        bar_sync(kBarCons, kCons);
        bar_arrive(kBarEmpty + buf, kPThreads);                    // the stage may be refilled from here on
        for (uint32_t k = 0; k < nq; ++k) {
            const int2 a = sm.part[first + k];
            h2 = min(min(h2, a.y), max(h1, a.x));
            h1 = min(h1, a.x);
        }

        // ---- decide ----
        if (live) {
            const float hi1 = dist_hi(h1, eband);
            if (h1 >= kInfF) flag = true;
            if (!flag && h2 < kInfF && !(hi1*(1.0f + eps) < dist_lo(h2, eband))) flag = true;
            if (flag) {
                const unsigned long long pos = atomicAdd(refine_count, 1ull);
                refine_list[pos] = p;
                store_res(res + me.idx, kUndecidedId, 0.0);
            } else {
                const uint32_t li = (uint32_t)(h1 & kIdxMask);
                const uint32_t slot = li < rc0 ? rs0 + li : (li < rc0 + rc1 ? rs1 + (li - rc0) : rs2 + (li - rc0 - rc1));
                const Rec w = load_rec(cand.rec + slot);
                const double pr = proxy_sph_small(me.x, me.y, me.c, w.x, w.y, w.c, g.sin_model);
                store_res(res + me.idx, (uint64_t)w.idx, proxy_to_true_small(pr));
            }
        }
        // the merge above read sm.part of this tile: the next tile's pooling must not overwrite it earlier
        bar_sync(kBarCons, kCons);
    }
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, o);
    if (lane == 0 && evals) atomicAdd(reverse ? &ctr->evals_rev : &ctr->evals_fwd, (unsigned long long)evals);
}

}  // namespace

// Sources per tile for catalogs of n_src / n_cand rows on grid g, 0 when the FP32 kernel does not apply:
// a tile's neighbourhood (3 columns x (rows + 2)) must fit the stage with room for fluctuations.
int nn32_tile_sources(const GridView& g, uint64_t n_src, uint64_t n_cand) {
    if (!fast_path_applicable(g)) return 0;
    const double ncell = (double)g.nra*(double)g.ndec;
    if (ncell <= 0 || n_src == 0 || n_cand == 0) return 0;
    const double a1 = (double)n_src/ncell, a2 = (double)n_cand/ncell;
    if (const char* e = getenv("XM_NN32_TILE")) {        // experiments: force a tile size (multiple of 32)
        const int t = atoi(e);
        if (t >= 64 && t <= kMaxT && t % 32 == 0) return t;
    }
    // staged candidates of a tile: mean 3*rows*a2 with rows = tile/a1 + 2; the spread comes from the
    // Poisson counts and from the number of rows a fixed number of sources spans. 4.5 sigma of head room.
    for (int tile = kMaxT; tile >= 64; tile -= 32) {
        const double rows = (double)tile/a1 + 2.0;
        const double mean = 3.0*rows*a2;
        const double sd = sqrt(mean) + 3.0*a2*sqrt((double)tile)/a1;
        if (rows + 4.5*sqrt((double)tile)/a1 <= (double)kRowsMax && mean + 4.5*sd <= (double)kCap) return tile;
    }
    return 0;
}

int nn32_persistent_tile() {
    const char* pe = getenv("XM_NN32_PERSIST");
    return (pe && pe[0] == '0') ? 0 : kCons;
}

size_t nn32_plan_bytes(uint64_t n_src, uint32_t nra, int tile) {
    const uint64_t ntiles = n_src/(uint64_t)tile + nra + 1;
    return (size_t)ntiles*sizeof(TilePlan32) + ((size_t)nra + 2)*sizeof(uint32_t) + bucket_scan_aux_bytes(nra + 1) + 64;
}

int32_t launch_ring_nn32(const GridView& g, const CatView& src, const CatView& cand, bool self, int tile,
                         uint64_t* id_out, double* d_out, void* result_scratch, void* plan_scratch,
                         uint32_t* refine_list, unsigned long long* refine_count, bool reverse,
                         SearchCounters* ctr, cudaStream_t s, cudaEvent_t filter_done, Error& err,
                         uint64_t* launches, cudaEvent_t k0, cudaEvent_t k1, cudaEvent_t after, bool defer_unpack) {
    if (src.n == 0) return XM_OK;
    const uint32_t ntiles = src.n/(uint32_t)tile + g.nra + 1;
    TilePlan32* plan = (TilePlan32*)plan_scratch;
    uint32_t* toff = (uint32_t*)(plan + ntiles);
    uint32_t* aux = toff + g.nra + 2;
    ResRec* res = (ResRec*)result_scratch;
    col_tiles_kernel<<<(g.nra + 1 + 255)/256, 256, 0, s>>>(g, src, (uint32_t)tile, toff);
    *launches += 1;
    launch_exclusive_scan_u32(toff, g.nra + 1, aux, s, launches);
    plan_tiles32_kernel<<<(ntiles + 255)/256, 256, 0, s>>>(g, src, cand, toff, (uint32_t)tile, ntiles, plan);
    if (tile == nn32_persistent_tile()) {
        // persistent, warp-specialised form: 4 CTAs per SM for the whole launch
        int dev = 0, sms = 148;
        XM_CUDA_TRY(cudaGetDevice(&dev));
        XM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const uint32_t grid = (uint32_t)(sms*kPCtas) < ntiles ? (uint32_t)(sms*kPCtas) : ntiles;
        // XM_NN32_PF=1: with the L2 prefetch of the next tile's records (measured: see DESIGN.md 4.2)
        static const bool pf = [] { const char* e = getenv("XM_NN32_PF"); return e && e[0] == '1'; }();
        auto kern = self ? (pf ? ring_nn32p_kernel<true, true> : ring_nn32p_kernel<true, false>)
                         : (pf ? ring_nn32p_kernel<false, true> : ring_nn32p_kernel<false, false>);
        XM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Persist32)));
        // `after`: the other direction's persistent launch, which fills every SM on its own -- running the two one
        // after the other costs nothing and lets the first lane's results leave for the host early
        if (after) XM_CUDA_TRY(cudaStreamWaitEvent(s, after, 0));
        if (k0) XM_CUDA_TRY(cudaEventRecord(k0, s));
        kern<<<grid, kPThreads, sizeof(Persist32), s>>>(g, src, cand, plan, toff + g.nra, res, refine_list, refine_count,
                                                         reverse ? 1 : 0, ctr);
        *launches += 2;
        if (k1) XM_CUDA_TRY(cudaEventRecord(k1, s));
        if (filter_done) XM_CUDA_TRY(cudaEventRecord(filter_done, s));
        if (defer_unpack) { XM_CUDA_TRY(cudaGetLastError()); return XM_OK; }      // the caller unpacks later
        int32_t rcp = launch_unpack_results(result_scratch, src.n, id_out, d_out, s, err, launches);
        if (rcp) return rcp;
        XM_CUDA_TRY(cudaGetLastError());
        return XM_OK;
    }
    // prefetch distance: about one wave of resident CTAs (XM_NN32_AHEAD overrides; 0 switches it off)
    uint32_t ahead = 148u*4u;
    if (const char* e = getenv("XM_NN32_AHEAD")) ahead = (uint32_t)atoi(e);
    if (ahead == 0) ahead = ntiles;
    if (k0) XM_CUDA_TRY(cudaEventRecord(k0, s));
    if (self) ring_nn32_kernel<true><<<ntiles, tile, 0, s>>>(g, src, cand, plan, res, refine_list, refine_count,
                                                              reverse ? 1 : 0, ctr, ntiles, ahead);
    else ring_nn32_kernel<false><<<ntiles, tile, 0, s>>>(g, src, cand, plan, res, refine_list, refine_count,
                                                          reverse ? 1 : 0, ctr, ntiles, ahead);
    *launches += 2;
    if (k1) XM_CUDA_TRY(cudaEventRecord(k1, s));
    if (filter_done) XM_CUDA_TRY(cudaEventRecord(filter_done, s));
    int32_t rc = launch_unpack_results(result_scratch, src.n, id_out, d_out, s, err, launches);
    if (rc) return rc;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm

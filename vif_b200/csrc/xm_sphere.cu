// Exact nth-nearest search on the sphere with a wrap- and pole-aware grid of our own.
//
// Why it exists: the reference's cell-binned path cannot run on full-sky catalogs (the hull area
// of an RA-360 bounding box is 0 -> 1-arcsec cells -> ~10^12 buckets), and
// its brute_force path is O(n1*n2). This kernel returns EXACTLY what brute_force returns
// (qxmatch.hpp:490-526: the nth smallest proxies, equal proxies in ascending catalog-2 row order,
// reverse match = smallest catalog-1 row on ties) at the cost of a neighbourhood search, so that
// BASELINE config 5 (1e9 x 1e8 full sky) has a GPU path whose oracle is the reference's brute force.
//
// Grid: Dec rows of height h, RA columns of width w = 360/ncols (one width for all rows), columns
// taken modulo ncols when the catalogs span more than half the circle (the proxy only sees
// sin^2(dRA/2), which is 2*pi-periodic, so wrapped neighbours need no special arithmetic).
// Search of one source: rows outward from its own (alternating above / below), inside a row the
// cells outward from its own column (alternating east / west). A direction is abandoned as soon as
// a rigorous lower bound of the proxy over the next cell,
//      lb = sin^2(gap_dec/2) + sin^2(gap_ra/2) * cos(dec1) * min_row cos(dec),
// reaches the current nth-best; rows touching a pole have min cos = 0 and are scanned whole, which
// is what makes the result exact there. Every surviving candidate is evaluated with the
// reference-order proxy (xm_math_exact.cuh) and inserted by (proxy, row) lexicographic order.
//
// nth = 1 on small cells runs sphere_nn_filter_kernel first (a cheap, rigorous decision from the
// 3x3 block around the source) and the exact kernel only on the rows that kernel cannot decide.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"
#include "xm_filter_common.cuh"

namespace xm {

namespace {

constexpr int kThreads = 128;
constexpr uint32_t kNone = 0xffffffffu;

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

__device__ __forceinline__ double sin2_lower(double h) {       // h >= 0
    if (h < 0.02) { const double hh = h*h; return hh*(1.0 - hh*(1.0/3.0)); }
    if (h <= 1.5707963267948966) { const double s = sin(h); return s*s; }
    return 1.0;                                                 // callers clamp gaps to <= pi
}

// K best (proxy, row) pairs in lexicographic order
template <int K>
struct LexList {
    double d[K];
    uint32_t i[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < K; ++k) { d[k] = dinf(); i[k] = kNone; }
    }
    __device__ __forceinline__ double kth() const { return d[K-1]; }
    __device__ __forceinline__ void insert(double sd, uint32_t row) {
        if (sd < d[K-1] || (sd == d[K-1] && row < i[K-1])) {
            d[K-1] = sd; i[K-1] = row;
#pragma unroll
            for (int k = K-2; k >= 0; --k) {
                if (d[k] > d[k+1] || (d[k] == d[k+1] && i[k] > i[k+1])) {
                    const double td = d[k]; d[k] = d[k+1]; d[k+1] = td;
                    const uint32_t ti = i[k]; i[k] = i[k+1]; i[k+1] = ti;
                }
            }
        }
    }
};

struct SphereGrid {
    GridView g;          // rra0/rdec0/dra/ddec/nra/ndec as for any cell index (degrees)
    int32_t  wrap;       // columns are taken modulo nra
    const double* cmin_row;   // lower bound of cos(dec) over each row, 0 where the row touches a pole
};

template <int K>
__global__ void __launch_bounds__(kThreads)
sphere_search_kernel(SphereGrid sg, CatView src, CatView cand, int self_mode, uint32_t nth_eff,
                     uint64_t plane_stride, uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                     const uint32_t* __restrict__ rows, const unsigned long long* __restrict__ nrows_dev,
                     int reverse, SearchCounters* ctr) {
    const GridView& g = sg.g;
    // rows == nullptr: thread t handles sorted slot t; otherwise a grid-stride walk over the row list
    // whose length lives on the device (rows the filter kernel could not decide)
    const uint32_t nrows = rows ? (uint32_t)min((unsigned long long)src.n, *nrows_dev) : src.n;
    const uint32_t stride = gridDim.x*blockDim.x;
    unsigned long long evals = 0;
    for (uint32_t t = blockIdx.x*blockDim.x + threadIdx.x; t < nrows; t += stride) {
        const uint32_t p = rows ? rows[t] : t;
        const Rec me = src.rec[p];
        const double x1 = me.x, y1 = me.y, c1 = me.c < 0.0 ? 0.0 : me.c;
        const int ncol = (int)g.nra, nrow = (int)g.ndec;
        const int x0 = (int)(me.key/g.ndec), y0 = (int)(me.key - (uint32_t)x0*g.ndec);
        const double wx = g.wx, wy = g.wy;           // cell widths in radians
        LexList<K> L;
        L.init();

        bool up_done = false, dn_done = false;
        for (int step = 0; !(up_done && dn_done); ++step) {
            // rows in the order 0, +1, -1, +2, -2, ...
            const int dy = step == 0 ? 0 : ((step & 1) ? (step + 1)/2 : -(step/2));
            if (dy > 0 && up_done) continue;
            if (dy < 0 && dn_done) continue;
            const int row = y0 + dy;
            if (row < 0) { dn_done = true; continue; }
            if (row >= nrow) { up_done = true; continue; }
            // declination gap to the row (with the binning slack), 0 for the own row
            const double r_lo = g.oy + (double)row*wy, r_hi = r_lo + wy;
            double gd = fmax(r_lo - y1, y1 - r_hi) - g.fzy; gd = gd > 0.0 ? gd : 0.0;
            const double A = sin2_lower(0.5*gd)*(1.0 - 1e-9);
            if (A >= L.kth() && dy != 0) {
                if (dy > 0) up_done = true; else dn_done = true;
                continue;
            }
            const double cm = sg.cmin_row[row];
            const double cc = c1*cm*(1.0 - 1e-9);
            // cells of this row outward from the own column: 0, +1, -1, +2, ...
            const int half = sg.wrap ? ncol/2 + 1 : ncol;
            bool e_done = false, w_done = false;
            for (int cs = 0; !(e_done && w_done); ++cs) {
                const int dx = cs == 0 ? 0 : ((cs & 1) ? (cs + 1)/2 : -(cs/2));
                if (dx > 0 && e_done) continue;
                if (dx < 0 && w_done) continue;
                if (dx > half) { e_done = true; continue; }
                if (dx < -half) { w_done = true; continue; }
                int col = x0 + dx;
                if (sg.wrap) {
                    // a full turn brings us back to visited cells: each side covers half the circle
                    if (dx > 0 && 2*dx > ncol) { e_done = true; continue; }
                    if (dx < 0 && -2*dx >= ncol) { w_done = true; continue; }
                    col %= ncol; if (col < 0) col += ncol;
                } else {
                    if (col < 0) { w_done = true; continue; }
                    if (col >= ncol) { e_done = true; continue; }
                }
                // RA gap from the source to the cell, measured the way we walked (|dx| cells away)
                double gr = 0.0;
                if (dx != 0) {
                    const double edge = dx > 0 ? (g.ox + (double)(x0 + dx)*wx) - x1 : x1 - (g.ox + (double)(x0 + dx + 1)*wx);
                    gr = edge - g.fzx; gr = gr > 0.0 ? gr : 0.0;
                    // sin^2(dRA/2) turns down past the antipode: a cell that reaches beyond it is
                    // bounded by whichever of its two edges is nearer (mod 2*pi)
                    const double two_pi = 6.283185307179586;
                    const double far = gr + wx + 2.0*g.fzx;
                    if (far > 3.141592653589793) { const double alt = two_pi - far; gr = fmin(gr, alt > 0.0 ? alt : 0.0); }
                }
                const double lb = A + sin2_lower(0.5*gr)*cc;
                // (near the antipode the bound stops growing with |dx|: never close a side there)
                if (lb >= L.kth() && dx != 0 && gr < 3.141592653589793 - 2.0*wx) {
                    if (dx > 0) e_done = true; else w_done = true;
                    continue;
                }
                const uint64_t ck = (uint64_t)col*g.ndec + (uint64_t)row;
                const uint32_t qs = cand.cstart[ck], qe = cand.cstart[ck + 1];
                const double kth4 = 4.0*L.kth()*(1.0 + 1e-12);
                for (uint32_t q = qs; q < qe; ++q) {
                    const Rec cd = cand.rec[q];
                    if (self_mode && cd.idx == me.idx) continue;
                    // Cheap rejection before the two libm-grade sines: sin^2(h) >= h^2 (1 - h^2/3) for
                    // every h, hence 4p >= Q (1 - M/12) with Q = dy^2 + dx^2 cos cos, M = max(dx^2, dy^2)
                    // (dx wrapped into [-pi, pi]: the proxy is 2*pi-periodic in it). A candidate whose
                    // bound is strictly above the current nth best cannot enter the list, ties included.
                    double dx = cd.x - x1;
                    const double dy = cd.y - y1;
                    if (sg.wrap) dx = dx > 3.141592653589793 ? dx - 6.283185307179586
                                                             : (dx < -3.141592653589793 ? dx + 6.283185307179586 : dx);
                    const double dx2 = dx*dx, dy2 = dy*dy;
                    const double Q = fma(dx2*cd.c, c1, dy2);
                    const double M = dx2 > dy2 ? dx2 : dy2;
                    if (me.c >= 0.0 && cd.c >= 0.0 && Q*(1.0 - M*(1.0/12.0) - 1e-12) >= kth4) continue;
                    L.insert(proxy_sph(x1, y1, me.c, cd.x, cd.y, cd.c, g.sin_model), cd.idx);
                }
                evals += qe - qs;
                if (ncol == 1) { e_done = w_done = true; }
            }
            if (nrow == 1) { up_done = dn_done = true; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if ((uint32_t)k < nth_eff) {
                id_out[(uint64_t)k*plane_stride + me.idx] = L.i[k] == kNone ? XM_NPOS : (uint64_t)L.i[k];
                d_out[(uint64_t)k*plane_stride + me.idx] = proxy_to_true(L.d[k], false);
            }
        }
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, o);
    if ((threadIdx.x & 31) == 0 && evals) atomicAdd(reverse ? &ctr->evals_rev : &ctr->evals_fwd, evals);
}

// ---- nearest-neighbour filter (nth = 1) ------------------------------------------------------------------
// Same contract as ring_nn_filter_kernel of the cell-binned path: rank the candidates of the 3x3 cell
// block with the cheap metric Q = dy^2 + dx^2 cos cos, tracked by its high word; accept the row only
// if (1) everything outside the block is provably farther than the winner and (2) the winner beats
// the runner-up by more than the metric's uncertainty (4p in [Q(1 - M/12), Q], M <= (2 max(w))^2
// inside the block); then evaluate the reference-order proxy of the winning pair once. Any other
// row -- grid edge, polar rows, the two columns at the RA seam, exact duplicates, incomplete blocks --
// is appended to the list the exact kernel above walks afterwards (1.4 % of the rows on a uniform
// full sky: the polar caps, where the fixed-width RA columns are physically narrow).
// What ncu said about the earlier forms (profiles/README.md): fully unrolled with the libm-grade sine
// it was 4700 SASS instructions and spent 58 % of its issue slots waiting for instruction fetches;
// rolled into one lockstep loop over the nine cells it kept 7.8 of 32 lanes busy and spent 29 % of
// its instructions on the FP64 selects of the RA wrap. Hence the shape below: polynomial bounds and
// the small-angle sine only, no wrap, and the three phases A / B / C.
using filt::kInfHi; using filt::hi_floor; using filt::hi_ceil; using filt::ResRec; using filt::store_res;

// cell order: own, N, S, E, W, NE, NW, SE, SW (sides before corners, so that the best shrinks early);
// two bits per entry hold offset + 1
constexpr uint32_t pack9(int a0, int a1, int a2, int a3, int a4, int a5, int a6, int a7, int a8) {
    return (uint32_t)(a0 + 1) | (uint32_t)(a1 + 1) << 2 | (uint32_t)(a2 + 1) << 4 | (uint32_t)(a3 + 1) << 6 |
           (uint32_t)(a4 + 1) << 8 | (uint32_t)(a5 + 1) << 10 | (uint32_t)(a6 + 1) << 12 |
           (uint32_t)(a7 + 1) << 14 | (uint32_t)(a8 + 1) << 16;
}
constexpr uint32_t kNbX = pack9(0, 0, 0, 1, -1, 1, -1, 1, -1);
constexpr uint32_t kNbY = pack9(0, 1, -1, 0, 0, 1, 1, -1, -1);

template <bool SELF, bool QUEUE>
__global__ void __launch_bounds__(kThreads)
sphere_nn_filter_kernel(SphereGrid sg, CatView src, CatView cand, double eps_lo, ResRec* __restrict__ res,
                        uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                        uint32_t* __restrict__ refine_list, unsigned long long* refine_count, int reverse,
                        SearchCounters* ctr) {
    const GridView& g = sg.g;
    const uint32_t p = blockIdx.x*blockDim.x + threadIdx.x;
    const bool live = p < src.n;
    uint32_t evals = 0;
    Rec me;
    me.x = me.y = 0.0; me.c = -1.0; me.idx = 0; me.key = 0;
    if (live) me = filt::load_rec(src.rec + p);
    const double x1 = me.x, y1 = me.y, c1 = me.c;
    const int ncol = (int)g.nra, nrow = (int)g.ndec;
    const int x0 = (int)(me.key/g.ndec), y0 = (int)(me.key - (uint32_t)x0*g.ndec);
    const double wx = g.wx, wy = g.wy;
    // grid edge, polar rows and the columns next to the RA seam (dx would need wrapping) are the
    // exact kernel's: 2/ncol of the sources on a full-sky grid
    bool flag = !(c1 >= 0.0) || y0 < 1 || y0 + 1 >= nrow || x0 < 1 || x0 + 1 >= ncol;
    int32_t h1 = kInfHi, h2 = kInfHi;
    uint32_t slot = kNone;
    double ge = 0, gw = 0, gn = 0, gs = 0, cm0 = 0, cmn = 0, cms = 0;
    const double shrink = 1.0 - eps_lo - 1e-9;
    uint32_t q = 0, qe = 0;

#define XM_EVAL()                                                                   \
    {                                                                               \
        const Rec cd = filt::load_rec(cand.rec + q);                                \
        const double dx = cd.x - x1, dy = cd.y - y1;                                \
        int32_t h = __double2hiint(fma(dx*dx*cd.c, c1, dy*dy));                     \
        if (SELF && cd.idx == me.idx) h = kInfHi;                                   \
        const bool better = h < h1;                                                 \
        h2 = min(h2, max(h1, h));                                                   \
        h1 = min(h1, h);                                                            \
        slot = better ? q : slot;                                                   \
        ++q;                                                                        \
    }

    // ---- A. own cell, lanes in lockstep ----------------------------------------------------------------
    if (!flag) {
        const double cx_lo = g.ox + (double)x0*wx, cy_lo = g.oy + (double)y0*wy;
        ge = cx_lo + wx - x1 - g.fzx; gw = x1 - cx_lo - g.fzx;
        gn = cy_lo + wy - y1 - g.fzy; gs = y1 - cy_lo - g.fzy;
        ge = ge > 0 ? ge : 0; gw = gw > 0 ? gw : 0; gn = gn > 0 ? gn : 0; gs = gs > 0 ? gs : 0;
        cm0 = sg.cmin_row[y0]*c1; cmn = sg.cmin_row[y0 + 1]*c1; cms = sg.cmin_row[y0 - 1]*c1;
        q = cand.cstart[me.key]; qe = cand.cstart[me.key + 1];
        evals += qe - q;
        while (q < qe) XM_EVAL()
    }
    // ---- B. which of the eight neighbours can hold a winner or a near-tie? (all lanes together) ---------
    uint32_t todo = 0;
    if (!flag) {
        const double thr = hi_ceil(h1)*(1.0 + 1e-9);
        const double Qe = ge*ge, Qw = gw*gw, Qn = gn*gn*shrink, Qs = gs*gs*shrink;
        const double s0 = cm0*shrink, sn = cmn*shrink, ss = cms*shrink;
        // bit e of kNbX/kNbY order: 1 N, 2 S, 3 E, 4 W, 5 NE, 6 NW, 7 SE, 8 SW
        todo = (Qn < thr ? 2u : 0u) | (Qs < thr ? 4u : 0u) | (Qe*s0 < thr ? 8u : 0u) | (Qw*s0 < thr ? 16u : 0u) |
               (fma(Qe, sn, Qn) < thr ? 32u : 0u) | (fma(Qw, sn, Qn) < thr ? 64u : 0u) |
               (fma(Qe, ss, Qs) < thr ? 128u : 0u) | (fma(Qw, ss, Qs) < thr ? 256u : 0u);
    }
    // ---- C. the surviving neighbours ---------------------------------------------------------------------
    if (QUEUE) {
        // Warp-level work queue: the (source, cell) pairs of the 32 sources -- 0.6 per source on
        // average, but 0..8 per lane -- are enumerated with ballots (cell by cell, lane by lane) and
        // dealt one per lane; the lane scans that cell for the owner's source (fetched with
        // shuffles) and the owner collects the partial results with shuffles. The warp runs for the
        // largest CELL, not for the lane with the most cells.
        __shared__ uint16_t s_item[kThreads/32][256];      // pair j of the warp -> owner lane | cell << 5
        uint16_t* items = s_item[threadIdx.x >> 5];
        const unsigned lane = threadIdx.x & 31u;
        const unsigned lt = (1u << lane) - 1u;
        unsigned bal[9], base[9];
        unsigned total = 0;
#pragma unroll
        for (int e = 1; e <= 8; ++e) {
            bal[e] = __ballot_sync(0xffffffffu, (todo >> e) & 1u);
            base[e] = total;
            if ((todo >> e) & 1u) items[total + __popc(bal[e] & lt)] = (uint16_t)(lane | ((unsigned)e << 5));
            total += __popc(bal[e]);
        }
        __syncwarp();
        for (unsigned r0 = 0; r0 < total; r0 += 32u) {
            const unsigned j = r0 + lane;
            int my_e = 0;
            unsigned owner = lane;
            if (j < total) { const unsigned it = items[j]; owner = it & 31u; my_e = (int)(it >> 5); }
            const double ox1 = __shfl_sync(0xffffffffu, x1, owner), oy1 = __shfl_sync(0xffffffffu, y1, owner);
            const double oc1 = __shfl_sync(0xffffffffu, c1, owner);
            const uint32_t okey = __shfl_sync(0xffffffffu, me.key, owner);
            const uint32_t oidx = __shfl_sync(0xffffffffu, me.idx, owner);
            int32_t a1 = kInfHi, a2 = kInfHi;
            uint32_t aslot = kNone, qq = 0, qqe = 0;
            if (my_e) {
                const int bx = (int)((kNbX >> (2*my_e)) & 3u) - 1, by = (int)((kNbY >> (2*my_e)) & 3u) - 1;
                const uint32_t k = okey + (uint32_t)(bx*(int32_t)g.ndec + by);
                qq = cand.cstart[k]; qqe = cand.cstart[k + 1];
                evals += qqe - qq;
            }
            for (; qq < qqe; ++qq) {
                const Rec cd = filt::load_rec(cand.rec + qq);
                const double dx = cd.x - ox1, dy = cd.y - oy1;
                int32_t h = __double2hiint(fma(dx*dx*cd.c, oc1, dy*dy));
                if (SELF && cd.idx == oidx) h = kInfHi;
                const bool better = h < a1;
                a2 = min(a2, max(a1, h));
                a1 = min(a1, h);
                aslot = better ? qq : aslot;
            }
#pragma unroll
            for (int e = 1; e <= 8; ++e) {
                const unsigned je = base[e] + __popc(bal[e] & lt);          // my pair for cell e, if I have one
                const bool mine = ((todo >> e) & 1u) && je >= r0 && je < r0 + 32u;
                const unsigned from = (je - r0) & 31u;
                const int32_t b1 = __shfl_sync(0xffffffffu, a1, from), b2 = __shfl_sync(0xffffffffu, a2, from);
                const uint32_t bs = __shfl_sync(0xffffffffu, aslot, from);
                if (mine) {
                    const bool better = b1 < h1;
                    h2 = min(min(h2, b2), max(h1, b1));
                    h1 = min(h1, b1);
                    slot = better ? bs : slot;
                }
            }
        }
    } else {
    // One flat loop over the candidates of the surviving neighbours:
    // a lane moves on to its next cell on its own, so the warp runs for the longest lane's total, not
    // for the sum of the per-cell maxima (walking the nine cells in lockstep kept 7.8 of 32 lanes busy).
    // Every lane of the warp stays in the loop until the last one is done: the pop branch and the
    // evaluation reconverge each iteration.
    bool active = todo != 0;
    q = qe = 0;
    while (__any_sync(0xffffffffu, active)) {
        if (active && q == qe) {
            bool found = false;
            while (todo) {
                const int e = __ffs(todo) - 1;
                todo &= todo - 1;
                const int bx = (int)((kNbX >> (2*e)) & 3u) - 1, by = (int)((kNbY >> (2*e)) & 3u) - 1;
                const double gx = bx > 0 ? ge : (bx < 0 ? gw : 0.0), gy = by > 0 ? gn : (by < 0 ? gs : 0.0);
                const double cm = by > 0 ? cmn : (by < 0 ? cms : cm0);
                // the best may have improved since step B
                if (!(fma(gx*gx, cm, gy*gy)*shrink < hi_ceil(h1)*(1.0 + 1e-9))) continue;
                const uint32_t k = me.key + (uint32_t)(bx*(int32_t)g.ndec + by);
                q = cand.cstart[k]; qe = cand.cstart[k + 1];
                evals += qe - q;
                if (q != qe) { found = true; break; }
            }
            active = found;
        }
        if (active) XM_EVAL()
    }
    }
#undef XM_EVAL

    if (live) {
        if (!flag) {
            // (1) completeness: proxy bounds of everything outside the 3x3 block, with
            // sin^2 h >= h^2 (1 - h^2/3) for every h; (2) winner vs runner-up
            const double oy = wy - 2.0*g.fzy, ox = wx - 2.0*g.fzx;      // binning slack of the far cells
            const double hy = 0.5*(fmin(gn, gs) + oy), hx = 0.5*fmin(fmin(ge, gw) + ox, 3.141592653589793);
            const double cmin3 = fmin(fmin(cm0, cmn), cms);
            const double out_p = fmin(hy*hy*(1.0 - hy*hy*(1.0/3.0)), hx*hx*(1.0 - hx*hx*(1.0/3.0))*cmin3)*(1.0 - 1e-9);
            const double q1 = hi_ceil(h1);
            if (!(0.25*q1*(1.0 + 1e-9) <= out_p)) flag = true;
            if (h2 < kInfHi && !(q1*(1.0 + 1e-12) < hi_floor(h2)*shrink)) flag = true;
            if (slot == kNone || h1 < 0) flag = true;
        }
        Rec w;
        if (!flag) {
            w = filt::load_rec(cand.rec + slot);
            // the small-angle sine covers |dx|/2 < 0.126
            if (!(fabs(w.x - x1) < 0.25 && fabs(w.y - y1) < 0.25)) flag = true;
        }
        if (flag) {
            const unsigned long long pos = atomicAdd(refine_count, 1ull);
            refine_list[pos] = p;
            if (res) store_res(res + me.idx, filt::kUndecidedId, 0.0);
        } else {
            const double dist = proxy_to_true_small(proxy_sph_small(x1, y1, c1, w.x, w.y, w.c, g.sin_model));
            if (res) store_res(res + me.idx, (uint64_t)w.idx, dist);
            else { id_out[me.idx] = (uint64_t)w.idx; d_out[me.idx] = dist; }
        }
    }
    __syncwarp();
    for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(0xffffffffu, evals, o);
    if ((threadIdx.x & 31) == 0 && evals) atomicAdd(reverse ? &ctr->evals_rev : &ctr->evals_fwd, (unsigned long long)evals);
}

__global__ void sphere_rows_kernel(GridView g, double* __restrict__ cmin) {
    const uint32_t r = blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= g.ndec) return;
    const double half_pi = 1.5707963267948966;
    double lo = g.oy + (double)r*g.wy - g.fzy, hi = lo + g.wy + 2.0*g.fzy;
    double v = 0.0;
    if (lo > -half_pi && hi < half_pi) {
        v = fmin(cos(lo), cos(hi))*(1.0 - 1e-12) - 1e-15;
        if (v < 0.0) v = 0.0;
    }
    cmin[r] = v;
}

}  // namespace

// Grid of our own for the exact search. bb = {ra_min, ra_max, dec_min, dec_max} of BOTH catalogs
// together [deg]; n_cand = rows of the candidate catalog. Returns false when the geometry is not
// one this path handles (RA spanning more than one turn).
bool sphere_grid(const double* bb, uint64_t n_cand, uint32_t nth, GridView* g, int* wrap) {
    const double pi = 3.141592653589793;
    double ra0 = bb[0], ra1 = bb[1], de0 = bb[2], de1 = bb[3];
    if (!(ra1 - ra0 <= 360.0)) return false;
    if (de0 < -90.0) de0 = -90.0;
    if (de1 > 90.0) de1 = 90.0;
    if (!(de1 >= de0)) { de0 = -90.0; de1 = 90.0; }
    *wrap = (ra1 - ra0) > 180.0 ? 1 : 0;
    const double span = *wrap ? 360.0 : fmax(ra1 - ra0, 1e-9);
    // solid angle of the box and a cell size aiming at ~8*nth candidates per cell
    const double omega = (span*pi/180.0)*fmax(sin(de1*pi/180.0) - sin(de0*pi/180.0), 1e-12);
    const char* dens = getenv("XM_SPHERE_DENSITY");
    const double per_cell = dens ? atof(dens) : 8.0;    // measured on C5 (1e9 x 1e8): 8 -> 204 ms, 4 -> 215, 3 -> 222, 2 -> 247
    double h = sqrt(omega*per_cell*(double)(nth ? nth : 1)/(double)(n_cand ? n_cand : 1))*180.0/pi;   // [deg]
    if (h < 1.0/3600.0) h = 1.0/3600.0;
    if (h > 10.0) h = 10.0;
    // rows: pad by one so that every source falls strictly inside
    double nrow_f = ceil((de1 - de0)/h) + 1.0, ncol_f = *wrap ? floor(360.0/h) : ceil(span/h) + 1.0;
    if (ncol_f < 1.0) ncol_f = 1.0;
    while (nrow_f*ncol_f > (double)XM_MAX_CELLS) { h *= 1.5; nrow_f = ceil((de1 - de0)/h) + 1.0; ncol_f = *wrap ? floor(360.0/h) : ceil(span/h) + 1.0; if (ncol_f < 1.0) ncol_f = 1.0; }
    memset(g, 0, sizeof(*g));
    g->nra = (uint32_t)ncol_f; g->ndec = (uint32_t)nrow_f;
    g->dra = *wrap ? 360.0/ncol_f : h;
    g->ddec = h;
    g->rra0 = ra0;
    g->rdec0 = de0 - 0.5*(nrow_f*h - (de1 - de0));               // centre the rows on the data
    g->cell_size = h*3600.0;
    const double unit = pi/180.0;
    g->ox = g->rra0*unit; g->oy = g->rdec0*unit; g->wx = g->dra*unit; g->wy = g->ddec*unit;
    g->fzx = 1e-12*fmax(fmax(fabs(ra0), fabs(ra1)), 360.0)*unit;
    g->fzy = 1e-12*90.0*unit;
    g->max_depth = g->nra > g->ndec ? g->nra : g->ndec;
    return true;
}

int32_t launch_sphere_rows(const GridView& g, double* cmin_row, cudaStream_t s, Error& err, uint64_t* launches) {
    sphere_rows_kernel<<<(g.ndec + 255)/256, 256, 0, s>>>(g, cmin_row);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

bool sphere_applicable(bool linear, uint32_t nth) { return !linear && nth >= 1 && nth <= 8; }

int32_t launch_sphere_search(const GridView& g, int wrap, const double* cmin_row, const CatView& src,
                             const CatView& cand, bool self, uint32_t nth_eff, uint64_t plane_stride,
                             uint64_t* id_out, double* d_out, uint32_t* refine_list,
                             unsigned long long* refine_count, void* result_records, bool reverse,
                             SearchCounters* ctr, cudaStream_t s, cudaStream_t side, cudaEvent_t ev_fork,
                             cudaEvent_t ev_join, Error& err, uint64_t* launches) {
    if (src.n == 0 || nth_eff == 0) return XM_OK;
    SphereGrid sg;
    sg.g = g; sg.wrap = wrap; sg.cmin_row = cmin_row;
    const dim3 full((src.n + kThreads - 1)/kThreads);
    // nth = 1 on small cells: filter kernel first, the exact kernel then walks the rows it flagged
    const double wmax = g.wx > g.wy ? g.wx : g.wy;
    const bool filter = nth_eff == 1 && refine_list && wmax <= 0.02 && g.nra >= 8 && g.ndec >= 3;
    const uint32_t* rows = nullptr;
    dim3 grid = full;
    if (filter) {
        const double eps_lo = (2.0*wmax)*(2.0*wmax)/12.0 + 1e-14;
        ResRec* res = (ResRec*)result_records;
        const int rv = reverse ? 1 : 0;
        // warp-level queue for the neighbour cells (C5 on one B200: 204 -> 174 ms); XM_SPHERE_QUEUE=0
        // selects the per-lane flat loop instead (kept for A/B runs)
        static const bool queue = [] { const char* e = getenv("XM_SPHERE_QUEUE"); return !(e && e[0] == '0'); }();
#define XM_FILT(S, Q) sphere_nn_filter_kernel<S, Q><<<full, kThreads, 0, s>>>(sg, src, cand, eps_lo, res, id_out, \
        d_out, refine_list, refine_count, rv, ctr)
        if (self) { if (queue) XM_FILT(true, true); else XM_FILT(true, false); }
        else { if (queue) XM_FILT(false, true); else XM_FILT(false, false); }
#undef XM_FILT
        XM_CUDA_TRY(cudaGetLastError());
        *launches += 1;
        rows = refine_list;
        if (grid.x > 148*8) grid.x = 148*8;
        if (res) {
            // the unpack pass skips the undecided rows (kUndecidedId), so the exact kernel can write
            // them from a side stream while it runs
            if (side) {
                XM_CUDA_TRY(cudaEventRecord(ev_fork, s));
                XM_CUDA_TRY(cudaStreamWaitEvent(side, ev_fork, 0));
            }
            int32_t rc = launch_unpack_results(res, src.n, id_out, d_out, s, err, launches);
            if (rc) return rc;
        } else {
            side = nullptr;       // rows were written in place by the filter: keep the order
        }
    } else {
        side = nullptr;
    }
    cudaStream_t sx = side ? side : s;
#define XM_CASE(KK) case KK: sphere_search_kernel<KK><<<grid, kThreads, 0, sx>>>(sg, src, cand, self ? 1 : 0, \
        nth_eff, plane_stride, id_out, d_out, rows, refine_count, reverse ? 1 : 0, ctr); break;
    switch (nth_eff) { XM_CASE(1) XM_CASE(2) XM_CASE(3) XM_CASE(4) XM_CASE(5) XM_CASE(6) XM_CASE(7) default: XM_CASE(8) }
#undef XM_CASE
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    if (side) {
        XM_CUDA_TRY(cudaEventRecord(ev_join, side));
        XM_CUDA_TRY(cudaStreamWaitEvent(s, ev_join, 0));
    }
    return XM_OK;
}

}  // namespace xm

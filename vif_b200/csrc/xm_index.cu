// Cell index construction: bounding box + finiteness, cell keys, CSR cell offsets and the
// gather of each catalog into cell order (32-byte records: x = ra [rad], y = dec [rad], c = cos(dec)).
//
// Replaces the reference's eight min/max passes (qxmatch.hpp:232-235), the finiteness checks
// (:146-149), the unit conversion (:170-181) and the vector-of-vectors bucket fill (:270-288).
// All of it is HBM-bound streaming work: coalesced loads, one pass per array.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"

namespace xm {

namespace {

constexpr int kThreads = 256;

// order-preserving map double -> u64 so that min/max can use integer atomics
__device__ __forceinline__ unsigned long long f2ord(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) {
    return a < b ? a : b;
}
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) {
    return a > b ? a : b;
}

// out5 = {ord(min ra), ord(max ra), ord(min dec), ord(max dec), nonfinite count}
__global__ void __launch_bounds__(kThreads)
bbox_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint64_t n,
            unsigned long long* __restrict__ out5) {
    unsigned long long lo_r = ~0ull, hi_r = 0ull, lo_d = ~0ull, hi_d = 0ull, bad = 0;
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {
        const double r = __ldg(ra + i), d = __ldg(dec + i);
        if (isfinite(r) && isfinite(d)) {
            const unsigned long long a = f2ord(r), b = f2ord(d);
            lo_r = umin64(lo_r, a); hi_r = umax64(hi_r, a);
            lo_d = umin64(lo_d, b); hi_d = umax64(hi_d, b);
        } else {
            ++bad;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo_r = umin64(lo_r, __shfl_xor_sync(0xffffffffu, lo_r, o));
        hi_r = umax64(hi_r, __shfl_xor_sync(0xffffffffu, hi_r, o));
        lo_d = umin64(lo_d, __shfl_xor_sync(0xffffffffu, lo_d, o));
        hi_d = umax64(hi_d, __shfl_xor_sync(0xffffffffu, hi_d, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ unsigned long long sm[5][kThreads/32];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sm[0][w] = lo_r; sm[1][w] = hi_r; sm[2][w] = lo_d; sm[3][w] = hi_d; sm[4][w] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kThreads/32; ++k) {
            lo_r = umin64(lo_r, sm[0][k]); hi_r = umax64(hi_r, sm[1][k]);
            lo_d = umin64(lo_d, sm[2][k]); hi_d = umax64(hi_d, sm[3][k]);
            bad += sm[4][k];
        }
        atomicMin(out5 + 0, lo_r); atomicMax(out5 + 1, hi_r);
        atomicMin(out5 + 2, lo_d); atomicMax(out5 + 3, hi_d);
        if (bad) atomicAdd(out5 + 4, bad);
    }
}

__global__ void bbox_init_kernel(unsigned long long* out5) {
    out5[0] = ~0ull; out5[1] = 0ull; out5[2] = ~0ull; out5[3] = 0ull; out5[4] = 0ull;
}

// qxmatch.hpp:278-279: idx = floor((ra - rra0)/dra), idy = floor((dec - rdec0)/ddec) on the raw
// inputs; IEEE subtract, divide and floor only, so the keys equal the reference's buckets.
__global__ void __launch_bounds__(kThreads)
keys_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint32_t n, double rra0,
            double rdec0, double dra, double ddec, uint32_t nra, uint32_t ndec,
            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, unsigned long long* off_grid) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double fx = floor(__ddiv_rn(__dsub_rn(__ldg(ra + i), rra0), dra));
    const double fy = floor(__ddiv_rn(__dsub_rn(__ldg(dec + i), rdec0), ddec));
    if (off_grid && !(fx >= 0 && fx < (double)nra && fy >= 0 && fy < (double)ndec)) atomicAdd(off_grid, 1ull);
    // with has_bbox1 a caller could hand us rows outside the box it promised: clamp instead of
    // writing out of bounds (rows inside the box are unaffected)
    uint32_t ix = !(fx >= 0) ? 0u : (fx >= (double)nra ? nra - 1 : (uint32_t)fx);      // NaN -> 0
    uint32_t iy = !(fy >= 0) ? 0u : (fy >= (double)ndec ? ndec - 1 : (uint32_t)fy);
    keys[i] = ix*ndec + iy;
    vals[i] = i;
}

// Gather into cell order fused with the unit conversion (:174-180): one 32-byte record per source.
__global__ void __launch_bounds__(kThreads)
cells_gather_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ perm,
                    const double* __restrict__ ra, const double* __restrict__ dec, uint32_t n,
                    int linear, int model, int sanitize, double rra0, double rdec0, double dra, double ddec,
                    uint32_t nra, uint32_t ndec, Rec* __restrict__ rec) {
    const uint32_t p = blockIdx.x*blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t i = perm[p];
    double r = __ldg(ra + i), d = __ldg(dec + i);
    Rec out;
    out.idx = i; out.key = skeys[p];
    if (sanitize) {   // rows that are not finite or off the grid: to the centre of the cell they were clamped to (see place_kernel)
        const double fx = floor(__ddiv_rn(__dsub_rn(r, rra0), dra)), fy = floor(__ddiv_rn(__dsub_rn(d, rdec0), ddec));
        if (!(fx >= 0 && fx < (double)nra && fy >= 0 && fy < (double)ndec)) {
            r = rra0 + ((double)(out.key/ndec) + 0.5)*dra;
            d = rdec0 + ((double)(out.key%ndec) + 0.5)*ddec;
        }
    }
    if (linear) {
        out.x = r; out.y = d; out.c = 1.0;
    } else {
        out.y = __dmul_rn(d, XM_D2R);
        out.x = __dmul_rn(r, XM_D2R);
        out.c = cos_ref(out.y, model);
    }
    rec[p] = out;
}

// CSR cell offsets: cstart[k] = number of sources whose key is < k (lower bound in the sorted
// keys), for k = 0..ncell. Monotone over empty cells too, so a run of consecutive cells of one
// grid column is one contiguous slot range -- what the search kernels stage with bulk copies.
__global__ void __launch_bounds__(kThreads)
cell_offsets_kernel(const uint32_t* __restrict__ skeys, uint32_t n, uint32_t ncell,
                    uint32_t* __restrict__ cstart) {
    const uint32_t k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k > ncell) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(skeys + mid) < k) lo = mid + 1; else hi = mid;
    }
    cstart[k] = lo;
}

constexpr unsigned long long kLost = 0x7fffffffffffffffull;       // INT64_MAX: loses every min

__global__ void merge_pack_kernel(unsigned long long* __restrict__ rid, const double* __restrict__ rd,
                                  uint64_t n, unsigned long long off, double* __restrict__ key) {
    const uint64_t j = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long r = rid[j];
    const bool ok = r != ~0ull;
    rid[j] = ok ? r + off : kLost;
    key[j] = ok ? rd[j] : __longlong_as_double(0x7ff0000000000000ll);
}

__global__ void merge_mask_kernel(unsigned long long* __restrict__ rid, const double* __restrict__ kl,
                                  const double* __restrict__ kb, uint64_t n) {
    const uint64_t j = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (!(kl[j] == kb[j])) rid[j] = kLost;
}

__global__ void merge_unpack_kernel(unsigned long long* __restrict__ rid, double* __restrict__ rd,
                                    const double* __restrict__ kb, uint64_t n, double unmatched) {
    const uint64_t j = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= n) return;
    const bool ok = rid[j] != kLost;
    if (!ok) rid[j] = ~0ull;
    rd[j] = ok ? kb[j] : unmatched;
}

// One-kernel reverse merge over NVLink peer memory: rank `rank` owns rows [lo, hi); for each it reads
// every rank's raw (rd, rid) through the peer mappings, keeps the smallest distance -- the smallest
// GLOBAL row id among equal distances (qxmatch.hpp:510-513) -- and stores the merged row into every
// rank's result buffer. Loads and stores are coalesced 8-byte streams per peer.
struct PeerSet {
    const double* exch[XM_MAX_PEERS];       // n distances, n local row ids (uint64), first row of the rank (uint64)
    double* out[XM_MAX_PEERS];              // n merged distances followed by n merged global ids
};

__global__ void __launch_bounds__(256)
merge_p2p_kernel(PeerSet P, uint64_t n, uint64_t lo, uint64_t hi, int world, double unmatched) {
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x;
    for (uint64_t j = lo + (uint64_t)blockIdx.x*blockDim.x + threadIdx.x; j < hi; j += stride) {
        double bk = __longlong_as_double(0x7ff0000000000000ll);
        unsigned long long br = kLost;
#pragma unroll 4
        for (int r = 0; r < world; ++r) {
            const double k = P.exch[r][j];
            const unsigned long long i = reinterpret_cast<const unsigned long long*>(P.exch[r])[n + j];
            if (i == ~0ull) continue;
            const unsigned long long gi = i + reinterpret_cast<const unsigned long long*>(P.exch[r])[2*n];
            if (k < bk || (k == bk && gi < br)) { bk = k; br = gi; }
        }
        const bool ok = br != kLost;
        const double od = ok ? bk : unmatched;
        const unsigned long long oi = ok ? br : ~0ull;
        for (int r = 0; r < world; ++r) {
            if (P.out[r] == nullptr) continue;          // sliced result: the merged rows stay with their rank
            P.out[r][j] = od;
            reinterpret_cast<unsigned long long*>(P.out[r])[n + j] = oi;
        }
    }
}

}  // namespace

int32_t launch_merge_p2p(const void* const* exch, void* const* out, uint64_t n, int rank, int world,
                         double unmatched, cudaStream_t s, Error& err) {
    PeerSet P;
    for (int r = 0; r < world; ++r) { P.exch[r] = (const double*)exch[r]; P.out[r] = (double*)out[r]; }
    const uint64_t lo = n*(uint64_t)rank/(uint64_t)world, hi = n*(uint64_t)(rank + 1)/(uint64_t)world;
    if (hi == lo) return XM_OK;
    uint64_t blocks = (hi - lo + 255)/256;
    if (blocks > 148*8) blocks = 148*8;
    merge_p2p_kernel<<<(unsigned)blocks, 256, 0, s>>>(P, n, lo, hi, world, unmatched);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_merge_step(int which, unsigned long long* rid, double* rd, const double* ka, const double* kb,
                          double* key_out, uint64_t n, unsigned long long off, double unmatched,
                          cudaStream_t s, Error& err) {
    if (n == 0) return XM_OK;
    const unsigned blocks = (unsigned)((n + 255)/256);
    if (which == 0) merge_pack_kernel<<<blocks, 256, 0, s>>>(rid, rd, n, off, key_out);
    else if (which == 1) merge_mask_kernel<<<blocks, 256, 0, s>>>(rid, ka, kb, n);
    else merge_unpack_kernel<<<blocks, 256, 0, s>>>(rid, rd, kb, n, unmatched);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_bbox(const double* ra, const double* dec, uint64_t n, unsigned long long* out5,
                    cudaStream_t s, Error& err, uint64_t* launches) {
    bbox_init_kernel<<<1, 1, 0, s>>>(out5);
    uint64_t blocks = (n + kThreads*8 - 1)/(kThreads*8);
    if (blocks > 148*16) blocks = 148*16;
    if (blocks < 1) blocks = 1;
    bbox_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(ra, dec, n, out5);
    *launches += 2;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

// in place: {ord(min ra), ord(max ra), ord(min dec), ord(max dec), bad} -> doubles {-min ra, max ra, -min dec, max dec, bad}
__global__ void bbox_neg_kernel(unsigned long long* io5) {
    const unsigned long long o0 = io5[0], o1 = io5[1], o2 = io5[2], o3 = io5[3], bad = io5[4];
    auto ord2d = [](unsigned long long o) {
        const unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
        return __longlong_as_double((long long)b);
    };
    double* out = reinterpret_cast<double*>(io5);
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const bool empty = o0 == ~0ull && o1 == 0ull;           // no finite row seen: neutral element of max
    out[0] = empty ? -inf : -ord2d(o0); out[1] = empty ? -inf : ord2d(o1);
    out[2] = empty ? -inf : -ord2d(o2); out[3] = empty ? -inf : ord2d(o3);
    out[4] = (double)bad;
}

int32_t launch_bbox_neg(const double* ra, const double* dec, uint64_t n, double* out5, cudaStream_t s, Error& err,
                        uint64_t* launches) {
    unsigned long long* io = reinterpret_cast<unsigned long long*>(out5);
    bbox_init_kernel<<<1, 1, 0, s>>>(io);
    if (n) {
        uint64_t blocks = (n + kThreads*8 - 1)/(kThreads*8);
        if (blocks > 148*16) blocks = 148*16;
        bbox_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(ra, dec, n, io);
    }
    bbox_neg_kernel<<<1, 1, 0, s>>>(io);
    *launches += 3;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_keys(const double* ra, const double* dec, uint32_t n, const GridView& g,
                    uint32_t* keys, uint32_t* vals, cudaStream_t s, Error& err, uint64_t* launches) {
    keys_kernel<<<(n + kThreads - 1)/kThreads, kThreads, 0, s>>>(ra, dec, n, g.rra0, g.rdec0, g.dra,
                                                                g.ddec, g.nra, g.ndec, keys, vals, g.off_grid);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_cells_and_gather(const uint32_t* skeys, const uint32_t* perm, const double* ra,
                                const double* dec, uint32_t n, uint32_t ncell, bool linear,
                                Rec* rec, uint32_t* cstart, cudaStream_t s, Error& err,
                                uint64_t* launches, int model, const GridView& g) {
    cell_offsets_kernel<<<(ncell + 1 + kThreads - 1)/kThreads, kThreads, 0, s>>>(skeys, n, ncell, cstart);
    cells_gather_kernel<<<(n + kThreads - 1)/kThreads, kThreads, 0, s>>>(skeys, perm, ra, dec, n,
                                                                        linear ? 1 : 0, model, g.off_grid ? 1 : 0, g.rra0, g.rdec0, g.dra,
                                                                        g.ddec, g.nra, g.ndec, rec);
    *launches += 2;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm

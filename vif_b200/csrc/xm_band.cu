// Declination-band partition of a catalog: the exchange step of the spatially decomposed multi-GPU
// search (vif_b200/distributed.py: qxmatch_banded). The reference's threaded driver cuts both
// catalogs into contiguous row chunks (qxmatch.hpp:422-454); across GPUs a row chunk would have to be
// searched against the WHOLE other catalog, so rows are regrouped by sky position instead: band b
// owns declinations [edge[b], edge[b+1]); a row within `halo` of an edge is also sent to the band on
// the other side of it (it is a candidate there, never a source).
//
// The partition is ORDER PRESERVING: the rows of every destination leave in ascending row order, so
// a destination that concatenates what rank 0, 1, ... sent it holds its rows in ascending GLOBAL row
// order and the searches break exact ties exactly as a single-GPU run does (smallest row).
//   1. emit:    per row, the number of destinations (0, 1 or 2) -> exclusive scan
//   2. pairs:   (destination, row) pairs at the scanned offsets (row order), histogram of destinations
//   3. sort:    one stable radix pass on the destination (xm_sort.cu)
//   4. gather:  (ra, dec, global row | primary flag) triples in destination order
#include "xm_common.cuh"

namespace xm {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBands = 64;

struct BandEdges { double e[kMaxBands + 1]; };

// primary band of dec (edges ascending, edge[0] <= dec < edge[nb] guaranteed by the caller's padding)
__device__ __forceinline__ int band_of(double dec, const BandEdges& E, int nb) {
    int lo = 0, hi = nb;                                   // E.e[lo] <= dec < E.e[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (dec >= E.e[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// destinations of a row: primary band b and, within `halo` of one of b's edges, the band beyond it
__device__ __forceinline__ int destinations(double dec, const BandEdges& E, int nb, double halo, int only, int* d) {
    const int b = band_of(dec, E, nb);
    int n = 0;
    if (b > 0 && dec < E.e[b] + halo) { if (only < 0 || only == b - 1) d[n++] = b - 1; }
    if (only < 0 || only == b) d[n++] = b;
    if (b + 1 < nb && dec >= E.e[b + 1] - halo) { if (only < 0 || only == b + 1) d[n++] = b + 1; }
    return n;
}

__global__ void __launch_bounds__(kThreads)
band_emit_kernel(const double* __restrict__ dec, uint32_t n, BandEdges E, int nb, double halo, int only,
                 uint32_t* __restrict__ cnt) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i > n) return;
    int d[3];
    cnt[i] = i < n ? (uint32_t)destinations(__ldg(dec + i), E, nb, halo, only, d) : 0u;   // cnt[n] = 0: the scan leaves the total there
}

__global__ void __launch_bounds__(kThreads)
band_pairs_kernel(const double* __restrict__ dec, uint32_t n, BandEdges E, int nb, double halo, int only,
                  const uint32_t* __restrict__ off, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                  unsigned long long* __restrict__ counts) {
    __shared__ uint32_t hist[kMaxBands];
    if (threadIdx.x < kMaxBands) hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) {
        int d[3];
        const int m = destinations(__ldg(dec + i), E, nb, halo, only, d);
        const uint32_t o = off[i];
        for (int k = 0; k < m; ++k) {
            keys[o + k] = (uint32_t)d[k]; vals[o + k] = i;
            atomicAdd(&hist[d[k]], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < nb && hist[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)hist[threadIdx.x]);
}

// out: m triples {ra, dec, bits}: bits = global row | (1 << 63 when this entry is the row's PRIMARY band)
__global__ void __launch_bounds__(kThreads)
band_gather_kernel(const double* __restrict__ ra, const double* __restrict__ dec, const uint32_t* __restrict__ keys,
                   const uint32_t* __restrict__ vals, uint32_t m, BandEdges E, int nb, uint64_t row0,
                   double* __restrict__ out) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t i = vals[j];
    const double r = __ldg(ra + i), d = __ldg(dec + i);
    const bool primary = band_of(d, E, nb) == (int)keys[j];
    const unsigned long long bits = (row0 + i) | (primary ? 0x8000000000000000ull : 0ull);
    out[3*(size_t)j] = r; out[3*(size_t)j + 1] = d; out[3*(size_t)j + 2] = __longlong_as_double((long long)bits);
}

// after the band x band search: local candidate index -> global row (bits of the candidate triples),
// and the completeness test: a primary row whose result is farther than the halo (or missing) could have
// a nearer candidate outside the band
__global__ void __launch_bounds__(kThreads)
band_finish_kernel(long long* __restrict__ id_io, const double* __restrict__ d, const long long* __restrict__ bits_src,
                   const long long* __restrict__ bits_cand, uint32_t n, double hlim, unsigned long long* __restrict__ bad) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    bool b = false;
    if (j < n) {
        const long long l = id_io[j];
        id_io[j] = l >= 0 ? (bits_cand[l] & 0x7fffffffffffffffll) : -1ll;
        b = bits_src[j] < 0 && !(d[j] < hlim);
    }
    const unsigned m = __ballot_sync(0xffffffffu, b);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(bad, (unsigned long long)__popc(m));
}

// results that travelled back to the rows' owner: got = {id bits, d} per sent triple, in send order
__global__ void __launch_bounds__(kThreads)
band_home_kernel(const double* __restrict__ got, const long long* __restrict__ sent_bits, uint32_t m, uint64_t row0,
                 long long* __restrict__ id_out, double* __restrict__ d_out) {
    const uint32_t j = blockIdx.x*blockDim.x + threadIdx.x;
    if (j >= m) return;
    const long long b = sent_bits[3*(size_t)j + 2];
    if (b >= 0) return;                                   // a halo copy: its result is not the row's
    const uint64_t r = (uint64_t)(b & 0x7fffffffffffffffll) - row0;
    id_out[r] = __double_as_longlong(got[2*(size_t)j]);
    d_out[r] = got[2*(size_t)j + 1];
}

}  // namespace

int32_t launch_band_finish(long long* id_io, const double* d, const long long* bits_src, const long long* bits_cand,
                           uint32_t n, double hlim, unsigned long long* bad, cudaStream_t s, Error& err) {
    if (n == 0) return XM_OK;
    band_finish_kernel<<<(n + kThreads - 1)/kThreads, kThreads, 0, s>>>(id_io, d, bits_src, bits_cand, n, hlim, bad);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_band_home(const double* got, const long long* sent, uint32_t m, uint64_t row0, long long* id_out,
                         double* d_out, cudaStream_t s, Error& err) {
    if (m == 0) return XM_OK;
    band_home_kernel<<<(m + kThreads - 1)/kThreads, kThreads, 0, s>>>(got, sent, m, row0, id_out, d_out);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

size_t band_scratch_bytes(uint64_t n, uint64_t cap) {
    return ((size_t)n + 1)*4 + bucket_scan_aux_bytes((uint32_t)n + 1) + (size_t)cap*16 + sort_scratch_bytes(cap) + 256;
}

// dec/ra: n rows; edges: nb + 1 ascending values with edges[0] <= every dec < edges[nb]; out: cap triples;
// counts (device, nb u64): rows per destination; *total (host): triples written. only >= 0: emit that
// destination alone (the order-preserving selection of one band + halo out of a replicated catalog).
int32_t launch_band_partition(const double* ra, const double* dec, uint32_t n, uint64_t row0, const double* edges,
                              int nb, double halo, int only, double* out, uint64_t cap, unsigned long long* counts,
                              void* scratch, uint64_t* total, cudaStream_t s, Error& err, uint64_t* launches) {
    if (nb < 1 || nb > kMaxBands) return fail(err, XM_ERR_INVALID_ARG, "1 to %d bands", kMaxBands);
    BandEdges E;
    for (int k = 0; k <= nb; ++k) E.e[k] = edges[k];
    XM_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)nb*sizeof(unsigned long long), s));
    *total = 0;
    if (n == 0) return XM_OK;
    uint32_t* cnt = (uint32_t*)scratch;
    uint32_t* aux = cnt + (size_t)n + 1;
    unsigned char* p = (unsigned char*)(aux) + bucket_scan_aux_bytes(n + 1);
    p = (unsigned char*)(((uintptr_t)p + 127) & ~(uintptr_t)127);
    uint32_t* keys = (uint32_t*)p;
    uint32_t* vals = keys + cap;
    uint32_t* tkeys = vals + cap;
    uint32_t* tvals = tkeys + cap;
    void* sscr = tvals + cap;
    band_emit_kernel<<<(n + 1 + kThreads - 1)/kThreads, kThreads, 0, s>>>(dec, n, E, nb, halo, only, cnt);
    *launches += 1;
    launch_exclusive_scan_u32(cnt, n + 1, aux, s, launches);
    uint32_t m = 0;
    XM_CUDA_TRY(cudaMemcpyAsync(&m, cnt + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    XM_CUDA_TRY(cudaStreamSynchronize(s));
    if ((uint64_t)m > cap) return fail(err, XM_ERR_INVALID_ARG, "band partition: %u entries do not fit the output (%llu)", m,
                                       (unsigned long long)cap);
    *total = m;
    if (m == 0) return XM_OK;
    band_pairs_kernel<<<(n + kThreads - 1)/kThreads, kThreads, 0, s>>>(dec, n, E, nb, halo, only, cnt, keys, vals, counts);
    *launches += 1;
    if (only < 0 && nb > 1) {
        int bits = 1;
        while ((1 << bits) < nb) ++bits;
        int32_t rc = launch_sort_pairs(keys, vals, tkeys, tvals, m, bits, sscr, s, err, launches);
        if (rc) return rc;
    }
    band_gather_kernel<<<(m + kThreads - 1)/kThreads, kThreads, 0, s>>>(ra, dec, keys, vals, m, E, nb, row0, out);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm

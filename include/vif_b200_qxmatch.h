/*
 * vif_b200_qxmatch.h -- C ABI of the B200-native replacement for vif::astro::qxmatch.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types. The entry points
 * are what a binding of the reference's only hot path would need:
 *
 *   reference interface replaced                               (file:line under /root/reference)
 *   ----------------------------------------------------------------------------------------
 *   qxmatch(ra1, dec1, ra2, dec2, qxmatch_params) -> qxmatch_res   include/vif/astro/qxmatch.hpp:133-616
 *   struct qxmatch_params {thread,nth,verbose,self,brute_force,no_mirror,linear}      qxmatch.hpp:31-39
 *   struct qxmatch_res {id (nth,n1) ; d (nth,n1) ; rid (n2) ; rd (n2)}                qxmatch.hpp:8-17
 *   callers: tools/qxmatch2/qxmatch2.cpp:130,148 ; include/vif/astro/catalog_merge.hpp:195
 *
 * The C++ drop-in that sits on top of this ABI (same namespace, same signatures as the reference)
 * is include/vif_b200/astro/qxmatch.hpp; the Python mirror is vif_b200/qxmatch.py.
 *
 * Conventions (identical to the reference unless stated):
 *   - inputs in degrees (any units when `linear`), outputs in arcsec (input units when `linear`);
 *   - `id`, `d` are plane-major: neighbour rank k of source i is at [k*n1 + i]  (qxmatch.hpp:157);
 *   - unmatched slots: id = XM_NPOS, d = NaN (inf when `linear`, or when either catalog is empty:
 *     qxmatch.hpp:165-167, 609-613);
 *   - cell-binned path with `self`: rid = XM_NPOS, rd = NaN (reverse pass skipped, qxmatch.hpp:395);
 *   - `thread` is accepted and ignored (the GPU path has no host worker threads);
 *   - EXTENSION: where the reference's bucket grid is infeasible (catalog 2 spanning RA 0..360 gives a
 *     hull area of 0 -> 1-arcsec cells over the whole box -> ~1e12 buckets, which the reference would
 *     try to allocate) the cell-binned call is answered by the exact spherical grid search instead:
 *     true nth nearest neighbours, equal distances in ascending row order, i.e. the brute_force result;
 *   - there is NO CPU fallback: every entry point that computes returns XM_ERR_CUDA when no
 *     CUDA device / kernel image is usable.
 *
 * Errors: the reference prints and calls exit(EXIT_FAILURE) (vif_check, core/error.hpp:330-335).
 * The ABI returns a non-zero code and writes a message into `errbuf`; the C++ shim turns that back
 * into print + exit. Calls are synchronous. A context (one per device per process, created lazily)
 * is protected by a mutex: concurrent calls on the same device are serialised.
 */
#ifndef VIF_B200_QXMATCH_H
#define VIF_B200_QXMATCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XM_NPOS UINT64_MAX            /* vif::npos, core/typedefs.hpp:13 */

enum {
    XM_OK = 0,
    XM_ERR_INVALID_ARG = 1,           /* NULL pointer, n >= 2^32-1, bad flag combination          */
    XM_ERR_NONFINITE_1 = 2,           /* "first RA and Dec coordinates contain invalid values"    */
    XM_ERR_NONFINITE_2 = 3,           /* "second RA and Dec coordinates contain invalid values"   */
    XM_ERR_GRID_TOO_LARGE = 4,        /* the reference's bucket grid exceeds XM_MAX_CELLS and the exact */
                                      /* spherical search does not apply (linear mode, nth > 8)          */
    XM_ERR_CUDA = 5,                  /* CUDA runtime failure or no device: message has details   */
    XM_ERR_NO_MEMORY = 6
};

#define XM_MAX_CELLS (1ull << 30)

/* Mirrors vif::astro::qxmatch_params (qxmatch.hpp:31-39) field for field, then GPU extras. */
typedef struct xm_params {
    uint64_t struct_size;             /* sizeof(xm_params) as compiled into the CALLER (XM_PARAMS_INIT sets */
                                      /* it). The library reads that many bytes and takes every field the   */
                                      /* caller does not know as zero, so callers built against an older    */
                                      /* header keep working; 0 or a size it does not know is rejected.     */
    uint64_t thread;                  /* ignored by the GPU path                                   */
    uint64_t nth;                     /* clamped to >= 1 (qxmatch.hpp:152)                         */
    uint8_t  verbose;
    uint8_t  self;
    uint8_t  brute_force;
    uint8_t  no_mirror;
    uint8_t  linear;
    /* ---- extras (zero-initialise for reference behaviour) ---- */
    uint8_t  has_bbox1;               /* 1: use bbox1[] as catalog 1's bounding box instead of the  */
                                      /* box of the rows passed (row-sharded multi-GPU runs must   */
                                      /* bin with the GLOBAL box of qxmatch.hpp:232-238)           */
    uint8_t  exact_only;              /* 1: disable the filter kernels, run reference-order maths   */
                                      /* on every pair (debug / parity triage)                     */
    uint8_t  exact_grid;              /* 1 (with brute_force): compute the brute_force result with the */
                                      /* exact spherical grid search instead of all n1*n2 pairs; chosen  */
                                      /* automatically above 4e12 pairs. Identical results.             */
    int32_t  device;                  /* CUDA ordinal; -1 = current device                         */
    int32_t  has_ranges;              /* 1: partitioned run (xm_qxmatch_dev, cell-binned path only): both  */
                                      /* catalogs are passed WHOLE (they are each other's candidates) but  */
                                      /* the forward match is computed for catalog-1 rows                  */
                                      /* [fwd_begin, fwd_begin+fwd_count) only -> id,d hold nth*fwd_count,  */
                                      /* and the reverse match for catalog-2 rows [rev_begin, +rev_count)   */
                                      /* -> rid,rd hold rev_count. The reference's threaded driver splits   */
                                      /* work the same way (chunks of i AND of j, qxmatch.hpp:422-454).    */
    double   bbox1[4];                /* ra_min, ra_max, dec_min, dec_max                          */
    uint64_t fwd_begin, fwd_count, rev_begin, rev_count;
    /* ---- appended in round 2 (older callers: struct_size): catalog 2 still on its way (xm_qxmatch_dev, cell-binned path) ---- */
    uint64_t has_bbox2;               /* 1 (together with has_bbox1): bbox2[] is catalog 2's bounding box. */
                                      /* The grid (qxmatch.hpp:232-267) is then known before a single row  */
                                      /* is read: no host round trip at the head of the call, catalog 1 is */
                                      /* indexed at once and catalog 2 as soon as it has arrived. Both     */
                                      /* boxes and the finiteness of the rows are still verified on the    */
                                      /* device; a violation fails the call at its end.                    */
    double   bbox2[4];
    void*    cat2_ready_event;        /* cudaEvent_t or NULL: nothing reads ra2/dec2 before this event     */
                                      /* (e.g. the end of an NCCL broadcast of catalog 2 on another        */
                                      /* stream): the broadcast overlaps the index build of catalog 1.     */
} xm_params;
/* xm_params p = XM_PARAMS_INIT;  -- every field zero (= reference behaviour) except the size */
#define XM_PARAMS_INIT { sizeof(xm_params) }

/* Grid geometry chosen for the cell-binned path (qxmatch.hpp:232-267), reported for inspection. */
typedef struct xm_grid {
    double   rra0, rra1, rdec0, rdec1; /* padded bounds                                            */
    double   cell_size, dra, ddec;     /* cell size [arcsec], cell widths [deg]                    */
    uint64_t nra, ndec, nc2;
    double   area;                     /* hull area of catalog 2's bounding box [deg^2]            */
} xm_grid;

/* Per-call statistics of the last call made on a device (device-side times from CUDA events). */
typedef struct xm_stats {
    uint64_t struct_size;              /* in: sizeof(xm_stats) of the caller; out: bytes the library filled */
    xm_grid  grid;
    double   ms_total;                 /* whole device pipeline                                    */
    double   ms_bbox, ms_keys, ms_sort, ms_index, ms_forward, ms_reverse;
    double   ms_search;                /* forward + reverse searches, wall (they overlap on two streams) */
    uint64_t launches;                 /* kernels launched by this library during the call         */
    uint64_t rows_refined_fwd;         /* rows the filter kernel handed to the exact kernel        */
    uint64_t rows_refined_rev;
    uint64_t rows_ambiguous;           /* rows whose decisive proxies differ by < 8 ulp            */
    uint64_t evals_fwd, evals_rev;     /* distance evaluations (filter + exact)                    */
    uint64_t sin_model;                /* 1 = glibc FMA Taylor model, 2 = non-FMA model            */
    double   ms_host;                  /* host wall clock of the whole call (incl. copies)         */
    double   ms_filter_fwd;            /* the nearest-neighbour filter kernel alone (one launch), forward   */
    double   ms_filter_rev;            /* ... reverse; the two launches overlap in time on two streams      */
    uint64_t cos_model;                /* 1/2 = glibc cos restated on the device (FMA / non-FMA build), 0 = */
                                       /* CUDA libm cos (host libm not recognised: see rows_ambiguous)      */
} xm_stats;

/* Order-preserving declination-band partition of a device-resident catalog: the exchange step of the
 * spatially decomposed multi-GPU search (vif_b200.distributed.qxmatch_banded; replaces the row chunks of
 * qxmatch.hpp:422-454 where a chunk would have to meet the whole other catalog). Band b owns
 * [edges[b], edges[b+1]) (nb + 1 ascending edges [deg] enclosing every dec); a row within `halo` [deg]
 * of an edge is also emitted for the band beyond it. out receives *total triples {ra, dec, bits} grouped
 * by destination band, rows ascending inside a band; bits (read the double as u64) = row0 + row, with
 * bit 63 set when the band is the row's own (primary) one. counts[nb] (device) = triples per band.
 * only >= 0 emits that band alone (selection of one band + halo out of a replicated catalog).
 * cap = triples `out` can hold (n + the halo rows; the call fails when it does not suffice). */
int32_t xm_band_partition_dev(const double* ra, const double* dec, uint64_t n, uint64_t row0, const double* edges,
                              int32_t nb, double halo, int32_t only, double* out, uint64_t cap, uint64_t* counts,
                              uint64_t* total, int32_t device, void* stream, char* errbuf, uint64_t errcap);

/* After the band x band search of qxmatch_banded. xm_band_finish_dev: id_io[j] (index into the band's
 * candidate array, -1 = none) becomes the candidate's GLOBAL row (bits_cand = third column of its triples,
 * contiguous); *bad (device u64, accumulated) counts primary rows (bit 63 of bits_src[j]) whose distance is
 * not below hlim [arcsec]: their neighbour may lie outside the band. xm_band_home_dev: results that
 * travelled back to the rows' owner rank -- got = m pairs {id bits, d} in the order of the m triples
 * `sent` this rank emitted; entries of primary triples are stored at row - row0 of id_out / d_out. */
int32_t xm_band_finish_dev(int64_t* id_io, const double* d, const int64_t* bits_src, const int64_t* bits_cand, uint64_t n,
                           double hlim, uint64_t* bad, int32_t device, void* stream, char* errbuf, uint64_t errcap);
int32_t xm_band_home_dev(const double* got, const double* sent, uint64_t m, uint64_t row0, int64_t* id_out, double* d_out,
                         int32_t device, void* stream, char* errbuf, uint64_t errcap);

/* The host libm's sin/cos decide the last bits of the reference's proxies (qxmatch.hpp:174-204); the
 * library restates glibc 2.39's algorithm for them (csrc/xm_libm_glibc.cuh) and xm_init() checks the host
 * libm against it. xm_libm_model: 1 = glibc's FMA build, 2 = its plain build, 0 = host libm not recognised
 * (the device then uses the CUDA libm and near ties are only counted: xm_stats.rows_ambiguous). Needs no
 * GPU. xm_libm_eval_host / _dev evaluate the restatement (which: 0 = sin, 1 = cos; |x| < 2.426, NaN
 * beyond) on the host / on n device values, for tests. */
int32_t xm_libm_model(void);
double  xm_libm_eval_host(int32_t which, int32_t model, double x);
int32_t xm_libm_eval_dev(int32_t which, int32_t model, const double* x, uint64_t n, double* out, int32_t device,
                         void* stream, char* errbuf, uint64_t errcap);

/* DFMA micro-benchmark: the FP64 roofline denominator of the brute-force path (BASELINE.md section 2:
 * "must be measured"). Runs `reps` launches of a kernel whose every thread carries 8 independent
 * chains of 4096 DFMAs (grid = 8 CTAs of 256 threads per SM), times them with CUDA events and
 * returns the best launch in *tflops (2 flop per DFMA). device = -1 -> current. */
int32_t xm_fp64_peak(int32_t device, int32_t reps, double* tflops, char* errbuf, uint64_t errcap);

/* Library / ABI version string, e.g. "vif_b200_qxmatch 0.1 (sm_100a)". Never fails. */
const char* xm_version(void);

/* Creates the per-device context (stream, memory pool, libm calibration). Optional: every entry
 * point calls it lazily. device = -1 -> current. Returns XM_OK or XM_ERR_CUDA. */
int32_t xm_init(int32_t device, char* errbuf, uint64_t errcap);

/* Releases all contexts and device memory owned by the library. */
void xm_shutdown(void);

/* Scratch memory is drawn from the device's stream-ordered pool and stays cached there between
 * calls (a 1e9-row call re-maps ~80 GB otherwise, four times the cost of its index build). This
 * hands the cached memory back to the driver; the environment variable XM_POOL_KEEP_GB caps what
 * the pool keeps instead. device = -1 -> current. Returns XM_OK or XM_ERR_CUDA. */
int32_t xm_release_memory(int32_t device);

/*
 * qxmatch on HOST buffers (the reference-facing call; replaces qxmatch.hpp:133-616).
 *   ra1,dec1: n1 doubles; ra2,dec2: n2 doubles.
 *   id, d:    max(nth,1)*n1 elements, caller-allocated, fully overwritten.
 *   rid, rd:  n2 elements, or NULL when params->no_mirror.
 * Copies inputs H2D, runs the device pipeline, copies results D2H. Blocking.
 */
int32_t xm_qxmatch_f64(const double* ra1, const double* dec1, uint64_t n1,
                       const double* ra2, const double* dec2, uint64_t n2,
                       const xm_params* params,
                       uint64_t* id, double* d, uint64_t* rid, double* rd,
                       char* errbuf, uint64_t errcap);

/*
 * Same, on DEVICE buffers already resident in HBM on params->device (or the current device).
 * `stream` is a cudaStream_t (NULL = the library's own stream); the call synchronises the stream
 * twice internally (bounding boxes -> host geometry; refine counts) and once before returning.
 */
int32_t xm_qxmatch_dev(const double* ra1, const double* dec1, uint64_t n1,
                       const double* ra2, const double* dec2, uint64_t n2,
                       const xm_params* params,
                       uint64_t* id, double* d, uint64_t* rid, double* rd,
                       void* stream, char* errbuf, uint64_t errcap);

/* Host-only: grid geometry from the two bounding boxes {ra_min,ra_max,dec_min,dec_max}
 * (restates qxmatch.hpp:232-267 with the host libm so that cell keys match the reference). */
int32_t xm_grid_from_bbox(const double* bbox1, const double* bbox2, uint64_t n2, uint64_t nth,
                          int32_t linear, xm_grid* out);

/* Device bounding box + finiteness of one catalog: out = {ra_min,ra_max,dec_min,dec_max},
 * *nonfinite = count of rows with a NaN/inf coordinate (qxmatch.hpp:146-149, 232-235). */
int32_t xm_bbox_dev(const double* ra, const double* dec, uint64_t n, int32_t device, void* stream,
                    double* out, uint64_t* nonfinite, char* errbuf, uint64_t errcap);

/* The same pass without a host round trip, in the layout one all-reduce(max) over the ranks of a row-sharded
 * run wants (qxmatch.hpp:232-238 over shards): out_dev (DEVICE memory, 5 doubles) receives
 * {-ra_min, ra_max, -dec_min, dec_max, number of non-finite rows}; an empty shard gives -inf four times.
 * Launches on `stream` and returns without synchronising. */
int32_t xm_bbox_neg_dev(const double* ra, const double* dec, uint64_t n, int32_t device, void* stream,
                        double* out_dev, char* errbuf, uint64_t errcap);

/* Statistics of the most recent qxmatch call on `device` (-1 = current). */
int32_t xm_last_stats(int32_t device, xm_stats* out);

/* Stable LSD radix sort of (key,value) u32 pairs on the device -- the sort the cell index is
 * built with; exported so tests can check it in isolation. key_bits in [1,32]. In place. */
int32_t xm_sort_pairs_dev(uint32_t* keys, uint32_t* vals, uint64_t n, int32_t key_bits,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap);

/*
 * Reverse-match merge for row-sharded multi-GPU runs (one exchange step; the
 * counterpart of the reference's per-thread merge, qxmatch.hpp:478-488 / 592-603). Each rank holds
 * (rid, rd) over ITS catalog-1 rows. The host interleaves three tiny kernels with two all-reduces:
 *   xm_rev_merge_pack   rid -> global row id (or INT64_MAX when unmatched), key = rd (or +inf)
 *   all-reduce(min) on a copy of key                                  -> best
 *   xm_rev_merge_mask   rid kept only where this rank's key equals best
 *   all-reduce(min) on rid (as int64)                                 -> smallest row among winners
 *   xm_rev_merge_unpack rid INT64_MAX -> XM_NPOS, rd = best (or `unmatched`: NaN / inf)
 * All pointers are device pointers on `device`; `stream` is a cudaStream_t or NULL.
 */
int32_t xm_rev_merge_pack(uint64_t* rid, const double* rd, uint64_t n, uint64_t row_offset, double* key,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap);
int32_t xm_rev_merge_mask(uint64_t* rid, const double* key_local, const double* key_best, uint64_t n,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap);
int32_t xm_rev_merge_unpack(uint64_t* rid, double* rd, const double* key_best, uint64_t n, double unmatched,
                            int32_t device, void* stream, char* errbuf, uint64_t errcap);

/*
 * The same merge as ONE kernel over NVLink peer memory (no NCCL collective on the data path).
 *   exch[r], r < world: peer-mapped device pointer to rank r's exchange buffer of 2n+1 8-byte
 *                       words: its n local rd (double), its n local rid (uint64, XM_NPOS =
 *                       unmatched) as xm_qxmatch_dev returned them, and its first catalog-1 row;
 *   out[r]:             peer-mapped pointer to rank r's result buffer: n rd, then n rid.
 * The calling rank merges rows [n*rank/world, n*(rank+1)/world) -- smallest distance, smallest
 * global row id among equal distances (qxmatch.hpp:510-513) -- and stores them into EVERY out[r]
 * that is not NULL (out[rank] is required; passing NULL for the peers leaves the merged reverse
 * match sliced over the ranks and halves the NVLink traffic of the step).
 * The caller brackets the call with barriers over the ranks: all exchange buffers filled before,
 * all stores landed after (vif_b200/distributed.py uses torch symmetric memory for both).
 * world <= XM_MAX_PEERS. Launches on `stream` and returns without synchronising.
 */
#define XM_MAX_PEERS 16
int32_t xm_rev_merge_p2p(const void* const* exch, void* const* out, uint64_t n, int32_t rank, int32_t world,
                         double unmatched, int32_t device, void* stream, char* errbuf, uint64_t errcap);

/*
 * Distance helpers of the same haversine family, device pointers on `device`:
 *   xm_angdist_dev       astro/astro.hpp:33-81: out[i] = angdist(ra1[i],dec1[i], ra2[k],dec2[k]) in arcsec,
 *                        k = i, or 0 when second_is_scalar (the (vec, vec, double, double) overload);
 *   xm_angdist_less_dev  astro/astro.hpp:83-103: out[i] = 1 where the point lies within `radius` arcsec
 *                        of (ra2, dec2) -- the proxy comparison, no asin;
 *   xm_qdist_dev         astro/qxmatch.hpp:684-717: n x n row-major matrix, out[j*n+i] = distance of
 *                        rows i and j in arcsec for j > i, 0 elsewhere.
 * Degrees in. Synchronous.
 */
int32_t xm_angdist_dev(const double* ra1, const double* dec1, const double* ra2, const double* dec2, uint64_t n,
                       int32_t second_is_scalar, double* out, int32_t device, void* stream, char* errbuf,
                       uint64_t errcap);
int32_t xm_angdist_less_dev(const double* ra1, const double* dec1, uint64_t n, double ra2, double dec2,
                            double radius, uint8_t* out, int32_t device, void* stream, char* errbuf,
                            uint64_t errcap);
int32_t xm_qdist_dev(const double* ra, const double* dec, uint64_t n, double* out, int32_t device, void* stream,
                     char* errbuf, uint64_t errcap);

/*
 * Mutual-best filter on the device (replaces xmatch_clean_best, qxmatch.hpp:641-663, while id/rid
 * are still resident): row i of catalog 1 is kept when rid[id[i]] == i. id = rank-0 plane (n1), rid
 * (n2). Outputs (caller-allocated, n1 elements each): id1/id2 = the kept pairs in ascending i, lost =
 * the other rows; counts[0] = pairs, counts[1] = lost (host pointer, filled before returning).
 */
int32_t xm_clean_best_dev(const uint64_t* id, const uint64_t* rid, uint64_t n1, uint64_t n2,
                          uint64_t* id1, uint64_t* id2, uint64_t* lost, uint64_t* counts,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap);

/* Host-only: entry list of ring `k` of the reference's depth_cache (qxmatch.hpp:44-128) for an
 * (nx, ny) grid, in visiting order, as the kernels enumerate it (closed form, no table). Writes up
 * to `cap` (bx, by) pairs, returns the entry count. dedupe != 0 drops the repeated axis cells. */
uint64_t xm_ring_entries(uint32_t k, uint32_t nx, uint32_t ny, int32_t dedupe,
                         int32_t* bx, int32_t* by, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* VIF_B200_QXMATCH_H */

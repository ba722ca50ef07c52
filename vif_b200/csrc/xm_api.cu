// C ABI + host orchestration of the device pipeline (see include/vif_b200_qxmatch.h).
//
// Pipeline of one call (cell-binned path), two stream lanes:
//   bbox(cat1), bbox(cat2)            -> 10 scalars D2H, host geometry (xm_geometry.cpp)
//   lane 1: index of catalog 1 (count -> scan -> place), lane 2: index of catalog 2
//   lane 1: forward search (filter kernel -> unpack || exact list kernel), lane 2: reverse search
// There is no CPU fallback anywhere in this file: without a usable device every entry point
// returns XM_ERR_CUDA.
#include <chrono>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>

#include "xm_common.cuh"
#include "xm_ring.cuh"

namespace xm {

namespace {
// Set by fail()/fail_cuda(): an entry point is unwinding with work possibly still in flight on the
// library's other streams. The first scratch buffer released after that waits for the whole device
// before it hands memory back to the pool (an error return must not let kernels of the
// second lane write into memory the pool may give to the next call).
thread_local bool t_error_pending = false;
thread_local cudaMemPool_t t_pool = nullptr;      // private pool of the context the thread is working in
}  // namespace

int32_t fail_cuda(Error& err, cudaError_t e, const char* what, const char* file, int line) {
    t_error_pending = true;
    err.code = XM_ERR_CUDA;
    snprintf(err.msg, sizeof(err.msg), "CUDA error %d (%s) in %s at %s:%d", (int)e,
             cudaGetErrorString(e), what, file, line);
    return err.code;
}

int32_t fail(Error& err, int32_t code, const char* fmt, ...) {
    t_error_pending = true;
    err.code = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err.msg, sizeof(err.msg), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

struct Context {
    int          device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // second lane: catalog-2 index, reverse search
    cudaEvent_t  sync_ev[4];
    std::mutex   mu;
    int          sin_model = 1;
    unsigned long long* d_small = nullptr;   // 16 u64: two bounding boxes + spare
    unsigned long long* h_small = nullptr;   // pinned mirror
    SearchCounters*     d_ctr = nullptr;
    SearchCounters*     h_ctr = nullptr;     // pinned
    xm_stats     last;
    cudaEvent_t  ev[10];
    cudaEvent_t  kev[4];                     // filter kernel alone: forward begin/end, reverse begin/end
    // host-buffer calls: results leave over PCIe as soon as their lane is done (copy stream), the
    // forward results while the reverse search is still running
    cudaStream_t copy = nullptr;
    cudaStream_t side[2] = {nullptr, nullptr};   // exact kernel of the undecided rows, one per lane
    cudaEvent_t  lane_ev[4];
    cudaMemPool_t pool = nullptr;            // private stream-ordered pool: scratch of every call
    int          cos_model = 0;
    struct HostOut { uint64_t* id = nullptr; double* d = nullptr; uint64_t* rid = nullptr; double* rd = nullptr; bool done = false; } host_out;
};

std::mutex g_mu;
std::map<int, std::unique_ptr<Context>> g_ctx;

int32_t get_context(int32_t device, Context** out, Error& err) {
    if (device < 0) XM_CUDA_TRY(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) { *out = it->second.get(); return XM_OK; }
    int count = 0;
    XM_CUDA_TRY(cudaGetDeviceCount(&count));
    if (device >= count) return fail(err, XM_ERR_CUDA, "CUDA device %d not present (%d visible)", device, count);
    int prev = 0;
    XM_CUDA_TRY(cudaGetDevice(&prev));
    XM_CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<Context> c(new Context);
    c->device = device;
    memset(&c->last, 0, sizeof(c->last));
    // the forward lane gets the higher priority: its CTAs are dispatched first while both searches are
    // in flight, so the forward results are complete (and on their way to the host) early
    int prio_lo = 0, prio_hi = 0;
    XM_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char* pe = getenv("XM_PRIORITY");
    if (pe && pe[0] == '0') prio_hi = prio_lo;
    XM_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    XM_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_lo));
    XM_CUDA_TRY(cudaStreamCreateWithPriority(&c->copy, cudaStreamNonBlocking, prio_hi));
    XM_CUDA_TRY(cudaStreamCreateWithPriority(&c->side[0], cudaStreamNonBlocking, prio_hi));
    XM_CUDA_TRY(cudaStreamCreateWithPriority(&c->side[1], cudaStreamNonBlocking, prio_lo));
    for (auto& e : c->lane_ev) XM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : c->sync_ev) XM_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    XM_CUDA_TRY(cudaMalloc(&c->d_small, 16*sizeof(unsigned long long)));
    XM_CUDA_TRY(cudaMallocHost(&c->h_small, 16*sizeof(unsigned long long)));
    XM_CUDA_TRY(cudaMalloc(&c->d_ctr, sizeof(SearchCounters)));
    XM_CUDA_TRY(cudaMallocHost(&c->h_ctr, sizeof(SearchCounters)));
    for (auto& e : c->ev) XM_CUDA_TRY(cudaEventCreate(&e));
    for (auto& e : c->kev) XM_CUDA_TRY(cudaEventCreate(&e));
    // Scratch comes from a pool of the library's own (not the device's default pool: its release
    // threshold is process-wide state other cudaMallocAsync users would inherit). Freed scratch
    // stays cached there: repeated calls never go back to the driver for memory (measured on
    // 1e9 x 1e8: index build 60 ms with the pool warm, 230 ms when it had to re-map its ~80 GB every
    // call). xm_release_memory() / XM_POOL_KEEP_GB bound the footprint; xm_shutdown destroys the pool.
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        XM_CUDA_TRY(cudaMemPoolCreate(&c->pool, &props));
        uint64_t keep = ~0ull;
        if (const char* e = getenv("XM_POOL_KEEP_GB")) keep = (uint64_t)strtoull(e, nullptr, 10) << 30;
        XM_CUDA_TRY(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    c->sin_model = calibrate_sin_model();
    XM_CUDA_TRY(cudaSetDevice(prev));
    *out = c.get();
    g_ctx[device] = std::move(c);
    return XM_OK;
}

// stream-ordered scratch buffer
struct Scratch {
    void* p = nullptr;
    cudaStream_t s = nullptr;
    ~Scratch() {
        if (!p) return;
        if (t_error_pending) { cudaDeviceSynchronize(); cudaGetLastError(); t_error_pending = false; }
        cudaFreeAsync(p, s);
    }
    cudaError_t alloc(size_t bytes, cudaStream_t st) {
        s = st;
        if (t_pool) return cudaMallocFromPoolAsync(&p, bytes ? bytes : 16, t_pool, st);
        return cudaMallocAsync(&p, bytes ? bytes : 16, st);
    }
    template <class T> T* as() const { return (T*)p; }
};

// per-device serialisation of the entry points + the thread's scratch pool / error state
struct CtxLock {
    std::lock_guard<std::mutex> lk;
    explicit CtxLock(Context* c) : lk(c->mu) { t_pool = c->pool; t_error_pending = false; }
    ~CtxLock() { t_pool = nullptr; }
};

// Callers compiled against an older header pass a shorter xm_params (struct_size): fields they do not
// know are zero = reference behaviour.
int32_t normalize_params(const xm_params* in, xm_params* out, Error& err) {
    if (!in) return fail(err, XM_ERR_INVALID_ARG, "params is NULL");
    const uint64_t sz = in->struct_size;
    if (sz < offsetof(xm_params, has_bbox1) || sz > sizeof(xm_params))
        return fail(err, XM_ERR_INVALID_ARG, "xm_params.struct_size = %llu: set it to sizeof(xm_params) (XM_PARAMS_INIT); "
                    "this library knows sizes %zu..%zu", (unsigned long long)sz, offsetof(xm_params, has_bbox1),
                    sizeof(xm_params));
    memset(out, 0, sizeof(*out));
    memcpy(out, in, (size_t)sz);
    out->struct_size = sizeof(xm_params);
    return XM_OK;
}

struct DeviceGuard {
    int prev = -1;
    bool armed = false;
    cudaError_t set(int dev) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev != dev) { e = cudaSetDevice(dev); armed = (e == cudaSuccess); }
        return e;
    }
    ~DeviceGuard() { if (armed) cudaSetDevice(prev); }
};

double ord2f(unsigned long long o) {
    unsigned long long b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
    double v;
    memcpy(&v, &b, sizeof(v));
    return v;
}

struct SortedCat {
    Scratch rec, cell;
    CatView view{};
};

// One catalog -> cell-ordered records + CSR offsets. use_bucket: counting-sort placement
// (xm_bucket.cu); otherwise keys -> stable radix sort -> gather (no limit on cell occupancy).
int32_t build_index(const double* ra, const double* dec, uint32_t n, const GridView& g_in,
                    uint32_t ncell, bool use_bucket, bool order_cells, uint32_t* overflow, SortedCat& out,
                    cudaStream_t s,
                    Error& err, uint64_t* launches, cudaEvent_t ev_keys, cudaEvent_t ev_sort,
                    unsigned long long* off_grid = nullptr) {
    GridView g = g_in;
    g.off_grid = off_grid;
    XM_CUDA_TRY(out.rec.alloc((size_t)n*sizeof(Rec), s));
    XM_CUDA_TRY(out.cell.alloc(((size_t)ncell + 1)*sizeof(uint32_t), s));
    int32_t rc;
    if (n == 0) {                     // empty block of a partitioned run: every cell is empty
        XM_CUDA_TRY(cudaMemsetAsync(out.cell.p, 0, ((size_t)ncell + 1)*sizeof(uint32_t), s));
    } else if (use_bucket) {
        Scratch tmp, cursor, aux;
        if (order_cells) XM_CUDA_TRY(tmp.alloc((size_t)n*sizeof(Rec), s));
        XM_CUDA_TRY(cursor.alloc(bucket_index_scratch_words(n, ncell)*sizeof(uint32_t), s));
        XM_CUDA_TRY(aux.alloc(bucket_scan_aux_bytes(ncell), s));
        rc = launch_bucket_index(ra, dec, n, g, ncell, order_cells, tmp.as<Rec>(), out.rec.as<Rec>(),
                                 out.cell.as<uint32_t>(), cursor.as<uint32_t>(), aux.as<uint32_t>(), overflow, s,
                                 err, launches);
        if (rc) return rc;
        if (ev_keys) XM_CUDA_TRY(cudaEventRecord(ev_keys, s));
        if (ev_sort) XM_CUDA_TRY(cudaEventRecord(ev_sort, s));
    } else {
        Scratch key, perm, tmpk, tmpv, sscr;
        XM_CUDA_TRY(key.alloc((size_t)n*4, s));
        XM_CUDA_TRY(perm.alloc((size_t)n*4, s));
        XM_CUDA_TRY(tmpk.alloc((size_t)n*4, s));
        XM_CUDA_TRY(tmpv.alloc((size_t)n*4, s));
        XM_CUDA_TRY(sscr.alloc(sort_scratch_bytes(n), s));
        rc = launch_keys(ra, dec, n, g, key.as<uint32_t>(), perm.as<uint32_t>(), s, err, launches);
        if (rc) return rc;
        if (ev_keys) XM_CUDA_TRY(cudaEventRecord(ev_keys, s));
        int bits = 1;
        while (bits < 32 && (1ull << bits) < (uint64_t)ncell) ++bits;
        rc = launch_sort_pairs(key.as<uint32_t>(), perm.as<uint32_t>(), tmpk.as<uint32_t>(),
                               tmpv.as<uint32_t>(), n, bits, sscr.p, s, err, launches);
        if (rc) return rc;
        if (ev_sort) XM_CUDA_TRY(cudaEventRecord(ev_sort, s));
        rc = launch_cells_and_gather(key.as<uint32_t>(), perm.as<uint32_t>(), ra, dec, n, ncell,
                                     g.linear != 0, out.rec.as<Rec>(), out.cell.as<uint32_t>(), s, err,
                                     launches, g.sin_model, g);
        if (rc) return rc;
    }
    out.view.rec = out.rec.as<Rec>(); out.view.cstart = out.cell.as<uint32_t>(); out.view.n = n;
    return XM_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
    return ms;
}

// Exact spherical grid search (xm_sphere.cu): brute-force semantics at neighbourhood-search cost.
int32_t run_sphere(Context* ctx, cudaStream_t s, const double* ra1, const double* dec1, uint64_t n1,
                   const double* ra2, const double* dec2, uint64_t n2, bool same_cat, const double* bb1,
                   const double* bb2, uint32_t nth, bool self, bool mirror, int sin_model, uint64_t* id,
                   double* d, uint64_t* rid, double* rd, cudaEvent_t* ev, uint64_t* launches, Error& err) {
    int32_t rc;
    const double bb[4] = {fmin(bb1[0], bb2[0]), fmax(bb1[1], bb2[1]), fmin(bb1[2], bb2[2]), fmax(bb1[3], bb2[3])};
    GridView g;
    int wrap = 0;
    if (!sphere_grid(bb, n2, nth, &g, &wrap))
        return fail(err, XM_ERR_INVALID_ARG, "right ascensions span more than 360 degrees");
    g.sin_model = sin_model;
    ctx->last.grid.rra0 = g.rra0; ctx->last.grid.rdec0 = g.rdec0; ctx->last.grid.dra = g.dra; ctx->last.grid.ddec = g.ddec;
    ctx->last.grid.nra = g.nra; ctx->last.grid.ndec = g.ndec; ctx->last.grid.cell_size = g.cell_size;
    const uint32_t ncell = g.nra*g.ndec;
    Scratch cmin;
    XM_CUDA_TRY(cmin.alloc((size_t)g.ndec*8, s));
    if ((rc = launch_sphere_rows(g, cmin.as<double>(), s, err, launches))) return rc;
    // two lanes, as in the cell-binned path: s = catalog-1 index + forward, s2 = catalog-2 index + reverse
    cudaStream_t s2 = ctx->stream2;
    cudaEvent_t* sev = ctx->sync_ev;
    XM_CUDA_TRY(cudaEventRecord(sev[0], s));
    XM_CUDA_TRY(cudaStreamWaitEvent(s2, sev[0], 0));
    SortedCat s1, s2cat;
    if ((rc = build_index(ra1, dec1, (uint32_t)n1, g, ncell, true, false, &ctx->d_ctr->index_overflow, s1, s, err,
                          launches, nullptr, nullptr))) return rc;
    const CatView* v2 = &s1.view;
    if (!same_cat) {
        if ((rc = build_index(ra2, dec2, (uint32_t)n2, g, ncell, true, false, &ctx->d_ctr->index_overflow, s2cat, s2,
                              err, launches, nullptr, nullptr))) return rc;
        v2 = &s2cat.view;
    }
    XM_CUDA_TRY(cudaEventRecord(sev[1], s));
    XM_CUDA_TRY(cudaEventRecord(sev[2], s2));
    XM_CUDA_TRY(cudaStreamWaitEvent(s, sev[2], 0));
    XM_CUDA_TRY(cudaStreamWaitEvent(s2, sev[1], 0));
    XM_CUDA_TRY(cudaEventRecord(ev[5], s));
    // nth = 1: undecided-row lists of the filter kernel and its sector-sized result records (the records
    // are optional: a catalog too large to afford 32 bytes per row scatters 8-byte stores instead)
    Scratch rl1, rl2, rr1, rr2;
    XM_CUDA_TRY(rl1.alloc((size_t)n1*4, s));
    if (nth == 1 && rr1.alloc(filter_result_bytes(n1), s) != cudaSuccess) { cudaGetLastError(); rr1.p = nullptr; }
    if ((rc = launch_sphere_search(g, wrap, cmin.as<double>(), s1.view, *v2, self, nth, n1, id, d, rl1.as<uint32_t>(),
                                   &ctx->d_ctr->refine_fwd, rr1.p, false, ctx->d_ctr, s, ctx->side[0], ctx->lane_ev[0],
                                   ctx->lane_ev[1], err, launches))) return rc;
    XM_CUDA_TRY(cudaEventRecord(ev[6], s));
    XM_CUDA_TRY(cudaEventRecord(ev[9], s2));
    if (mirror) {
        XM_CUDA_TRY(rl2.alloc((size_t)n2*4, s2));
        if (rr2.alloc(filter_result_bytes(n2), s2) != cudaSuccess) { cudaGetLastError(); rr2.p = nullptr; }
        if ((rc = launch_sphere_search(g, wrap, cmin.as<double>(), *v2, s1.view, self, 1u, n2, rid, rd, rl2.as<uint32_t>(),
                                       &ctx->d_ctr->refine_rev, rr2.p, true, ctx->d_ctr, s2, ctx->side[1],
                                       ctx->lane_ev[2], ctx->lane_ev[3], err, launches))) return rc;
    }
    XM_CUDA_TRY(cudaEventRecord(ev[7], s2));
    XM_CUDA_TRY(cudaEventRecord(sev[3], s2));
    XM_CUDA_TRY(cudaStreamWaitEvent(s, sev[3], 0));
    XM_CUDA_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(SearchCounters), cudaMemcpyDeviceToHost, s));
    XM_CUDA_TRY(cudaEventRecord(ev[8], s));
    XM_CUDA_TRY(cudaStreamSynchronize(s));
    XM_CUDA_TRY(cudaStreamSynchronize(s2));
    return XM_OK;
}

int32_t run_device_once(Context* ctx, cudaStream_t s, const double* ra1, const double* dec1, uint64_t n1,
                        const double* ra2, const double* dec2, uint64_t n2, const xm_params& P,
                        uint64_t* id, double* d, uint64_t* rid, double* rd, bool use_bucket, Error& err) {
    const uint64_t nth = P.nth < 1 ? 1 : P.nth;
    const bool mirror = !P.no_mirror;
    const bool linear = P.linear != 0;
    const bool self = P.self != 0;
    const double dnan = __builtin_nan(""), dinf = __builtin_inf();
    xm_stats& st = ctx->last;
    const int sin_model = ctx->sin_model;
    memset(&st, 0, sizeof(st));
    st.sin_model = (uint64_t)sin_model;
    st.cos_model = (uint64_t)sin_model;        // one libm model serves sin and cos (xm_libm_glibc.cuh)
    uint64_t launches = 0;
    int32_t rc;

    if (n1 >= 0xfffffffeull || n2 >= 0xfffffffeull)
        return fail(err, XM_ERR_INVALID_ARG, "catalogs of 2^32-2 rows or more are not supported");
    if (nth > 0xffffffffull) return fail(err, XM_ERR_INVALID_ARG, "nth is too large");

    // partitioned run: outputs cover row ranges of the (whole) catalogs
    const bool ranged = P.has_ranges != 0;
    if (ranged) {
        if (P.brute_force || P.linear)
            return fail(err, XM_ERR_INVALID_ARG, "row ranges are only supported by the spherical cell-binned path");
        if (P.self && !(ra1 == ra2 && dec1 == dec2 && n1 == n2))
            return fail(err, XM_ERR_INVALID_ARG, "row ranges with self need catalog 2 to be catalog 1 (same pointers)");
        if (P.fwd_begin + P.fwd_count > n1 || P.rev_begin + P.rev_count > n2)
            return fail(err, XM_ERR_INVALID_ARG, "row range outside the catalog");
    }
    const uint64_t n1o = ranged ? P.fwd_count : n1, n2o = ranged ? P.rev_count : n2;   // rows the outputs cover

    cudaEvent_t* ev = ctx->ev;
    XM_CUDA_TRY(cudaEventRecord(ev[0], s));

    // qxmatch.hpp:157-167: outputs start as npos / inf and an empty catalog returns them as is
    if (n1 == 0 || n2 == 0) {
        if ((rc = launch_fill_outputs(id, d, nth*n1o, dinf, s, err, &launches))) return rc;
        if (mirror && (rc = launch_fill_outputs(rid, rd, n2o, dinf, s, err, &launches))) return rc;
        XM_CUDA_TRY(cudaStreamSynchronize(s));
        st.launches = launches;
        return XM_OK;
    }

    // bounding boxes + finiteness (qxmatch.hpp:146-149, 232-235)
    const bool same_cat = (ra1 == ra2 && dec1 == dec2 && n1 == n2);
    cudaEvent_t cat2_ready = (cudaEvent_t)P.cat2_ready_event;
    // Both boxes promised by the caller (row-sharded multi-GPU runs: the boxes come out of one all-reduce
    // while catalog 2 is still being broadcast): the grid is known without reading a row, so the call does
    // not start with a bounding-box pass and a host round trip, and lane s2 alone waits for catalog 2. That
    // the promise held -- every row finite and on the grid -- is checked by the count pass of the index build
    // (SearchCounters::off_grid1/2) and reported when the call ends. Only where the reference's grid is feasible.
    bool promised = P.has_bbox1 && P.has_bbox2 && !P.brute_force && !same_cat && !ranged;
    if (promised) {
        xm_grid Gp;
        grid_from_bbox(P.bbox1, P.bbox2, n2, nth, linear, &Gp);
        if (Gp.nra == 0 || Gp.ndec == 0 || Gp.nra > 0x7fffffffull || Gp.ndec > 0x7fffffffull || Gp.nra*Gp.ndec > XM_MAX_CELLS)
            promised = false;
    }
    double bb1[4], bb2[4];
    if (promised) {
        XM_CUDA_TRY(cudaEventRecord(ev[1], s));
        for (int k = 0; k < 4; ++k) { bb1[k] = P.bbox1[k]; bb2[k] = P.bbox2[k]; }
    } else {
        if (cat2_ready) XM_CUDA_TRY(cudaStreamWaitEvent(s, cat2_ready, 0));
        if (!same_cat) {
            // the two passes are independent streams of loads: side by side on two streams they share the HBM
            // bandwidth instead of each leaving a third of it unused (2 x 37 us -> 50 us on 1e7-row catalogs)
            XM_CUDA_TRY(cudaEventRecord(ctx->sync_ev[0], s));
            XM_CUDA_TRY(cudaStreamWaitEvent(ctx->stream2, ctx->sync_ev[0], 0));
            if ((rc = launch_bbox(ra2, dec2, n2, ctx->d_small + 5, ctx->stream2, err, &launches))) return rc;
            XM_CUDA_TRY(cudaEventRecord(ctx->sync_ev[1], ctx->stream2));
        }
        if ((rc = launch_bbox(ra1, dec1, n1, ctx->d_small, s, err, &launches))) return rc;
        if (!same_cat) XM_CUDA_TRY(cudaStreamWaitEvent(s, ctx->sync_ev[1], 0));
        XM_CUDA_TRY(cudaMemcpyAsync(ctx->h_small, ctx->d_small, 10*sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, s));
        XM_CUDA_TRY(cudaEventRecord(ev[1], s));
        XM_CUDA_TRY(cudaStreamSynchronize(s));
        if (same_cat) memcpy(ctx->h_small + 5, ctx->h_small, 5*sizeof(unsigned long long));
        const unsigned long long* h = ctx->h_small;
        if (h[4])
            return fail(err, XM_ERR_NONFINITE_1, "first RA and Dec coordinates contain invalid values (infinite or NaN)");
        if (h[9])
            return fail(err, XM_ERR_NONFINITE_2, "second RA and Dec coordinates contain invalid values (infinite or NaN)");
        for (int k = 0; k < 4; ++k) { bb1[k] = ord2f(h[k]); bb2[k] = ord2f(h[5 + k]); }
        if (P.has_bbox1) {
            // the promised box must contain the rows passed
            if (P.bbox1[0] > bb1[0] || P.bbox1[1] < bb1[1] || P.bbox1[2] > bb1[2] || P.bbox1[3] < bb1[3])
                return fail(err, XM_ERR_INVALID_ARG, "bbox1 does not contain the catalog-1 rows passed");
            for (int k = 0; k < 4; ++k) bb1[k] = P.bbox1[k];
        }
    }

    XM_CUDA_TRY(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(SearchCounters), s));

    GridView g;
    memset(&g, 0, sizeof(g));
    g.linear = linear ? 1 : 0;
    g.sin_model = sin_model;

    bool sphere_done = false;
    if (P.brute_force && sphere_applicable(linear, (uint32_t)nth) &&
        (P.exact_grid || (double)n1*(double)n2 > 4e12)) {
        // ---- brute_force result through the exact spherical grid search ----
        if ((rc = run_sphere(ctx, s, ra1, dec1, n1, ra2, dec2, n2, same_cat, bb1, bb2, (uint32_t)nth, self, mirror,
                             sin_model, id, d, rid, rd, ev, &launches, err))) return rc;
        sphere_done = true;
        st.ms_bbox = ev_ms(ev[0], ev[1]);
        st.ms_index = ev_ms(ev[1], ev[5]);
    } else if (P.brute_force) {
        // ---- brute force (qxmatch.hpp:490-526) ----
        Scratch x1, y1, c1, x2, y2, c2;
        XM_CUDA_TRY(x1.alloc(n1*8, s)); XM_CUDA_TRY(y1.alloc(n1*8, s)); XM_CUDA_TRY(c1.alloc(n1*8, s));
        if ((rc = launch_prep_unsorted(ra1, dec1, (uint32_t)n1, linear, x1.as<double>(), y1.as<double>(),
                                       c1.as<double>(), s, err, &launches, sin_model))) return rc;
        const double *px2 = x1.as<double>(), *py2 = y1.as<double>(), *pc2 = c1.as<double>();
        if (!same_cat) {
            XM_CUDA_TRY(x2.alloc(n2*8, s)); XM_CUDA_TRY(y2.alloc(n2*8, s)); XM_CUDA_TRY(c2.alloc(n2*8, s));
            if ((rc = launch_prep_unsorted(ra2, dec2, (uint32_t)n2, linear, x2.as<double>(),
                                           y2.as<double>(), c2.as<double>(), s, err, &launches, sin_model))) return rc;
            px2 = x2.as<double>(); py2 = y2.as<double>(); pc2 = c2.as<double>();
        }
        const bool fast = !P.exact_only && brute_fast_applicable(linear, (uint32_t)nth);
        Scratch u1, u2, rlist, rthr;
        const void *pu1 = nullptr, *pu2 = nullptr;
        if (fast) {
            // unit vectors of both catalogs for the dot-product filter (xm_brute_fast.cu)
            XM_CUDA_TRY(u1.alloc(brute_uvec_bytes(n1), s));
            if ((rc = launch_unit_vectors(x1.as<double>(), y1.as<double>(), c1.as<double>(), (uint32_t)n1,
                                          u1.p, s, err, &launches))) return rc;
            pu1 = u1.p; pu2 = u1.p;
            if (!same_cat) {
                XM_CUDA_TRY(u2.alloc(brute_uvec_bytes(n2), s));
                if ((rc = launch_unit_vectors(px2, py2, pc2, (uint32_t)n2, u2.p, s, err, &launches))) return rc;
                pu2 = u2.p;
            }
            XM_CUDA_TRY(rlist.alloc((size_t)(n1 > n2 ? n1 : n2)*4, s));
            XM_CUDA_TRY(rthr.alloc((size_t)(n1 > n2 ? n1 : n2)*8, s));
        }
        XM_CUDA_TRY(cudaEventRecord(ev[5], s));
        if (fast) {
            if ((rc = launch_brute_fast(sin_model, pu1, (uint32_t)n1, pu2, (uint32_t)n2, x1.as<double>(),
                                        y1.as<double>(), c1.as<double>(), px2, py2, pc2, self, (uint32_t)nth,
                                        n1, id, d, rlist.as<uint32_t>(), rthr.as<double>(),
                                        &ctx->d_ctr->refine_fwd, ctx->d_ctr, s, err, &launches))) return rc;
        } else {
            if ((rc = launch_brute_exact(g, x1.as<double>(), y1.as<double>(), c1.as<double>(), (uint32_t)n1,
                                         px2, py2, pc2, (uint32_t)n2, self, (uint32_t)nth, n1, id, d,
                                         nullptr, (uint32_t)n1, false, ctx->d_ctr, s, err, &launches))) return rc;
        }
        XM_CUDA_TRY(cudaEventRecord(ev[6], s));
        if (mirror) {
            // reverse match = the same search with the roles swapped: ascending i, strict '<',
            // i.e. the thread=1 rule (smallest i wins exact ties, qxmatch.hpp:510-513)
            if (fast) {
                if ((rc = launch_brute_fast(sin_model, pu2, (uint32_t)n2, pu1, (uint32_t)n1, px2, py2, pc2,
                                            x1.as<double>(), y1.as<double>(), c1.as<double>(), self, 1u, n2,
                                            rid, rd, rlist.as<uint32_t>(), rthr.as<double>(),
                                            &ctx->d_ctr->refine_rev, ctx->d_ctr, s, err, &launches))) return rc;
            } else {
                if ((rc = launch_brute_exact(g, px2, py2, pc2, (uint32_t)n2, x1.as<double>(), y1.as<double>(),
                                             c1.as<double>(), (uint32_t)n1, self, 1u, n2, rid, rd, nullptr,
                                             (uint32_t)n2, true, ctx->d_ctr, s, err, &launches))) return rc;
            }
        }
        XM_CUDA_TRY(cudaEventRecord(ev[7], s));
        XM_CUDA_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(SearchCounters), cudaMemcpyDeviceToHost, s));
        XM_CUDA_TRY(cudaEventRecord(ev[8], s));
        XM_CUDA_TRY(cudaStreamSynchronize(s));
        st.ms_bbox = ev_ms(ev[0], ev[1]);
        st.ms_index = ev_ms(ev[1], ev[5]);
    } else {
        // ---- cell-binned path (qxmatch.hpp:230-489) ----
        xm_grid G;
        grid_from_bbox(bb1, bb2, n2, nth, linear, &G);
        st.grid = G;
        const bool too_large = G.nra == 0 || G.ndec == 0 || G.nra > 0x7fffffffull || G.ndec > 0x7fffffffull ||
                               G.nra*G.ndec > XM_MAX_CELLS;
        if (too_large && ranged)
            return fail(err, XM_ERR_GRID_TOO_LARGE, "row ranges are not supported where the reference's bucket grid is infeasible");
        if (too_large && sphere_applicable(linear, (uint32_t)nth)) {
            // The reference cannot run here (it would allocate the grid). Answer with the exact
            // neighbours instead -- what its binned search returns wherever that search is exact.
            if ((rc = run_sphere(ctx, s, ra1, dec1, n1, ra2, dec2, n2, same_cat, bb1, bb2, (uint32_t)nth, self,
                                 mirror && !self, sin_model, id, d, rid, rd, ev, &launches, err))) return rc;
            if (mirror && self) {     // binned + self: the reverse pass is skipped (qxmatch.hpp:395)
                if ((rc = launch_fill_outputs(rid, rd, n2, dnan, s, err, &launches))) return rc;
                XM_CUDA_TRY(cudaStreamSynchronize(s));
            }
            sphere_done = true;
            st.ms_bbox = ev_ms(ev[0], ev[1]);
            st.ms_index = ev_ms(ev[1], ev[5]);
            goto finish;
        }
        if (too_large)
            return fail(err, XM_ERR_GRID_TOO_LARGE,
                        "bucket grid of %llu x %llu cells (cell size %.3g) is too large; the catalogs "
                        "span too wide a field for the cell-binned path (use brute_force)",
                        (unsigned long long)G.nra, (unsigned long long)G.ndec, G.cell_size);
        const uint32_t ncell = (uint32_t)(G.nra*G.ndec);
        g.rra0 = G.rra0; g.rdec0 = G.rdec0; g.dra = G.dra; g.ddec = G.ddec; g.cell_size = G.cell_size;
        g.nra = (uint32_t)G.nra; g.ndec = (uint32_t)G.ndec;
        g.max_depth = g.nra > g.ndec ? g.nra : g.ndec;
        const double unit = linear ? 1.0 : XM_D2R;
        g.ox = G.rra0*unit; g.oy = G.rdec0*unit; g.wx = G.dra*unit; g.wy = G.ddec*unit;
        // binning rounds (v - origin)/width: edges are uncertain by a few ulp of the largest
        // coordinate magnitude; 1e-12 relative leaves three orders of magnitude of margin
        const double mx = fmax(fmax(fabs(G.rra0), fabs(G.rra1)), G.dra);
        const double my = fmax(fmax(fabs(G.rdec0), fabs(G.rdec1)), G.ddec);
        g.fzx = 1e-12*mx*unit; g.fzy = 1e-12*my*unit;

        g.cs_rad = G.cell_size*(XM_DPI/180.0/3600.0);
        {   // filter-kernel bounds (xm_search_fast.cu): H = largest half coordinate difference
            const double H = 1.5*fmax(g.wx, g.wy);
            g.eps_lo = H*H/3.0 + 1e-14;
            g.eps_hi = 1e-14;
            g.stop_delta = g.cs_rad*(1e-9 + 2.0*H*H);
        }
        Scratch cmin, ccen;
        XM_CUDA_TRY(cmin.alloc((size_t)g.ndec*8, s));
        XM_CUDA_TRY(ccen.alloc((size_t)g.ndec*8, s));
        g.cmin_row = cmin.as<double>();
        g.ccen_row = ccen.as<double>();
        if (!linear && (rc = launch_row_tables(g, cmin.as<double>(), ccen.as<double>(), s, err, &launches))) return rc;

        // qxmatch.hpp:383-387: nth is lowered to n2 after allocation; planes >= n2 keep npos / NaN
        const uint32_t nth_eff = (uint32_t)(n2 < nth ? n2 : nth);
        // partitioned self match: the searched block is indexed on its own, so "itself" is told by
        // the original row (block row + block offset), not by the slot
        const int self_mode = self ? ((same_cat && !ranged) ? 1 : 2) : 0;
        g.self_off = (self && ranged) ? (uint32_t)P.fwd_begin : 0u;
        // nth = 1 on small cells away from the poles: filter kernel first, exact kernel only on
        // the rows it flags (list length stays on the device: no host round trip)
        const bool fast = !P.exact_only && fast_path_applicable(g) && self_mode != 2;
        // Nearest-neighbour-only calls never depend on the order inside a cell (the filter flags
        // every tie, the K=1 exact kernel breaks in-cell ties by row), so the bucket index may
        // leave cells in arrival order.
        const bool unordered = use_bucket && fast && nth_eff == 1;
        // XM_FILTER=v1 keeps the FP64 filter kernel (xm_search_fast.cu) for every call; by default it only
        // serves the catalogs whose density does not suit the FP32 kernel's stage (xm_search_nn32.cu)
        const char* fenv = getenv("XM_FILTER");
        const bool filter_v1 = fenv && strcmp(fenv, "v1") == 0;
        // Two lanes: `s` indexes catalog 1 and runs the forward search, `s2` indexes catalog 2 and
        // runs the reverse search. The lanes only meet where one needs the other's index.
        cudaStream_t s2 = ctx->stream2;
        cudaEvent_t* sev = ctx->sync_ev;
        XM_CUDA_TRY(cudaEventRecord(sev[0], s));                       // inputs / row tables ready
        XM_CUDA_TRY(cudaStreamWaitEvent(s2, sev[0], 0));
        if (promised && cat2_ready) XM_CUDA_TRY(cudaStreamWaitEvent(s2, cat2_ready, 0));   // lane s2 alone waits for catalog 2
        SortedCat s1, s2cat;
        if ((rc = build_index(ra1, dec1, (uint32_t)n1, g, ncell, use_bucket, !unordered,
                              &ctx->d_ctr->index_overflow, s1, s, err, &launches, ev[2], ev[3],
                              &ctx->d_ctr->off_grid1))) return rc;
        const CatView* v2 = &s1.view;
        if (!same_cat) {
            if ((rc = build_index(ra2, dec2, (uint32_t)n2, g, ncell, use_bucket, !unordered,
                                  &ctx->d_ctr->index_overflow, s2cat, s2, err, &launches, nullptr, nullptr,
                                  &ctx->d_ctr->off_grid2))) return rc;
            v2 = &s2cat.view;
        }
        // partitioned run: the searched rows are blocks of the catalogs, indexed on their own; the
        // whole catalogs stay the candidates. Block records carry block-local rows, so results land
        // at [0, count) of the (block-sized) outputs; rid values come from the whole catalog 1.
        SortedCat blk1, blk2;
        const CatView* srcF = &s1.view;
        const CatView* srcR = v2;
        if (ranged) {
            if ((rc = build_index(ra1 + P.fwd_begin, dec1 + P.fwd_begin, (uint32_t)P.fwd_count, g, ncell, use_bucket,
                                  !unordered, &ctx->d_ctr->index_overflow, blk1, s, err, &launches, nullptr, nullptr))) return rc;
            if ((rc = build_index(ra2 + P.rev_begin, dec2 + P.rev_begin, (uint32_t)P.rev_count, g, ncell, use_bucket,
                                  !unordered, &ctx->d_ctr->index_overflow, blk2, s2, err, &launches, nullptr, nullptr))) return rc;
            srcF = &blk1.view; srcR = &blk2.view;
        }
        XM_CUDA_TRY(cudaEventRecord(sev[1], s));                       // index of catalog 1 ready
        XM_CUDA_TRY(cudaEventRecord(sev[2], s2));                      // index of catalog 2 ready
        XM_CUDA_TRY(cudaStreamWaitEvent(s, sev[2], 0));
        XM_CUDA_TRY(cudaStreamWaitEvent(s2, sev[1], 0));
        XM_CUDA_TRY(cudaEventRecord(ev[5], s));

        if (nth_eff < nth) {
            if ((rc = launch_fill_outputs(id, d, nth*n1o, linear ? dinf : dnan, s, err, &launches))) return rc;
        }
        // ---- forward search (lane s) ----
        bool timed_f = false, timed_r = false, fwd_deferred = false;
        int tile_f = 0;
        Scratch refine, resrec, plans;
        // what follows the forward filter: (unpack, when deferred) || exact kernel of the undecided rows on a side
        // stream -> join
        auto fwd_post = [&]() -> int32_t {
            if (fwd_deferred) {
                XM_CUDA_TRY(cudaStreamWaitEvent(s, ctx->lane_ev[2], 0));             // the reverse filter has run
                int32_t r = launch_unpack_results(resrec.p, srcF->n, id, d, s, err, &launches);
                if (r) return r;
                XM_CUDA_TRY(cudaStreamWaitEvent(ctx->side[0], ctx->lane_ev[2], 0));
            }
            XM_CUDA_TRY(cudaStreamWaitEvent(ctx->side[0], ctx->lane_ev[0], 0));
            int32_t r = launch_ring_search_exact(g, *srcF, *v2, self_mode, 1u, n1o, id, d, refine.as<uint32_t>(),
                                                 (uint32_t)n1o, &ctx->d_ctr->refine_fwd, false, unordered, ctx->d_ctr,
                                                 ctx->side[0], err, &launches);
            if (r) return r;
            XM_CUDA_TRY(cudaEventRecord(ctx->lane_ev[1], ctx->side[0]));
            XM_CUDA_TRY(cudaStreamWaitEvent(s, ctx->lane_ev[1], 0));
            return XM_OK;
        };
        if (srcF->n == 0) {
            // empty block: nothing to match
        } else if (fast && nth_eff == 1) {
            XM_CUDA_TRY(refine.alloc((size_t)n1o*4, s));
            XM_CUDA_TRY(resrec.alloc(filter_result_bytes(n1o), s));
            // filter -> (unpack on s || exact kernel of the undecided rows on a side stream) -> join
            const int tile = filter_v1 ? 0 : nn32_tile_sources(g, srcF->n, v2->n);
            if (tile) {
                XM_CUDA_TRY(plans.alloc(nn32_plan_bytes(n1o, g.nra, tile), s));
                // Device-resident results, both directions through the persistent filter: the forward lane's unpack
                // pass and exact list wait until the reverse filter has run. Beside it they cost it 0.13 ms (0.64
                // against 0.50 ms per launch) -- more than they take themselves -- while after it they overlap with the
                // reverse lane's own unpack pass. With host outputs the order stays: the forward results must leave
                // for the host as early as possible. XM_NN32_NODEFER=1 keeps the round-1 order.
                fwd_deferred = !ctx->host_out.id && mirror && n2o && !self && tile == nn32_persistent_tile() &&
                               nn32_tile_sources(g, srcR->n, s1.view.n) == tile && !getenv("XM_NN32_NODEFER") &&
                               !getenv("XM_NN32_CONCURRENT");
                if ((rc = launch_ring_nn32(g, *srcF, *v2, self, tile, id, d, resrec.p, plans.p, refine.as<uint32_t>(),
                                           &ctx->d_ctr->refine_fwd, false, ctx->d_ctr, s, ctx->lane_ev[0], err,
                                           &launches, ctx->kev[0], ctx->kev[1], nullptr, fwd_deferred))) return rc;
                timed_f = true;
                tile_f = tile;
            } else {
                XM_CUDA_TRY(plans.alloc(filter_plan_bytes(n1o), s));
                if ((rc = launch_ring_nn_filter(g, *srcF, *v2, self, id, d, resrec.p, plans.p, refine.as<uint32_t>(),
                                                &ctx->d_ctr->refine_fwd, false, ctx->d_ctr, s, ctx->lane_ev[0], err,
                                                &launches))) return rc;
            }
            if (!fwd_deferred && (rc = fwd_post())) return rc;
        } else if (!P.exact_only && fast_path_applicable(g) && topk_filter_applicable(nth_eff)) {
            // nth = 2..8: top-K filter kernel, exact kernel on the rows it flags (cells are in row order)
            XM_CUDA_TRY(refine.alloc((size_t)n1o*4, s));
            if ((rc = launch_ring_topk_filter(g, *srcF, *v2, self, nth_eff, n1o, id, d, refine.as<uint32_t>(),
                                              &ctx->d_ctr->refine_fwd, ctx->d_ctr, s, err, &launches))) return rc;
            if ((rc = launch_ring_search_exact(g, *srcF, *v2, self_mode, nth_eff, n1o, id, d, refine.as<uint32_t>(),
                                               (uint32_t)n1o, &ctx->d_ctr->refine_fwd, false, false, ctx->d_ctr, s,
                                               err, &launches))) return rc;
        } else {
            if ((rc = launch_ring_search_exact(g, *srcF, *v2, self_mode, nth_eff, n1o, id, d, nullptr,
                                               (uint32_t)n1o, nullptr, false, false, ctx->d_ctr, s, err, &launches))) return rc;
        }
        if (!fwd_deferred) XM_CUDA_TRY(cudaEventRecord(ev[6], s));
        Context::HostOut& ho = ctx->host_out;
        if (ho.id) {
            XM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy, ev[6], 0));
            XM_CUDA_TRY(cudaMemcpyAsync(ho.id, id, nth*n1o*8, cudaMemcpyDeviceToHost, ctx->copy));
            XM_CUDA_TRY(cudaMemcpyAsync(ho.d, d, nth*n1o*8, cudaMemcpyDeviceToHost, ctx->copy));
        }
        // ---- reverse search (lane s2) ----
        Scratch refine2, resrec2, plans2;
        XM_CUDA_TRY(cudaEventRecord(ev[9], s2));
        if (mirror && n2o) {
            if (!self) {
                if (fast) {
                    XM_CUDA_TRY(refine2.alloc((size_t)n2o*4, s2));
                    XM_CUDA_TRY(resrec2.alloc(filter_result_bytes(n2o), s2));
                    const int tile = filter_v1 ? 0 : nn32_tile_sources(g, srcR->n, s1.view.n);
                    if (tile) {
                        XM_CUDA_TRY(plans2.alloc(nn32_plan_bytes(n2o, g.nra, tile), s2));
                        if ((rc = launch_ring_nn32(g, *srcR, s1.view, false, tile, rid, rd, resrec2.p, plans2.p,
                                                   refine2.as<uint32_t>(), &ctx->d_ctr->refine_rev, true, ctx->d_ctr,
                                                   s2, ctx->lane_ev[2], err, &launches, ctx->kev[2], ctx->kev[3],
                                                   (timed_f && tile_f == tile && !getenv("XM_NN32_CONCURRENT")) ? ctx->kev[1] : nullptr))) return rc;
                        timed_r = true;
                        if (fwd_deferred) {
                            if ((rc = fwd_post())) return rc;
                            XM_CUDA_TRY(cudaEventRecord(ev[6], s));
                        }
                    } else {
                        XM_CUDA_TRY(plans2.alloc(filter_plan_bytes(n2o), s2));
                        if ((rc = launch_ring_nn_filter(g, *srcR, s1.view, false, rid, rd, resrec2.p, plans2.p,
                                                        refine2.as<uint32_t>(), &ctx->d_ctr->refine_rev, true, ctx->d_ctr,
                                                        s2, ctx->lane_ev[2], err, &launches))) return rc;
                    }
                    XM_CUDA_TRY(cudaStreamWaitEvent(ctx->side[1], ctx->lane_ev[2], 0));
                    if ((rc = launch_ring_search_exact(g, *srcR, s1.view, 0, 1u, n2o, rid, rd,
                                                       refine2.as<uint32_t>(), (uint32_t)n2o,
                                                       &ctx->d_ctr->refine_rev, true, unordered, ctx->d_ctr,
                                                       ctx->side[1], err, &launches))) return rc;
                    XM_CUDA_TRY(cudaEventRecord(ctx->lane_ev[3], ctx->side[1]));
                    XM_CUDA_TRY(cudaStreamWaitEvent(s2, ctx->lane_ev[3], 0));
                } else {
                    if ((rc = launch_ring_search_exact(g, *srcR, s1.view, 0, 1u, n2o, rid, rd, nullptr,
                                                       (uint32_t)n2o, nullptr, true, false, ctx->d_ctr, s2, err, &launches))) return rc;
                }
            } else {
                // reverse pass skipped (qxmatch.hpp:395): rid = npos, rd = proxy_to_true(inf)
                if ((rc = launch_fill_outputs(rid, rd, n2o, linear ? dinf : dnan, s2, err, &launches))) return rc;
            }
        }
        XM_CUDA_TRY(cudaEventRecord(ev[7], s2));
        if (ho.id) {
            if (mirror && n2o && ho.rid) {
                XM_CUDA_TRY(cudaStreamWaitEvent(ctx->copy, ev[7], 0));
                XM_CUDA_TRY(cudaMemcpyAsync(ho.rid, rid, n2o*8, cudaMemcpyDeviceToHost, ctx->copy));
                XM_CUDA_TRY(cudaMemcpyAsync(ho.rd, rd, n2o*8, cudaMemcpyDeviceToHost, ctx->copy));
            }
            ho.done = true;
        }
        XM_CUDA_TRY(cudaEventRecord(sev[3], s2));
        XM_CUDA_TRY(cudaStreamWaitEvent(s, sev[3], 0));                // join
        XM_CUDA_TRY(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, sizeof(SearchCounters), cudaMemcpyDeviceToHost, s));
        XM_CUDA_TRY(cudaEventRecord(ev[8], s));
        XM_CUDA_TRY(cudaStreamSynchronize(s));
        XM_CUDA_TRY(cudaStreamSynchronize(s2));
        if (ho.done) XM_CUDA_TRY(cudaStreamSynchronize(ctx->copy));
        st.ms_bbox = ev_ms(ev[0], ev[1]);
        st.ms_keys = ev_ms(ev[1], ev[2]);
        st.ms_sort = ev_ms(ev[2], ev[3]);
        st.ms_index = ev_ms(ev[1], ev[5]);
        st.ms_forward = ev_ms(ev[5], ev[6]);
        st.ms_reverse = ev_ms(ev[9], ev[7]);           // runs concurrently with the forward search
        st.ms_search = ev_ms(ev[5], ev[8]);
        if (timed_f) st.ms_filter_fwd = ev_ms(ctx->kev[0], ctx->kev[1]);
        if (timed_r) st.ms_filter_rev = ev_ms(ctx->kev[2], ctx->kev[3]);
    }
finish:
    if (sphere_done) {
        st.ms_forward = ev_ms(ev[5], ev[6]);
        st.ms_reverse = ev_ms(ev[9], ev[7]);           // concurrent with the forward search
        st.ms_search = ev_ms(ev[5], ev[8]);
    } else if (P.brute_force) {
        st.ms_forward = ev_ms(ev[5], ev[6]);
        st.ms_reverse = ev_ms(ev[6], ev[7]);
        st.ms_search = ev_ms(ev[5], ev[7]);
    }
    st.ms_total = ev_ms(ev[0], ev[8]);
    st.launches = launches;
    st.evals_fwd = ctx->h_ctr->evals_fwd; st.evals_rev = ctx->h_ctr->evals_rev;
    if (P.brute_force && !sphere_done) { // every pair is evaluated once per direction
        st.evals_fwd = n1*n2 - (self ? (n1 < n2 ? n1 : n2) : 0);
        st.evals_rev = mirror ? st.evals_fwd : 0;
    }
    st.rows_ambiguous = ctx->h_ctr->ambiguous;
    st.rows_refined_fwd = ctx->h_ctr->refine_fwd; st.rows_refined_rev = ctx->h_ctr->refine_rev;
    if (promised) {
        // did the promise hold? (counted by the index build, copied with the other counters)
        if (ctx->h_ctr->off_grid1)
            return fail(err, XM_ERR_INVALID_ARG, "catalog 1: %llu rows are not finite or lie outside the promised bounding box (bbox1)",
                        (unsigned long long)ctx->h_ctr->off_grid1);
        if (ctx->h_ctr->off_grid2)
            return fail(err, XM_ERR_INVALID_ARG, "catalog 2: %llu rows are not finite or lie outside the promised bounding box (bbox2)",
                        (unsigned long long)ctx->h_ctr->off_grid2);
    }
    if (P.verbose) {
        printf("[vif_b200] qxmatch n1=%llu n2=%llu nth=%llu %s: %.3f ms on device (index %.3f, forward %.3f, "
               "reverse %.3f), %llu kernel launches\n", (unsigned long long)n1, (unsigned long long)n2,
               (unsigned long long)nth, P.brute_force ? "brute" : "binned", st.ms_total, st.ms_index,
               st.ms_forward, st.ms_reverse, (unsigned long long)launches);
        fflush(stdout);
    }
    return XM_OK;
}

// The bucket index orders each cell with a bounded in-cell sort; a catalog with a pathologically
// crowded cell (> 4096 sources) reports an overflow and the call is redone with the radix-sort
// index, which has no such bound. XM_INDEX=radix forces the latter (tests).
int32_t run_device(Context* ctx, cudaStream_t s, const double* ra1, const double* dec1, uint64_t n1,
                   const double* ra2, const double* dec2, uint64_t n2, const xm_params& P,
                   uint64_t* id, double* d, uint64_t* rid, double* rd, Error& err) {
    const char* env = getenv("XM_INDEX");
    const bool want_bucket = !(env && strcmp(env, "radix") == 0);
    ctx->h_ctr->index_overflow = 0;
    int32_t rc = run_device_once(ctx, s, ra1, dec1, n1, ra2, dec2, n2, P, id, d, rid, rd, want_bucket, err);
    if (rc == XM_OK && want_bucket && ctx->h_ctr->index_overflow != 0) {
        const double ms_first = ctx->last.ms_total;
        rc = run_device_once(ctx, s, ra1, dec1, n1, ra2, dec2, n2, P, id, d, rid, rd, false, err);
        ctx->last.ms_total += ms_first;
    }
    return rc;
}

void copy_err(const Error& err, char* errbuf, uint64_t errcap) {
    if (errbuf && errcap) {
        strncpy(errbuf, err.msg, errcap - 1);
        errbuf[errcap - 1] = 0;
    }
}

int32_t check_args(const double* ra1, const double* dec1, uint64_t n1, const double* ra2,
                   const double* dec2, uint64_t n2, const xm_params* p, uint64_t* id, double* d,
                   uint64_t* rid, double* rd, Error& err) {
    if ((n1 && (!ra1 || !dec1)) || (n2 && (!ra2 || !dec2)))
        return fail(err, XM_ERR_INVALID_ARG, "NULL coordinate pointer");
    const uint64_t n1o = p->has_ranges ? p->fwd_count : n1, n2o = p->has_ranges ? p->rev_count : n2;   // rows the outputs cover
    if (n1o && (!id || !d)) return fail(err, XM_ERR_INVALID_ARG, "NULL id/d output pointer");
    if (!p->no_mirror && n2o && (!rid || !rd))
        return fail(err, XM_ERR_INVALID_ARG, "NULL rid/rd output pointer while no_mirror is false");
    return XM_OK;
}

// shared shape of the small distance entry points (xm_dist.cu)
template <class F>
int32_t dist_entry(int32_t device, void* stream, char* errbuf, uint64_t errcap, bool bad_args, F&& f) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = bad_args ? fail(err, XM_ERR_INVALID_ARG, "bad argument") : XM_OK;
    if (!rc) rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            CtxLock lk(ctx);
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            int32_t r = f(ctx, s, err);
            if (r) return r;
            XM_CUDA_TRY(cudaStreamSynchronize(s));
            return XM_OK;
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

}  // namespace
}  // namespace xm

using namespace xm;

extern "C" {

const char* xm_version(void) { return "vif_b200_qxmatch 0.1 (sm_100a)"; }

int32_t xm_init(int32_t device, char* errbuf, uint64_t errcap) {
    Error err;
    Context* c = nullptr;
    int32_t rc = get_context(device, &c, err);
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

void xm_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_ctx) {
        Context* c = kv.second.get();
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        for (auto& e : c->ev) cudaEventDestroy(e);
        for (auto& e : c->kev) cudaEventDestroy(e);
        for (auto& e : c->sync_ev) cudaEventDestroy(e);
        cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2);
        cudaStreamSynchronize(c->copy); cudaStreamDestroy(c->copy);
        for (auto& q : c->side) { cudaStreamSynchronize(q); cudaStreamDestroy(q); }
        for (auto& e : c->lane_ev) cudaEventDestroy(e);
        cudaFree(c->d_small); cudaFreeHost(c->h_small);
        cudaFree(c->d_ctr); cudaFreeHost(c->h_ctr);
        cudaStreamDestroy(c->stream);
        if (c->pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(c->pool); }
    }
    g_ctx.clear();
    stage_shutdown();
}

int32_t xm_release_memory(int32_t device) {
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return XM_ERR_CUDA; }
    if (device < 0) device = prev;
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_ctx.find(device);
        if (it != g_ctx.end()) pool = it->second->pool;
    }
    if (!pool) return XM_OK;                  // nothing allocated on that device yet
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess &&
              cudaMemPoolTrimTo(pool, 0) == cudaSuccess;
    if (!ok) cudaGetLastError();
    cudaSetDevice(prev);
    return ok ? XM_OK : XM_ERR_CUDA;
}

int32_t xm_band_partition_dev(const double* ra, const double* dec, uint64_t n, uint64_t row0, const double* edges,
                              int32_t nb, double halo, int32_t only, double* out, uint64_t cap, uint64_t* counts,
                              uint64_t* total, int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    const bool bad = (n && (!ra || !dec)) || !edges || !counts || !total || (cap && !out) || n >= 0xfffffff0ull ||
                     cap >= 0xfffffff0ull;
    return dist_entry(device, stream, errbuf, errcap, bad, [&](Context*, cudaStream_t s, Error& err) -> int32_t {
        Scratch scr;
        XM_CUDA_TRY(scr.alloc(band_scratch_bytes(n, cap), s));
        uint64_t launches = 0;
        return launch_band_partition(ra, dec, (uint32_t)n, row0, edges, nb, halo, only, out, cap,
                                     (unsigned long long*)counts, scr.p, total, s, err, &launches);
    });
}

int32_t xm_band_finish_dev(int64_t* id_io, const double* d, const int64_t* bits_src, const int64_t* bits_cand, uint64_t n,
                           double hlim, uint64_t* bad, int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    const bool badarg = (n && (!id_io || !d || !bits_src || !bits_cand)) || !bad || n >= 0xfffffff0ull;
    return dist_entry(device, stream, errbuf, errcap, badarg, [&](Context*, cudaStream_t s, Error& err) {
        return launch_band_finish((long long*)id_io, d, (const long long*)bits_src, (const long long*)bits_cand, (uint32_t)n,
                                  hlim, (unsigned long long*)bad, s, err);
    });
}

int32_t xm_band_home_dev(const double* got, const double* sent, uint64_t m, uint64_t row0, int64_t* id_out, double* d_out,
                         int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    const bool badarg = (m && (!got || !sent || !id_out || !d_out)) || m >= 0xfffffff0ull;
    return dist_entry(device, stream, errbuf, errcap, badarg, [&](Context*, cudaStream_t s, Error& err) {
        return launch_band_home(got, (const long long*)sent, (uint32_t)m, row0, (long long*)id_out, d_out, s, err);
    });
}

int32_t xm_libm_model(void) { return calibrate_sin_model(); }

double xm_libm_eval_host(int32_t which, int32_t model, double x) { return libm_host_eval(which, model, x); }

int32_t xm_libm_eval_dev(int32_t which, int32_t model, const double* x, uint64_t n, double* out, int32_t device,
                         void* stream, char* errbuf, uint64_t errcap) {
    const bool bad = (n && (!x || !out)) || model < 1 || model > 2 || which < 0 || which > 1;
    return dist_entry(device, stream, errbuf, errcap, bad, [&](Context*, cudaStream_t s, Error& err) {
        return launch_libm_eval(which, model, x, n, out, s, err);
    });
}

int32_t xm_fp64_peak(int32_t device, int32_t reps, double* tflops, char* errbuf, uint64_t errcap) {
    return dist_entry(device, nullptr, errbuf, errcap, !tflops,
                      [&](Context*, cudaStream_t s, Error& err) { return launch_fp64_peak(reps, tflops, s, err); });
}

int32_t xm_qxmatch_dev(const double* ra1, const double* dec1, uint64_t n1, const double* ra2,
                       const double* dec2, uint64_t n2, const xm_params* params, uint64_t* id,
                       double* d, uint64_t* rid, double* rd, void* stream, char* errbuf,
                       uint64_t errcap) {
    Error err;
    xm_params pn;
    int32_t rc = normalize_params(params, &pn, err);
    params = &pn;
    if (!rc) rc = check_args(ra1, dec1, n1, ra2, dec2, n2, params, id, d, rid, rd, err);
    Context* ctx = nullptr;
    if (!rc) rc = get_context(params->device, &ctx, err);
    if (!rc) {
        CtxLock lk(ctx);
        DeviceGuard dg;
        cudaError_t e = dg.set(ctx->device);
        if (e != cudaSuccess) rc = fail_cuda(err, e, "cudaSetDevice", __FILE__, __LINE__);
        else {
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            const auto t0 = std::chrono::steady_clock::now();
            rc = run_device(ctx, s, ra1, dec1, n1, ra2, dec2, n2, *params, id, d, rid, rd, err);
            ctx->last.ms_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (rc != XM_OK) { cudaDeviceSynchronize(); cudaGetLastError(); }   // nothing of this call stays in flight
        }
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

int32_t xm_qxmatch_f64(const double* ra1, const double* dec1, uint64_t n1, const double* ra2,
                       const double* dec2, uint64_t n2, const xm_params* params, uint64_t* id,
                       double* d, uint64_t* rid, double* rd, char* errbuf, uint64_t errcap) {
    Error err;
    xm_params pn;
    int32_t rc = normalize_params(params, &pn, err);
    params = &pn;
    if (!rc) rc = check_args(ra1, dec1, n1, ra2, dec2, n2, params, id, d, rid, rd, err);
    if (!rc && params->has_ranges) rc = fail(err, XM_ERR_INVALID_ARG, "row ranges need the device entry point (xm_qxmatch_dev)");
    Context* ctx = nullptr;
    if (!rc) rc = get_context(params->device, &ctx, err);
    if (rc) { copy_err(err, errbuf, errcap); return rc; }

    auto body = [&]() -> int32_t {
        CtxLock lk(ctx);
        DeviceGuard dg;
        XM_CUDA_TRY(dg.set(ctx->device));
        cudaStream_t s = ctx->stream;
        const auto t0 = std::chrono::steady_clock::now();
        const uint64_t nth = params->nth < 1 ? 1 : params->nth;
        const bool mirror = !params->no_mirror;
        const bool same_cat = (ra1 == ra2 && dec1 == dec2 && n1 == n2);
        Scratch dra1, ddec1, dra2, ddec2, did, dd, drid, drd;
        XM_CUDA_TRY(dra1.alloc(n1*8, s)); XM_CUDA_TRY(ddec1.alloc(n1*8, s));
        const double *pra2 = dra1.as<double>(), *pdec2 = ddec1.as<double>();
        if (!same_cat) {
            XM_CUDA_TRY(dra2.alloc(n2*8, s)); XM_CUDA_TRY(ddec2.alloc(n2*8, s));
            pra2 = dra2.as<double>(); pdec2 = ddec2.as<double>();
        }
        // inputs: pinned buffers go straight to the DMA engine, pageable ones (std::vector storage
        // behind the drop-in header) through the chunked staging pipeline of xm_staging.cu
        {
            StageJob in[4] = {{dra1.p, (void*)ra1, (size_t)n1*8}, {ddec1.p, (void*)dec1, (size_t)n1*8},
                              {dra2.p, (void*)ra2, (size_t)n2*8}, {ddec2.p, (void*)dec2, (size_t)n2*8}};
            StageJob staged[4];
            int ns = 0;
            for (int k = 0; k < (same_cat ? 2 : 4); ++k) {
                if (in[k].bytes == 0) continue;
                if (host_pointer_is_pageable(in[k].host)) staged[ns++] = in[k];
                else XM_CUDA_TRY(cudaMemcpyAsync(in[k].dev, in[k].host, in[k].bytes, cudaMemcpyHostToDevice, s));
            }
            if (ns) { int32_t r = stage_h2d(ctx->device, staged, ns, s, err); if (r) return r; }
        }
        XM_CUDA_TRY(did.alloc(nth*n1*8, s)); XM_CUDA_TRY(dd.alloc(nth*n1*8, s));
        if (mirror) { XM_CUDA_TRY(drid.alloc(n2*8, s)); XM_CUDA_TRY(drd.alloc(n2*8, s)); }
        StageJob outj[4] = {{did.p, id, (size_t)(nth*n1*8)}, {dd.p, d, (size_t)(nth*n1*8)},
                            {drid.p, rid, mirror ? (size_t)n2*8 : 0}, {drd.p, rd, mirror ? (size_t)n2*8 : 0}};
        bool out_pinned = true;
        for (int k = 0; k < 4; ++k) if (outj[k].bytes && host_pointer_is_pageable(outj[k].host)) out_pinned = false;
        if (out_pinned) {       // results leave lane by lane from inside the cell-binned pipeline
            ctx->host_out.id = id; ctx->host_out.d = d; ctx->host_out.rid = mirror ? rid : nullptr;
            ctx->host_out.rd = mirror ? rd : nullptr;
        }
        ctx->host_out.done = false;
        int32_t r = run_device(ctx, s, dra1.as<double>(), ddec1.as<double>(), n1, pra2, pdec2, n2, *params,
                               did.as<uint64_t>(), dd.as<double>(), drid.as<uint64_t>(), drd.as<double>(), err);
        const bool copied = ctx->host_out.done;
        ctx->host_out = Context::HostOut();
        if (r) { cudaStreamSynchronize(ctx->copy); return r; }
        if (!copied) {
            StageJob staged[4];
            int ns = 0;
            for (int k = 0; k < 4; ++k) {
                if (outj[k].bytes == 0) continue;
                if (host_pointer_is_pageable(outj[k].host)) staged[ns++] = outj[k];
                else XM_CUDA_TRY(cudaMemcpyAsync(outj[k].host, outj[k].dev, outj[k].bytes, cudaMemcpyDeviceToHost, s));
            }
            if (ns) { int32_t r2 = stage_d2h(ctx->device, staged, ns, s, err); if (r2) return r2; }
        }
        XM_CUDA_TRY(cudaStreamSynchronize(s));
        ctx->last.ms_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return XM_OK;
    };
    rc = body();
    if (rc == XM_ERR_CUDA) { cudaDeviceSynchronize(); cudaGetLastError(); }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

int32_t xm_grid_from_bbox(const double* bbox1, const double* bbox2, uint64_t n2, uint64_t nth,
                          int32_t linear, xm_grid* out) {
    if (!bbox1 || !bbox2 || !out) return XM_ERR_INVALID_ARG;
    grid_from_bbox(bbox1, bbox2, n2, nth < 1 ? 1 : nth, linear != 0, out);
    return XM_OK;
}

int32_t xm_bbox_dev(const double* ra, const double* dec, uint64_t n, int32_t device, void* stream,
                    double* out, uint64_t* nonfinite, char* errbuf, uint64_t errcap) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = (!ra || !dec || !out || n == 0) ? fail(err, XM_ERR_INVALID_ARG, "bad argument") : XM_OK;
    if (!rc) rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            CtxLock lk(ctx);
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            uint64_t launches = 0;
            int32_t r = launch_bbox(ra, dec, n, ctx->d_small, s, err, &launches);
            if (r) return r;
            XM_CUDA_TRY(cudaMemcpyAsync(ctx->h_small, ctx->d_small, 5*sizeof(unsigned long long),
                                        cudaMemcpyDeviceToHost, s));
            XM_CUDA_TRY(cudaStreamSynchronize(s));
            for (int k = 0; k < 4; ++k) out[k] = ord2f(ctx->h_small[k]);
            if (nonfinite) *nonfinite = ctx->h_small[4];
            return XM_OK;
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

int32_t xm_bbox_neg_dev(const double* ra, const double* dec, uint64_t n, int32_t device, void* stream,
                        double* out_dev, char* errbuf, uint64_t errcap) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = (!out_dev || (n && (!ra || !dec))) ? fail(err, XM_ERR_INVALID_ARG, "bad argument") : XM_OK;
    if (!rc) rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            uint64_t launches = 0;
            return launch_bbox_neg(ra, dec, n, out_dev, s, err, &launches);
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

int32_t xm_last_stats(int32_t device, xm_stats* out) {
    if (!out) return XM_ERR_INVALID_ARG;
    Error err;
    Context* ctx = nullptr;
    int32_t rc = get_context(device, &ctx, err);
    if (rc) return rc;
    CtxLock lk(ctx);
    const uint64_t want = out->struct_size;
    if (want < sizeof(uint64_t) || want > (1u << 20)) return XM_ERR_INVALID_ARG;   // caller did not set struct_size
    ctx->last.struct_size = want < sizeof(xm_stats) ? want : sizeof(xm_stats);
    memcpy(out, &ctx->last, (size_t)ctx->last.struct_size);
    return XM_OK;
}

int32_t xm_sort_pairs_dev(uint32_t* keys, uint32_t* vals, uint64_t n, int32_t key_bits,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = (!keys || !vals || n >= 0xfffffffeull) ? fail(err, XM_ERR_INVALID_ARG, "bad argument") : XM_OK;
    if (!rc) rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            CtxLock lk(ctx);
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            Scratch tk, tv, sc;
            XM_CUDA_TRY(tk.alloc(n*4, s)); XM_CUDA_TRY(tv.alloc(n*4, s));
            XM_CUDA_TRY(sc.alloc(sort_scratch_bytes(n), s));
            uint64_t launches = 0;
            int32_t r = launch_sort_pairs(keys, vals, tk.as<uint32_t>(), tv.as<uint32_t>(), (uint32_t)n,
                                          key_bits, sc.p, s, err, &launches);
            if (r) return r;
            XM_CUDA_TRY(cudaStreamSynchronize(s));
            return XM_OK;
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

static int32_t merge_entry(int which, uint64_t* rid, double* rd, const double* ka, const double* kb, double* key,
                           uint64_t n, uint64_t off, double unmatched, int32_t device, void* stream,
                           char* errbuf, uint64_t errcap) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            int32_t r = launch_merge_step(which, (unsigned long long*)rid, rd, ka, kb, key, n, off, unmatched, s, err);
            if (r) return r;
            XM_CUDA_TRY(cudaStreamSynchronize(s));       // tiny kernels: keep the host-side ordering simple
            return XM_OK;
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

int32_t xm_rev_merge_pack(uint64_t* rid, const double* rd, uint64_t n, uint64_t row_offset, double* key,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    return merge_entry(0, rid, const_cast<double*>(rd), nullptr, nullptr, key, n, row_offset, 0.0, device, stream, errbuf, errcap);
}

int32_t xm_rev_merge_mask(uint64_t* rid, const double* key_local, const double* key_best, uint64_t n,
                          int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    return merge_entry(1, rid, nullptr, key_local, key_best, nullptr, n, 0, 0.0, device, stream, errbuf, errcap);
}

int32_t xm_rev_merge_unpack(uint64_t* rid, double* rd, const double* key_best, uint64_t n, double unmatched,
                            int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    return merge_entry(2, rid, rd, nullptr, key_best, nullptr, n, 0, unmatched, device, stream, errbuf, errcap);
}

int32_t xm_rev_merge_p2p(const void* const* exch, void* const* out, uint64_t n, int32_t rank, int32_t world,
                         double unmatched, int32_t device, void* stream, char* errbuf, uint64_t errcap) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = (!exch || !out || world < 1 || world > XM_MAX_PEERS || rank < 0 || rank >= world)
                     ? fail(err, XM_ERR_INVALID_ARG, "bad argument") : XM_OK;
    for (int r = 0; !rc && r < world; ++r)
        if (n && (!exch[r] || (r == rank && !out[r]))) rc = fail(err, XM_ERR_INVALID_ARG, "null peer pointer");
    if (!rc) rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            return launch_merge_p2p(exch, out, n, rank, world, unmatched, s, err);
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

int32_t xm_angdist_dev(const double* ra1, const double* dec1, const double* ra2, const double* dec2, uint64_t n,
                       int32_t second_is_scalar, double* out, int32_t device, void* stream, char* errbuf,
                       uint64_t errcap) {
    return dist_entry(device, stream, errbuf, errcap, n && (!ra1 || !dec1 || !ra2 || !dec2 || !out),
                      [&](Context* ctx, cudaStream_t s, Error& err) {
                          return launch_angdist(ra1, dec1, ra2, dec2, n, second_is_scalar != 0, ctx->sin_model, out, s, err);
                      });
}

int32_t xm_angdist_less_dev(const double* ra1, const double* dec1, uint64_t n, double ra2, double dec2,
                            double radius, uint8_t* out, int32_t device, void* stream, char* errbuf,
                            uint64_t errcap) {
    return dist_entry(device, stream, errbuf, errcap, n && (!ra1 || !dec1 || !out),
                      [&](Context* ctx, cudaStream_t s, Error& err) {
                          return launch_angdist_less(ra1, dec1, n, ra2, dec2, radius, ctx->sin_model, out, s, err);
                      });
}

int32_t xm_qdist_dev(const double* ra, const double* dec, uint64_t n, double* out, int32_t device, void* stream,
                     char* errbuf, uint64_t errcap) {
    return dist_entry(device, stream, errbuf, errcap, n && (!ra || !dec || !out),
                      [&](Context* ctx, cudaStream_t s, Error& err) -> int32_t {
                          Scratch scr;
                          XM_CUDA_TRY(scr.alloc(qdist_scratch_bytes(n), s));
                          return launch_qdist(ra, dec, n, ctx->sin_model, scr.p, out, s, err);
                      });
}

int32_t xm_clean_best_dev(const uint64_t* id, const uint64_t* rid, uint64_t n1, uint64_t n2, uint64_t* id1,
                          uint64_t* id2, uint64_t* lost, uint64_t* counts, int32_t device, void* stream,
                          char* errbuf, uint64_t errcap) {
    Error err;
    Context* ctx = nullptr;
    int32_t rc = (!counts || n1 >= 0xfffffffeull || (n1 && (!id || !id1 || !id2 || !lost)) || (n2 && !rid))
                     ? fail(err, XM_ERR_INVALID_ARG, "bad argument") : XM_OK;
    if (!rc) rc = get_context(device, &ctx, err);
    if (!rc) {
        auto body = [&]() -> int32_t {
            CtxLock lk(ctx);
            DeviceGuard dg;
            XM_CUDA_TRY(dg.set(ctx->device));
            cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
            Scratch scr, dcnt;
            XM_CUDA_TRY(scr.alloc(clean_best_scratch_bytes((uint32_t)n1), s));
            XM_CUDA_TRY(dcnt.alloc(2*sizeof(uint64_t), s));
            uint64_t launches = 0;
            int32_t r = launch_clean_best(id, rid, (uint32_t)n1, n2, id1, id2, lost, dcnt.as<uint64_t>(), scr.p, s,
                                          err, &launches);
            if (r) return r;
            XM_CUDA_TRY(cudaMemcpyAsync(counts, dcnt.p, 2*sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
            XM_CUDA_TRY(cudaStreamSynchronize(s));
            return XM_OK;
        };
        rc = body();
    }
    if (rc) copy_err(err, errbuf, errcap);
    return rc;
}

uint64_t xm_ring_entries(uint32_t k, uint32_t nx, uint32_t ny, int32_t dedupe, int32_t* bx,
                         int32_t* by, uint64_t cap) {
    uint64_t n = 0;
    if (nx == 0 || ny == 0) return 0;
    for_each_ring_entry(k, nx, ny, dedupe != 0, [&](int x, int y) {
        if (n < cap && bx && by) { bx[n] = x; by[n] = y; }
        ++n;
    });
    return n;
}

}  // extern "C"

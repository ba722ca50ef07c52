// Host <-> device transfers for callers whose buffers are ordinary (pageable) memory -- the normal
// case behind the drop-in header: vif's vec<1,double> is std::vector storage. A cudaMemcpyAsync from
// pageable memory is staged by the driver through one thread at memcpy speed; here the copy is cut
// into 4 MB chunks that a few host threads move through a ring of pinned bounce buffers while the
// DMA engine drains / fills the ring, so the transfer runs near the PCIe rate. Pinned (or registered)
// buffers bypass all this and go straight to cudaMemcpyAsync.
#include <atomic>
#include <thread>
#include <vector>
#include <cstring>
#include <cstdlib>

#include "xm_common.cuh"

namespace xm {

namespace {

constexpr size_t kChunk = 4u << 20;
constexpr int kSlots = 12;

struct Ring {
    char* buf[kSlots] = {nullptr};
    cudaEvent_t ev[kSlots];
    std::atomic<long long> issued[kSlots];      // 1 + index of the last chunk whose DMA was enqueued on the slot
    std::atomic<long long> drained[kSlots];     // 1 + index of the last chunk copied out of the slot (D2H)
    bool ready = false;
};

Ring g_ring[16];

int32_t ring_for(int device, Ring** out, Error& err) {
    if (device < 0 || device >= 16) return fail(err, XM_ERR_INVALID_ARG, "device index out of range");
    Ring& r = g_ring[device];
    if (!r.ready) {
        for (int k = 0; k < kSlots; ++k) {
            XM_CUDA_TRY(cudaHostAlloc((void**)&r.buf[k], kChunk, cudaHostAllocDefault));
            XM_CUDA_TRY(cudaEventCreateWithFlags(&r.ev[k], cudaEventDisableTiming));
        }
        r.ready = true;
    }
    *out = &r;
    return XM_OK;
}

int stage_threads() {
    const char* e = getenv("XM_STAGE_THREADS");
    int t = e ? atoi(e) : 0;
    if (t <= 0) {
        const unsigned hc = std::thread::hardware_concurrency();
        t = hc >= 16 ? 6 : (hc >= 8 ? 4 : 2);
    }
    return t > 16 ? 16 : t;
}

struct Chunk { char* dev; char* host; size_t bytes; };

std::vector<Chunk> cut(const StageJob* jobs, int njobs) {
    std::vector<Chunk> c;
    for (int j = 0; j < njobs; ++j)
        for (size_t off = 0; off < jobs[j].bytes; off += kChunk)
            c.push_back({(char*)jobs[j].dev + off, (char*)jobs[j].host + off,
                         jobs[j].bytes - off < kChunk ? jobs[j].bytes - off : kChunk});
    return c;
}

inline void pause_briefly() { std::this_thread::yield(); }

}  // namespace

bool host_pointer_is_pageable(const void* p) {
    static const bool off = [] { const char* e = getenv("XM_STAGE"); return e && e[0] == '0'; }();
    if (off) return false;          // XM_STAGE=0: leave pageable buffers to the driver's own staging (A/B runs)
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Enqueues host -> device copies of pageable buffers on `s`. Returns when every chunk has been
// handed to the DMA engine (the source buffers may be reused; the copies complete in stream order).
int32_t stage_h2d(int device, const StageJob* jobs, int njobs, cudaStream_t s, Error& err) {
    Ring* rp = nullptr;
    int32_t rc = ring_for(device, &rp, err);
    if (rc) return rc;
    Ring& r = *rp;
    const std::vector<Chunk> chunks = cut(jobs, njobs);
    const long long n = (long long)chunks.size();
    if (n == 0) return XM_OK;
    for (int k = 0; k < kSlots; ++k) r.issued[k].store(0, std::memory_order_relaxed);
    std::atomic<long long> next(0);
    std::atomic<int> cuda_err(0);
    auto worker = [&]() {
        cudaSetDevice(device);
        for (;;) {
            const long long i = next.fetch_add(1);
            if (i >= n) return;
            const int slot = (int)(i % kSlots);
            // the slot's previous chunk (i - kSlots) must have been enqueued, then its DMA must be over
            if (i >= kSlots) while (r.issued[slot].load(std::memory_order_acquire) != i - kSlots + 1) pause_briefly();
            cudaError_t e = cudaEventSynchronize(r.ev[slot]);
            const Chunk& c = chunks[(size_t)i];
            memcpy(r.buf[slot], c.host, c.bytes);
            if (e == cudaSuccess) e = cudaMemcpyAsync(c.dev, r.buf[slot], c.bytes, cudaMemcpyHostToDevice, s);
            if (e == cudaSuccess) e = cudaEventRecord(r.ev[slot], s);
            if (e != cudaSuccess) cuda_err.store((int)e);
            r.issued[slot].store(i + 1, std::memory_order_release);
        }
    };
    const int nt = (int)(n < stage_threads() ? n : stage_threads());
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    if (cuda_err.load()) return fail_cuda(err, (cudaError_t)cuda_err.load(), "staged host-to-device copy", __FILE__, __LINE__);
    return XM_OK;
}

// Device -> host copies into pageable buffers, enqueued on `s` after whatever `s` already holds.
// Returns when the data is in the destination buffers.
int32_t stage_d2h(int device, const StageJob* jobs, int njobs, cudaStream_t s, Error& err) {
    Ring* rp = nullptr;
    int32_t rc = ring_for(device, &rp, err);
    if (rc) return rc;
    Ring& r = *rp;
    const std::vector<Chunk> chunks = cut(jobs, njobs);
    const long long n = (long long)chunks.size();
    if (n == 0) return XM_OK;
    for (int k = 0; k < kSlots; ++k) {
        // bounce buffers may still feed an earlier host-to-device copy
        cudaError_t e = cudaEventSynchronize(r.ev[k]);
        if (e != cudaSuccess) return fail_cuda(err, e, "cudaEventSynchronize", __FILE__, __LINE__);
        r.issued[k].store(0, std::memory_order_relaxed);
        r.drained[k].store(0, std::memory_order_relaxed);
    }
    std::atomic<long long> next(0);
    std::atomic<int> cuda_err(0);
    auto consumer = [&]() {
        cudaSetDevice(device);
        for (;;) {
            const long long i = next.fetch_add(1);
            if (i >= n) return;
            const int slot = (int)(i % kSlots);
            while (r.issued[slot].load(std::memory_order_acquire) != i + 1) {
                if (cuda_err.load()) return;
                pause_briefly();
            }
            cudaError_t e = cudaEventSynchronize(r.ev[slot]);
            if (e != cudaSuccess) cuda_err.store((int)e);
            const Chunk& c = chunks[(size_t)i];
            memcpy(c.host, r.buf[slot], c.bytes);
            r.drained[slot].store(i + 1, std::memory_order_release);
        }
    };
    const int nt = (int)(n < stage_threads() ? n : stage_threads());
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(consumer);
    for (long long i = 0; i < n && !cuda_err.load(); ++i) {
        const int slot = (int)(i % kSlots);
        if (i >= kSlots)
            while (r.drained[slot].load(std::memory_order_acquire) != i - kSlots + 1 && !cuda_err.load()) pause_briefly();
        const Chunk& c = chunks[(size_t)i];
        cudaError_t e = cudaMemcpyAsync(r.buf[slot], c.dev, c.bytes, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaEventRecord(r.ev[slot], s);
        if (e != cudaSuccess) cuda_err.store((int)e);
        r.issued[slot].store(i + 1, std::memory_order_release);
    }
    for (auto& t : pool) t.join();
    if (cuda_err.load()) return fail_cuda(err, (cudaError_t)cuda_err.load(), "staged device-to-host copy", __FILE__, __LINE__);
    return XM_OK;
}

void stage_shutdown() {
    for (Ring& r : g_ring) {
        if (!r.ready) continue;
        for (int k = 0; k < kSlots; ++k) { cudaFreeHost(r.buf[k]); cudaEventDestroy(r.ev[k]); }
        r.ready = false;
    }
}

}  // namespace xm

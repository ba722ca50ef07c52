// DFMA micro-benchmark: measures the FP64 FMA throughput the brute-force roofline is quoted against
// (the FP64 peak "must be measured with a DFMA micro-benchmark").
// Every thread carries 8 independent accumulator chains (enough ILP to cover the FP64 pipe latency
// with 8 CTAs of 256 threads per SM) through kIters fused multiply-adds each.
#include "xm_common.cuh"

namespace xm {

namespace {

constexpr int kChains = 8;
constexpr int kIters = 4096;

__global__ void __launch_bounds__(256)
dfma_peak_kernel(double a, double b, double* __restrict__ sink) {
    double acc[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) acc[k] = (double)(threadIdx.x + k)*1e-3;
#pragma unroll 1
    for (int it = 0; it < kIters; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int k = 0; k < kChains; ++k) acc[k] = fma(acc[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += acc[k];
    if (s == 12345.678) sink[0] = s;        // never true: keeps the chains alive
}

}  // namespace

int32_t launch_fp64_peak(int reps, double* tflops, cudaStream_t s, Error& err) {
    int dev = 0, sms = 0;
    XM_CUDA_TRY(cudaGetDevice(&dev));
    XM_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    XM_CUDA_TRY(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    XM_CUDA_TRY(cudaEventCreate(&e0));
    XM_CUDA_TRY(cudaEventCreate(&e1));
    const int blocks = sms*8*4;                 // 4 waves of 8 resident CTAs per SM
    const double flops = 2.0*(double)blocks*256.0*(double)kChains*(double)kIters;
    double best = 0.0;
    for (int r = 0; r < (reps < 1 ? 1 : reps) + 1; ++r) {       // launch 0 warms up
        cudaEventRecord(e0, s);
        dfma_peak_kernel<<<blocks, 256, 0, s>>>(0.999999, 1e-9, sink);
        cudaEventRecord(e1, s);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { cudaFree(sink); return fail_cuda(err, e, "dfma_peak_kernel", __FILE__, __LINE__); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms > 0.f) { const double t = flops/(ms*1e-3)/1e12; if (t > best) best = t; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return XM_OK;
}

}  // namespace xm

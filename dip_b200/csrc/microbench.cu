// microbench.cu -- the shared-memory roofline denominator, measured on the device.
//
// BASELINE.md section 2 asks the builder to replace the derived figure
// (148 SM x 128 B/clk x f_SM) by a measurement.  Every SM runs two 1024-thread
// CTAs that stream 64-bit shared-memory loads and stores (the access width of
// the knapsack DP row, knapsack.cu); bytes moved / CUDA-event time is reported.
#include "common.cuh"

namespace dipgpu {

constexpr int kMbThreads = 1024;
constexpr int kMbUnroll = 8;

__global__ void __launch_bounds__(kMbThreads) smem_copy_kernel(int iters, double *sink)
{
    extern __shared__ __align__(16) double mb[];
    const int t = threadIdx.x;
    const int half = kMbThreads * kMbUnroll;
    for (int j = t; j < 2 * half; j += kMbThreads) mb[j] = (double)j;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(mb);
    const uint32_t ld = base + (uint32_t)t * 8u;
    const uint32_t st = base + (uint32_t)(half + ((t + 32) & (kMbThreads - 1))) * 8u;
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        double v[kMbUnroll];
#pragma unroll
        for (int k = 0; k < kMbUnroll; ++k)
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[k]) : "r"(ld + (uint32_t)(k * kMbThreads * 8)));
#pragma unroll
        for (int k = 0; k < kMbUnroll; ++k)
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(st + (uint32_t)(k * kMbThreads * 8)), "d"(v[k]));
        acc += v[i & (kMbUnroll - 1)];
    }
    if (acc == -1.0) sink[0] = acc;  // keeps the loads alive, never true
}

int measureSmemGbs(Ctx *c, double *gbs)
{
    const int iters = 4096;
    const size_t smem = sizeof(double) * 2 * kMbThreads * kMbUnroll;  // 128 KiB
    const int grid = (c->smCount > 0 ? c->smCount : 148);
    double *sink = nullptr;
    DIPGPU_CUDA(c, cudaMalloc(&sink, 64));
    {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(smem_copy_kernel), smem);
        if (rc) { cudaFree(sink); return rc; }
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        smem_copy_kernel<<<grid, kMbThreads, smem, c->stream>>>(iters, sink);
        cudaEventRecord(e1, c->stream);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            cudaFree(sink);
            return cudaFail(c, e, "smem_copy_kernel");
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
        c->launches++;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double bytes = (double)grid * kMbThreads * (double)iters * kMbUnroll * 16.0;
    *gbs = bytes / ((double)best * 1e-3) / 1e9;
    return DIPGPU_OK;
}

}  // namespace dipgpu

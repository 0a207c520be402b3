// Launch-to-first-instruction and end-to-host latencies of a kernel shaped like the fused GAP-E round (80 CTAs x 192
// threads, ~100 KB dynamic shared memory), by launch kind (plain / cooperative), parameter size (64 B / 4 KB) and L2
// state (warm / flushed): what an end-to-end round pays around the kernel itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/launch_lat scripts/micro/launch_lat.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

struct Small { volatile unsigned *startFlag, *endFlag; unsigned val; int spinNs; };
struct Big { volatile unsigned *startFlag, *endFlag; unsigned val; int spinNs; int tab[1000]; };

template <class P>
__global__ void __launch_bounds__(192) k(const __grid_constant__ P p)
{
    extern __shared__ unsigned char sm[];
    if (blockIdx.x == 0 && threadIdx.x == 0) { *p.startFlag = p.val; }
    // a body of p.spinNs on every CTA
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while ((long long)(t1 - t0) < p.spinNs);
    sm[threadIdx.x] = (unsigned char)t1;
    __syncthreads();
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { __threadfence_system(); *p.endFlag = p.val + sm[1] * 0; }
}

static double us(std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b)
{
    return std::chrono::duration<double, std::micro>(b - a).count();
}

static char *gThrash = nullptr;         // host-cache thrash between launches ("cold" launch call, as after an LP solve)
static size_t gThrashBytes = 0;
static volatile long gSink = 0;

template <class P>
static void run(const char *name, bool coop, bool flush, int spinNs, cudaStream_t st, unsigned *hFlags, unsigned *dFlags, void *flushBuf, size_t flushBytes)
{
    P p;
    memset(&p, 0, sizeof(p));
    p.startFlag = dFlags;
    p.endFlag = dFlags + 16;
    p.spinNs = spinNs;
    const size_t smem = 100 * 1024;
    cudaFuncSetAttribute(k<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    double a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    const int reps = 200;
    volatile unsigned *hs = hFlags, *he = hFlags + 16;
    for (int it = 0; it < reps + 10; ++it) {
        if (flush) cudaMemsetAsync(flushBuf, it & 0xff, flushBytes, st);
        cudaStreamSynchronize(st);
        if (gThrash) {
            long acc = 0;
            for (size_t k = 0; k < gThrashBytes; k += 64) acc += gThrash[k]++;
            gSink += acc;
        }
        p.val = (unsigned)(it + 1) + 1000u * (unsigned)spinNs;
        const auto t0 = std::chrono::steady_clock::now();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(80, 1, 1);
        cfg.blockDim = dim3(192, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = coop ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k<P>, p);
        const auto t1 = std::chrono::steady_clock::now();
        while (*hs != p.val) {}
        const auto t2 = std::chrono::steady_clock::now();
        while (*he != p.val) {}
        const auto t3 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(st);
        const auto t4 = std::chrono::steady_clock::now();
        if (it >= 10) { a1 += us(t0, t1); a2 += us(t0, t2); a3 += us(t0, t3); a4 += us(t0, t4); }
    }
    printf("%-34s launch returned %5.2f | first CTA seen %5.2f | end flag seen %5.2f | stream synchronised %5.2f  (us after the call, body %d us)\n",
           name, a1 / reps, a2 / reps, a3 / reps, a4 / reps, spinNs / 1000);
}

int main()
{
    cudaSetDevice(0);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    unsigned *hFlags, *dFlags;
    cudaHostAlloc((void **)&hFlags, 256, cudaHostAllocMapped);
    memset(hFlags, 0, 256);
    cudaHostGetDevicePointer((void **)&dFlags, hFlags, 0);
    void *flushBuf;
    const size_t flushBytes = 256u << 20;
    cudaMalloc(&flushBuf, flushBytes);
    for (int cold = 0; cold < 2; ++cold) {
    if (cold) {
        gThrashBytes = 96u << 20;
        gThrash = (char *)malloc(gThrashBytes);
        memset(gThrash, 1, gThrashBytes);
        printf("---- host caches thrashed (96 MiB touched) before every launch call\n");
    }
    for (int spin : {0, 30000}) {
        for (int flush = 0; flush < 2; ++flush) {
            run<Small>(flush ? "plain, 64 B params, L2 flushed" : "plain, 64 B params, warm", false, flush, spin, st, hFlags, dFlags, flushBuf, flushBytes);
            run<Big>(flush ? "plain, 4 KB params, L2 flushed" : "plain, 4 KB params, warm", false, flush, spin, st, hFlags, dFlags, flushBuf, flushBytes);
            run<Small>(flush ? "cooperative, 64 B params, flushed" : "cooperative, 64 B params, warm", true, flush, spin, st, hFlags, dFlags, flushBuf, flushBytes);
            run<Big>(flush ? "cooperative, 4 KB params, flushed" : "cooperative, 4 KB params, warm", true, flush, spin, st, hFlags, dFlags, flushBuf, flushBytes);
        }
    }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

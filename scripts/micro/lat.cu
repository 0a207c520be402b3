// Dependent-issue latencies on the target GPU (single warp), to ground the DP latency model.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void k_dadd(double *out, double x, long long *cyc) {
    double a = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + x;
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd_sel(double *out, double x, long long *cyc) {
    double a = out[threadIdx.x], m = out[threadIdx.x + 32];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) { double c = x + a; bool t = c > m; m = t ? c : m; a = m; }
    long long t1 = clock64();
    out[threadIdx.x] = m; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fadd(float *out, float x, long long *cyc) {
    float a = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + x;
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(int *out, long long *cyc) {
    __shared__ int s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 33) & 1023;
    __syncthreads();
    int a = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = s[a];
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sts_bar_lds(double *out, long long *cyc) {
    __shared__ double s[2][1024];
    s[0][threadIdx.x] = threadIdx.x; s[1][threadIdx.x] = 0;
    __syncthreads();
    double a = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        s[i & 1][threadIdx.x] = a;
        __syncthreads();
        a = s[i & 1][(threadIdx.x + 7) % blockDim.x];
    }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_step(double *out, double x, long long *cyc) {  // the full DP step chain
    __shared__ double s[2][1024];
    s[0][threadIdx.x] = 0; s[1][threadIdx.x] = 0;
    __syncthreads();
    double m = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        double sh = s[i & 1][(threadIdx.x + 7) % blockDim.x];
        double c = x + sh; bool t = c > m; m = t ? c : m;
        s[(i & 1) ^ 1][threadIdx.x] = m;
        __syncthreads();
    }
    long long t1 = clock64();
    out[threadIdx.x] = m; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double *out, long long *cyc) {
    double a = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = __shfl_up_sync(0xffffffffu, a, 1);
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *d; long long *c, h; float *f; int *ii;
    cudaMalloc(&d, 8192 * 8); cudaMemset(d, 0, 8192 * 8); cudaMalloc(&c, 8); cudaMalloc(&f, 8192 * 4); cudaMalloc(&ii, 8192 * 4);
    cudaMemset(f, 0, 8192 * 4);
#define RUN(name, call) for (int r = 0; r < 2; ++r) { call; cudaDeviceSynchronize(); } cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-28s %7.1f cycles/iter\n", name, (double)h / N);
    RUN("DADD chain (1 warp)", (k_dadd<<<1, 32>>>(d, 1.0, c)));
    RUN("DADD+DSETP+FSEL chain", (k_dadd_sel<<<1, 32>>>(d, 1.0, c)));
    RUN("FADD chain", (k_fadd<<<1, 32>>>(f, 1.0f, c)));
    RUN("LDS chain", (k_lds<<<1, 32>>>(ii, c)));
    RUN("SHFL.64 chain", (k_shfl<<<1, 32>>>(d, c)));
    RUN("STS+BAR+LDS 1 warp", (k_sts_bar_lds<<<1, 32>>>(d, c)));
    RUN("STS+BAR+LDS 6 warps", (k_sts_bar_lds<<<1, 192>>>(d, c)));
    RUN("STS+BAR+LDS 32 warps", (k_sts_bar_lds<<<1, 1024>>>(d, c)));
    RUN("DP step 1 warp", (k_step<<<1, 32>>>(d, 1.0, c)));
    RUN("DP step 6 warps", (k_step<<<1, 192>>>(d, 1.0, c)));
    RUN("DP step 14 warps", (k_step<<<1, 448>>>(d, 1.0, c)));
    // contention: 80 CTAs like GAP-E
    RUN("DP step 6 warps x80 CTAs", (k_step<<<80, 192>>>(d, 1.0, c)));
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}

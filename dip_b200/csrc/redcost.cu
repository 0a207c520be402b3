// redcost.cu -- stage (a) of the pricing round: redCostX = c - A''^T u.
//
// Replaces DecompAlgo::generateVarsAdjustDuals + generateVarsCalcRedCost
// (Dip/src/DecompAlgo.cpp:4550-4619, 4506-4547).  The reference scatters rows of
// the row-ordered A'' (CoinPackedMatrix::transposeTimes); here A'' is kept
// transposed on the device (CSC = CSR of A''^T) so each output column is a
// gather-and-reduce with no atomics, and the convexity-row gap of the master
// dual vector is folded into the stored row ids (rowMaster), so the "adjusted"
// dual vector of :4717-4722 is never materialised.
//
// HBM-bound: algorithmic bytes = 12*nnz + 20*N + 8*R.
#include "common.cuh"

namespace dipgpu {

__device__ __forceinline__ double ldgStream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldgStream(const int32_t *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// L lanes cooperate on one column; columns are consecutive across the warp so a
// warp streams one contiguous span of (val, rowMaster).
template <int L>
__global__ void __launch_bounds__(256)
redcost_csc_kernel(int N, const int64_t *__restrict__ colStart,
                   const int32_t *__restrict__ rowMaster, const double *__restrict__ val,
                   const double *__restrict__ u, const double *__restrict__ c, int phase1,
                   double *__restrict__ out)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t col = gtid / L;
    const int sub = (int)(gtid % L);
    double sum = 0.0;
    if (col < N) {
        const int64_t beg = colStart[col], end = colStart[col + 1];
        for (int64_t k = beg + sub; k < end; k += L) {
            const double a = ldgStream(val + k);
            const int r = ldgStream(rowMaster + k);
            sum += __ldg(u + r) * a;
        }
    }
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (col < N && sub == 0) {
        // DecompAlgo.cpp:4538-4545
        out[col] = phase1 ? -sum : c[col] - sum;
    }
}

int launchRedCost(Ctx *c, const double *dU, int phase, cudaStream_t st)
{
    if (c->N == 0) return DIPGPU_OK;
    int L = c->spmvLanes;
    if (L == 0) {
        const double avg = c->N > 0 ? (double)c->nnz / (double)c->N : 0.0;
        L = avg <= 2.5 ? 1 : avg <= 6 ? 4 : avg <= 24 ? 8 : avg <= 96 ? 16 : 32;
    }
    const int phase1 = phase == DIPGPU_PHASE_PRICE1 ? 1 : 0;
    const int threads = 256;
    const int64_t total = (int64_t)c->N * L;
    const unsigned grid = (unsigned)((total + threads - 1) / threads);
#define LAUNCH(LL)                                                                              \
    redcost_csc_kernel<LL><<<grid, threads, 0, st>>>(c->N, c->dColStart, c->dRowMaster, c->dVal, \
                                                      dU, c->dObj, phase1, c->dRedCost)
    switch (L) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        default: LAUNCH(32); break;
    }
#undef LAUNCH
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

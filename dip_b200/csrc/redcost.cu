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
//
// Main kernel (redcost_stream_kernel): the columns are cut on the host into
// tiles of at most kSpmvChunk stored entries; a CTA pulls its tile's (value,
// row) stream into shared memory with two TMA bulk copies, multiplies by the
// gathered duals in place, and then one thread per column adds that column's
// products SEQUENTIALLY in storage order.  Entries are stored per column in
// DESCENDING row order and the arithmetic is an unfused multiply followed by an
// add, i.e. exactly the order and rounding of the reference's row-major scatter
// (rows R-1..0, y[col] += u[i]*a), so the reduced costs are bit-identical to the
// CPU restatement, not merely within 1e-12.  Matrices with a column longer than
// one tile fall back to the lane-cooperative kernel below (tolerance parity).
#include "common.cuh"

namespace dipgpu {

__device__ __forceinline__ uint32_t rcSmemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kSpmvThreads)
redcost_stream_kernel(const int32_t *__restrict__ tileCol, const int32_t *__restrict__ colStart32,
                      const int32_t *__restrict__ rowMaster, const double *__restrict__ val,
                      const double *__restrict__ u, const double *__restrict__ c, int phase1,
                      double *__restrict__ out)
{
    __shared__ __align__(16) double sVal[kSpmvChunk + 4];
    __shared__ __align__(16) int32_t sRow[kSpmvChunk + 4];
    __shared__ __align__(8) uint64_t mbar;
    const int t = threadIdx.x;
    const int c0 = tileCol[blockIdx.x], c1 = tileCol[blockIdx.x + 1];
    const int k0 = colStart32[c0], k1 = colStart32[c1];
    const int a0 = k0 & ~3;  // 16-byte aligned start for both streams
    const int len = k1 - a0;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rcSmemAddr(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (len > 0) {
            const uint32_t bytesVal = (uint32_t)(((len + 1) & ~1) * 8);
            const uint32_t bytesRow = (uint32_t)(((len + 3) & ~3) * 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rcSmemAddr(&mbar)),
                         "r"(bytesVal + bytesRow)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    rcSmemAddr(sVal)),
                "l"(val + a0), "r"(bytesVal), "r"(rcSmemAddr(&mbar))
                : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    rcSmemAddr(sRow)),
                "l"(rowMaster + a0), "r"(bytesRow), "r"(rcSmemAddr(&mbar))
                : "memory");
        }
    }
    __syncthreads();
    if (len > 0) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "WAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra DONE_%=;\n\t"
            "bra WAIT_%=;\n\t"
            "DONE_%=:\n\t"
            "}" ::"r"(rcSmemAddr(&mbar)),
            "r"(0)
            : "memory");
        // products in place: u[i] * a  (unfused, as the reference's  y += u*a)
        for (int k = (k0 - a0) + t; k < len; k += kSpmvThreads) sVal[k] = __dmul_rn(__ldg(u + sRow[k]), sVal[k]);
    }
    __syncthreads();
    for (int col = c0 + t; col < c1; col += kSpmvThreads) {
        const int b = colStart32[col] - a0, e = colStart32[col + 1] - a0;
        double sum = 0.0;
        for (int k = b; k < e; ++k) sum = __dadd_rn(sum, sVal[k]);
        out[col] = phase1 ? -sum : c[col] - sum;  // DecompAlgo.cpp:4538-4545
    }
}

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
    if (c->nSpmvTiles > 0 && c->spmvLanes == 0) {
        redcost_stream_kernel<<<c->nSpmvTiles, kSpmvThreads, 0, st>>>(
            c->dTileCol, c->dColStart32, c->dRowMaster, c->dVal, dU, c->dObj,
            phase == DIPGPU_PHASE_PRICE1 ? 1 : 0, c->dRedCost);
        c->launches++;
        DIPGPU_CUDA(c, cudaGetLastError());
        return DIPGPU_OK;
    }
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

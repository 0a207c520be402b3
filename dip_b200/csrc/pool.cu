// pool.cu -- the two dual-vector passes that surround a pricing round in DecompAlgoPC:
//
//  * re-pricing of the waiting columns, DecompVarPool::setReducedCosts ->
//    DecompWaitingCol::setReducedCost (Dip/src/DecompVarPool.cpp:185-203, :22-40; called once per
//    price call, Dip/src/DecompAlgo.cpp:1890-1908):  redCost_s = origCost_s - col_s . u  (feasible
//    master) or -col_s . u (dual ray), col_s = the waiting column in MASTER row space;
//  * Wentges smoothing of the master duals, DecompAlgoPC::adjustMasterDualSolution
//    (Dip/src/DecompAlgoPC.cpp:126-212):  dualST = alpha*dual + (1-alpha)*dualRM, for one alpha or
//    for the whole ladder alpha, 0.9 alpha, ... that addVarsToPool walks when every new column is
//    a duplicate (Dip/src/DecompAlgo.cpp:5642-5655).
//
// Both are HBM streams: the pool pass reads 12 B per stored entry + 24 B per column, the
// smoothing reads 16 B and writes 8 B per row and vector.
#include <algorithm>

#include "common.cuh"

namespace dipgpu {

constexpr int kPoolWarps = 8;
constexpr int kPoolWarpsBitmap = 16;  // CTA width of the variant that keeps the dual-support bitmap in shared memory
constexpr int kPoolChunk = 512;       // products staged per warp and pass (4 KiB)

// One warp per 32 consecutive waiting columns.  Their stored entries are one contiguous span of
// the pool; the warp walks it from the END in chunks: all lanes form the products val*u[row]
// (coalesced loads, unfused multiply) into shared memory, then lane l adds the products of ITS
// column, last entry first -- the order of CoinPackedVectorBase::dotProduct as the oracle restates
// it (dp = 0; for i = n-1..0: dp += els[i]*u[ind[i]]), so the values are bit-equal to the oracle.
//
// BITMAP: a dual support is declared (dipgpu_set_dual_support) and its bitmap (one bit per master row)
// fits shared memory.  Most entries of a master column sit in rows that cannot carry a dual (GAP: the
// free branch rows, 2/3 of every column): the gather of the dual -- one 32-byte sector per 8 useful
// bytes, what bounds this pass -- is skipped for them, a shared-memory bit test decides.  The product
// is still formed (val * 0), so the result is the same bit for bit.
template <bool BITMAP>
__global__ void __launch_bounds__(kPoolWarpsBitmap * 32)
pool_redcost_kernel(long long nCols, const int64_t *__restrict__ colStart, const int32_t *__restrict__ row,
                    const double *__restrict__ val, const double *__restrict__ origCost,
                    const double *__restrict__ u, int feasible, double *__restrict__ out, int32_t *flag,
                    const uint32_t *__restrict__ supportBits, int bitWords)
{
    extern __shared__ __align__(16) unsigned char poolSm[];
    const int warpsPerCta = blockDim.x >> 5;
    double *sPall = reinterpret_cast<double *>(poolSm);
    const uint32_t *sBits = reinterpret_cast<const uint32_t *>(sPall + (size_t)warpsPerCta * kPoolChunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (BITMAP) {
        uint32_t *w = const_cast<uint32_t *>(sBits);
        for (int k = threadIdx.x; k < bitWords; k += blockDim.x) w[k] = __ldg(supportBits + k);
        __syncthreads();
    }
    const long long c0 = ((long long)blockIdx.x * warpsPerCta + warp) * 32;
    if (c0 >= nCols) return;
    const long long col = c0 + lane;
    const long long cEnd = min(nCols, c0 + 32);
    const int64_t B = colStart[c0], E = colStart[cEnd];
    int64_t b = 0, e = 0;
    double oc = 0.0;
    if (col < nCols) {
        b = colStart[col];
        e = colStart[col + 1];
        oc = origCost[col];
    }
    double sum = 0.0;
    double *p = sPall + (size_t)warp * kPoolChunk;
    for (int64_t hi = E; hi > B;) {
        const int64_t lo = max(B, hi - kPoolChunk);
        // batches of 8 independent (row, value) loads, then their 8 dependent gathers of the dual: the
        // round trips overlap instead of serialising (a warp issues in order)
        for (int64_t base = lo + lane; base < hi; base += 256) {
            int r[8];
            double v[8], g[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t k = base + 32 * j;
                r[j] = k < hi ? __ldg(row + k) : 0;
                v[j] = k < hi ? __ldg(val + k) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bool on = base + 32 * j < hi;
                if (BITMAP) on = on && ((sBits[r[j] >> 5] >> (r[j] & 31)) & 1u);
                g[j] = on ? __ldg(u + r[j]) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (base + 32 * j < hi) p[base + 32 * j - lo] = __dmul_rn(v[j], g[j]);
        }
        __syncwarp();
        const int64_t kHi = min(hi, e), kLo = max(lo, b);
        if (kHi > kLo) {
            // the next product is fetched while the current one is added (one load ahead of the DADD chain)
            int q = (int)(kHi - 1 - lo);
            const int qLo = (int)(kLo - lo);
            double cur = p[q];
            while (q > qLo) {
                const double nxt = p[q - 1];
                sum = __dadd_rn(sum, cur);
                cur = nxt;
                --q;
            }
            sum = __dadd_rn(sum, cur);
        }
        __syncwarp();
        hi = lo;
    }
    if (col < nCols) {
        const double rc = feasible ? oc - sum : -sum;  // DecompVarPool.cpp:29 / :36
        out[col] = rc;
        if (rc <= -0.0000000001) *flag = 1;            // :30, :37 -> found_negrc_var (:199)
    }
}

int launchPoolRedCost(Ctx *c, const double *dU, int feasible, cudaStream_t st)
{
    DIPGPU_CUDA(c, cudaMemsetAsync(c->dPoolFlag, 0, sizeof(int32_t), st));
    if (c->poolCols == 0) return DIPGPU_OK;
    const long long warps = (c->poolCols + 31) / 32;
    // the bitmap variant pays when the pool is large (the bitmap is loaded once per CTA) and the support small
    const int bitWords = (c->R + c->numConvex + 31) / 32;
    const bool bitmap = c->dSupportBits != nullptr && !c->hSupport.empty() && bitWords * 4 <= 64 * 1024 &&
                        c->poolCols >= 32 * 1024 && (int64_t)c->hSupport.size() * 4 <= (int64_t)c->R + c->numConvex;
    const int wpc = bitmap ? kPoolWarpsBitmap : kPoolWarps;
    const size_t smem = sizeof(double) * (size_t)wpc * kPoolChunk + (bitmap ? sizeof(uint32_t) * (size_t)bitWords : 0);
    const unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
    auto fn = bitmap ? pool_redcost_kernel<true> : pool_redcost_kernel<false>;
    if (c->poolFnSmem[bitmap ? 1 : 0] < smem) {
        DIPGPU_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->poolFnSmem[bitmap ? 1 : 0] = smem;
    }
    fn<<<grid, wpc * 32, smem, st>>>(c->poolCols, c->dPoolStart, c->dPoolRow, c->dPoolVal, c->dPoolOrigCost, dU, feasible,
                                    c->dPoolRedCost, c->dPoolFlag, c->dSupportBits, bitWords);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// dualST_v[r] = (alpha_v * dual[r]) + ((1 - alpha_v) * dualRM[r])   (DecompAlgoPC.cpp:153-154, :179-181)
// for nVec smoothing parameters at once; two unfused multiplies and one add, as the reference.
__global__ void __launch_bounds__(256)
smooth_duals_kernel(const double *__restrict__ uRM, const double *__restrict__ center,
                    const double *__restrict__ alphas, int nVec, int uLen, double *__restrict__ out, long long outStride)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= uLen) return;
    const double d = center[r], rm = uRM[r];
    for (int v = 0; v < nVec; ++v) {
        const double alpha = alphas[v];
        const double alpha1 = __dsub_rn(1.0, alpha);
        out[(long long)v * outStride + r] = __dadd_rn(__dmul_rn(alpha, d), __dmul_rn(alpha1, rm));
    }
}

int launchSmoothDuals(Ctx *c, const double *dURM, const double *dCenter, const double *dAlphas, int nVec, int uLen,
                      double *dOut, int64_t outStride, cudaStream_t st)
{
    if (uLen <= 0 || nVec <= 0) return DIPGPU_OK;
    smooth_duals_kernel<<<(uLen + 255) / 256, 256, 0, st>>>(dURM, dCenter, dAlphas, nVec, uLen, dOut, outStride);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// dst_v[idx[k]] = vals_v[k] (GATHER: vals_v[idx[k]] -- vals is the caller's whole dual vector in mapped host
// memory): the support of a sparse dual upload (dipgpu_set_dual_support)
template <bool GATHER>
__global__ void __launch_bounds__(256)
scatter_duals_kernel(int n, const int32_t *__restrict__ idx, const double *__restrict__ vals, long long valStride,
                     double *__restrict__ dst, long long dstStride)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int i = idx[k];
        dst[(long long)blockIdx.y * dstStride + i] = vals[(long long)blockIdx.y * valStride + (GATHER ? i : k)];
    }
}

int launchScatterDuals(Ctx *c, int n, const int32_t *dIdx, const double *dVals, int64_t valStride, bool gather, double *dDst,
                       int64_t dstStride, int nVec, cudaStream_t st)
{
    if (n <= 0 || nVec <= 0) return DIPGPU_OK;
    const dim3 grid((n + 255) / 256, nVec, 1);
    if (gather) scatter_duals_kernel<true><<<grid, 256, 0, st>>>(n, dIdx, dVals, valStride, dDst, dstStride);
    else scatter_duals_kernel<false><<<grid, 256, 0, st>>>(n, dIdx, dVals, valStride, dDst, dstStride);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// ------------------------------------------------------------------ pool upkeep
//
// Columns enter the pool either from host arrays (DecompVarPool::push_back of a waiting column
// the host built) or straight from the packed master columns of the last expansion, and leave it
// when addVarsFromPool moves them into the master (Dip/src/DecompAlgo.cpp:5678-5900).

struct PoolCopyDesc {
    int64_t src;  // first entry in the source
    int64_t dst;  // first entry in the pool
    int32_t len;
    int32_t pad;
};

// srcEntries != NULL: packed master columns (MColEntry); else (srcRow, srcVal).
__global__ void __launch_bounds__(256)
pool_copy_kernel(int n, const PoolCopyDesc *__restrict__ d, const MColEntry *__restrict__ srcEntries,
                 const int32_t *__restrict__ srcRow, const double *__restrict__ srcVal, int32_t *__restrict__ dstRow,
                 double *__restrict__ dstVal)
{
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= n) return;
    const PoolCopyDesc dk = d[k];
    for (int q = lane; q < dk.len; q += 32) {
        if (srcEntries != nullptr) {
            const MColEntry en = srcEntries[dk.src + q];
            dstRow[dk.dst + q] = en.row;
            dstVal[dk.dst + q] = en.val;
        } else {
            dstRow[dk.dst + q] = srcRow[dk.src + q];
            dstVal[dk.dst + q] = srcVal[dk.src + q];
        }
    }
}

int launchPoolCopy(Ctx *c, int n, const void *dDesc, const void *srcEntries, const int32_t *srcRow, const double *srcVal,
                   int32_t *dstRow, double *dstVal, cudaStream_t st)
{
    if (n <= 0) return DIPGPU_OK;
    pool_copy_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, static_cast<const PoolCopyDesc *>(dDesc),
                                                  static_cast<const MColEntry *>(srcEntries), srcRow, srcVal, dstRow, dstVal);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

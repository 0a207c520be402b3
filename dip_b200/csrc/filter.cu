// filter.cu -- stage (c) of the pricing round: keep the candidate columns with
// negative reduced cost and pack them for the host.
//
// Replaces the loop of DecompAlgo::generateVars over potentialVars
// (Dip/src/DecompAlgo.cpp:5151-5232):
//     mostNegRCvec[b] = min(mostNegRCvec[b](=0, :4761), rc_b)        (:5212-5214)
//     keep the column iff rc_b < -RedCostEpsilon                     (:5216-5226)
// The per-block min(0, rc_b) is produced for EVERY block, kept or not, because it
// feeds the lower bound (updateObjBound, :3757).  Kept columns are compacted in
// block order with a warp-shuffle scan (ballot + popc within a warp, shuffle
// scan of the per-warp totals), so the output order equals the reference's list
// order and is deterministic.
#include "common.cuh"

namespace dipgpu {

struct FilterArgs {
    ResultHeader *hdr;
    const double *blockRedCost;
    double *blockMostNeg;
    const int32_t *blockNnz;
    int64_t *keptStart;
    int32_t *keptBlock;
    int32_t *keptInd;
    const int32_t *candInd;
    const int32_t *blockColStart;
    const int32_t *capacity;
    const int32_t *nItems;
    int firstBlock, nLocal;
    int64_t indCap;
    double eps;
    int fuseCopy;
};

constexpr int kFilterThreads = 1024;

__device__ __forceinline__ void copyColumn(const FilterArgs &a, int k, int lane)
{
    const int lb = a.keptBlock[k];
    const int cs = a.blockColStart[a.firstBlock + lb];
    const int64_t dst = a.keptStart[k];
    const int n = a.blockNnz[lb];
    for (int q = lane; q < n; q += 32) a.keptInd[dst + q] = a.candInd[cs + q];
}

__global__ void __launch_bounds__(kFilterThreads) filter_scan_kernel(const FilterArgs a)
{
    __shared__ int sCnt[32];
    __shared__ long long sNnz[32];
    __shared__ long long sCells[32];
    __shared__ int sCarryCnt;
    __shared__ long long sCarryNnz;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        sCarryCnt = 0;
        sCarryNnz = 0;
    }
    long long cells = 0;
    __syncthreads();
    for (int base = 0; base < a.nLocal; base += kFilterThreads) {
        const int lb = base + t;
        bool keep = false;
        int nnz = 0;
        if (lb < a.nLocal) {
            const double rc = a.blockRedCost[lb];
            a.blockMostNeg[lb] = rc < 0.0 ? rc : 0.0;  // min(0, rc)
            keep = rc < -a.eps;                        // strict, :5216
            nnz = a.blockNnz[lb];
            cells += (long long)a.nItems[lb] * ((long long)a.capacity[a.firstBlock + lb] + 1);
        }
        // warp-shuffle compaction
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        long long v = keep ? nnz : 0;
        long long incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += up;
        }
        if (lane == 31) {
            sCnt[warp] = __popc(bal);
            sNnz[warp] = incl;
        }
        __syncthreads();
        int cntOff = sCarryCnt;
        long long nnzOff = sCarryNnz;
        int totCnt = 0;
        long long totNnz = 0;
        for (int wv = 0; wv < kFilterThreads / 32; ++wv) {
            const int cw = sCnt[wv];
            const long long nw = sNnz[wv];
            if (wv < warp) {
                cntOff += cw;
                nnzOff += nw;
            }
            totCnt += cw;
            totNnz += nw;
        }
        if (keep) {
            const int pos = cntOff + wpre;
            a.keptBlock[pos] = lb;
            a.keptStart[pos] = nnzOff + (incl - v);
        }
        __syncthreads();
        if (t == 0) {
            sCarryCnt += totCnt;
            sCarryNnz += totNnz;
        }
        __syncthreads();
    }
    // total DP cells of the shard
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, off);
    if (lane == 0) sCells[warp] = cells;
    __syncthreads();
    if (t == 0) {
        long long tot = 0;
        for (int wv = 0; wv < kFilterThreads / 32; ++wv) tot += sCells[wv];
        a.keptStart[sCarryCnt] = sCarryNnz;
        a.hdr->magic = kResultMagic;
        a.hdr->m = a.nLocal;
        a.hdr->firstBlock = a.firstBlock;
        a.hdr->nKept = sCarryCnt;
        a.hdr->nInd = sCarryNnz;
        a.hdr->cells = tot;
        a.hdr->indCap = a.indCap;
    }
    if (a.fuseCopy) {
        __threadfence_block();
        __syncthreads();
        const int nKept = sCarryCnt;
        for (int k = warp; k < nKept; k += kFilterThreads / 32) copyColumn(a, k, lane);
    }
}

__global__ void __launch_bounds__(256) filter_copy_kernel(const FilterArgs a)
{
    const int lane = threadIdx.x & 31;
    const int nKept = a.hdr->nKept;
    const int warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < nKept; k += warpsTotal)
        copyColumn(a, k, lane);
}

int launchFilter(Ctx *c, double redCostEps, cudaStream_t st)
{
    char *res = static_cast<char *>(c->dResult);
    FilterArgs a;
    a.hdr = reinterpret_cast<ResultHeader *>(res);
    a.blockRedCost = reinterpret_cast<double *>(res + c->layout.offRedCost);
    a.blockMostNeg = reinterpret_cast<double *>(res + c->layout.offMostNeg);
    a.blockNnz = reinterpret_cast<int32_t *>(res + c->layout.offNnz);
    a.keptStart = reinterpret_cast<int64_t *>(res + c->layout.offKeptStart);
    a.keptBlock = reinterpret_cast<int32_t *>(res + c->layout.offKeptBlock);
    a.keptInd = reinterpret_cast<int32_t *>(res + c->layout.offInd);
    a.candInd = c->dCandInd;
    a.blockColStart = c->dBlockColStart;
    a.capacity = c->dCapacity;
    a.nItems = c->dNItems;
    a.firstBlock = c->firstBlock;
    a.nLocal = c->nLocal;
    a.indCap = c->indCap;
    a.eps = redCostEps;
    a.fuseCopy = c->nLocal <= 2048 ? 1 : 0;
    filter_scan_kernel<<<1, kFilterThreads, 0, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    if (!a.fuseCopy) {
        const int grid = c->smCount > 0 ? c->smCount * 4 : 592;
        filter_copy_kernel<<<grid, 256, 0, st>>>(a);
        c->launches++;
        DIPGPU_CUDA(c, cudaGetLastError());
    }
    return DIPGPU_OK;
}

}  // namespace dipgpu

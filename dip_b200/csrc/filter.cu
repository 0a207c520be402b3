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
#include "filter.cuh"

namespace dipgpu {

constexpr int kFilterThreads = 1024;

__global__ void __launch_bounds__(kFilterThreads) filter_scan_kernel(const __grid_constant__ FilterArgs a)
{
    filterScanBody(a, kFilterThreads);
}

__global__ void __launch_bounds__(256) filter_copy_kernel(const __grid_constant__ FilterArgs a)
{
    const int lane = threadIdx.x & 31;
    const int nKept = a.hdr->nKept;
    const int warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < nKept; k += warpsTotal)
        copyColumn(a, k, lane);
}

FilterArgs makeFilterArgs(Ctx *c, double redCostEps)
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
    a.blockColStart = c->blockKind == 2 ? c->dCandStart : c->dBlockColStart;  // (where block b's candidates start)
    a.pairs = c->blockKind == 2 ? 1 : 0;
    a.capacity = c->haveFix ? c->dCapEff : c->dCapacity;
    a.nItems = c->dNItems;
    a.firstBlock = c->firstBlock;
    a.nLocal = c->nLocal;
    a.indCap = c->indCap;
    a.eps = redCostEps;
    a.fuseCopy = c->nLocal <= kFuseFilterMaxBlocks ? 2 : c->nLocal <= 2048 ? 1 : 0;
    return a;
}

int launchFilter(Ctx *c, double redCostEps, cudaStream_t st)
{
    FilterArgs a = makeFilterArgs(c, redCostEps);
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

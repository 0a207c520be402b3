// mckp.cu -- multiple-choice knapsack blocks: the MCKP0 relaxation of DIP's MMKP example
// (Dip/examples/MMKP/MMKP_DecompApp.cpp:493-557 -> MMKP_MCKnap::solveMCKnap,
// Dip/examples/MMKP/MMKP_MCKnap.cpp:172-383).  Block b = a run of groups, group g = a run of
// contiguous x-space columns; EXACTLY one column per group, sum of weights <= capacity, minimise
// sum redCost.  The reference's solver (Pisinger's mcknap behind a lossy integer scaling of the
// costs) is not vendored; the exact fp64 DP over the groups takes its place:
//     c[g][j] = min over the group's columns k with w_k <= j of redCost_k + c[g-1][j - w_k]
// first minimal column wins, backtrack from (last group, capacity).
//
// Shape: one CTA per block, groups sequential (one barrier per group: every column of a group reads
// the PREVIOUS group's row, so there is no dependency inside a group), cells strided over the
// threads, the two rows in shared memory, the argmin of every (group, cell) streamed to HBM as
// uint16 for the backtrack.  Shared-memory traffic per cell and column: one 8-byte read; the
// columns' reduced costs and weights are warp-uniform L1 reads.
#include <algorithm>

#include "filter.cuh"

namespace dipgpu {

constexpr int kMckpThreads = 256;

struct MckpArgs {
    const double *redCost;
    const double *obj;
    const double *alpha;            // indexed by GLOBAL block
    const int32_t *blockGroupStart; // global block -> first group
    const int32_t *groupColStart;   // group -> first x-space column
    const int32_t *blockColStart;   // global block -> first column
    const int32_t *weight;
    const int32_t *capacity;
    const int64_t *choiceOff;       // local block -> first entry of its (group, cell) argmin table
    uint16_t *choice;
    int32_t *candInd;
    int32_t *nItems;
    double *blockObj, *blockRedCost, *blockOrigCost;
    int32_t *blockNnz;
    int firstBlock, nLocal;
};

__global__ void __launch_bounds__(kMckpThreads) mckp_dp_kernel(const __grid_constant__ MckpArgs a)
{
    extern __shared__ __align__(16) unsigned char mckpSm[];
    const int lb = blockIdx.x, b = a.firstBlock + lb, t = threadIdx.x;
    const int cap = a.capacity[b];
    const int g0 = a.blockGroupStart[b], g1 = a.blockGroupStart[b + 1];
    const int W = cap + 1;
    double *rowA = reinterpret_cast<double *>(mckpSm), *rowB = rowA + W;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int j = t; j < W; j += kMckpThreads) rowA[j] = 0.0;  // c[-1][.] = 0
    __syncthreads();
    double *prev = rowA, *cur = rowB;
    uint16_t *choice = a.choice + a.choiceOff[lb];
    for (int g = g0; g < g1; ++g) {
        const int cs = a.groupColStart[g], ce = a.groupColStart[g + 1];
        uint16_t *ch = choice + (size_t)(g - g0) * W;
        for (int j = t; j < W; j += kMckpThreads) {
            double best = inf;
            int arg = 0xffff;
            for (int k = cs; k < ce; ++k) {
                const int w = __ldg(a.weight + k);
                if (w <= j) {
                    const double cand = __dadd_rn(__ldg(a.redCost + k), prev[j - w]);
                    if (cand < best) {  // first minimal column
                        best = cand;
                        arg = k - cs;
                    }
                }
            }
            cur[j] = best;
            ch[j] = (uint16_t)arg;
        }
        __syncthreads();
        double *x = prev;
        prev = cur;
        cur = x;
    }
    if (t == 0) {
        a.blockObj[lb] = g1 > g0 ? prev[cap] : 0.0;
        a.nItems[lb] = a.blockColStart[b + 1] - a.blockColStart[b];  // cells = columns x (cap + 1)
    }
}

// One warp per block: walk the argmin table from (last group, capacity), emit the column (one index
// per group, ascending), sums in ascending order, rc - alpha (DecompAlgo.cpp:6350-6356).
__global__ void __launch_bounds__(128) mckp_backtrack_kernel(const __grid_constant__ MckpArgs a)
{
    const int lane = threadIdx.x & 31;
    const int lb = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (lb >= a.nLocal) return;
    const int b = a.firstBlock + lb;
    const int g0 = a.blockGroupStart[b], g1 = a.blockGroupStart[b + 1], G = g1 - g0;
    const int cap = a.capacity[b], W = cap + 1;
    const int cs0 = a.blockColStart[b];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const bool feasible = G == 0 || __ldcg(a.blockObj + lb) < inf;
    if (!feasible) {
        if (lane == 0) {
            a.blockRedCost[lb] = inf;  // no column: the filter keeps nothing, mostNeg = 0
            a.blockOrigCost[lb] = 0.0;
            a.blockNnz[lb] = 0;
        }
        return;
    }
    if (lane == 0) {
        const uint16_t *choice = a.choice + a.choiceOff[lb];
        int j = cap;
        for (int g = G - 1; g >= 0; --g) {
            const int col = a.groupColStart[g0 + g] + (int)__ldcg(choice + (size_t)g * W + j);
            a.candInd[cs0 + g] = col;
            j -= a.weight[col];
        }
        double rcs = 0.0, ocs = 0.0;
        for (int g = 0; g < G; ++g) {
            const int col = a.candInd[cs0 + g];
            rcs += a.redCost[col];
            ocs += a.obj[col];
        }
        a.blockRedCost[lb] = rcs - a.alpha[b];
        a.blockOrigCost[lb] = ocs;
        a.blockNnz[lb] = G;
    }
}

int launchMckp(Ctx *c, const double *dRedCost, const double *dAlpha, cudaStream_t st)
{
    if (c->nLocal == 0) return DIPGPU_OK;
    char *res = static_cast<char *>(c->dResult);
    MckpArgs a;
    a.redCost = dRedCost;
    a.obj = c->dObj;
    a.alpha = dAlpha;
    a.blockGroupStart = c->dBlockGroupStart;
    a.groupColStart = c->dGroupColStart;
    a.blockColStart = c->dBlockColStart;
    a.weight = c->dWeight;
    a.capacity = c->dCapacity;
    a.choiceOff = c->dBitsOff;
    a.choice = reinterpret_cast<uint16_t *>(c->dBits);
    a.candInd = c->dCandInd;
    a.nItems = c->dNItems;
    a.blockObj = reinterpret_cast<double *>(res + c->layout.offObj);
    a.blockRedCost = reinterpret_cast<double *>(res + c->layout.offRedCost);
    a.blockOrigCost = reinterpret_cast<double *>(res + c->layout.offOrigCost);
    a.blockNnz = reinterpret_cast<int32_t *>(res + c->layout.offNnz);
    a.firstBlock = c->firstBlock;
    a.nLocal = c->nLocal;
    const size_t smem = sizeof(double) * 2 * ((size_t)c->maxCapacity + 1);
    if (smem > 227 * 1024) return fail(c, DIPGPU_ERR_UNSUPPORTED, "capacity %d needs %zu bytes of DP rows", c->maxCapacity, smem);
    if (c->mckpSmem < smem) {
        DIPGPU_CUDA(c, cudaFuncSetAttribute(mckp_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->mckpSmem = smem;
    }
    mckp_dp_kernel<<<c->nLocal, kMckpThreads, smem, st>>>(a);
    c->launches++;
    mckp_backtrack_kernel<<<(c->nLocal + 3) / 4, 128, 0, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

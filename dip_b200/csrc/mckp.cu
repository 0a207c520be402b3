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
// the PREVIOUS group's row, so there is no dependency inside a group), up to four cells per thread,
// the two rows in shared memory behind a guard of +inf cells, the argmin of every (group, cell)
// streamed to HBM as uint16 for the backtrack.  Shared-memory traffic per cell and column: one
// 8-byte read; the columns' reduced costs and weights are warp-uniform L1 reads, once per thread
// and four cells.
#include <algorithm>

#include "filter.cuh"

namespace dipgpu {

constexpr int kMckpThreads = 256;
constexpr int kMckpStage = 256;  // columns staged in shared memory at a time (whole groups)

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
    int guard;                      // +inf cells in front of each DP row: min(max weight, max capacity + 1)
};

// The groups of one block, C cells per thread: the column's weight and reduced cost are read once
// per thread and used for C cells (j = base + t + i*256), and "does the column fit" is not a
// branch -- each row is preceded by `guard` cells of +inf, so a column that does not fit reads
// +inf and never wins (guard >= every usable weight; heavier columns are clamped onto it).
template <int OFF>
__device__ __forceinline__ double ldsF64At(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
    return v;
}

// One group of one block, C cells per thread: the column's weight and reduced cost are read once per
// thread and used for C cells (j = base + t + i*256), and "does the column fit" is not a branch --
// each row is preceded by `guard` cells of +inf, so a column that does not fit reads +inf and never
// wins (guard >= every usable weight; heavier columns are clamped onto it).  STAGED: the group's
// columns sit in shared memory as (redCost, 8 * min(weight, guard)) pairs, one 16-byte broadcast
// read per column; otherwise (a group larger than the staging buffer) they are read through L1.
template <int C, bool STAGED>
__device__ __forceinline__ void mckpGroup(const MckpArgs &a, uint32_t prevA, uint32_t curA, uint16_t *ch, int cs, int ce, int W,
                                          int guard, uint32_t itemA)
{
    const int t = threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int base = 0; base < W; base += C * kMckpThreads) {
        // cells past the row read whatever follows it (the other row, or the pad behind both) and are not stored
        const uint32_t src = prevA + 8u * base;
        double best[C];
        int arg[C];
#pragma unroll
        for (int i = 0; i < C; ++i) {
            best[i] = inf;
            arg[i] = 0xffff;
        }
#pragma unroll 2
        for (int k = cs; k < ce; ++k) {
            double r;
            uint32_t at;
            if (STAGED) {
                unsigned long long lo, hi;
                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(itemA + 16u * (uint32_t)(k - cs)));
                r = __longlong_as_double((long long)lo);
                at = src - (uint32_t)hi;
            } else {
                at = src - 8u * (uint32_t)min(__ldg(a.weight + k), guard);
                r = __ldg(a.redCost + k);
            }
            double v[C];
            v[0] = ldsF64At<0>(at);
            if (C > 1) v[1] = ldsF64At<8 * kMckpThreads>(at);
            if (C > 2) v[2] = ldsF64At<16 * kMckpThreads>(at);
            if (C > 3) v[3] = ldsF64At<24 * kMckpThreads>(at);
#pragma unroll
            for (int i = 0; i < C; ++i) {
                const double cand = __dadd_rn(r, v[i]);
                if (cand < best[i]) {  // first minimal column
                    best[i] = cand;
                    arg[i] = k - cs;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const int j = base + t + i * kMckpThreads;
            if (j < W) {
                asm volatile("st.shared.f64 [%0], %1;" ::"r"(curA + 8u * (base + i * kMckpThreads)), "d"(best[i]) : "memory");
                ch[j] = (uint16_t)arg[i];
            }
        }
    }
    __syncthreads();
}

// The groups of one block: runs of whole groups are staged into `items` (kMckpStage columns) and
// solved from there; one barrier per group.  Returns the shared address of the final row.
template <int C>
__device__ __forceinline__ uint32_t mckpGroups(const MckpArgs &a, uint32_t rowA, uint32_t rowB, ulonglong2 *items, uint16_t *choice,
                                               int g0, int g1, int W, int guard)
{
    const int t = threadIdx.x;
    uint32_t prevA = rowA + 8u * t, curA = rowB + 8u * t;
    const uint32_t itemA = (uint32_t)__cvta_generic_to_shared(items);
    int g = g0;
    while (g < g1) {
        const int c0 = a.groupColStart[g];
        int gEnd = g, cEnd = c0;
        while (gEnd < g1) {
            const int e = a.groupColStart[gEnd + 1];
            if (e - c0 > kMckpStage) break;
            cEnd = e;
            ++gEnd;
        }
        if (gEnd == g) {  // one group larger than the staging buffer
            mckpGroup<C, false>(a, prevA, curA, choice + (size_t)(g - g0) * W, c0, a.groupColStart[g + 1], W, guard, 0u);
            const uint32_t x = prevA;
            prevA = curA;
            curA = x;
            ++g;
            continue;
        }
        // (the barrier that ended the previous group also released the staging buffer)
        for (int k = c0 + t; k < cEnd; k += kMckpThreads)
            items[k - c0] = make_ulonglong2((unsigned long long)__double_as_longlong(__ldg(a.redCost + k)),
                                            (unsigned long long)(8u * (uint32_t)min(__ldg(a.weight + k), guard)));
        __syncthreads();
        for (; g < gEnd; ++g) {
            const int cs = a.groupColStart[g], ce = a.groupColStart[g + 1];
            mckpGroup<C, true>(a, prevA, curA, choice + (size_t)(g - g0) * W, cs, ce, W, guard, itemA + 16u * (uint32_t)(cs - c0));
            const uint32_t x = prevA;
            prevA = curA;
            curA = x;
        }
    }
    return prevA - 8u * t;
}

__global__ void __launch_bounds__(kMckpThreads) mckp_dp_kernel(const __grid_constant__ MckpArgs a)
{
    extern __shared__ __align__(16) unsigned char mckpSm[];
    const int lb = blockIdx.x, b = a.firstBlock + lb, t = threadIdx.x;
    const int cap = a.capacity[b];
    const int g0 = a.blockGroupStart[b], g1 = a.blockGroupStart[b + 1];
    const int W = cap + 1, guard = a.guard;
    ulonglong2 *items = reinterpret_cast<ulonglong2 *>(mckpSm);
    double *rowA = reinterpret_cast<double *>(mckpSm + sizeof(ulonglong2) * kMckpStage) + guard, *rowB = rowA + W + guard;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int j = t; j < guard; j += kMckpThreads) {
        rowA[j - guard] = inf;
        rowB[j - guard] = inf;
    }
    for (int j = t; j < W; j += kMckpThreads) rowA[j] = 0.0;  // c[-1][.] = 0
    __syncthreads();
    uint16_t *choice = a.choice + a.choiceOff[lb];
    const uint32_t ra = (uint32_t)__cvta_generic_to_shared(rowA), rb = (uint32_t)__cvta_generic_to_shared(rowB);
    uint32_t last;
    if (W <= kMckpThreads) last = mckpGroups<1>(a, ra, rb, items, choice, g0, g1, W, guard);
    else if (W <= 2 * kMckpThreads) last = mckpGroups<2>(a, ra, rb, items, choice, g0, g1, W, guard);
    else last = mckpGroups<4>(a, ra, rb, items, choice, g0, g1, W, guard);
    if (t == 0) {
        a.blockObj[lb] = g1 > g0 ? (last == ra ? rowA[cap] : rowB[cap]) : 0.0;
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
    a.guard = std::min(c->mckpMaxWeight, c->maxCapacity + 1);
    // the staging buffer, two rows, each behind its guard, and a pad so that the last (partial) chunk of four cells per thread reads
    // allocated memory
    const size_t wMax = (size_t)c->maxCapacity + 1, chunk = 4 * kMckpThreads;
    const size_t smem = sizeof(ulonglong2) * kMckpStage + sizeof(double) * (2 * (wMax + a.guard) + ((wMax + chunk - 1) / chunk * chunk - wMax));
    if (smem > 227 * 1024) return fail(c, DIPGPU_ERR_UNSUPPORTED, "capacity %d needs %zu bytes of DP rows", c->maxCapacity, smem);
    if (c->mckpSmem < smem) {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(mckp_dp_kernel), smem);
        if (rc) return rc;
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

// bigdp.cu -- knapsack blocks whose DP row does not fit one CTA's shared memory / registers (more than
// 12 800 cells): the same recurrence, item order and tie rule as knapsack.cu -- knapsack01(),
// Dip/src/dippy/examples/gap/gap_func.py:160-181; the reference prices ANY capacity
// (Dip/examples/GAP/GAP_KnapPisinger.h:243-249 hands the items to combo whatever c is) -- with the two DP rows
// in global memory (L2-resident: 8 B per cell and row) instead of shared memory.
//
// One CTA per knapsack, items sequential, cells parallel: thread t owns the cells t*4 .. t*4+3 of every
// 4096-cell tile.  Per cell-update: c[i-1][j] and c[i-1][j-w] read, c[i][j] written (24 B through L2), one
// DADD, one strict compare -- the reference's operations on the reference's operands, so optimum and
// decisions are bit-identical.  Decisions go to HBM as one bit per (item, cell), laid out [item][cell / 32]
// (a warp's ballot is one word), and the backtrack kernel of knapsack.cu walks them one item at a time.
#include <algorithm>

#include "common.cuh"

namespace dipgpu {

struct BigDpArgs {
    const double *redCost;         // reduced costs the items are taken from (node bounds: fixed columns masked to +0)
    const int32_t *weight;
    const int32_t *blockColStart;
    const int32_t *capacity;
    const int64_t *bitsOff;        // local block -> word offset into bits
    uint32_t *bits;                // [item][cellWord]
    const int64_t *rowOff;         // local block -> first cell of its two row buffers (each rowStride cells)
    double *rows;
    double *itemP;                 // slot blockColStart[b]+k: profit of active item k
    int32_t *itemW, *itemCol, *nItems;
    double *blockObj;
    int firstBlock;
    int profitMode;
    const int32_t *scale, *shortcut;
    const double *zShort;
    const uint8_t *fix;            // node bounds (else NULL): a fixed column is no item
};

constexpr int kBigThreads = 1024;

__global__ void __launch_bounds__(kBigThreads, 1) knap_dp_big_kernel(const __grid_constant__ BigDpArgs a)
{
    __shared__ int sWarpCnt[kBigThreads / 32];
    __shared__ int sTotal;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int lb = blockIdx.x, b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int shortCode = a.profitMode == DIPGPU_PROFIT_PISINGER_INT ? a.shortcut[lb] : 0;
    const int cap = a.capacity[b];
    const int n = (shortCode != 0 || cap < 0) ? 0 : a.blockColStart[b + 1] - cs;
    const double pScale = a.profitMode == DIPGPU_PROFIT_PISINGER_INT ? (double)a.scale[lb] : 0.0;
    const long long cells = (long long)cap + 1;
    const long long rowStride = (cells + 3) & ~3ll;
    double *row0 = a.rows + a.rowOff[lb], *row1 = row0 + rowStride;
    const long long cellWords = (cells + 31) >> 5;
    uint32_t *bits = a.bits + a.bitsOff[lb];
    // ---- order-preserving compaction of the active items (gap_func.py:102-105) into global arrays
    int nIt = 0;
    for (int base = 0; base < n; base += kBigThreads) {
        const int cl = base + t;
        bool flag = false;
        double rc = 0.0;
        int w = 0;
        if (cl < n) {
            rc = a.redCost[cs + cl];
            w = a.weight[cs + cl];
            flag = rc < 0.0 && w <= cap && (a.fix == nullptr || a.fix[cs + cl] == 0);  // w > cap can never be taken (:161)
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) sWarpCnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int v = 0; v < kBigThreads / 32; ++v) {
            off += v < warp ? sWarpCnt[v] : 0;
            tot += sWarpCnt[v];
        }
        if (flag) {
            const int pos = cs + nIt + off + __popc(bal & ((1u << lane) - 1u));
            a.itemP[pos] = pScale != 0.0 ? (double)(long long)floor((-rc) * pScale + 0.5) : -rc;
            a.itemW[pos] = w;
            a.itemCol[pos] = cl;
        }
        nIt += tot;
        __syncthreads();
    }
    if (t == 0) {
        a.nItems[lb] = nIt;
        sTotal = nIt;
    }
    // c[-1][.] = 0 (gap_func.py:156)
    for (long long j = t; j < cells; j += kBigThreads) row0[j] = 0.0;
    __threadfence_block();
    __syncthreads();
    double *prev = row0, *next = row1;
    for (int i = 0; i < nIt; ++i) {
        const int w = __ldcg(a.itemW + cs + i);
        const double p = __ldcg(a.itemP + cs + i);
        uint32_t *bitRow = bits + (long long)i * cellWords;
        // a warp's 32 lanes take 32 CONSECUTIVE cells, so that their ballot is one decision word
        for (long long j0 = (long long)warp * 32; j0 < cells; j0 += kBigThreads) {
            const long long j = j0 + lane;
            bool take = false;
            if (j < cells) {
                const double keep = __ldcg(prev + j);
                double best = keep;
                if (j >= w) {
                    const double cand = p + __ldcg(prev + j - w);  // :164
                    take = cand > keep;                            // :166 strict
                    if (take) best = cand;
                }
                next[j] = best;
            }
            const unsigned word = __ballot_sync(0xffffffffu, take);
            if (lane == 0) bitRow[j0 >> 5] = word;                 // added[i][j]  (:168)
        }
        __threadfence_block();
        __syncthreads();
        double *tmp = prev;
        prev = next;
        next = tmp;
    }
    if (t == 0) {
        // z = c[n-1][capacity]  (gap_func.py:183)
        double z;
        if (cap < 0) z = __longlong_as_double((long long)0xfff0000000000000ULL);
        else if (shortCode != 0) z = a.zShort[lb];
        else z = __ldcg(prev + cap);
        a.blockObj[lb] = z;
    }
    (void)sTotal;
}

int launchBigDp(Ctx *c, const double *dRedCost, cudaStream_t st)
{
    if (c->nLocal == 0) return DIPGPU_OK;
    BigDpArgs a;
    a.redCost = dRedCost;
    a.weight = c->dWeight;
    a.blockColStart = c->dBlockColStart;
    a.capacity = c->haveFix ? c->dCapEff : c->dCapacity;
    a.bitsOff = c->dBitsOff;
    a.bits = c->dBits;
    a.rowOff = c->dBigRowOff;
    a.rows = c->dBigRows;
    a.itemP = c->dBigItemP;
    a.itemW = c->dItemW;
    a.itemCol = c->dItemCol;
    a.nItems = c->dNItems;
    a.blockObj = reinterpret_cast<double *>(static_cast<char *>(c->dResult) + c->layout.offObj);
    a.firstBlock = c->firstBlock;
    a.profitMode = c->profitMode;
    a.scale = c->dScale;
    a.shortcut = c->dShortcut;
    a.zShort = c->dZShort;
    a.fix = c->haveFix ? c->dFix : nullptr;
    knap_dp_big_kernel<<<c->nLocal, kBigThreads, 0, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

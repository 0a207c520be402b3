// pisinger.cu -- the pre-processing of GAP_KnapPisinger::solve
// (Dip/examples/GAP/GAP_KnapPisinger.h:146-273) for DIPGPU_PROFIT_PISINGER_INT:
//   * calcScaleFactor (:89-143): the smallest power of ten (<= 10^4) that makes the 1e-4-truncated
//     fractional parts of the profits -redCost integral.  The reference's loop is sequential over
//     the items (the running scale feeds the next item's test), so it is evaluated here as an
//     ORDERED composition of per-item transition tables over the states {1,10,100,1000,too big}:
//     each item's table comes from the very floating-point operations of :124-139, and composing
//     tables is associative, hence parallel and still exact.
//   * the trivial cases decided before combo() is called (:203-238): one item, everything fits,
//     two items (enumerated with the reference's preference order on strict '>').
// One CTA per knapsack; the results steer knap_dp_kernel (profits, early exit) and the emit step.
#include "common.cuh"

namespace dipgpu {

// Dip/src/UtilMacros.h:521-531 (UtilEpsilon = 1e-6) and :548-552
__device__ __forceinline__ double utilFracPart(double x)
{
    const double fl = floor(x);
    const double flp = floor(x + 0.5);
    if (fabs(flp - x) < 1.0e-6 * (fabs(flp) + 1.0)) return 0.0;
    return x - fl;
}
__device__ __forceinline__ bool utilIsZero(double x) { return fabs(x) < 1.0e-8; }

// Transition of one item with truncated fractional part wv, entered with scale 10^s (s = 0..3):
// GAP_KnapPisinger.h:124-139.  Returns the exit state, 4 = "factor too big" (scale 10^4, loop ends).
__device__ __forceinline__ int scaleStep(double wv, int s)
{
    int scale = s == 0 ? 1 : s == 1 ? 10 : s == 2 ? 100 : 1000;
    double work = wv * (double)scale;
    while (!utilIsZero(utilFracPart(work))) {
        scale *= 10;
        ++s;
        if ((double)scale >= 1.0 / 1.0e-4) return 4;
        work *= 10.0;
    }
    return s;
}

// tables: 4 entries of 3 bits (entry s = exit state when entered in state s); state 4 absorbs
__device__ __forceinline__ uint32_t tabIdentity() { return 0u | (1u << 3) | (2u << 6) | (3u << 9); }
__device__ __forceinline__ int tabGet(uint32_t t, int s) { return s >= 4 ? 4 : (int)((t >> (3 * s)) & 7u); }
// first `a`, then `b`
__device__ __forceinline__ uint32_t tabCompose(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) r |= (uint32_t)tabGet(b, tabGet(a, s)) << (3 * s);
    return r;
}

struct First2 {
    int c0, c1;  // first two active columns (local ids), -1 = none
};
__device__ __forceinline__ First2 first2Merge(First2 a, First2 b)
{
    First2 r = a;
    if (r.c0 < 0) return b;
    if (r.c1 < 0) r.c1 = b.c0;
    return r;
}

constexpr int kPrepThreads = 128;

__global__ void __launch_bounds__(kPrepThreads) knap_prep_kernel(const __grid_constant__ PrepArgs a)
{
    __shared__ uint32_t sTab[kPrepThreads / 32];
    __shared__ int sCnt[kPrepThreads / 32];
    __shared__ long long sW[kPrepThreads / 32];
    __shared__ First2 sF[kPrepThreads / 32];
    __shared__ double sZ[kPrepThreads / 32];
    __shared__ int sScale, sCntAll;
    __shared__ long long sWAll;
    __shared__ First2 sFAll;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int lb = blockIdx.x;
    const int b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int n = a.blockColStart[b + 1] - cs;
    const int cap = a.capacity[b];
    const double *redCostV = a.redCost + (long long)blockIdx.y * a.rcStride;
    const int vlb = (int)blockIdx.y * (int)gridDim.x + lb;
    // each thread owns a CONTIGUOUS run of columns so that thread order == item order
    const int per = (n + kPrepThreads - 1) / kPrepThreads;
    const int j0 = min(n, t * per), j1 = min(n, j0 + per);
    uint32_t tab = tabIdentity();
    int cnt = 0;
    long long wsum = 0;
    First2 f2{-1, -1};
    // node bounds (DecompAlgo::setSubProbBounds): a fixed column is no item -- the scale factor, the counts and the
    // trivial cases are those of the FREE columns against the capacity left after the columns fixed to 1
    auto isItem = [&](int j) { return redCostV[cs + j] < 0 && (a.fix == nullptr || a.fix[cs + j] == 0); };
    for (int j = j0; j < j1; ++j) {
        const double rc = redCostV[cs + j];
        if (!isItem(j)) continue;  // :109, :173
        ++cnt;
        wsum += a.weight[cs + j];
        if (f2.c0 < 0) f2.c0 = j; else if (f2.c1 < 0) f2.c1 = j;
        double frac = utilFracPart(-rc);
        if (!utilIsZero(frac)) {
            frac *= 1.0 / 1.0e-4;
            const double wv = (double)(int)frac * 1.0e-4;  // :118-119
            uint32_t ti = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) ti |= (uint32_t)scaleStep(wv, s) << (3 * s);
            tab = tabCompose(tab, ti);
        }
    }
    // ordered reduction inside the warp (lane order == item order), then over the warps
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t ot = __shfl_down_sync(0xffffffffu, tab, off);
        const int oc = __shfl_down_sync(0xffffffffu, cnt, off);
        const long long ow = __shfl_down_sync(0xffffffffu, wsum, off);
        First2 of;
        of.c0 = __shfl_down_sync(0xffffffffu, f2.c0, off);
        of.c1 = __shfl_down_sync(0xffffffffu, f2.c1, off);
        if ((lane & (2 * off - 1)) == 0) {
            tab = tabCompose(tab, ot);
            cnt += oc;
            wsum += ow;
            f2 = first2Merge(f2, of);
        }
    }
    if (lane == 0) {
        sTab[warp] = tab;
        sCnt[warp] = cnt;
        sW[warp] = wsum;
        sF[warp] = f2;
    }
    __syncthreads();
    if (t == 0) {
        uint32_t tt = tabIdentity();
        int c = 0;
        long long w = 0;
        First2 f{-1, -1};
        for (int v = 0; v < kPrepThreads / 32; ++v) {
            tt = tabCompose(tt, sTab[v]);
            c += sCnt[v];
            w += sW[v];
            f = first2Merge(f, sF[v]);
        }
        const int st = tabGet(tt, 0);  // the loop starts with scaleFactor = 1
        sScale = st == 0 ? 1 : st == 1 ? 10 : st == 2 ? 100 : st == 3 ? 1000 : 10000;
        sCntAll = c;
        sWAll = w;
        sFAll = f;
    }
    __syncthreads();
    const double scale = (double)sScale;
    const int nAll = sCntAll;
    // scaled integer profit of a column, :174-177
    auto profit = [&](int j) { return (double)(long long)floor((-redCostV[cs + j]) * scale + 0.5); };
    int code = 0, sel0 = -1, sel1 = -1;
    double z = 0.0;
    if (cap < 0) {
        code = 2;  // the columns fixed to 1 do not fit: the block returns no column (the DP kernel sees capacity < 0)
    } else if (nAll == 0) {
        code = 2;
    } else if (nAll == 1) {
        // :203-205 (the reference asserts w <= capacity; an item that does not fit is left out)
        code = 2;
        if (a.weight[cs + sFAll.c0] <= cap) {
            sel0 = sFAll.c0;
            z = profit(sel0);
        }
    } else if (sWAll <= (long long)cap) {
        // :206-212 everything fits: take every item with negative reduced cost
        code = 1;
        double part = 0.0;
        for (int j = j0; j < j1; ++j)
            if (isItem(j)) part += profit(j);  // integers < 2^53: exact in any order
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0) sZ[warp] = part;
        __syncthreads();
        for (int v = 0; v < kPrepThreads / 32; ++v) z += sZ[v];
    } else if (nAll == 2) {
        // :213-238, preference order {1}, {0}, {0,1} on strict '>'
        code = 2;
        const int c0 = sFAll.c0, c1 = sFAll.c1;
        const int w0 = a.weight[cs + c0], w1 = a.weight[cs + c1];
        const double p0 = profit(c0), p1 = profit(c1);
        double best = 0.0;
        int x0 = 0, x1 = 0;
        if (w1 <= cap && p1 > best) { best = p1; x0 = 0; x1 = 1; }
        if (w0 <= cap && p0 > best) { best = p0; x0 = 1; x1 = 0; }
        if (w0 + w1 <= cap && p0 + p1 > best) { best = p0 + p1; x0 = 1; x1 = 1; }
        z = best;
        if (x0 && x1) { sel0 = c0; sel1 = c1; }
        else if (x0) sel0 = c0;
        else if (x1) sel0 = c1;
    }
    if (t == 0) {
        a.scale[vlb] = sScale;
        a.shortcut[vlb] = code;
        a.sel[2 * vlb] = sel0;
        a.sel[2 * vlb + 1] = sel1;
        a.zShort[vlb] = z;
    }
}

static int launchPrep(Ctx *c, const double *dRedCost, int64_t rcStride, int nVec, bool batch, cudaStream_t st)
{
    if (c->nLocal == 0 || nVec <= 0 || c->profitMode != DIPGPU_PROFIT_PISINGER_INT) return DIPGPU_OK;
    PrepArgs a;
    a.redCost = dRedCost;
    a.weight = c->dWeight;
    a.blockColStart = c->dBlockColStart;
    a.capacity = c->haveFix ? c->dCapEff : c->dCapacity;
    a.fix = c->haveFix ? c->dFix : nullptr;
    a.firstBlock = c->firstBlock;
    a.rcStride = rcStride;
    a.scale = batch ? c->dScaleB : c->dScale;
    a.shortcut = batch ? c->dShortcutB : c->dShortcut;
    a.sel = batch ? c->dSelB : c->dSel;
    a.zShort = batch ? c->dZShortB : c->dZShort;
    knap_prep_kernel<<<dim3(c->nLocal, nVec, 1), kPrepThreads, 0, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

int launchKnapsackPrep(Ctx *c, const double *dRedCost, cudaStream_t st)
{
    return launchPrep(c, dRedCost, 0, 1, false, st);
}

int launchKnapsackPrepBatch(Ctx *c, const double *dRedCost, int64_t rcStride, int nVec, cudaStream_t st)
{
    return launchPrep(c, dRedCost, rcStride, nVec, true, st);
}

}  // namespace dipgpu

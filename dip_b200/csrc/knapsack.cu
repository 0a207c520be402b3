// knapsack.cu -- stage (b) of the pricing round: the independent block
// subproblems as batched 0-1 knapsack dynamic programs.
//
// Replaces the per-block DecompApp::solveRelaxed calls
// (Dip/src/DecompAlgo.cpp:4815-4848 -> GAP_DecompApp::solveRelaxed,
// Dip/examples/GAP/GAP_DecompApp.cpp:235-285).  The recurrence, item order and
// tie rule are those of the reference's own DP, knapsack01()
// (Dip/src/dippy/examples/gap/gap_func.py:160-181):
//     items  = columns of the block with redCost < 0, ascending   (:102)
//     profit = -redCost (fp64)                                     (:104)
//     take item i at capacity j  iff  p_i + c[i-1][j-w_i] > c[i-1][j]   (:163-167)
//     backtrack from (last item, capacity)                         (:173-181)
// One fp64 add and one fp64 compare per cell, exactly as the reference, so the
// optimum and the decision table are bit-identical whatever the parallel shape.
//
// Shape: one CTA per knapsack, items sequential, capacity parallel.  Thread t of
// T owns cells t, t+T, ... (S of them) IN REGISTERS; shared memory only carries
// the shifted row c[i-1][j-w] between threads (double-buffered, one barrier per
// item): 16 B of shared-memory traffic per cell instead of the 24 B of a
// row-in-smem formulation.  The block's reduced costs and weights are staged by
// TMA bulk copies (cp.async.bulk + mbarrier) and compacted in shared memory.
// Decisions are bit-packed per CELL over 32 consecutive items (one register per
// owned cell) and streamed to HBM as coalesced 32-bit words, laid out
// [itemWord][cell] so that the backtrack can skip 1024 items per warp load.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dipgpu {

// ------------------------------------------------------------------ PTX helpers

__device__ __forceinline__ uint32_t smemAddr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbarInit(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smemAddr(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (1-D), completion counted on an mbarrier.
__device__ __forceinline__ void tmaLoad1d(void *dstSmem, const void *srcGlobal, uint32_t bytes,
                                          uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smemAddr(dstSmem)),
        "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
        : "memory");
}

// ------------------------------------------------------------------- DP kernel

// Shifted reads below cell 0 see -inf, so "p + c[i-1][j-w] > c[i-1][j]" is false
// there, which is the reference's "w > j: keep c[i-1][j]" branch.
#define kNegInf (__longlong_as_double((long long)0xfff0000000000000ULL))

struct KnapArgs {
    const double *redCost;         // x-space reduced costs (16 B aligned, padded)
    const int32_t *weight;         // x-space weights (padded)
    const int32_t *blockColStart;  // global block -> first column
    const int32_t *capacity;       // global block -> capacity
    const int64_t *bitsOff;        // local block -> word offset into bits
    uint32_t *bits;                // decisions [itemWord][cell]
    int32_t *itemW;                // slot blockColStart[b]+k: weight of active item k
    int32_t *itemCol;              // slot blockColStart[b]+k: local column of active item k
    int32_t *nItems;               // local block -> number of active items
    double *blockObj;              // local block -> knapsack optimum
    int firstBlock;
    int chunkCols;
};

__host__ __device__ inline size_t dpSmemBytes(int rowCells, int chunkCols)
{
    size_t b = 0;
    b += sizeof(double) * 2 * (size_t)rowCells;        // double-buffered row
    b += sizeof(double) * ((size_t)chunkCols + 2);      // raw reduced costs (TMA dst)
    b += sizeof(double) * ((size_t)chunkCols + 2);      // compacted profits (+1 prefetch pad)
    b += sizeof(int32_t) * ((size_t)chunkCols + 4);     // raw weights (TMA dst)
    b += sizeof(int32_t) * ((size_t)chunkCols + 4);     // compacted weights (+1 prefetch pad)
    b += sizeof(int32_t) * 32;                          // per-warp counts
    b += 16;                                            // mbarrier
    return b;
}

template <int T, int S>
__global__ void __launch_bounds__(T) knap_dp_kernel(const KnapArgs a)
{
    constexpr int RC = T * S;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int CH = a.chunkCols;
    double *row = reinterpret_cast<double *>(smemRaw);
    double *rawRc = row + 2 * RC;
    double *itP = rawRc + (CH + 2);
    int32_t *rawW = reinterpret_cast<int32_t *>(itP + (CH + 2));
    int32_t *itW = rawW + (CH + 4);
    int32_t *warpCnt = itW + (CH + 4);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(warpCnt + 32);

    const int t = threadIdx.x;
    const int lane = t & 31;
    const int warp = t >> 5;
    const int lb = blockIdx.x;
    const int b = a.firstBlock + lb;
    const int colStart = a.blockColStart[b];
    const int nCols = a.blockColStart[b + 1] - colStart;
    const int cap = a.capacity[b];
    uint32_t *bitsBase = a.bits + a.bitsOff[lb];

    if (t == 0) mbarInit(mbar, 1);
    for (int j = t; j < RC; j += T) row[j] = 0.0;  // c[-1][.] = 0  (gap_func.py:156)
    __syncthreads();

    double mine[S];
    uint32_t acc[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        mine[s] = 0.0;
        acc[s] = 0u;
    }

    int cur = 0;        // buffer holding c[i-1][.]
    int kGlobal = 0;    // active items processed so far
    uint32_t parity = 0;

    for (int chunkBase = 0; chunkBase < nCols; chunkBase += CH) {
        const int len = min(CH, nCols - chunkBase);
        const int g0 = colStart + chunkBase;
        const int shiftRc = g0 & 1, shiftW = g0 & 3;
        // ---- stage the chunk's reduced costs and weights with TMA bulk copies
        if (t == 0) {
            const uint32_t bytesRc = (uint32_t)(((shiftRc + len + 1) & ~1) * 8);
            const uint32_t bytesW = (uint32_t)(((shiftW + len + 3) & ~3) * 4);
            mbarExpectTx(mbar, bytesRc + bytesW);
            tmaLoad1d(rawRc, a.redCost + (g0 - shiftRc), bytesRc, mbar);
            tmaLoad1d(rawW, a.weight + (g0 - shiftW), bytesW, mbar);
        }
        mbarWait(mbar, parity);
        parity ^= 1u;

        // ---- order-preserving compaction of the active items (gap_func.py:102-105)
        int nIt = 0;
        for (int base = 0; base < len; base += T) {
            const int cl = base + t;
            bool flag = false;
            double rc = 0.0;
            int w = 0;
            if (cl < len) {
                rc = rawRc[shiftRc + cl];
                w = rawW[shiftW + cl];
                flag = (rc < 0.0) && (w <= cap);  // w > cap can never be taken (:161)
            }
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            const int wpre = __popc(bal & ((1u << lane) - 1u));
            int off = 0, tot = __popc(bal);
            if (T > 32) {
                if (lane == 0) warpCnt[warp] = tot;
                __syncthreads();
                tot = 0;
#pragma unroll
                for (int v = 0; v < T / 32; ++v) {
                    const int cv = warpCnt[v];
                    off += v < warp ? cv : 0;
                    tot += cv;
                }
                __syncthreads();
            }
            if (flag) {
                const int pos = nIt + off + wpre;
                itP[pos] = -rc;
                itW[pos] = w;
                const int slot = colStart + kGlobal + pos;
                a.itemW[slot] = w;
                a.itemCol[slot] = chunkBase + cl;
            }
            nIt += tot;
        }
        __syncthreads();

        // ---- the DP proper: items sequential, cells parallel
        int wNext = nIt > 0 ? itW[0] : 0;
        double pNext = nIt > 0 ? itP[0] : 0.0;
        for (int i = 0; i < nIt; ++i) {
            const int w = wNext;
            const double p = pNext;
            const double *rd = row + cur * RC + (t - w);
            double *wr = row + (cur ^ 1) * RC + t;
            const int bitpos = (kGlobal + i) & 31;
            const uint32_t bm = 1u << bitpos;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int src = t + s * T - w;
                // c[i-1][j-w]; j < w keeps c[i-1][j]  (gap_func.py:161-162)
                const double sh = src >= 0 ? rd[s * T] : kNegInf;
                const double cand = p + sh;           // :164
                const bool take = cand > mine[s];     // :166 strict
                mine[s] = take ? cand : mine[s];
                acc[s] |= take ? bm : 0u;             // added[i][j]  (:168)
                wr[s * T] = mine[s];
            }
            // prefetch the next item before the barrier (itP/itW are padded by one)
            wNext = itW[i + 1];
            pNext = itP[i + 1];
            if (bitpos == 31) {
                uint32_t *dst = bitsBase + (size_t)((kGlobal + i) >> 5) * RC + t;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    dst[s * T] = acc[s];
                    acc[s] = 0u;
                }
            }
            cur ^= 1;
            __syncthreads();
        }
        kGlobal += nIt;
    }

    if ((kGlobal & 31) != 0) {
        uint32_t *dst = bitsBase + (size_t)(kGlobal >> 5) * RC + t;
#pragma unroll
        for (int s = 0; s < S; ++s) dst[s * T] = acc[s];
    }
    // z = c[n-1][capacity]  (gap_func.py:183)
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (t + s * T == cap) a.blockObj[lb] = mine[s];
    }
    if (t == 0) a.nItems[lb] = kGlobal;
}

// --------------------------------------------------------------- geometry table

struct GeomEntry {
    int T, S;
    void (*fn)(const KnapArgs);
};

#define GE(T, S) {T, S, knap_dp_kernel<T, S>}
static const GeomEntry kGeoms[] = {
    GE(32, 1),   GE(32, 2),   GE(32, 3),   GE(32, 4),   GE(32, 5),   GE(32, 6),   GE(32, 7),  GE(32, 8),
    GE(64, 2),   GE(64, 3),   GE(64, 4),   GE(64, 5),   GE(64, 6),   GE(64, 7),   GE(64, 8),
    GE(128, 3),  GE(128, 4),  GE(128, 5),  GE(128, 6),  GE(128, 7),  GE(128, 8),
    GE(256, 4),  GE(256, 5),  GE(256, 6),  GE(256, 7),  GE(256, 8),
    GE(512, 4),  GE(512, 5),  GE(512, 6),  GE(512, 7),  GE(512, 8),
    GE(1024, 4), GE(1024, 5), GE(1024, 6), GE(1024, 7), GE(1024, 8), GE(1024, 10), GE(1024, 12),
};
#undef GE
static const int kNumGeoms = (int)(sizeof(kGeoms) / sizeof(kGeoms[0]));
static const size_t kMaxSmem = 227 * 1024;

static int chunkColsFor(int T) { return T <= 64 ? 512 : T <= 256 ? 1024 : 2048; }

int chooseDpGeom(Ctx *c, DpGeom *g)
{
    const int cells = c->maxCapacity + 1;
    const GeomEntry *best = nullptr;
    if (c->optThreads > 0 || c->optCellsPerThread > 0) {
        for (int i = 0; i < kNumGeoms; ++i) {
            const GeomEntry &e = kGeoms[i];
            if (c->optThreads > 0 && e.T != c->optThreads) continue;
            if (c->optCellsPerThread > 0 && e.S != c->optCellsPerThread) continue;
            if (e.T * e.S < cells) continue;
            if (!best || e.T * e.S < best->T * best->S) best = &e;
        }
        if (!best)
            return fail(c, DIPGPU_ERR_INVALID, "no DP geometry with threads=%d cellsPerThread=%d covers %d cells",
                        c->optThreads, c->optCellsPerThread, cells);
    } else {
        // smallest thread count whose row fits with S <= 8 (S <= 12 at T = 1024),
        // then the smallest S that covers the row
        for (int i = 0; i < kNumGeoms && !best; ++i)
            if (kGeoms[i].T * kGeoms[i].S >= cells) best = &kGeoms[i];
        if (!best)
            return fail(c, DIPGPU_ERR_UNSUPPORTED,
                        "capacity %d needs %d DP cells; one CTA stages at most %d", c->maxCapacity, cells,
                        1024 * 12);
    }
    g->T = best->T;
    g->S = best->S;
    g->rowCells = best->T * best->S;
    g->chunkCols = chunkColsFor(best->T);
    g->smemBytes = dpSmemBytes(g->rowCells, g->chunkCols);
    if (g->smemBytes > kMaxSmem)
        return fail(c, DIPGPU_ERR_UNSUPPORTED, "DP row of %d cells needs %zu B shared memory (> %zu)",
                    g->rowCells, g->smemBytes, kMaxSmem);
    return DIPGPU_OK;
}

int launchKnapsackDp(Ctx *c, const double *dRedCost, cudaStream_t st)
{
    if (c->nLocal == 0) return DIPGPU_OK;
    const DpGeom &g = c->geom;
    const GeomEntry *e = nullptr;
    for (int i = 0; i < kNumGeoms; ++i)
        if (kGeoms[i].T == g.T && kGeoms[i].S == g.S) e = &kGeoms[i];
    if (!e) return fail(c, DIPGPU_ERR_STATE, "DP geometry not initialised");
    KnapArgs a;
    a.redCost = dRedCost;
    a.weight = c->dWeight;
    a.blockColStart = c->dBlockColStart;
    a.capacity = c->dCapacity;
    a.bitsOff = c->dBitsOff;
    a.bits = c->dBits;
    a.itemW = c->dItemW;
    a.itemCol = c->dItemCol;
    a.nItems = c->dNItems;
    a.blockObj = reinterpret_cast<double *>(static_cast<char *>(c->dResult) + c->layout.offObj);
    a.firstBlock = c->firstBlock;
    a.chunkCols = g.chunkCols;
    DIPGPU_CUDA(c, cudaFuncSetAttribute(e->fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)g.smemBytes));
    e->fn<<<c->nLocal, g.T, g.smemBytes, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// --------------------------------------------------------- backtrack + emit
//
// One warp per knapsack walks the decision table from (last item, capacity)
// (gap_func.py:173-181).  With bits laid out [itemWord][cell], the lanes of a
// warp fetch the 32 item-words below the current item for the current cell in
// one gather, so every iteration of the walk lands on the next TAKEN item.
// The column is emitted with ascending x-space indices; varRedCost/varOrigCost
// are accumulated in that order (GAP_KnapPisinger.h:263-272), then alpha is
// subtracted (DecompAlgo.cpp:6350-6356).

struct BackArgs {
    const double *redCost;
    const double *obj;
    const double *alpha;  // indexed by GLOBAL block
    const int32_t *blockColStart;
    const int32_t *capacity;
    const int64_t *bitsOff;
    const uint32_t *bits;
    const int32_t *itemW;
    const int32_t *itemCol;
    const int32_t *nItems;
    int32_t *scratch;
    int32_t *candInd;
    double *blockRedCost;
    double *blockOrigCost;
    int32_t *blockNnz;
    int firstBlock, nLocal, rowCells;
};

__global__ void __launch_bounds__(128) knap_backtrack_kernel(const BackArgs a)
{
    const int lane = threadIdx.x & 31;
    const int lb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (lb >= a.nLocal) return;
    const int b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int RC = a.rowCells;
    const uint32_t *bitsBase = a.bits + a.bitsOff[lb];
    int j = a.capacity[b];
    int i = a.nItems[lb] - 1;
    int cnt = 0;
    while (i >= 0) {
        const int kw = i >> 5;
        const int wv = kw - lane;
        uint32_t v = wv >= 0 ? __ldg(bitsBase + (size_t)wv * RC + j) : 0u;
        if (lane == 0) v &= 0xffffffffu >> (31 - (i & 31));
        const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
        if (nz == 0u) {
            i = ((kw - 32) << 5) | 31;  // nothing taken in these 1024 items
            continue;
        }
        const int L = __ffs(nz) - 1;
        const uint32_t vL = __shfl_sync(0xffffffffu, v, L);
        const int item = ((kw - L) << 5) + (31 - __clz(vL));
        if (lane == 0) a.scratch[cs + cnt] = a.itemCol[cs + item];
        ++cnt;
        j -= __ldg(a.itemW + cs + item);
        i = item - 1;
    }
    __syncwarp();
    double rcs = 0.0, ocs = 0.0;
    for (int base = 0; base < cnt; base += 32) {
        const int k = base + lane;
        double r = 0.0, o = 0.0;
        if (k < cnt) {
            const int col = cs + a.scratch[cs + cnt - 1 - k];
            a.candInd[cs + k] = col;
            r = a.redCost[col];
            o = a.obj[col];
        }
        const int nn = min(32, cnt - base);
        for (int q = 0; q < nn; ++q) {
            rcs += __shfl_sync(0xffffffffu, r, q);
            ocs += __shfl_sync(0xffffffffu, o, q);
        }
    }
    if (lane == 0) {
        a.blockRedCost[lb] = rcs - a.alpha[b];
        a.blockOrigCost[lb] = ocs;
        a.blockNnz[lb] = cnt;
    }
}

int launchBacktrackEmit(Ctx *c, const double *dRedCost, const double *dAlpha, cudaStream_t st)
{
    if (c->nLocal == 0) return DIPGPU_OK;
    char *res = static_cast<char *>(c->dResult);
    BackArgs a;
    a.redCost = dRedCost;
    a.obj = c->dObj;
    a.alpha = dAlpha;
    a.blockColStart = c->dBlockColStart;
    a.capacity = c->dCapacity;
    a.bitsOff = c->dBitsOff;
    a.bits = c->dBits;
    a.itemW = c->dItemW;
    a.itemCol = c->dItemCol;
    a.nItems = c->dNItems;
    a.scratch = c->dScratch;
    a.candInd = c->dCandInd;
    a.blockRedCost = reinterpret_cast<double *>(res + c->layout.offRedCost);
    a.blockOrigCost = reinterpret_cast<double *>(res + c->layout.offOrigCost);
    a.blockNnz = reinterpret_cast<int32_t *>(res + c->layout.offNnz);
    a.firstBlock = c->firstBlock;
    a.nLocal = c->nLocal;
    a.rowCells = c->geom.rowCells;
    const int warpsPerCta = 4;
    const int grid = (c->nLocal + warpsPerCta - 1) / warpsPerCta;
    knap_backtrack_kernel<<<grid, warpsPerCta * 32, 0, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

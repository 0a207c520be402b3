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
#include <utility>

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

// ------------------------------------------------------------------------------------------------------------------
// Rows up to 16 x 13 312 cells: ONE THREAD-BLOCK CLUSTER per knapsack, the row distributed over the shared memories of
// its CTAs (distributed shared memory) -- the same recurrence at shared-memory speed instead of L2 speed.
//
// CTA r of a cluster of C owns the cells [r L, (r+1) L), L = S * 1024; inside it thread t owns the cells
// r L + s 1024 + t (s = 0..S-1), their values in REGISTERS.  For a fixed s the 32 lanes of a warp hold 32 consecutive
// cells: the shifted read c[i-1][j-w] is a conflict-free row of shared memory -- of this CTA or of a lower-ranked one,
// read through the cluster window (mapa + ld.shared::cluster) --, and the warp's ballot is one decision word of the
// [item][cell / 32] table the backtrack kernel walks.  Per item: S shifted reads, S adds, S strict compares, S stores
// of the new cells into the CTA's other row buffer, ONE cluster barrier (release / acquire: the stores are visible to
// the next item's remote reads, and nobody overwrites a buffer somebody still reads).  Items, order, operands and tie
// rule are those of knapsack01 (gap_func.py:160-181): optimum and decisions are bit-identical to every other DP
// kernel here.
constexpr int kClusterThreads = 1024;
constexpr int kClusterMaxS = 13;   // 2 x 13 x 1024 x 8 B = 213 KB of row buffers per CTA
constexpr int kClusterStage = 8;   // items whose decision words are staged in shared memory between two flushes

struct ClusterDpArgs {
    BigDpArgs b;
    int S;            // cells per thread
    int clusterSize;  // CTAs per knapsack
    int dbg;          // timing experiments only (WRONG results): 1 = no cluster barrier per item, 2 = every shifted read from the own slice
};

__device__ __forceinline__ unsigned clusterCtaRank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void clusterSync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ldClusterF64(uint32_t localAddr, unsigned rank)
{
    uint32_t remote;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(localAddr), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}

template <int S, int T>
__global__ void __launch_bounds__(T, 1) knap_dp_cluster_kernel(const __grid_constant__ ClusterDpArgs ca)
{
    extern __shared__ __align__(16) unsigned char clSm[];
    __shared__ int sWarpCnt[T / 32];
    const BigDpArgs &a = ca.b;
    constexpr int L = S * T;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int C = ca.clusterSize;
    const int r = (int)clusterCtaRank();
    const int lb = blockIdx.x / C, b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int shortCode = a.profitMode == DIPGPU_PROFIT_PISINGER_INT ? a.shortcut[lb] : 0;
    const int cap = a.capacity[b];
    const int n = (shortCode != 0 || cap < 0) ? 0 : a.blockColStart[b + 1] - cs;
    const double pScale = a.profitMode == DIPGPU_PROFIT_PISINGER_INT ? (double)a.scale[lb] : 0.0;
    const long long cells = (long long)cap + 1;
    const long long cellWords = (cells + 31) >> 5;
    uint32_t *bits = a.bits + a.bitsOff[lb];
    double *buf0 = reinterpret_cast<double *>(clSm);
    // decision words of the last kClusterStage items, [item][s * 32 + warp]: written to HBM eight items at a time, so
    // that the GPU-scope fence inside every cluster barrier (MEMBAR.ALL.GPU) rarely finds a global store in flight
    uint32_t *sWords = reinterpret_cast<uint32_t *>(clSm + (size_t)2 * L * sizeof(double));
    constexpr int kWordsPerItem = S * (T / 32);
    auto flushWords = [&](int iFirst, int count) {
        for (int idx = t; idx < count * kWordsPerItem; idx += T) {
            const int k = idx / kWordsPerItem, wi = idx - k * kWordsPerItem;
            const long long gw = (long long)r * kWordsPerItem + wi;
            if (gw < cellWords) bits[(long long)(iFirst + k) * cellWords + gw] = sWords[idx];
        }
    };
    // ---- order-preserving compaction of the active items (gap_func.py:102-105): every CTA of the cluster counts,
    // rank 0 writes the global arrays (the backtrack kernel reads them too)
    int nIt = 0;
    for (int base = 0; base < n; base += T) {
        const int cl = base + t;
        bool flag = false;
        double rc = 0.0;
        int w = 0;
        if (cl < n) {
            rc = a.redCost[cs + cl];
            w = a.weight[cs + cl];
            flag = rc < 0.0 && w <= cap && (a.fix == nullptr || a.fix[cs + cl] == 0);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) sWarpCnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
        for (int v = 0; v < T / 32; ++v) {
            off += v < warp ? sWarpCnt[v] : 0;
            tot += sWarpCnt[v];
        }
        if (flag && r == 0) {
            const int pos = cs + nIt + off + __popc(bal & ((1u << lane) - 1u));
            a.itemP[pos] = pScale != 0.0 ? (double)(long long)floor((-rc) * pScale + 0.5) : -rc;
            a.itemW[pos] = w;
            a.itemCol[pos] = cl;
        }
        nIt += tot;
        __syncthreads();
    }
    if (r == 0 && t == 0) a.nItems[lb] = nIt;
    // c[-1][.] = 0 (gap_func.py:156)
    double mine[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        mine[s] = 0.0;
        buf0[s * T + t] = 0.0;
    }
    __threadfence();   // rank 0's item arrays, before the other CTAs of the cluster read them
    clusterSync();
    const uint32_t bufA = (uint32_t)__cvta_generic_to_shared(buf0);
    int cur = 0;
    int w = nIt > 0 ? __ldcg(a.itemW + cs) : 0;
    double p = nIt > 0 ? __ldcg(a.itemP + cs) : 0.0;
    for (int i = 0; i < nIt; ++i) {
        // the next item's weight and profit travel under this item's work
        const int wNext = i + 1 < nIt ? __ldcg(a.itemW + cs + i + 1) : 0;
        const double pNext = i + 1 < nIt ? __ldcg(a.itemP + cs + i + 1) : 0.0;
        // cell j - w of this thread's cell (s = 0): CTA rq0 of the cluster, offset off0 of its slice; the cells s > 0 lie
        // s T further on, in the next CTA once the offset passes L
        const int wq = w / L, wr = w - wq * L;
        int off0 = t - wr, rq0 = r - wq;
        if (off0 < 0) {
            off0 += L;
            rq0 -= 1;
        }
        const uint32_t rdBase = bufA + (uint32_t)cur * (uint32_t)(L * 8);
        double *wr_ = buf0 + (cur ^ 1) * L;
        uint32_t myWord = 0u;
        // the shifted cell of (s, t): this CTA's own slice (plain ld.shared -- the cluster window or a generic load of
        // the own slice runs at the remote rate: 343 instead of 443 GCUPS), a lower CTA's (mapa + ld.shared::cluster), or
        // none (j < w: -inf never wins, gap_func.py:161)
        auto loadShift = [&](int off, int rq) -> double {
            double v = __longlong_as_double((long long)0xfff0000000000000ULL);
            const uint32_t addr = rdBase + (uint32_t)off * 8u;
            if (rq == r || ((ca.dbg & 2) && rq >= 0)) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
            else if (rq >= 0) v = ldClusterF64(addr, (unsigned)rq);
            return v;
        };
        constexpr int kBatch = S > 7 ? (S + 1) / 2 : S;   // shifted reads in flight per thread (registers: 1 024 threads)
#pragma unroll
        for (int s0 = 0; s0 < S; s0 += kBatch) {
            double sh[kBatch];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int s = s0 + q;
                if (s < S) {
                    int off = off0 + s * T, rq = rq0;
                    if (off >= L) {
                        off -= L;
                        rq += 1;
                    }
                    sh[q] = loadShift(off, rq);
                }
            }
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int s = s0 + q;
                if (s < S) {
                    const double keep = mine[s];
                    const double cand = p + sh[q];                 // gap_func.py:164
                    const bool take = cand > keep;                 // :166 strict
                    if (take) mine[s] = cand;
                    wr_[s * T + t] = mine[s];
                    const unsigned word = __ballot_sync(0xffffffffu, take);
                    if (lane == s) myWord = word;                  // added[i][j]  (:168): lane s keeps the word of the cells s T + warp 32 ..
                }
            }
        }
        if (lane < S) sWords[(i % kClusterStage) * kWordsPerItem + lane * (T / 32) + warp] = myWord;
        if ((i % kClusterStage) == kClusterStage - 1) {
            // (before the arrival: a warp that leaves the barrier may overwrite the staging area at once)
            __syncthreads();
            flushWords(i - (kClusterStage - 1), kClusterStage);
        }
        // (DIPGPU_CLUSTER_RELAXED: a CTA-scope fence and a relaxed arrival instead of the cluster-scope release, which is a
        // MEMBAR.ALL.GPU in SASS -- measured 553 instead of 520 GCUPS, not used: the PTX memory model asks for the release)
#ifdef DIPGPU_CLUSTER_RELAXED
        asm volatile("fence.acq_rel.cta;\n\tbarrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#else
        if (!(ca.dbg & 1)) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#endif
        cur ^= 1;
        w = wNext;
        p = pNext;
        if (!(ca.dbg & 1)) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        else __syncthreads();
    }
    if (nIt % kClusterStage != 0) {
        __syncthreads();
        flushWords(nIt - nIt % kClusterStage, nIt % kClusterStage);
    }
    if (ca.dbg & 1) clusterSync();   // (timing experiment without the per-item barrier: nobody leaves while a peer may still read its slice)
    // z = c[n-1][capacity]  (gap_func.py:183)
    if (cap < 0 || shortCode != 0) {
        if (r == 0 && t == 0)
            a.blockObj[lb] = cap < 0 ? __longlong_as_double((long long)0xfff0000000000000ULL) : a.zShort[lb];
    } else {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if ((long long)r * L + s * T + t == (long long)cap) a.blockObj[lb] = mine[s];
    }
}

typedef void (*ClusterFn)(const ClusterDpArgs);
template <int T, int... Ss>
static ClusterFn clusterFnTable(int S, std::integer_sequence<int, Ss...>)
{
    ClusterFn fn = nullptr;
    ((S == Ss + 1 ? (fn = knap_dp_cluster_kernel<Ss + 1, T>) : nullptr), ...);
    return fn;
}
static ClusterFn clusterFnFor(int S, int T)
{
    if (T == 512) return clusterFnTable<512>(S, std::make_integer_sequence<int, 2 * kClusterMaxS>{});
    return clusterFnTable<1024>(S, std::make_integer_sequence<int, kClusterMaxS>{});
}

// Cluster shape for rows of `cells` cells: the smallest power-of-two cluster whose CTAs hold them with <= 13 cells per
// thread, then the fewest cells per thread that still cover the row.  0 = no cluster covers it (rows in global memory).
void planClusterDp(int64_t cells, int threads, int maxCells, int *clusterSize, int *S)
{
    *clusterSize = 0;
    *S = 0;
    int maxS = kClusterMaxS * (kClusterThreads / threads);   // the same 213 KB of row buffers per CTA
    if (maxCells > 0) maxS = std::min(maxS, maxCells);
    for (int C = 2; C <= 16; C *= 2) {
        if ((int64_t)C * maxS * threads >= cells) {
            *clusterSize = C;
            *S = (int)((cells + (int64_t)C * threads - 1) / ((int64_t)C * threads));
            return;
        }
    }
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
    if (c->geom.cluster > 0 && !c->clusterUnavailable) {
        ClusterDpArgs ca;
        ca.b = a;
        ca.S = c->geom.S;
        ca.clusterSize = c->geom.cluster;
        ca.dbg = c->optClusterDebug;
        const int T = c->geom.T;
        const ClusterFn fn = clusterFnFor(c->geom.S, T);
        const size_t smem = (size_t)2 * c->geom.S * T * sizeof(double) + (size_t)kClusterStage * c->geom.S * (T / 32) * sizeof(uint32_t);
        if (c->clusterFnConfigured != (void *)fn) {
            int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(fn), smem);
            if (rc) return rc;
            if (c->geom.cluster > 8)
                DIPGPU_CUDA(c, cudaFuncSetAttribute(reinterpret_cast<const void *>(fn), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)c->nLocal * (unsigned)c->geom.cluster, 1, 1);
        cfg.blockDim = dim3((unsigned)T, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)c->geom.cluster;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (c->clusterFnConfigured != (void *)fn) {
            // can the device place one such cluster at all (a 16-CTA cluster of whole-SM CTAs needs 16 free SMs in one GPC)?
            int nClusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nClusters, fn, &cfg) != cudaSuccess || nClusters < 1) {
                (void)cudaGetLastError();
                c->clusterUnavailable = true;
            }
            c->clusterFnConfigured = (void *)fn;
        }
        if (!c->clusterUnavailable) {
            DIPGPU_CUDA(c, cudaLaunchKernelEx(&cfg, fn, ca));
            c->launches++;
            DIPGPU_CUDA(c, cudaGetLastError());
            return DIPGPU_OK;
        }
    }
    knap_dp_big_kernel<<<c->nLocal, kBigThreads, 0, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

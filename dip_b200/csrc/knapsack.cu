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
// T owns the S consecutive cells [t*S, t*S+S) IN REGISTERS; shared memory only
// carries the shifted row c[i-1][j-w] between threads (double-buffered, one
// barrier per item or per two items): 16 B of shared-memory traffic per cell
// instead of the 24 B of a row-in-smem formulation.  The block's reduced costs and weights are staged by
// TMA bulk copies (cp.async.bulk + mbarrier) and compacted in shared memory.
// Decisions are bit-packed per CELL over 32 consecutive items (one register per
// owned cell) and streamed to HBM as coalesced 32-bit words, laid out
// [itemWord][cell] so that the backtrack can skip 1024 items per warp load.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <type_traits>

#include "filter.cuh"

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
    unsigned int *doneCounter;  // non-NULL: the last CTA to finish runs the filter (stage c)
    const int32_t *shortcut;    // Pisinger mode (else NULL): per local block 0 = DP, 1 = all, 2 = sel
    const int32_t *sel;         // 2 per local block
    const uint8_t *fix;         // node bounds (else NULL): per column bit 0 = fixed to 1, bit 1 = fixed to 0
    int bigLayout;              // 1: bits are [item][cell / 32] (bigdp.cu), rowCells is unused
};

constexpr int kBackThreads = 128;

// Pisinger-mode trivial cases (GAP_KnapPisinger.h:203-238), decided by knap_prep_kernel: the
// column is either every item with negative reduced cost (code 1) or an explicit set of at most
// two columns (code 2).  Emitted ascending with the sums of GAP_KnapPisinger.h:263-272.
__device__ __forceinline__ int emitShortcut(const BackArgs &a, int lb, int lane, int code)
{
    // fills candInd[cs ..) with the chosen columns, ascending; returns their number (the caller merges the columns
    // the node bounds fix to 1 and forms the sums)
    const int b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int n = a.blockColStart[b + 1] - cs;
    int cnt = 0;
    if (code == 1) {
        for (int base = 0; base < n; base += 32) {
            const int j = base + lane;
            const bool act = j < n && a.redCost[cs + j] < 0.0 && (a.fix == nullptr || a.fix[cs + j] == 0);
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (act) a.candInd[cs + cnt + __popc(bal & ((1u << lane) - 1u))] = cs + j;
            cnt += __popc(bal);
        }
    } else {
        for (int k = 0; k < 2; ++k) {
            const int j = a.sel[2 * lb + k];
            if (j >= 0) {
                if (lane == 0) a.candInd[cs + cnt] = cs + j;
                ++cnt;
            }
        }
    }
    __syncwarp();
    return cnt;
}

// nItems = number of active items of the block.  Loads are L2-coherent (ld.global.cg): in the
// fused path the decision bits were written earlier by this very kernel.
__device__ __forceinline__ void backtrackOne(const BackArgs &a, int lb, int lane, int nItems)
{
    const int b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int RC = a.rowCells;
    const uint32_t *bitsBase = a.bits + a.bitsOff[lb];
    int j = a.capacity[b];
    if (j < 0) {
        // the columns fixed to 1 by the node bounds do not fit: no column for this block
        if (lane == 0) {
            a.blockRedCost[lb] = __longlong_as_double(0x7ff0000000000000LL);
            a.blockOrigCost[lb] = 0.0;
            a.blockNnz[lb] = 0;
        }
        return;
    }
    const int code = a.shortcut != nullptr ? a.shortcut[lb] : 0;
    int cnt = 0;
    if (code != 0) {
        // Pisinger-mode trivial case: the chosen columns are already ascending in candInd
        cnt = emitShortcut(a, lb, lane, code);
    } else if (a.bigLayout) {
        // rows in global memory (bigdp.cu): bit (item, cell) at word item * cellWords + cell / 32.  The lanes fetch the
        // 32 items below the current one at the current cell together, the walk lands on the next TAKEN item.
        const long long cellWords = ((long long)j + 1 + 31) >> 5;
        int i = nItems - 1;
        while (i >= 0) {
            const int it = i - lane;
            const bool taken = it >= 0 && ((__ldcg(bitsBase + (long long)it * cellWords + (j >> 5)) >> (j & 31)) & 1u);
            const unsigned nz = __ballot_sync(0xffffffffu, taken);
            if (nz == 0u) {
                i -= 32;
                continue;
            }
            const int item = i - (__ffs(nz) - 1);
            if (lane == 0) a.scratch[cs + cnt] = __ldcg(a.itemCol + cs + item);
            ++cnt;
            j -= __ldcg(a.itemW + cs + item);
            i = item - 1;
        }
        __syncwarp();
        if (a.fix != nullptr)
            for (int k = lane; k < cnt; k += 32) a.candInd[cs + k] = cs + __ldcg(a.scratch + cs + cnt - 1 - k);
        __syncwarp();
    } else {
        int i = nItems - 1;
        while (i >= 0) {
            const int kw = i >> 5;
            const int wv = kw - lane;
            uint32_t v = wv >= 0 ? __ldcg(bitsBase + (size_t)wv * RC + j) : 0u;
            if (lane == 0) v &= 0xffffffffu >> (31 - (i & 31));
            const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
            if (nz == 0u) {
                i = ((kw - 32) << 5) | 31;  // nothing taken in these 1024 items
                continue;
            }
            const int L = __ffs(nz) - 1;
            const uint32_t vL = __shfl_sync(0xffffffffu, v, L);
            const int item = ((kw - L) << 5) + (31 - __clz(vL));
            if (lane == 0) a.scratch[cs + cnt] = __ldcg(a.itemCol + cs + item);
            ++cnt;
            j -= __ldcg(a.itemW + cs + item);
            i = item - 1;
        }
        __syncwarp();
        // chosen items ascending
        if (a.fix != nullptr)
            for (int k = lane; k < cnt; k += 32) a.candInd[cs + k] = cs + __ldcg(a.scratch + cs + cnt - 1 - k);
        __syncwarp();
    }
    if (a.fix != nullptr) {
        // node bounds: the columns fixed to 1 join the column (they were kept out of the DP and their
        // weight out of the capacity): a backward merge of the chosen columns (ascending in candInd) with the
        // fixed ones (ascending by construction) -- large shards only, the fused kernel merges in parallel.
        const int n = a.blockColStart[b + 1] - cs;
        int nf = 0;
        for (int base = 0; base < n; base += 32) {
            const int q = base + lane;
            const bool f = q < n && (a.fix[cs + q] & 1u);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (f) a.scratch[cs + nf + __popc(bal & ((1u << lane) - 1u))] = cs + q;
            nf += __popc(bal);
        }
        __syncwarp();
        if (lane == 0 && nf > 0) {
            __threadfence_block();
            int x = cnt - 1, y = nf - 1, o = cnt + nf - 1;
            while (y >= 0) {
                const int fy = __ldcg(a.scratch + cs + y);
                if (x >= 0 && __ldcg(a.candInd + cs + x) > fy) a.candInd[cs + o--] = __ldcg(a.candInd + cs + x--);
                else { a.candInd[cs + o--] = fy; --y; }
            }
        }
        cnt += nf;
        __syncwarp();
    }
    // varRedCost / varOrigCost in ascending index order (GAP_KnapPisinger.h:263-272), then alpha (DecompAlgo.cpp:6350-6356)
    const bool inCand = a.fix != nullptr || code != 0;  // else the walk's descending list is still in scratch
    double rcs = 0.0, ocs = 0.0;
    for (int base = 0; base < cnt; base += 32) {
        const int k = base + lane;
        double r = 0.0, o = 0.0;
        if (k < cnt) {
            int col;
            if (inCand) {
                col = __ldcg(a.candInd + cs + k);
            } else {
                col = cs + __ldcg(a.scratch + cs + cnt - 1 - k);
                a.candInd[cs + k] = col;
            }
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

// Tail shared by the backtrack kernel and the fused DP kernel: every CTA publishes its blocks
// (fence + atomic ticket); the CTA that draws the last ticket runs the filter of stage (c) over
// the whole shard, saving a launch on the latency-bound GAP rounds.
__device__ __forceinline__ void ticketThenFilter(unsigned int *doneCounter, const FilterArgs &f, int nThreads)
{
    __shared__ int sLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(doneCounter, 1u);
        sLast = (ticket == gridDim.x - 1) ? 1 : 0;
        if (sLast) *doneCounter = 0u;  // ready for the next round (same stream)
    }
    __syncthreads();
    if (!sLast) return;
    __threadfence();
    filterScanBody(f, nThreads);
}

__global__ void __launch_bounds__(kBackThreads) knap_backtrack_kernel(const __grid_constant__ BackArgs a, const __grid_constant__ FilterArgs f)
{
    const int lane = threadIdx.x & 31;
    const int lb = blockIdx.x * (kBackThreads >> 5) + (threadIdx.x >> 5);
    if (lb < a.nLocal) backtrackOne(a, lb, lane, a.nItems[lb]);
    if (a.doneCounter != nullptr) ticketThenFilter(a.doneCounter, f, kBackThreads);
}

// ------------------------------------------------------------------- DP kernel

__device__ __forceinline__ void dbgStamp(unsigned long long *buf, int slot, int lb, int vec)
{
    if (buf != nullptr && threadIdx.x == 0 && vec == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        buf[(size_t)lb * 8 + slot] = t;
    }
}

// Publication word of a fused block: valid (1) | epoch (26) | kept (1) | items (12) | evaluated items (12) | nnz (12).
// In the fused shapes a block has <= 2048 columns.  `valid` keeps a never-written (zero) word from matching
// any epoch; the host clears the words every 2^24 launches (dipgpu_api.cu: pubEpochGuard), so a word a
// launch did not rewrite (a vector beyond the launch's nVec) can never carry the current epoch again.
constexpr unsigned int kPubEpochMask = (1u << 26) - 1u;
constexpr int kPubEpochShift = 37, kPubKeptBit = 36, kPubItemsShift = 24, kPubEvalShift = 12;

#define kNegInf (__longlong_as_double((long long)0xfff0000000000000ULL))
constexpr int kPruneBuckets = 1024;   // efficiency histogram of the reduction pre-pass
constexpr int kPruneMinItems = 256;   // (columns of the block) below this the DP is about as short as the pre-pass
constexpr int kPruneTopFill = 2;      // pass 0 runs over the most efficient items worth this many capacities of weight
constexpr int kSumsRcOff = 256, kSumsObjOff = 512;  // doubles into itP: prefetched reduced costs / objective entries
constexpr int kPass0Words = 7;        // ... and over at most 32 * kPass0Words (its decision words sit in front of the raw arrays)
constexpr size_t kPruneScratchBytes = sizeof(int) * 3 * kPruneBuckets + sizeof(double) * 6 + 16 + 64;


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
    int chunkCols;                 // columns staged per TMA chunk (multiple of 4)
    int guard;                     // -inf cells in front of each row buffer (even; 0 = predicated reads)
    int profitMode;                // DIPGPU_PROFIT_*
    const int32_t *scale;          // Pisinger mode: local block -> scale factor
    const int32_t *shortcut;       // Pisinger mode: local block -> trivial-case code (0 = run the DP)
    const double *zShort;          // Pisinger mode: local block -> optimum of a trivial case
    unsigned long long *dbgTimes;  // diagnostics: 8 globaltimer stamps per CTA (NULL = off)
    unsigned long long *pubWord;   // FUSE: per (vector, local block) publication word (layout: kPub* below)
    unsigned long long *pubCells;  // FUSE: per local block DP cells
    // FUSE: where a CTA waits for other CTAs (the publish/poll tail), the launch guarantees that those CTAs run:
    //   * ticket == NULL: a COOPERATIVE launch -- the whole grid is co-resident (the runtime refuses the launch
    //     otherwise), CTA (blockIdx.x, blockIdx.y) is (block, vector);
    //   * ticket != NULL (grids too large to be co-resident): every CTA draws a ticket on entry and takes it as
    //     its LOGICAL (vector, block) id, so it only ever waits for CTAs that started before it, whatever the
    //     order the hardware dispatches blockIdx in.  The CTA that draws the last ticket resets the counter.
    // epochCtr: the launch epoch the publication words carry, read at kernel start; the last CTA to EXIT
    // (doneCtr) advances it.  Nothing is counted or reset by the host: a captured graph can replay the launch.
    unsigned int *ticket;
    unsigned int *epochCtr, *doneCtr;
    // FUSE, reduced costs formed by THIS kernel (rcIdx != NULL; DecompAlgo::generateVarsCalcRedCost,
    // Dip/src/DecompAlgo.cpp:4506-4547): column j = obj[j] - sum_e uVec[rcIdx[e]] * rcVal[e] over the entries
    // e in [rcStart[j], rcStart[j+1]) -- stored in the reference's scatter order (descending rows), multiplied
    // and added unfused, one after the other: the bits of redcost_stream_kernel and of the oracle.  (rcIdx, uVec)
    // is either (master row, the dense dual vector) or (slot, u[support] compacted): rows that cannot carry a
    // dual are not stored in the second form.  The reduced costs are also written to `redCost` (the tail's
    // sums and the host read them there).  alpha_b = uVec[alphaIdx[b]].
    const int32_t *rcStart;
    const int32_t *rcIdx;
    const double *rcVal;
    const double *uVec;
    const int32_t *alphaIdx;
    long long uStride;
    int phase1;
    // FUSE, cooperative launches only (pullSrc != NULL): the grid itself brings u[support] in from HOST memory.  The
    // CTAs of a vector each fetch a slice of the support -- pullIdx == NULL: pullSrc is staging that already holds
    // u[support] contiguously; else pullSrc is the caller's whole (page-locked) dual vector and pullIdx the support's
    // rows -- into uVec (vector v: source + v*pullStride), then meet at a per-vector counter (pullCtr[v], reset by the
    // last CTA to exit) before anybody gathers duals.  No upload in front of the kernel, no host pass over the vector.
    int vecBase;  // first vector of this launch (a batch priced in chunks: vector = vecBase + the CTA's vector coordinate)
    const double *pullSrc;
    const int32_t *pullIdx;
    long long pullStride;
    int pullN;
    unsigned int *pullCtr;
    // FUSE, batched dual vectors (gridDim.y = vectors): vector v = blockIdx.y reads redCost + v*rcStride
    // and alpha + v*alphaStride and writes the packed result at byte offset v*resStride; the per-block
    // arrays pubWord / scale / shortcut / sel / zShort hold nLocal entries per vector
    long long rcStride, alphaStride, resStride;
    int prune;                     // FUSE: drop items that cannot be in an optimal solution before the DP
    int pdl;                       // FUSE: launched with programmatic stream serialization (griddepcontrol.wait before the reduced costs are read)
    BackArgs back;
    FilterArgs filt;
    // FUSE, single vector: byte distance from the packed result in HBM to its mirror in mapped pinned host
    // memory (0: none).  Every result store is repeated there, so the host needs no device-to-host copy after
    // the kernel -- the stores cross PCIe while the other blocks still run.
    long long hostOff;
    // non-NULL: a word of mapped host memory that receives doneVal once EVERY CTA's stores (device and host mirror)
    // are visible system-wide -- the host waits on it instead of synchronising the stream
    unsigned int *doneFlag;
    unsigned int doneVal;
    // FUSE (<= kFuseFilterMaxBlocks blocks): per LOCAL block {first column, capacity, first stored entry of its
    // columns in the CSC of the in-block reduced costs}, entry nLocal closes the ranges.  In the kernel
    // parameters: a CTA knows its block without a (cold) global load in front of everything else.
    int32_t tab[kFuseFilterMaxBlocks + 1][3];
    int useTab;  // 0: more blocks than the table holds (a forced fused launch): the global arrays are read instead
};

// Fused shapes: the TMA destinations of the raw reduced costs / weights are dead once the items are
// compacted, so they share their bytes with the decision bits (written by the DP afterwards), and the
// taken list of the backtrack lives in the compacted-profit array (dead after the DP): ~26 KB less
// per CTA on GAP-E, i.e. three resident CTAs per SM instead of two when dual vectors are batched.
__host__ __device__ inline bool dpPruneLayout(int chunkCols, bool fuse, bool prune)
{
    // the reduction pre-pass needs its histograms to fit the (not yet written) compacted-item arrays
    return fuse && prune && (size_t)(chunkCols + 4) * 12 >= kPruneScratchBytes;
}

__host__ __device__ inline size_t dpSmemBytes(int rowCells, int guard, int chunkCols, bool fuse, bool prune)
{
    size_t b = 0;
    b += sizeof(double) * 2 * ((size_t)rowCells + (size_t)guard);  // two row buffers, each behind its guard
    b += sizeof(double) * ((size_t)chunkCols + 4);      // compacted profits (+ prefetch pad)
    b += sizeof(int32_t) * ((size_t)chunkCols + 4);     // compacted weights (+ prefetch pad)
    b += sizeof(int32_t) * 32;                          // per-warp counts
    b += 16;                                            // mbarrier
    size_t raw = sizeof(double) * ((size_t)chunkCols + 2) + sizeof(int32_t) * ((size_t)chunkCols + 4);
    if (fuse) {
        b += sizeof(int32_t) * ((size_t)chunkCols + 4);                                          // item columns
        const size_t bits = sizeof(uint32_t) * (size_t)((chunkCols + 31) / 32) * (size_t)rowCells;  // decision bits
        // with the reduction, the raw arrays must outlive pass 0's decision words (its survivor test and
        // pass 1 read them): they sit behind the first kPass0Words words instead of at the region's start
        if (dpPruneLayout(chunkCols, fuse, prune)) raw += sizeof(uint32_t) * (size_t)kPass0Words * (size_t)rowCells;
        b += bits > raw ? bits : raw;
    } else {
        b += raw;
    }
    return b;
}

// In-block reduced costs (fused shapes): how many stored entries of a block fit the staging area of their values
// and products, the not yet written itP | itW.  The host asks for in-block reduced costs only when every block of
// the shard fits.
__host__ __device__ inline int rcStageCap(int chunkCols)
{
    return (int)(((size_t)(chunkCols + 4) * 12) / 8) - 2;
}

int rcFuseStageCap(const Ctx *c)
{
    return rcStageCap(c->geom.chunkCols);
}

// One CTA per knapsack; T = blockDim.x threads (a multiple of 32), thread t owns
// the S consecutive cells [t*S, t*S+S) of the DP row in registers.  S is odd, so
// the 64-bit shared-memory accesses of a half-warp (stride S doubles) hit 16
// distinct bank pairs.  GUARD: each row buffer is preceded by `guard` cells of
// -inf so that the shifted read c[i-1][j-w] needs no j >= w predicate.
// ITEMS == 2 advances two items per barrier (latency-bound shapes): the
// intermediate row c[i][.] is rebuilt in registers at j and j-w' from three
// shifted reads of c[i-1][.], with exactly the additions and strict compares
// the one-item step performs, so values and decisions are unchanged.
// FUSE (small shards, every block a single chunk): the decision bits, the compacted items and
// the taken list stay in shared memory, warp 0 walks the decisions as soon as the DP is done,
// and the last CTA of the grid runs the filter -- stages (b) and (c) in one launch.
template <int S, bool GUARD, int ITEMS, int MAXT, bool FUSE>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256 ? 2 : 1)) knap_dp_kernel(const __grid_constant__ KnapArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int T = blockDim.x;
    const int RC = T * S;
    const int G = a.guard;
    const int CH = a.chunkCols;
    double *rowBase = reinterpret_cast<double *>(smemRaw);
    const int bufStride = RC + G;
    // non-fused: rows | rawRc | itP | rawW | itW | counts | mbarrier
    // fused:     rows | itP | itW | counts | mbarrier | itC | { rawRc, rawW } = decision bits
    // (the TMA destinations are dead once the items are compacted, the bits are written afterwards)
    double *rawRc = FUSE ? nullptr : rowBase + 2 * bufStride;
    double *itP = FUSE ? rowBase + 2 * bufStride : rawRc + (CH + 2);
    int32_t *rawW = FUSE ? nullptr : reinterpret_cast<int32_t *>(itP + (CH + 4));
    int32_t *itW = FUSE ? reinterpret_cast<int32_t *>(itP + (CH + 4)) : rawW + (CH + 4);
    int32_t *warpCnt = itW + (CH + 4);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(warpCnt + 32);
    int32_t *itC = reinterpret_cast<int32_t *>(mbar + 2);                       // FUSE only
    uint32_t *sBits = reinterpret_cast<uint32_t *>(itC + (CH + 4));             // FUSE only
    const bool pruneLayout = FUSE && dpPruneLayout(CH, true, a.prune != 0);
    if (FUSE) {
        rawRc = reinterpret_cast<double *>(sBits + (pruneLayout ? (size_t)kPass0Words * RC : 0));
        rawW = reinterpret_cast<int32_t *>(rawRc + (CH + 2));
    }
    int32_t *sTaken = reinterpret_cast<int32_t *>(itP);                         // FUSE only: itP is dead after the DP

    const int t = threadIdx.x;
    const int lane = t & 31;
    const int warp = t >> 5;
    const int nWarps = T >> 5;
    // FUSE: the CTA's (block, vector): its grid coordinates (cooperative launch), or the ticket it draws
    __shared__ unsigned int sTicket;
    unsigned int tkIdx = 0u;
    if (FUSE && a.ticket != nullptr) {
        if (t == 0) {
            sTicket = atomicAdd(a.ticket, 1u);
            if (sTicket == gridDim.x * gridDim.y - 1u) atomicExch(a.ticket, 0u);  // every ticket is drawn: next launch
        }
        __syncthreads();
        tkIdx = sTicket;
    }
    const bool ticketed = FUSE && a.ticket != nullptr;
    const int lb = ticketed ? (int)(tkIdx % gridDim.x) : (int)blockIdx.x;
    const int b = a.firstBlock + lb;
    // batched dual vectors (fused shapes only; gridDim.y == 1 otherwise): this CTA's vector
    const long long vec = FUSE ? (long long)a.vecBase + (ticketed ? (long long)(tkIdx / gridDim.x) : (long long)blockIdx.y) : 0;
    const int vlb = (int)vec * (int)gridDim.x + lb;  // slot of (vector, block) in the per-vector block arrays
    // the launch epoch: needed at the publication only, the load travels under everything before it
    unsigned int epochRaw = 0u;
    if (FUSE) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(epochRaw) : "l"(a.epochCtr) : "memory");
    const bool rcFused = FUSE && a.rcIdx != nullptr;
    const double *redCostV = a.redCost + vec * a.rcStride;
    const long long resOff = vec * a.resStride;
#define RESV(ptr) (reinterpret_cast<decltype(ptr)>(reinterpret_cast<char *>(const_cast<void *>(static_cast<const void *>(ptr))) + resOff))
    // a result store: the packed result in HBM and, when asked for, its host mirror
#define RES_PUT(ptr, idx, val)                                                                                  \
    do {                                                                                                        \
        auto *p__ = RESV(ptr) + (idx);                                                                          \
        const auto v__ = (val);                                                                                 \
        *p__ = v__;                                                                                             \
        if (FUSE && a.hostOff != 0) *reinterpret_cast<decltype(p__)>(reinterpret_cast<char *>(p__) + a.hostOff) = v__; \
    } while (0)
    const bool tab = FUSE && a.useTab != 0;
    const int colStart = tab ? a.tab[lb][0] : a.blockColStart[b];
    const int colEnd = tab ? a.tab[lb + 1][0] : a.blockColStart[b + 1];
    const int rcE0 = rcFused ? a.tab[lb][2] : 0, rcE1 = rcFused ? a.tab[lb + 1][2] : 0;  // the block's run of stored entries
    const int shortCode = a.profitMode == DIPGPU_PROFIT_PISINGER_INT ? a.shortcut[vlb] : 0;
    // a trivial case of GAP_KnapPisinger::solve needs no DP: behave like an empty block here
    const int capRaw = tab ? a.tab[lb][1] : a.capacity[b];  // < 0: the columns fixed to 1 do not fit (no column for this block)
    const int nCols = (shortCode != 0 || capRaw < 0) ? 0 : colEnd - colStart;
    const double pScale = a.profitMode == DIPGPU_PROFIT_PISINGER_INT ? (double)a.scale[vlb] : 0.0;
    const int cap = capRaw;
    // fetched now so that the (possibly cold) load is long done when the tail needs it
    const bool pulls = rcFused && a.pullSrc != nullptr;
    double alphaB = !FUSE ? 0.0 : rcFused ? (pulls ? 0.0 : a.uVec[vec * a.uStride + a.alphaIdx[b]]) : a.back.alpha[vec * a.alphaStride + b];
    uint32_t *bitsBase = FUSE ? sBits : a.bits + a.bitsOff[lb];
    const int j0 = t * S;  // first owned cell

    dbgStamp(a.dbgTimes, 0, lb, (int)vec);
    unsigned long long touch = 0ull;
    if (FUSE && t == 32 % T) {
        // touch, now, the pages the tail will write and poll (publication words, packed result): a first
        // access from this SM otherwise pays its translation miss (~4 us, seen on a random third of the
        // blocks) on the critical path of the tail instead of under the DP
        unsigned long long x0, x1, x2;
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x0) : "l"(a.pubWord + (size_t)vlb * kPubWordStride) : "memory");
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x1) : "l"(RESV(a.filt.hdr)) : "memory");
        asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x2) : "l"(RESV(a.filt.keptInd) + (size_t)(colStart - a.blockColStart[a.firstBlock]) / 2 * 2) : "memory");
        touch = x0 ^ x1 ^ x2;  // consumed in the tail, so that nothing waits for these loads here
        if (rcFused && !pulls) {  // ... and the first page of the dual vector, which the gathers of the reduced costs read
            unsigned long long x3;
            asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x3) : "l"(a.uVec + vec * a.uStride) : "memory");
            touch ^= x3;
        }
    }
    if (t == 0) mbarInit(mbar, 1);
    // ---- in-block reduced costs (stage (a) by the block itself, see the chunk loop): everything is requested NOW --
    // the bulk copies of the weights, the objective entries (into rawRc: c_j - sum is formed in place), the columns'
    // entry offsets (into the not yet written item-column array) and the entries' values (into the not yet written
    // item arrays), and, by every thread, the dual indices of ITS entries straight from memory followed by the
    // gathers of those duals -- so the two dependent round trips of the gathers run under the bulk copies.
    // entries per thread whose duals are gathered up front (GAP-E: 1 632 entries, 192 threads); the 1024-thread
    // build (64 registers; picked for large batches, where occupancy counts) gathers them after the bulk copies
    constexpr int kRcRegs = MAXT <= 256 ? 9 : 1;
    constexpr int kRcPre = MAXT <= 256 ? 9 : 0;
    const int rcLen = FUSE ? nCols : 0;
    const int rcNE = rcE1 - rcE0;
    const bool rcStaged = rcFused && rcLen > 0 && rcNE <= rcStageCap(CH);
    const double *uV = rcFused ? a.uVec + vec * a.uStride : nullptr;
    double rcG[kRcRegs];
    if (rcFused && rcLen > 0) {
        if (t == 0) {
            const int shRc = colStart & 1, shW = colStart & 3, shS = colStart & 3, shV = rcE0 & 1;
            const uint32_t bytesObj = a.phase1 ? 0u : (uint32_t)(((shRc + rcLen + 1) & ~1) * 8);
            const uint32_t bytesW = (uint32_t)(((shW + rcLen + 3) & ~3) * 4);
            const uint32_t bytesS = (uint32_t)(((shS + rcLen + 1 + 3) & ~3) * 4);
            const uint32_t bytesV = rcStaged ? (uint32_t)(((shV + rcNE + 1) & ~1) * 8) : 0u;
            mbarExpectTx(mbar, bytesObj + bytesW + bytesS + bytesV);
            tmaLoad1d(itC, a.rcStart + (colStart - shS), bytesS, mbar);
            tmaLoad1d(rawW, a.weight + (colStart - shW), bytesW, mbar);
            if (!a.phase1) tmaLoad1d(rawRc, a.back.obj + (colStart - shRc), bytesObj, mbar);
            if (bytesV) tmaLoad1d(itP, a.rcVal + (rcE0 - shV), bytesV, mbar);
        }
        int ix[kRcRegs];
        if (rcStaged) {
#pragma unroll
            for (int q = 0; q < kRcPre; ++q) {
                const int k = t + q * T;
                ix[q] = k < rcNE ? a.rcIdx[rcE0 + k] : -1;
            }
        }
        if (!pulls && rcStaged) {
#pragma unroll
            for (int q = 0; q < kRcPre; ++q) rcG[q] = ix[q] >= 0 ? uV[ix[q]] : 0.0;
        }
    }
    if (pulls) {
        // ---- this CTA's slice of u[support], straight out of host memory (every block of the vector takes part,
        // also an empty one), then the vector's CTAs meet: the grid is co-resident (cooperative launch)
        const int per = (a.pullN + (int)gridDim.x - 1) / (int)gridDim.x;
        const int k0 = min(a.pullN, lb * per), k1 = min(a.pullN, k0 + per);
        const double *src = a.pullSrc + vec * a.pullStride;
        double *dst = const_cast<double *>(a.uVec) + vec * a.uStride;
        // (four fetches in flight per thread: a slice of more than T entries -- few blocks per device -- must not pay
        // one PCIe / NVLink round trip per pass)
        for (int kb = k0 + t; kb < k1; kb += 4 * T) {
            int ix4[4];
            double v4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = kb + q * T;
                ix4[q] = k < k1 ? (a.pullIdx != nullptr ? a.pullIdx[k] : k) : -1;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) v4[q] = ix4[q] >= 0 ? src[ix4[q]] : 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (ix4[q] >= 0) dst[kb + q * T] = v4[q];
        }
        __threadfence();
        __syncthreads();
        if (t == 0) {
            atomicAdd(a.pullCtr + vec, 1u);
            while (atomicAdd(a.pullCtr + vec, 0u) < gridDim.x) {}
            __threadfence();
        }
        __syncthreads();
        alphaB = __ldcg(a.uVec + vec * a.uStride + a.alphaIdx[b]);
        if (rcStaged) {
            int ix2[kRcRegs];
#pragma unroll
            for (int q = 0; q < kRcPre; ++q) {
                const int k = t + q * T;
                ix2[q] = k < rcNE ? a.rcIdx[rcE0 + k] : -1;   // (L1-resident by now; keeps the indices out of the wait)
            }
#pragma unroll
            for (int q = 0; q < kRcPre; ++q) rcG[q] = ix2[q] >= 0 ? __ldcg(uV + ix2[q]) : 0.0;
        }
    }
    // c[-1][.] = 0 (gap_func.py:156); guards = -inf
    for (int j = t; j < 2 * bufStride; j += T) {
        const int r = j >= bufStride ? j - bufStride : j;
        rowBase[j] = r < G ? kNegInf : 0.0;
    }
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
    int nItRef = 0;     // FUSE: items the reference's DP runs over (before the reduction)
    bool smemSums = false;  // FUSE: the items' reduced costs / objective entries were prefetched into shared memory
    uint32_t parity = 0;

    for (int chunkBase = 0; chunkBase < nCols; chunkBase += CH) {
        const int len = min(CH, nCols - chunkBase);
        const int g0 = colStart + chunkBase;
        const int shiftRc = g0 & 1, shiftW = g0 & 3;
        // node bounds (DecompAlgo::setSubProbBounds): a column fixed to 0 or to 1 stays out of the DP
        // (fused shapes test the flags here; the large-shard path is fed reduced costs in which the fixed
        // columns are already +0, so that its hot kernels stay exactly as they are without bounds)
        const uint8_t *fixB = (FUSE && a.back.fix != nullptr) ? a.back.fix + g0 : nullptr;
        const bool pruneLayoutHere = FUSE && dpPruneLayout(CH, true, a.prune != 0) && len >= kPruneMinItems;
        if (pruneLayoutHere && !rcFused) {
            // the histograms of the reduction pre-pass (in the not yet written itP | itW), zeroed before the
            // reduced costs are there
            int *h0 = reinterpret_cast<int *>(itP);
            for (int k = t; k < 3 * kPruneBuckets; k += T) h0[k] = 0;
        }
        // Programmatic dependent launch: this kernel may have started while the kernel that writes the
        // reduced costs (the SpMV) is still running; everything above -- row buffers, histograms, page
        // touches, the convexity dual -- overlapped with it.  From here on its output is read.
        if (FUSE && a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (!rcFused) {
            // ---- stage the chunk's reduced costs and weights with TMA bulk copies
            if (t == 0) {
                const uint32_t bytesRc = (uint32_t)(((shiftRc + len + 1) & ~1) * 8);
                const uint32_t bytesW = (uint32_t)(((shiftW + len + 3) & ~3) * 4);
                mbarExpectTx(mbar, bytesRc + bytesW);
                tmaLoad1d(rawRc, redCostV + (g0 - shiftRc), bytesRc, mbar);
                tmaLoad1d(rawW, a.weight + (g0 - shiftW), bytesW, mbar);
            }
            if (pruneLayoutHere) __syncthreads();  // the zeroed histograms are visible to every warp
            mbarWait(mbar, parity);
            parity ^= 1u;
        } else {
            // ---- stage (a) for this block's columns, by the block itself (no SpMV launch, no round trip of the
            // reduced costs through HBM before they are used).  The block's stored entries are ONE contiguous run
            // [rcE0, rcE1) of the CSC; the bulk copies and the gathers of the duals were requested at the top of
            // the kernel.  Out of shared memory and registers now:
            //   1. flat over the entries: product u*a, unfused, written over the value;
            //   2. one thread per column adds ITS products one after the other in storage order (descending
            //      rows: CoinPackedMatrix::transposeTimes' scatter order) and subtracts from c_j.
            const int32_t *st = itC + (g0 & 3);  // st[cl] = first entry of column cl
            const int nE = rcNE;
            const bool staged = rcStaged;
            mbarWait(mbar, parity);
            parity ^= 1u;
            double *rcOut = const_cast<double *>(redCostV);
            if (staged) {
                double *vp = itP + (rcE0 & 1);
#pragma unroll
                for (int q = 0; q < kRcPre; ++q) {
                    const int k = t + q * T;
                    if (k < nE) vp[k] = __dmul_rn(rcG[q], vp[k]);
                }
                for (int k = t + kRcPre * T; k < nE; k += T) vp[k] = __dmul_rn(__ldcg(uV + a.rcIdx[rcE0 + k]), vp[k]);  // (longer blocks)
                __syncthreads();
                // four columns per thread and step: their (short) addition chains run side by side
                for (int base = t; base < len; base += 4 * T) {
                    int k0[4], k1[4];
                    double sum[4], cj[4];
                    int maxLen = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cl = base + q * T;
                        const bool in = cl < len;
                        k0[q] = in ? st[cl] - rcE0 : 0;
                        k1[q] = in ? st[cl + 1] - rcE0 : 0;
                        cj[q] = (in && !a.phase1) ? rawRc[shiftRc + cl] : 0.0;
                        sum[q] = 0.0;
                        maxLen = max(maxLen, k1[q] - k0[q]);
                    }
                    for (int k = 0; k < maxLen; ++k) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (k0[q] + k < k1[q]) sum[q] = __dadd_rn(sum[q], vp[k0[q] + k]);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cl = base + q * T;
                        if (cl < len) {
                            const double rc = a.phase1 ? -sum[q] : cj[q] - sum[q];  // DecompAlgo.cpp:4538-4545
                            rawRc[shiftRc + cl] = rc;
                            rcOut[g0 + cl] = rc;
                        }
                    }
                }
            } else {
                // (not taken when the host sized the call; kept correct: thread t forms columns t, t + T, ... four at a
                // time straight from memory)
                for (int base = t; base < len; base += 4 * T) {
                    int e0[4], e1[4];
                    double sum[4];
                    int maxLen = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cl = base + q * T;
                        const bool in = cl < len;
                        e0[q] = in ? st[cl] : 0;
                        e1[q] = in ? st[cl + 1] : 0;
                        sum[q] = 0.0;
                        maxLen = max(maxLen, e1[q] - e0[q]);
                    }
                    for (int k = 0; k < maxLen; ++k) {
                        int ix[4];
                        double v[4], g[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const bool on = e0[q] + k < e1[q];
                            ix[q] = on ? a.rcIdx[e0[q] + k] : -1;
                            v[q] = on ? a.rcVal[e0[q] + k] : 0.0;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) g[q] = ix[q] >= 0 ? __ldcg(uV + ix[q]) : 0.0;
                        // (a column past its end adds 0*0 = +0, which leaves the sum unchanged bit for bit: it is never -0)
#pragma unroll
                        for (int q = 0; q < 4; ++q) sum[q] = __dadd_rn(sum[q], __dmul_rn(g[q], v[q]));
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cl = base + q * T;
                        if (cl < len) {
                            const double rc = a.phase1 ? -sum[q] : rawRc[shiftRc + cl] - sum[q];
                            rawRc[shiftRc + cl] = rc;
                            rcOut[g0 + cl] = rc;
                        }
                    }
                }
            }
            __syncthreads();  // the reduced costs are visible to every warp, the staging areas are free again
            if (pruneLayoutHere) {
                int *h0 = reinterpret_cast<int *>(itP);
                for (int k = t; k < 3 * kPruneBuckets; k += T) h0[k] = 0;
                __syncthreads();
            }
        }
        dbgStamp(a.dbgTimes, 1, lb, (int)vec);

        // ---- reduction, fused shapes (what Pisinger's combo does before its DP): an item is left out
        // when NO optimal solution can contain it, so the DP over the remaining items -- same order, same
        // additions, same strict compares -- returns the same optimum bit for bit and the same column,
        // ties included (a dropped item could only ever have raised the losing side of a compare).
        //   upper bound, any multiplier L >= 0 (Lagrangian relaxation of the capacity row):
        //       z(x_j = 1) <= L*cap + sum_k max(0, p_k - L w_k) + min(0, p_j - L w_j)
        //     with L at both edges of the efficiency bucket where the greedy fill crosses the capacity;
        //   lower bound: the optimum of the DP over the ~64 most efficient items IN ITEM ORDER -- the
        //     value the full DP assigns to that very column, so it needs no rounding allowance.
        // Pass 0 runs the DP over those items; if no other item survives the bound (the usual case) its
        // decision table is the final one, else pass 1 runs the DP over the survivors.  An item is
        // dropped only if a bound is below the lower bound by a relative 1e-9, far more than the
        // rounding of the bound's own sums.
        auto isItem = [&](int cl) -> bool {
            return rawRc[shiftRc + cl] < 0.0 && rawW[shiftW + cl] <= cap && (fixB == nullptr || fixB[cl] == 0);
        };
        // profit: -redCost (gap_func.py:104), or its scaled integer value
        // floor(-rc*scale + 0.5) held exactly in fp64 (GAP_KnapPisinger.h:174-177)
        auto profitOf = [&](int cl) -> double {
            const double rc = rawRc[shiftRc + cl];
            return pScale != 0.0 ? (double)(long long)floor((-rc) * pScale + 0.5) : -rc;
        };
        bool prune = false;
        int thrBucket = 0;
        double lamLo = 0.0, lamHi = 0.0, uLo = 0.0, uHi = 0.0, cut = 0.0, slack = 0.0;
        // bucket of an item: monotone in its efficiency p/w -- the float's exponent and top five mantissa
        // bits = 32 buckets per octave over 32 octaves from 2^-12 (times the Pisinger scale); weight 0 =
        // infinite efficiency = top bucket.  No integer division, no min/max pass: with 1.5 warps per
        // scheduler every instruction of this pre-pass is paid at its full latency.
        const int bucket0 = ((127 - 12) << 5) + (pScale >= 1000.0 ? (10 << 5) : pScale >= 10.0 ? (3 << 5) : 0);
        auto bucketOf = [&](int cl) -> int {
            const int w = rawW[shiftW + cl];
            if (w <= 0) return kPruneBuckets - 1;
            const int q = (int)(__float_as_uint(__fdividef((float)profitOf(cl), (float)w)) >> 18) - bucket0;
            return min(max(q, 0), kPruneBuckets - 1);
        };
        // per column, filled by the histogram pass: its bucket, or 0xffff when it is no item.  Lives in the
        // UPPER half of the item-column array, which pass 0 (<= 32*kPass0Words items) never reaches.
        uint16_t *meta = reinterpret_cast<uint16_t *>(itC) + (CH + 4);
        if (pruneLayoutHere) {
            int *histW = reinterpret_cast<int *>(itP);          // itP | itW are contiguous and not written yet
            int *histLo = histW + kPruneBuckets;                // profit sums, fixed point, low 20 bits ...
            int *histHi = histLo + kPruneBuckets;               // ... and the bits above
            double *sBound = reinterpret_cast<double *>(histHi + kPruneBuckets);  // lamLo, lamHi, uLo, uHi, sum of all profits
            int *sCross = reinterpret_cast<int *>(sBound + 6);                    // crossing bucket, threshold bucket
            // profits in units of 2^-12 (integers in Pisinger mode), rounded UP: the bound stays an upper
            // bound, and the integer sums do not depend on the order of the atomics
            const double fix = pScale != 0.0 ? 1.0 : 4096.0;
            bool tooBig = false;
            // four columns per thread and step, their loads issued together: with 1.5 warps per scheduler a
            // column's chain (load, compare, convert, divide, convert, three atomics) is paid at full latency,
            // so independent chains are interleaved by hand
            for (int base = t; base < len; base += 4 * T) {
                double rcv[4];
                int wv[4];
                bool it[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cl = base + q * T;
                    const bool in = cl < len;
                    rcv[q] = in ? rawRc[shiftRc + cl] : 0.0;
                    wv[q] = in ? rawW[shiftW + cl] : 0;
                    it[q] = in && rcv[q] < 0.0 && wv[q] <= cap && (fixB == nullptr || fixB[cl] == 0);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cl = base + q * T;
                    unsigned int m = 0xffffu;
                    if (it[q]) {
                        const double pk = pScale != 0.0 ? (double)(long long)floor((-rcv[q]) * pScale + 0.5) : -rcv[q];
                        tooBig |= !(pk < 268435456.0);
                        int bkt = kPruneBuckets - 1;
                        if (wv[q] > 0) {
                            const int qb = (int)(__float_as_uint(__fdividef((float)pk, (float)wv[q])) >> 18) - bucket0;
                            bkt = min(max(qb, 0), kPruneBuckets - 1);
                        }
                        const unsigned long long v = (unsigned long long)ceil(pk * fix);
                        atomicAdd(&histW[bkt], wv[q]);
                        atomicAdd(&histLo[bkt], (int)(v & 0xfffffull));
                        atomicAdd(&histHi[bkt], (int)(v >> 20));
                        m = (unsigned int)bkt;
                    }
                    if (cl < len) meta[cl] = (uint16_t)m;
                }
            }
            const int anyBig = __syncthreads_or(tooBig ? 1 : 0);
            const double unfix = pScale != 0.0 ? 1.0 : 1.0 / 4096.0;
            {
                // From the most efficient bucket down (greedy fill).  With L = the lower edge of bucket c,
                //     U(L) = L*cap + sum_{buckets >= c} (p - L w)      (every item there has p/w >= L,
                // every item below has p/w < L), i.e. the bound needs only the buckets' sums.
                // Up to four warps scan: a thread sums its run of buckets, a shuffle suffix-scan and the
                // warps' totals give the sums above it, then it walks its run looking for the crossing.
                auto histP = [&](int bkt) -> unsigned long long {
                    return ((unsigned long long)(unsigned int)histHi[bkt] << 20) + (unsigned long long)(unsigned int)histLo[bkt];
                };
                const int nScanW = nWarps >= 4 ? 4 : nWarps >= 2 ? 2 : 1;
                const int perThr = kPruneBuckets / (32 * nScanW);
                long long *sScanW = reinterpret_cast<long long *>(sCross + 4);
                unsigned long long *sScanP = reinterpret_cast<unsigned long long *>(sScanW + 4);
                long long wSum = 0, aboveW = 0;
                unsigned long long pSum = 0ull, aboveP = 0ull;
                if (t == 0) {
                    sCross[0] = -1;  // everything fits: nothing to drop
                    sCross[1] = 0;   // less weight in all than pass 0 may take: every item
                }
                if (warp < nScanW) {
                    for (int q = 0; q < perThr; ++q) {
                        wSum += histW[t * perThr + q];
                        pSum += histP(t * perThr + q);
                    }
                    long long sw = wSum;
                    unsigned long long sp = pSum;
#pragma unroll
                    for (int off2 = 1; off2 < 32; off2 <<= 1) {
                        const long long ow = __shfl_down_sync(0xffffffffu, sw, off2);
                        const unsigned long long op = __shfl_down_sync(0xffffffffu, sp, off2);
                        if (lane + off2 < 32) {
                            sw += ow;
                            sp += op;
                        }
                    }
                    if (lane == 0) {
                        sScanW[warp] = sw;
                        sScanP[warp] = sp;
                    }
                    aboveW = sw - wSum;
                    aboveP = sp - pSum;
                }
                __syncthreads();
                if (warp < nScanW) {
                    for (int v = warp + 1; v < nScanW; ++v) {
                        aboveW += sScanW[v];
                        aboveP += sScanP[v];
                    }
                    auto edgeOf = [&](int bkt) -> double {  // the smallest efficiency that falls in bucket bkt
                        return bkt <= 0 ? 0.0 : (double)__uint_as_float((unsigned int)(bkt + bucket0) << 18);
                    };
                    long long cumW = aboveW;
                    unsigned long long cumP = aboveP;
                    for (int q = perThr - 1; q >= 0; --q) {
                        const int bkt = t * perThr + q;
                        const long long beforeW = cumW;
                        const unsigned long long beforeP = cumP;
                        cumW += histW[bkt];
                        cumP += histP(bkt);
                        if (cumW > (long long)cap && beforeW <= (long long)cap) {
                            const double lLo = edgeOf(bkt), lHi = edgeOf(bkt + 1);
                            sBound[0] = lLo;
                            sBound[1] = lHi;
                            sBound[2] = lLo * (double)cap + ((double)cumP * unfix - lLo * (double)cumW);
                            sBound[3] = lHi * (double)cap + ((double)beforeP * unfix - lHi * (double)beforeW);
                            sCross[0] = bkt;
                        }
                        // pass 0 takes the buckets down to kPruneTopFill capacities' worth of weight
                        if (cumW >= kPruneTopFill * (long long)cap && beforeW < kPruneTopFill * (long long)cap) sCross[1] = bkt;
                    }
                    if (t == 0) sBound[4] = (double)cumP * unfix;  // thread 0 ends at bucket 0: every item
                }
            }
            __syncthreads();
            const int cross = sCross[0];
            if (cross >= 0 && !anyBig) {
                thrBucket = min(cross, sCross[1]);
                lamLo = sBound[0];
                lamHi = sBound[1];
                uLo = sBound[2];
                uHi = sBound[3];
                // an item whose efficiency is within float rounding (2^-21 relative) of a bucket edge may sit on
                // the wrong side of it; each such item moves a bound by at most 2^-21 of its profit
                slack = 2e-6 * sBound[4];
                prune = true;
            }
            __syncthreads();  // the scratch is consumed: itP / itW may be written now
        }
        auto survives = [&](int cl) -> bool {  // valid once `cut` is set from a lower bound
            const double pk = profitOf(cl), wk = (double)rawW[shiftW + cl];
            const double dLo = pk - lamLo * wk, dHi = pk - lamHi * wk;
            return !((dLo < 0.0 && uLo + dLo < cut) || (dHi < 0.0 && uHi + dHi < cut));
        };
        constexpr int kPasses = FUSE ? 2 : 1;
        int nIt = 0;
        for (int pass = 0; pass < kPasses; ++pass) {
        // ---- order-preserving compaction of the active items (gap_func.py:102-105): thread t owns
        // the contiguous run [t*per, (t+1)*per) of the chunk's columns (thread order == item order),
        // counts its active items, one block-wide exclusive scan places them.
        nIt = 0;
        if (FUSE) {
            auto takes = [&](int cl) -> bool {
                if (prune && pass == 0) {
                    const int m = meta[cl];  // bucket, 0xffff = no item
                    return m != 0xffff && m >= thrBucket;
                }
                if (!isItem(cl)) return false;
                return prune ? survives(cl) : true;
            };
            const int per = (len + T - 1) / T;
            const int c0 = min(len, t * per), c1 = min(len, c0 + per);
            int mineCnt = 0, mineRef = 0;
            // (runs of up to 32 columns: the taken columns are remembered as a bit mask, so that the write pass
            // below visits only them instead of testing the whole run again)
            unsigned tkMask = 0u;
            const bool useMask = per <= 32;
#pragma unroll 4
            for (int cl = c0; cl < c1; ++cl) {
                mineRef += ((prune && pass == 0) ? meta[cl] != 0xffff : isItem(cl)) ? 1 : 0;
                const bool tk = takes(cl);
                mineCnt += tk ? 1 : 0;
                tkMask |= (tk && useMask ? 1u : 0u) << ((cl - c0) & 31);
            }
            // warp-inclusive scan, then the warps' totals (low half: taken items, high half: all items)
            int incl = mineCnt | (mineRef << 16);
#pragma unroll
            for (int off2 = 1; off2 < 32; off2 <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, off2);
                if (lane >= off2) incl += up;
            }
            if (lane == 31) warpCnt[warp] = incl;
            __syncthreads();
            int off = 0;
            nItRef = 0;
            for (int v = 0; v < nWarps; ++v) {
                const int cv = warpCnt[v];
                off += v < warp ? (cv & 0xffff) : 0;
                nIt += cv & 0xffff;
                nItRef += cv >> 16;
            }
            if (prune && pass == 0 && nIt > 32 * kPass0Words) {
                // more items tie around the threshold than pass 0 has decision words for: plain DP over all
                prune = false;
                __syncthreads();
                continue;
            }
            // pass 0 (<= 32*kPass0Words items): the items' true reduced costs and objective entries are
            // fetched NOW into the unused upper part of itP (cp.async, consumed by the tail's sums), so the
            // tail does not wait for two dependent global gathers
            smemSums = prune && pass == 0 && a.back.fix == nullptr;
            int pos = off + (incl & 0xffff) - mineCnt;
            auto place = [&](int cl) {  // w > cap can never be taken (:161)
                itP[pos] = profitOf(cl);
                itW[pos] = rawW[shiftW + cl];
                itC[pos] = cl;
                if (smemSums) {
                    if (rcFused) itP[kSumsRcOff + pos] = rawRc[shiftRc + cl];
                    else
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(itP + kSumsRcOff + pos)),
                                     "l"(redCostV + colStart + cl)
                                     : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(itP + kSumsObjOff + pos)),
                                 "l"(a.back.obj + colStart + cl)
                                 : "memory");
                }
                ++pos;
            };
            if (useMask) {
                for (unsigned mk = tkMask; mk; mk &= mk - 1) place(c0 + __ffs(mk) - 1);  // ascending
            } else {
                for (int cl = c0; cl < c1; ++cl)
                    if (takes(cl)) place(cl);
            }
            if (smemSums) asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            // throughput shapes: strided ballot compaction (one column per thread and round)
            for (int base = 0; base < len; base += T) {
                const int cl = base + t;
                bool flag = false;
                double rc = 0.0;
                int w = 0;
                if (cl < len) {
                    rc = rawRc[shiftRc + cl];
                    w = rawW[shiftW + cl];
                    flag = (rc < 0.0) && (w <= cap) && (fixB == nullptr || fixB[cl] == 0);  // w > cap can never be taken (:161)
                }
                const unsigned bal = __ballot_sync(0xffffffffu, flag);
                const int wpre = __popc(bal & ((1u << lane) - 1u));
                int off = 0, tot = __popc(bal);
                if (nWarps > 1) {
                    if (lane == 0) warpCnt[warp] = tot;
                    __syncthreads();
                    tot = 0;
                    for (int v = 0; v < nWarps; ++v) {
                        const int cv = warpCnt[v];
                        off += v < warp ? cv : 0;
                        tot += cv;
                    }
                    __syncthreads();
                }
                if (flag) {
                    const int pos = nIt + off + wpre;
                    itP[pos] = pScale != 0.0 ? (double)(long long)floor((-rc) * pScale + 0.5) : -rc;
                    itW[pos] = w;
                    const int slot = colStart + kGlobal + pos;
                    a.itemW[slot] = w;
                    a.itemCol[slot] = chunkBase + cl;
                }
                nIt += tot;
            }
        }
        __syncthreads();
        dbgStamp(a.dbgTimes, 2, lb, (int)vec);

        // ---- the DP proper: items sequential, cells parallel.  (w0,p0),(w1,p1) hold
        // items i and i+1, fetched one step ahead (itP/itW are padded by four entries).
        int i = 0;
        if (ITEMS == 3 && S <= 3 && GUARD) {
            // THREE items per barrier (the fused GAP shapes): 7 shifted reads of c[i-1][.] rebuild, in
            // registers, c[i][.] at j, j-wB, j-wC, j-wC-wB and c[i+1][.] at j, j-wC -- every one with
            // the same single DADD and strict compare the one-item step would have applied to the
            // same operands, so values and decisions are unchanged while the barrier + shared-memory
            // round trip (52 of the 76 cycles of a one-item step, scripts/micro/lat.cu) is paid once
            // per three items.  A 32-item word = 10 triples + 1 pair = 11 steps (odd: `cur` flips).
            while (((kGlobal + i) & 31) == 0 && i + 32 <= nIt) {
                double *bufA = rowBase + cur * bufStride + G + j0;        // holds c[i-1][.]
                double *bufB = rowBase + (cur ^ 1) * bufStride + G + j0;
                int wA = itW[i], wB = itW[i + 1], wC = itW[i + 2];
                double pA = itP[i], pB = itP[i + 1], pC = itP[i + 2];
#pragma unroll
                for (int st = 0; st < 11; ++st) {
                    const int q = 3 * st;  // bit of the step's first item
                    const double *rdBuf = (st & 1) ? bufB : bufA;
                    double *wrBuf = (st & 1) ? bufA : bufB;
                    if (st < 10) {
                        const double *rA = rdBuf - wA, *rB = rdBuf - wB, *rC = rdBuf - wC;
                        const double *rAB = rA - wB, *rAC = rA - wC, *rBC = rB - wC, *rABC = rAB - wC;
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const double c0A = rA[s], c0B = rB[s], c0C = rC[s];
                            const double c0AB = rAB[s], c0AC = rAC[s], c0BC = rBC[s], c0ABC = rABC[s];
                            // item A (row i) at j, j-wB, j-wC, j-wC-wB
                            const double a0 = pA + c0A;
                            const bool takeA = a0 > mine[s];
                            const double c1 = takeA ? a0 : mine[s];
                            const double a1 = pA + c0AB;
                            const double c1B = a1 > c0B ? a1 : c0B;
                            const double a2 = pA + c0AC;
                            const double c1C = a2 > c0C ? a2 : c0C;
                            const double a3 = pA + c0ABC;
                            const double c1BC = a3 > c0BC ? a3 : c0BC;
                            // item B (row i+1) at j, j-wC
                            const double b0 = pB + c1B;
                            const bool takeB = b0 > c1;
                            const double c2 = takeB ? b0 : c1;
                            const double b1 = pB + c1BC;
                            const double c2C = b1 > c1C ? b1 : c1C;
                            // item C (row i+2) at j
                            const double d0 = pC + c2C;
                            const bool takeC = d0 > c2;
                            mine[s] = takeC ? d0 : c2;
                            if (takeA) acc[s] |= 1u << q;
                            if (takeB) acc[s] |= 2u << q;
                            if (takeC) acc[s] |= 4u << q;
                            wrBuf[s] = mine[s];
                        }
                    } else {  // the word's last two items
                        const double *rA = rdBuf - wA, *rB = rdBuf - wB;
                        const double *rAB = rA - wB;
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const double c0A = rA[s], c0B = rB[s], c0AB = rAB[s];
                            const double a0 = pA + c0A;
                            const bool takeA = a0 > mine[s];
                            const double c1 = takeA ? a0 : mine[s];
                            const double a1 = pA + c0AB;
                            const double c1B = a1 > c0B ? a1 : c0B;
                            const double b0 = pB + c1B;
                            const bool takeB = b0 > c1;
                            mine[s] = takeB ? b0 : c1;
                            if (takeA) acc[s] |= 1u << q;
                            if (takeB) acc[s] |= 2u << q;
                            wrBuf[s] = mine[s];
                        }
                    }
                    // the next step's items, fetched before the barrier
                    if (st < 10) {
                        wA = itW[i + q + 3];
                        pA = itP[i + q + 3];
                        wB = itW[i + q + 4];
                        pB = itP[i + q + 4];
                        if (st < 9) {
                            wC = itW[i + q + 5];
                            pC = itP[i + q + 5];
                        }
                    }
                    __syncthreads();
                }
                {
                    uint32_t *dst = bitsBase + (size_t)((kGlobal + i) >> 5) * RC + j0;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        dst[s] = acc[s];
                        acc[s] = 0u;
                    }
                }
                i += 32;
                cur ^= 1;  // 11 steps
            }
        } else if (S <= 3 && GUARD) {
            // Latency-bound shapes (few cells per thread): whole 32-item words are processed by a
            // fully unrolled body -- constant bit masks, compile-time buffer roles, item data
            // fetched eight steps ahead -- so that a step is little more than its dependent chain
            // barrier -> shifted load -> DADD -> DSETP -> select -> store -> barrier.
            while (((kGlobal + i) & 31) == 0 && i + 32 <= nIt) {
                double *bufA = rowBase + cur * bufStride + G + j0;        // holds c[i-1][.]
                double *bufB = rowBase + (cur ^ 1) * bufStride + G + j0;
                int wq[8];
                double pq[8];
#pragma unroll
                for (int q = 0; q < 32; q += ITEMS) {
                    if ((q & 7) == 0) {
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            wq[r] = itW[i + q + r];
                            pq[r] = itP[i + q + r];
                        }
                    }
                    // ITEMS == 1: buffers alternate every step; ITEMS == 2: every pair
                    const bool even = ((q / ITEMS) & 1) == 0;
                    const double *rdBuf = even ? bufA : bufB;
                    double *wrBuf = even ? bufB : bufA;
                    if (ITEMS == 1) {
                        const int w = wq[q & 7];
                        const double p = pq[q & 7];
                        const double *rd = rdBuf - w;
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const double cand = p + rd[s];        // gap_func.py:164
                            const bool take = cand > mine[s];     // :166 strict
                            mine[s] = take ? cand : mine[s];
                            if (take) acc[s] |= 1u << q;          // added[i][j]  (:168)
                            wrBuf[s] = mine[s];
                        }
                    } else {
                        const int wA = wq[q & 7], wB = wq[(q & 7) + 1];
                        const double pA = pq[q & 7], pB = pq[(q & 7) + 1];
                        const double *rdA = rdBuf - wA;
                        const double *rdB = rdBuf - wB;
                        const double *rdAB = rdA - wB;
#pragma unroll
                        for (int s = 0; s < S; ++s) {
                            const double shA = rdA[s], shB = rdB[s], shAB = rdAB[s];
                            const double candA = pA + shA;
                            const bool takeA = candA > mine[s];
                            const double r1 = takeA ? candA : mine[s];
                            const double candAB = pA + shAB;
                            const double r1B = candAB > shB ? candAB : shB;
                            const double candB = pB + r1B;
                            const bool takeB = candB > r1;
                            mine[s] = takeB ? candB : r1;
                            if (takeA) acc[s] |= 1u << q;
                            if (takeB) acc[s] |= 2u << q;
                            wrBuf[s] = mine[s];
                        }
                    }
                    __syncthreads();
                }
                // 32 / ITEMS steps: an even number of buffer swaps, `cur` is unchanged
                {
                    uint32_t *dst = bitsBase + (size_t)((kGlobal + i) >> 5) * RC + j0;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        dst[s] = acc[s];
                        acc[s] = 0u;
                    }
                }
                i += 32;
            }
        }
        int w0 = itW[i], w1 = itW[i + 1];
        double p0 = itP[i], p1 = itP[i + 1];
        while (i < nIt) {
            const int bitpos = (kGlobal + i) & 31;
            const double *base = rowBase + cur * bufStride + G + j0;
            double *wr = rowBase + (cur ^ 1) * bufStride + G + j0;
            if (ITEMS >= 2 && i + 1 < nIt && (bitpos & 1) == 0) {
                const int wA = w0, wB = w1;
                const double pA = p0, pB = p1;
                const double *rdA = base - wA;
                const double *rdB = base - wB;
                const double *rdAB = rdA - wB;
                const uint32_t bmA = 1u << bitpos, bmB = 2u << bitpos;
                double shA[S], shB[S], shAB[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const int j = j0 + s;
                    shA[s] = (GUARD || j >= wA) ? rdA[s] : kNegInf;
                    shB[s] = (GUARD || j >= wB) ? rdB[s] : kNegInf;
                    shAB[s] = (GUARD || j >= wA + wB) ? rdAB[s] : kNegInf;
                }
                w0 = itW[i + 2];
                p0 = itP[i + 2];
                w1 = itW[i + 3];
                p1 = itP[i + 3];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    // item A at cell j  (gap_func.py:164-168)
                    const double candA = pA + shA[s];
                    const bool takeA = candA > mine[s];
                    const double r1 = takeA ? candA : mine[s];
                    // item A at cell j - wB, i.e. c[i][j-wB], which item B reads
                    const double candAB = pA + shAB[s];
                    const double r1B = candAB > shB[s] ? candAB : shB[s];
                    // item B at cell j
                    const double candB = pB + r1B;
                    const bool takeB = candB > r1;
                    mine[s] = takeB ? candB : r1;
                    acc[s] |= (takeA ? bmA : 0u) | (takeB ? bmB : 0u);
                    wr[s] = mine[s];
                }
                i += 2;
            } else {
                const int w = w0;
                const double p = p0;
                const double *rd = base - w;
                const uint32_t bm = 1u << bitpos;
                double sh[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    // c[i-1][j-w]; j < w keeps c[i-1][j]  (gap_func.py:161-162)
                    sh[s] = (GUARD || j0 + s >= w) ? rd[s] : kNegInf;
                }
                w0 = w1;
                p0 = p1;
                w1 = itW[i + 2];
                p1 = itP[i + 2];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const double cand = p + sh[s];        // :164
                    const bool take = cand > mine[s];     // :166 strict
                    mine[s] = take ? cand : mine[s];
                    acc[s] |= take ? bm : 0u;             // added[i][j]  (:168)
                    wr[s] = mine[s];
                }
                i += 1;
            }
            if (((kGlobal + i) & 31) == 0) {  // a 32-item word is complete
                uint32_t *dst = bitsBase + (size_t)((kGlobal + i - 1) >> 5) * RC + j0;
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    dst[s] = acc[s];
                    acc[s] = 0u;
                }
            }
            cur ^= 1;
            __syncthreads();
        }
        if (!FUSE || !prune || pass == 1) break;
        {
            // pass 0 is done: its optimum is the lower bound; does any item outside pass 0 survive it?
            double *sLB = reinterpret_cast<double *>(warpCnt);  // 32 ints: one double + the survivor counts
            int *sExtra = warpCnt;
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2)
                if (j0 + s2 == cap) *sLB = mine[s2];
            __syncthreads();
            const double lb = *sLB;
            cut = lb - 1e-9 * (fabs(lb) + fmax(uLo, uHi) + 1.0) - slack;  // (the bound's own sums are rounded UP)
            int extra = 0;
            for (int cl = t; cl < len; cl += T)
                if (meta[cl] < thrBucket && survives(cl)) ++extra;  // an item (0xffff is no bucket) outside pass 0
            extra = __reduce_add_sync(0xffffffffu, extra);
            __syncthreads();  // sLB is read by everyone before the counts overwrite it
            if (lane == 0) sExtra[warp] = extra;
            __syncthreads();
            extra = 0;
            for (int v = 0; v < nWarps; ++v) extra += sExtra[v];
            __syncthreads();
            if (extra == 0) break;  // no item outside pass 0 can be in an optimal solution
            smemSums = false;
            // pass 1 over the survivors: the DP starts again from c[-1][.] = 0
            for (int j = t; j < 2 * bufStride; j += T) {
                const int r = j >= bufStride ? j - bufStride : j;
                rowBase[j] = r < G ? kNegInf : 0.0;
            }
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) {
                mine[s2] = 0.0;
                acc[s2] = 0u;
            }
            cur = 0;
            __syncthreads();
        }
        }  // pass
        kGlobal += nIt;
    }

    if ((kGlobal & 31) != 0) {
        uint32_t *dst = bitsBase + (size_t)(kGlobal >> 5) * RC + j0;
#pragma unroll
        for (int s = 0; s < S; ++s) dst[s] = acc[s];
    }
    // z = c[n-1][capacity]  (gap_func.py:183)
    if (cap < 0) {
        if (t == 0) RES_PUT(a.blockObj, lb, kNegInf);
    } else if (shortCode != 0) {
        if (t == 0) RES_PUT(a.blockObj, lb, a.zShort[vlb]);
    } else {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (j0 + s == cap) RES_PUT(a.blockObj, lb, mine[s]);
        }
    }
    if (t == 0 && vec == 0) a.nItems[lb] = kGlobal;
    dbgStamp(a.dbgTimes, 3, lb, (int)vec);
    if (FUSE) {
        // ---- stage (b) continued: the column of this block, out of shared memory
        __shared__ int sCnt;
        __shared__ double sRcSum, sOcSum;
        int32_t *sFinal = itC;  // ascending x-space indices of the column (itC is free after the walk)
        int32_t *sTakenIdx = sTaken + 256;  // item of each taken entry (pass-0 shapes: <= 32*kPass0Words)
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (touch == 0x9e3779b97f4a7c15ull) sCnt = -1;  // never true in practice: keeps the page-touching loads alive
        __syncthreads();
        if (warp == 0) {
            int cnt = 0;
            if (shortCode == 1) {
                // GAP_KnapPisinger.h:206-212: every item with negative reduced cost
                const int nAll = colEnd - colStart;
                for (int base = 0; base < nAll; base += 32) {
                    const int j = base + lane;
                    const bool act = j < nAll && redCostV[colStart + j] < 0.0 && (a.back.fix == nullptr || a.back.fix[colStart + j] == 0);
                    const unsigned bal = __ballot_sync(0xffffffffu, act);
                    if (act) sFinal[cnt + __popc(bal & ((1u << lane) - 1u))] = colStart + j;
                    cnt += __popc(bal);
                }
            } else if (shortCode == 2) {
                for (int k = 0; k < 2; ++k) {
                    const int j = a.back.sel[2 * vlb + k];
                    if (j >= 0) {
                        if (lane == 0) sFinal[cnt] = colStart + j;
                        ++cnt;
                    }
                }
            } else {
                // backtrack from (last item, capacity), gap_func.py:173-181
                int j = cap;
                int i = kGlobal - 1;
                if (kGlobal <= 96) {
                    // few items (the usual case after the reduction): ONE lane walks, without the ballot /
                    // shuffle round trip per taken item.  A single warp pays every dependent instruction at
                    // full latency, so the step is kept short: the cell's three decision words are fetched
                    // together and masked to the items at or below i, selects pick the highest set bit
                    // (~155 cycles per taken item, two dependent shared-memory loads of ~33 among them;
                    // GAP-E: 47 taken items per block, 3.8 us).
                    if (lane == 0) {
                        // words 1 and 2 are read unconditionally (three independent loads, no branch in the step);
                        // where they do not exist the pointer falls back to word 0 and the mask is zero anyway
                        const uint32_t *b0 = sBits, *b1 = kGlobal > 32 ? sBits + RC : sBits, *b2 = kGlobal > 64 ? sBits + 2 * RC : sBits;
                        // the three masks depend on i only, which is known one shared-memory load before j is:
                        // they are formed in the shadow of the weight fetch, without a branch
                        auto low = [](int n) { return (uint32_t)((1ull << max(0, min(32, n))) - 1ull); };  // n low bits
                        uint32_t m0 = low(i + 1), m1 = low(i - 31), m2 = low(i - 63);
                        while (i >= 0) {  // (left by the break below)
                            const uint32_t v0 = b0[j] & m0, v1 = b1[j] & m1, v2 = b2[j] & m2;
                            const int i0 = 31 - __clz(v0), i1 = 63 - __clz(v1), i2 = 95 - __clz(v2);
                            const int item = v2 ? i2 : v1 ? i1 : v0 ? i0 : -1;
                            if (item < 0) break;
                            j -= itW[item];
                            m0 = low(item);
                            m1 = low(item - 32);
                            m2 = low(item - 64);
                            sTaken[cnt] = itC[item];
                            if (smemSums) sTakenIdx[cnt] = item;
                            ++cnt;
                        }
                    }
                    cnt = __shfl_sync(0xffffffffu, cnt, 0);
                    i = -1;
                }
                while (i >= 0) {
                    const int kw = i >> 5;
                    const int wv = kw - lane;
                    uint32_t v = wv >= 0 ? sBits[(size_t)wv * RC + j] : 0u;
                    if (lane == 0) v &= 0xffffffffu >> (31 - (i & 31));
                    const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
                    if (nz == 0u) {
                        i = ((kw - 32) << 5) | 31;
                        continue;
                    }
                    const int L = __ffs(nz) - 1;
                    const uint32_t vL = __shfl_sync(0xffffffffu, v, L);
                    const int item = ((kw - L) << 5) + (31 - __clz(vL));
                    if (lane == 0) {
                        sTaken[cnt] = itC[item];
                        if (smemSums) sTakenIdx[cnt] = item;
                    }
                    ++cnt;
                    j -= itW[item];
                    i = item - 1;
                }
                __syncwarp();
                // the walk found the items last-to-first: reverse into ascending x-space indices
                for (int k = lane; k < cnt; k += 32) sFinal[k] = colStart + sTaken[cnt - 1 - k];
            }
            __syncwarp();
            if (a.back.fix != nullptr && cap >= 0) {
                // node bounds: the columns fixed to 1 join the column.  Both lists ascend; every element
                // finds its place by a binary search in the other list (no column is in both).
                const int nAll = colEnd - colStart;
                int32_t *sForced = sTaken;  // the taken list is consumed
                int nf = 0;
                for (int base = 0; base < nAll; base += 32) {
                    const int q = base + lane;
                    const bool f = q < nAll && (a.back.fix[colStart + q] & 1u);
                    const unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (f) sForced[nf + __popc(bal & ((1u << lane) - 1u))] = colStart + q;
                    nf += __popc(bal);
                }
                __syncwarp();
                if (nf > 0) {
                    int32_t *sMerged = itW;  // the compacted weights are consumed
                    for (int k = lane; k < cnt + nf; k += 32) {
                        const bool fromChosen = k < cnt;
                        const int me = fromChosen ? sFinal[k] : sForced[k - cnt];
                        const int32_t *other = fromChosen ? sForced : sFinal;
                        int lo = 0, hi = fromChosen ? nf : cnt;  // elements of `other` below `me`
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (other[mid] < me) lo = mid + 1; else hi = mid;
                        }
                        sMerged[(fromChosen ? k : k - cnt) + lo] = me;
                    }
                    __syncwarp();
                    for (int k = lane; k < cnt + nf; k += 32) sFinal[k] = sMerged[k];
                    cnt += nf;
                    __syncwarp();
                }
            }
            // varRedCost / varOrigCost: added in ascending index order (GAP_KnapPisinger.h:263-272).
            // All lanes fetch, one lane adds sequentially out of shared memory.
            if (smemSums && shortCode == 0) {
                // the values are already in shared memory (prefetched during the compaction); the walk found
                // the items last to first, the sums run first to last
                if (lane < 2) {
                    const double *v = itP + (lane == 0 ? kSumsRcOff : kSumsObjOff);
                    double acc2 = 0.0;
                    for (int k = cnt - 1; k >= 0; --k) acc2 += v[sTakenIdx[k]];
                    if (lane == 0) sRcSum = acc2; else sOcSum = acc2;
                }
            } else {
                // Both gathers are issued before either sum (one round trip), lanes 0 and 1 add side by side.
                double *sValsRc = itP;                                  // free after the DP
                double *sValsOc = reinterpret_cast<double *>(sBits);    // the decision bits are consumed
                for (int k = lane; k < cnt; k += 32) {
                    const int col = sFinal[k];
                    const double r = redCostV[col], o = a.back.obj[col];
                    sValsRc[k] = r;
                    sValsOc[k] = o;
                }
                __syncwarp();
                if (lane < 2) {
                    const double *v = lane == 0 ? sValsRc : sValsOc;
                    double acc2 = 0.0;
                    for (int k = 0; k < cnt; ++k) acc2 += v[k];
                    if (lane == 0) sRcSum = acc2; else sOcSum = acc2;
                }
            }
            __syncwarp();
            if (lane == 0) sCnt = cnt;
        }
        __syncthreads();
        dbgStamp(a.dbgTimes, 4, lb, (int)vec);

        // ---- stage (c), DecompAlgo.cpp:5151-5232, without a second launch and without a serial
        // tail: every CTA publishes (kept?, nnz, cells) and waits only for the LOWER-numbered
        // blocks to learn where its column goes.  Block ids are TICKETS drawn at CTA entry, so a lower
        // block has always started before this one (whatever order the hardware dispatches CTAs in,
        // however many fit on the GPU at once) and block 0 never waits.  Words carry the launch epoch,
        // so nothing has to be reset between rounds.
        const int cnt = sCnt;
        // DecompAlgo.cpp:6350-6356; a block whose fixed columns do not fit has no column: rc = +inf
        const double rcB = cap < 0 ? -kNegInf : sRcSum - alphaB;
        const bool keep = rcB < -a.filt.eps;          // strict, :5216
        // cells of the reference's DP (items before the reduction) and cells actually evaluated
        const unsigned long long myCells = (unsigned long long)nItRef * (unsigned long long)(cap + 1);  // 0 when cap < 0
        const unsigned long long myEval = (unsigned long long)kGlobal * (unsigned long long)(cap + 1);
        if (t == 0) {
            // one 64-bit word per block (layout: kPub* above).  Every block publishes in every launch that
            // prices its vector.  Published FIRST: the word is all the other blocks need, the result stores
            // below are read by the host only.
            const unsigned long long word = (1ull << 63) | ((unsigned long long)(epochRaw & kPubEpochMask) << kPubEpochShift) |
                                            (keep ? 1ull << kPubKeptBit : 0ull) |
                                            ((unsigned long long)(cap < 0 ? 0 : nItRef) << kPubItemsShift) |
                                            ((unsigned long long)kGlobal << kPubEvalShift) | (unsigned long long)cnt;
            // relaxed: the word itself is everything the other blocks read (no other store has to be ordered before it)
            // an atomic goes straight to the L2 slice (a plain store may wait in the SM's write path)
            atomicExch(a.pubWord + (size_t)vlb * kPubWordStride, word);
            RES_PUT(a.back.blockRedCost, lb, rcB);
            RES_PUT(a.back.blockOrigCost, lb, sOcSum);
            RES_PUT(a.back.blockNnz, lb, cnt);
            RES_PUT(a.filt.blockMostNeg, lb, rcB < 0.0 ? rcB : 0.0);  // min(0, rc): for EVERY block (:5212)
        }
        // prefix over the lower blocks: each thread polls its share, warp-shuffle reduction, then
        // the per-warp partials (shared-memory 64-bit atomics would serialise through CAS loops)
        __shared__ int sKeptW[32];
        __shared__ unsigned long long sNnzW[32], sCellsW[32], sEvalW[32];
        __shared__ int sKept;
        __shared__ unsigned long long sNnz, sCells, sEval;
        {
            int kc = 0;
            unsigned long long nz = 0ull, cl = 0ull, ev = 0ull;
            for (int pB = t; pB < lb; pB += T) {
                unsigned long long wv;
                // (fetched before the spin, so that its latency hides behind the wait)
                const unsigned long long rowB = (unsigned long long)max(tab ? a.tab[pB][1] : a.capacity[a.firstBlock + pB], -1) + 1ull;
                const unsigned long long *src = a.pubWord + (size_t)(vlb - lb + pB) * kPubWordStride;
                for (;;) {
                    // an atomic is performed at the line's home L2 slice: a plain (even .relaxed.gpu) load
                    // was seen to return the old word for ~4.5 us on part of the SMs (the other die's L2)
                    wv = atomicAdd(const_cast<unsigned long long *>(src), 0ull);
                    if ((wv >> kPubEpochShift) == ((1ull << 26) | (unsigned long long)(epochRaw & kPubEpochMask))) break;  // valid bit + epoch
                    // a short pause between polls (__nanosleep's granularity here is microseconds): 80 blocks
                    // x up to 79 threads polling back to back congest the words' L2 slices, and the stores that
                    // follow the wait queue up behind the polls
                    const long long t0 = clock64();
                    while (clock64() - t0 < 400) {}
                }
                if (wv & (1ull << kPubKeptBit)) {
                    ++kc;
                    nz += wv & 0xfffull;  // nnz <= 2048
                }
                cl += ((wv >> kPubItemsShift) & 0xfffull) * rowB;
                ev += ((wv >> kPubEvalShift) & 0xfffull) * rowB;
            }
#pragma unroll
            for (int off2 = 16; off2 > 0; off2 >>= 1) {
                kc += __shfl_xor_sync(0xffffffffu, kc, off2);
                nz += __shfl_xor_sync(0xffffffffu, nz, off2);
                cl += __shfl_xor_sync(0xffffffffu, cl, off2);
                ev += __shfl_xor_sync(0xffffffffu, ev, off2);
            }
            if (lane == 0) {
                sKeptW[warp] = kc;
                sNnzW[warp] = nz;
                sCellsW[warp] = cl;
                sEvalW[warp] = ev;
            }
        }
        __syncthreads();
        dbgStamp(a.dbgTimes, 6, lb, (int)vec);
        if (t == 0) {
            int kc = 0;
            unsigned long long nz = 0ull, cl = 0ull, ev = 0ull;
            for (int v = 0; v < nWarps; ++v) {
                kc += sKeptW[v];
                nz += sNnzW[v];
                cl += sCellsW[v];
                ev += sEvalW[v];
            }
            sKept = kc;
            sNnz = nz;
            sCells = cl;
            sEval = ev;
        }
        __syncthreads();
        dbgStamp(a.dbgTimes, 7, lb, (int)vec);
        const int pos = sKept;
        const long long off = (long long)sNnz;
        if (keep) {
            if (t == 0) {
                RES_PUT(a.filt.keptBlock, pos, lb);
                RES_PUT(a.filt.keptStart, pos, off);
            }
            for (int k = t; k < cnt; k += T) RES_PUT(a.filt.keptInd, off + k, sFinal[k]);
        }
        if (lb == a.back.nLocal - 1 && t == 0) {
            // the highest block has seen every other block: it closes the packed result
            const int nKept = pos + (keep ? 1 : 0);
            const long long nInd = off + (keep ? cnt : 0);
            RES_PUT(a.filt.keptStart, nKept, nInd);
            ResultHeader h;
            h.magic = kResultMagic;
            h.m = a.back.nLocal;
            h.firstBlock = a.firstBlock;
            h.nKept = nKept;
            h.nInd = nInd;
            h.cells = (long long)(sCells + myCells);
            h.indCap = a.filt.indCap;
            h.reserved[0] = (long long)(sEval + myEval);  // cells evaluated after the reduction
            h.reserved[1] = h.reserved[2] = 0;
            RES_PUT(a.filt.hdr, 0, h);
        }
        dbgStamp(a.dbgTimes, 5, lb, (int)vec);
        // the last CTA to get here has seen every CTA of the launch read the epoch: the next launch gets a new one
        if (a.doneFlag != nullptr) __syncthreads();  // every thread's mirrored stores are ordered before thread 0's fence
        if (t == 0) {
            if (a.doneFlag != nullptr) __threadfence_system();
            if (atomicAdd(a.doneCtr, 1u) == gridDim.x * gridDim.y - 1u) {
                *a.doneCtr = 0u;
                if (pulls)
                    for (unsigned int v = 0; v < gridDim.y; ++v) a.pullCtr[a.vecBase + v] = 0u;
                *a.epochCtr = epochRaw + 1u;
                if (a.doneFlag != nullptr) {
                    // every CTA's stores were performed system-wide (its fence.sys above) before its arrival was
                    // counted, and the count was read before this store is issued: a device-scope fence orders the
                    // rest (a second fence.sys would only wait once more for the link)
                    __threadfence();
                    *reinterpret_cast<volatile unsigned int *>(a.doneFlag) = a.doneVal;
                }
            }
        }
    }
#undef RES_PUT
#undef RESV
}

// --------------------------------------------------------------- geometry table

typedef void (*KnapFn)(const KnapArgs);
struct GeomEntry {
    int S, maxT;
    KnapFn fn[2][3];      // [guard][items-1]; three items per barrier only with guard cells, S <= 3
    KnapFn fnFuse[2][3];  // same, fused tail (NULL where not built)
    // the fused kernel built for <= 256 threads (the GAP shapes: a row of a few hundred cells): 85 registers per
    // thread instead of the 64 a 1024-thread build is held to -- no spills, room to keep loads in flight
    KnapFn fnFuseSmall[2][3];
};

#define GE_FNS(S, MT, F)                                                                         \
    {                                                                                            \
        {knap_dp_kernel<S, false, 1, MT, F>, knap_dp_kernel<S, false, 2, MT, F>, nullptr},       \
        {knap_dp_kernel<S, true, 1, MT, F>, knap_dp_kernel<S, true, 2, MT, F>, nullptr}          \
    }
#define GE_FNS3(S, MT, F)                                                                        \
    {                                                                                            \
        {knap_dp_kernel<S, false, 1, MT, F>, knap_dp_kernel<S, false, 2, MT, F>, nullptr},       \
        {knap_dp_kernel<S, true, 1, MT, F>, knap_dp_kernel<S, true, 2, MT, F>,                   \
         knap_dp_kernel<S, true, 3, MT, F>}                                                      \
    }
#define GE_NULL {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}}
#define GE(S, MT) {S, MT, GE_FNS(S, MT, false), GE_NULL, GE_NULL}
#define GEF(S, MT) {S, MT, GE_FNS(S, MT, false), GE_FNS(S, MT, true), GE_FNS(S, 256, true)}
#define GEF3(S, MT) {S, MT, GE_FNS3(S, MT, false), GE_FNS3(S, MT, true), GE_FNS3(S, 256, true)}
static const GeomEntry kGeoms[] = {
    GEF3(1, 1024), GEF3(3, 1024), GEF(5, 1024), GEF(7, 1024), GE(9, 1024), GE(11, 1024),
    GE(13, 512),   GE(15, 512),   GE(17, 512),  GE(19, 512),  GE(21, 512), GE(25, 512),
};
#undef GE
#undef GEF
#undef GEF3
#undef GE_FNS
#undef GE_FNS3
#undef GE_NULL
static_assert(kPruneBuckets % 32 == 0, "one lane scans kPruneBuckets / 32 buckets");
static const int kNumGeoms = (int)(sizeof(kGeoms) / sizeof(kGeoms[0]));
static const size_t kMaxSmem = 227 * 1024 - 6 * 1024;  // the fused tail's static shared memory comes off the 227 KiB

static KnapFn pickKnapFn(const GeomEntry *e, const DpGeom &g, bool latency = true)
{
    const int gi = g.guard > 0 ? 1 : 0, ii = g.items - 1;
    // latency-bound launches (a round, a short ladder: co-resident grids): the <= 256-thread build with registers to
    // spare; large batches: the 64-register build, three CTAs per SM
    if (latency && g.fuse && g.T <= 256 && e->fnFuseSmall[gi][ii] != nullptr) return e->fnFuseSmall[gi][ii];
    return (g.fuse ? e->fnFuse : e->fn)[gi][ii];
}

static const GeomEntry *findGeom(int S);

static const GeomEntry *findGeom(int S)
{
    for (int i = 0; i < kNumGeoms; ++i)
        if (kGeoms[i].S == S) return &kGeoms[i];
    return nullptr;
}

// Fills the derived fields of a (T, S) candidate; returns false if it cannot run.
static bool fitGeom(const Ctx *c, int T, int S, int items, bool fuse, DpGeom *g)
{
    const GeomEntry *e = findGeom(S);
    if (!e || T < 32 || T > e->maxT || (T & 31)) return false;
    if (fuse && (e->fnFuse[0][0] == nullptr || c->maxBlockCols > 2048)) return false;
    if (items == 3 && (fuse ? e->fnFuse : e->fn)[1][2] == nullptr) return false;
    g->fuse = fuse ? 1 : 0;
    if ((int64_t)T * S < (int64_t)c->maxCapacity + 1) return false;
    g->T = T;
    g->S = S;
    g->items = items;
    g->rowCells = T * S;
    int ch = (c->maxBlockCols + 3) & ~3;
    ch = std::max(4, std::min(ch, 2048));
    g->chunkCols = ch;
    // guard: the largest usable weight (x2 when two items advance per barrier)
    int guard = ((c->maxUsableWeight * items) + 1) & ~1;
    g->guard = guard;
    g->smemBytes = dpSmemBytes(g->rowCells, guard, ch, fuse, c->optNoPrune == 0);
    if (g->smemBytes > kMaxSmem || c->optNoGuard) {
        g->guard = 0;  // predicated shifted reads instead
        g->smemBytes = dpSmemBytes(g->rowCells, 0, ch, fuse, c->optNoPrune == 0);
    }
    while (!fuse && g->smemBytes > kMaxSmem && g->chunkCols > 256) {
        g->chunkCols = (g->chunkCols / 2 + 3) & ~3;
        g->smemBytes = dpSmemBytes(g->rowCells, g->guard, g->chunkCols, false, false);
    }
    if (items == 3 && g->guard == 0) return false;  // the three-item step is built with guard cells only
    return g->smemBytes <= kMaxSmem;
}

// Cost model (cycles per item step of one CTA, times the number of CTA waves):
// a step is bounded below by the dependent chain of one warp (shared-memory
// load -> DADD -> DSETP -> select -> store -> barrier), by the issue slots of
// the warps resident on a scheduler, and by the shared-memory crossbar
// (128 B/clk/SM, 16 B per cell).  Constants fitted to B200 measurements
// (profiles/r01_dp_geometry_sweep.md).
static double geomCost(const Ctx *c, const DpGeom &g)
{
    const int sm = c->smCount > 0 ? c->smCount : 148;
    const int warps = g.T / 32;
    int ctasPerSm = std::min(32 / warps, (int)(kMaxSmem / g.smemBytes));  // 1024 threads at <= 64 regs
    if (g.S > 11) ctasPerSm = std::min(ctasPerSm, 16 / warps);           // 512-thread kernels use <= 128 regs
    ctasPerSm = std::max(ctasPerSm, 1);
    const int64_t concurrent = (int64_t)sm * ctasPerSm;
    const int64_t n = std::max(c->nLocal, 1);
    const double waves = std::ceil((double)n / (double)concurrent);
    const double residentCtas = std::min<double>(ctasPerSm, std::ceil((double)n / sm));
    // per item: dependent chain of one warp (GAP-E sweep: ~120 + 22 S cycles; the two-item
    // step issues 1.5x the instructions and loses in this regime)
    double chain = 120.0 + 22.0 * g.S + 2.0 * warps;
    // the unrolled two-item step (S <= 3 with guard cells) halves the barriers: 68 vs 76 us on
    // GAP-E; the generic two-item step costs 1.3x (profiles/r01_dp_geometry_sweep.txt)
    if (g.items == 2) chain *= (g.S <= 3 && g.guard > 0) ? 0.9 : 1.3;
    // three items per barrier: measured SLOWER than two on GAP-E (70 vs 60 us) -- 7 DADD + 7 DSETP per
    // cell and step make the FP64 pipe, not the barrier, the limiter; kept as a tested option only
    if (g.items == 3) chain *= 1.15;
    // throughput regime (CSP sweep): shared-memory floor times the measured inefficiency
    const double ineff = g.S <= 11 ? (g.items == 2 ? 1.52 : 1.61) : g.S <= 21 ? (g.items == 2 ? 1.40 : 1.53)
                                                                              : (g.items == 2 ? 1.69 : 1.81);
    const double instrWarp = (g.items == 2 ? 10.0 + 10.5 * g.S : 14.0 + 8.5 * g.S);
    const double issue = instrWarp * warps * residentCtas / 4.0;
    const double smem = (double)g.rowCells * 16.0 * residentCtas / 128.0 * ineff;
    return waves * std::max(chain, std::max(issue, smem));
}

int chooseDpGeom(Ctx *c, DpGeom *g, bool wantFuse)
{
    const int cells = c->maxCapacity + 1;
    DpGeom best{};
    double bestCost = 0.0;
    bool have = false;
    for (int pass = wantFuse ? 0 : 1; pass < 2 && !have; ++pass)  // fused candidates first, if wanted
    for (int i = 0; i < kNumGeoms; ++i) {
        const bool fuse = pass == 0;
        const int S = kGeoms[i].S;
        if (c->optCellsPerThread > 0 && S != c->optCellsPerThread) continue;
        int T = ((cells + S - 1) / S + 31) & ~31;
        if (c->optThreads > 0) {
            if (c->optThreads < T) continue;
            T = c->optThreads;
        }
        for (int items = 1; items <= 3; ++items) {
            if (c->optItemsPerStep > 0 && items != c->optItemsPerStep) continue;
            DpGeom cand{};
            if (!fitGeom(c, T, S, items, fuse, &cand)) continue;
            const double cost = geomCost(c, cand);
            if (!have || cost < bestCost) {
                best = cand;
                bestCost = cost;
                have = true;
            }
        }
    }
    if (!have) {
        if (c->optThreads > 0 || c->optCellsPerThread > 0)
            return fail(c, DIPGPU_ERR_INVALID, "no DP geometry with threads=%d cellsPerThread=%d covers %d cells",
                        c->optThreads, c->optCellsPerThread, cells);
        // more cells than one CTA holds in registers / shared memory: rows in global memory (bigdp.cu)
        DpGeom big{};
        big.T = 1024;
        big.items = 1;
        big.big = 1;
        // ... or, up to 16 x 13 312 cells, distributed over the shared memories of a thread-block cluster
        if (!c->optNoCluster) {
            planClusterDp((int64_t)cells, c->optClusterThreads, c->optClusterMaxCells, &big.cluster, &big.S);
            if (big.cluster > 0) big.T = c->optClusterThreads;
        }
        *g = big;
        return DIPGPU_OK;
    }
    *g = best;
    return DIPGPU_OK;
}

int prepareKnapsackDp(Ctx *c)
{
    const DpGeom &g = c->geom;
    const GeomEntry *e = findGeom(g.S);
    if (!e) return fail(c, DIPGPU_ERR_STATE, "DP geometry not initialised");
    KnapFn fn = pickKnapFn(e, g);
    if (c->dpFnConfigured != (void *)fn || c->dpFnSmem != g.smemBytes) {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(fn), g.smemBytes);
        if (rc) return rc;
        int perSm = 0;
        c->dpCoResident = 0;
        if (g.fuse && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, g.T, g.smemBytes) == cudaSuccess)
            c->dpCoResident = perSm * std::max(c->smCount, 1);
        (void)cudaGetLastError();
        // the build large batches run (tickets): how many of its CTAs are resident at once
        c->dpResidentBatch = 0;
        if (g.fuse) {
            KnapFn fb = pickKnapFn(e, g, false);
            int rcB = ensureSmemLimit(c, reinterpret_cast<const void *>(fb), g.smemBytes);
            if (rcB) return rcB;
            perSm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fb, g.T, g.smemBytes) == cudaSuccess)
                c->dpResidentBatch = perSm * std::max(c->smCount, 1);
            (void)cudaGetLastError();
        }
        c->dpFnConfigured = (void *)fn;
        c->dpFnSmem = g.smemBytes;
    }
    return DIPGPU_OK;
}

static void fillBackArgs(Ctx *c, const double *dRedCost, const double *dAlpha, BackArgs *a)
{
    char *res = static_cast<char *>(c->dResult);
    a->redCost = dRedCost;
    a->obj = c->dObj;
    a->alpha = dAlpha;
    a->blockColStart = c->dBlockColStart;
    a->capacity = c->dCapacity;
    a->bitsOff = c->dBitsOff;
    a->bits = c->dBits;
    a->itemW = c->dItemW;
    a->itemCol = c->dItemCol;
    a->nItems = c->dNItems;
    a->scratch = c->dScratch;
    a->candInd = c->dCandInd;
    a->blockRedCost = reinterpret_cast<double *>(res + c->layout.offRedCost);
    a->blockOrigCost = reinterpret_cast<double *>(res + c->layout.offOrigCost);
    a->blockNnz = reinterpret_cast<int32_t *>(res + c->layout.offNnz);
    a->firstBlock = c->firstBlock;
    a->nLocal = c->nLocal;
    a->rowCells = c->geom.rowCells;
    a->doneCounter = c->dDoneCounter;
    const bool pis = c->profitMode == DIPGPU_PROFIT_PISINGER_INT;
    a->shortcut = pis ? c->dShortcut : nullptr;
    a->sel = pis ? c->dSel : nullptr;
    a->fix = c->haveFix ? c->dFix : nullptr;
    a->bigLayout = c->geom.big;
    if (c->haveFix) a->capacity = c->dCapEff;  // capacity left once the columns fixed to 1 are in
}

static int launchDp(Ctx *c, const double *dRedCost, int64_t rcStride, const double *dAlpha, int64_t alphaStride,
                    double redCostEps, int nVec, bool batch, cudaStream_t st, const RcFuse *rf, int vecBase = 0);

// Large-shard path under node bounds: the DP kernels read a copy of the reduced costs in which every
// fixed column is +0 (so it is not an item); the backtrack kernel still sums the true reduced costs.
__global__ void __launch_bounds__(256) mask_fixed_kernel(int n, const double *__restrict__ rc, const uint8_t *__restrict__ fix,
                                                         double *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = fix[j] ? 0.0 : rc[j];
}

int launchKnapsackDp(Ctx *c, const double *dRedCost, const double *dAlpha, double redCostEps, cudaStream_t st, const RcFuse *rf)
{
    return launchDp(c, dRedCost, 0, dAlpha, 0, redCostEps, 1, false, st, rf);
}

// Batched dual vectors: the fused kernel only (gridDim.y = vectors); the per-vector packed results,
// publication words and Pisinger-mode block arrays live in the context's batch buffers.
int launchKnapsackDpBatch(Ctx *c, const double *dRedCost, int64_t rcStride, const double *dAlpha, int64_t alphaStride,
                          double redCostEps, int nVec, cudaStream_t st, const RcFuse *rf, int vecBase)
{
    return launchDp(c, dRedCost, rcStride, dAlpha, alphaStride, redCostEps, nVec, true, st, rf, vecBase);
}

static int launchDp(Ctx *c, const double *dRedCost, int64_t rcStride, const double *dAlpha, int64_t alphaStride,
                    double redCostEps, int nVec, bool batch, cudaStream_t st, const RcFuse *rf, int vecBase)
{
    if (c->nLocal == 0 || nVec <= 0) return DIPGPU_OK;
    if (batch && !c->geom.fuse) return fail(c, DIPGPU_ERR_STATE, "batched DP launch needs the fused kernel shape");
    const DpGeom &g = c->geom;
    const GeomEntry *e = findGeom(g.S);
    if (!e) return fail(c, DIPGPU_ERR_STATE, "DP geometry not initialised");
    {
        int rc = prepareKnapsackDp(c);  // (the latency build: shared-memory limit, co-resident CTA count)
        if (rc) return rc;
    }
    // co-resident grids (a GAP round, a short ladder): cooperative launch, CTA ids = grid coordinates; larger ones
    // (batches): tickets, and the build that fits three CTAs per SM
    const long long totalCtas = (long long)c->nLocal * nVec;
    const bool coop = g.fuse && !c->optPdl && !c->optNoCoop && c->dpCoResident > 0 && totalCtas <= c->dpCoResident;
    KnapFn fn = pickKnapFn(e, g, coop || !g.fuse || totalCtas <= c->dpCoResident);
    if (fn != (KnapFn)c->dpFnConfigured) {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(fn), g.smemBytes);
        if (rc) return rc;
    }
    KnapArgs a;
    a.redCost = dRedCost;
    if (c->haveFix && !g.fuse) {
        const int n = c->nBlockCols;
        if (!c->dRcEff) {
            DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dRcEff), sizeof(double) * ((size_t)c->N + 16)));
            DIPGPU_CUDA(c, cudaMemset(c->dRcEff, 0, sizeof(double) * ((size_t)c->N + 16)));
        }
        mask_fixed_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, dRedCost, c->dFix, c->dRcEff);
        c->launches++;
        a.redCost = c->dRcEff;
    }
    a.weight = c->dWeight;
    a.blockColStart = c->dBlockColStart;
    a.capacity = c->haveFix ? c->dCapEff : c->dCapacity;
    a.bitsOff = c->dBitsOff;
    a.bits = c->dBits;
    a.itemW = c->dItemW;
    a.itemCol = c->dItemCol;
    a.nItems = c->dNItems;
    char *resBase = static_cast<char *>(batch ? c->dResultBatch : c->dResult);
    a.blockObj = reinterpret_cast<double *>(resBase + c->layout.offObj);
    a.prune = c->optNoPrune ? 0 : 1;
    a.rcStride = rcStride;
    a.alphaStride = alphaStride;
    a.resStride = batch ? (long long)c->resStride : 0;
    a.firstBlock = c->firstBlock;
    a.chunkCols = g.chunkCols;
    a.guard = g.guard;
    a.dbgTimes = c->dDbgTimes;
    a.pubWord = batch ? c->dPubWordBatch : c->dPubWord;
    a.pubCells = c->dPubCells;
    a.ticket = coop ? nullptr : c->dSyncWords;
    a.epochCtr = c->dSyncWords + 1;
    a.doneCtr = c->dSyncWords + 2;
    a.rcStart = rf ? rf->start : nullptr;
    a.rcIdx = rf ? rf->idx : nullptr;
    a.rcVal = rf ? rf->val : nullptr;
    a.uVec = rf ? rf->uVec : nullptr;
    a.alphaIdx = rf ? rf->alphaIdx : nullptr;
    a.uStride = rf ? (long long)rf->uStride : 0;
    a.phase1 = rf ? rf->phase1 : 0;
    // in-kernel dual upload: a co-resident grid, or a ticketed one in which at least one whole vector's CTAs are
    // resident at once (tickets are drawn in start order and a CTA leaves only after its vector's CTAs have met: the
    // lowest unfinished vector always has all its CTAs started)
    const bool ticketPull = rf && rf->pullSrc && !coop && batch && g.fuse && c->dpResidentBatch >= c->nLocal && c->dPullCtrBatch != nullptr;
    a.pullSrc = (rf && (coop || ticketPull)) ? rf->pullSrc : nullptr;
    a.pullIdx = rf ? rf->pullIdx : nullptr;
    a.pullStride = rf ? (long long)rf->pullStride : 0;
    a.pullN = rf ? rf->pullN : 0;
    a.pullCtr = (batch && c->dPullCtrBatch) ? c->dPullCtrBatch : c->dPullCtr;
    a.vecBase = vecBase;
    if (rf && rf->pullSrc && !coop && !ticketPull) return fail(c, DIPGPU_ERR_STATE, "in-kernel dual upload needs a co-resident grid");
    if (rf && !g.fuse) return fail(c, DIPGPU_ERR_STATE, "reduced costs can only be formed inside the fused DP kernel");
    a.useTab = (g.fuse && c->nLocal <= kFuseFilterMaxBlocks) ? 1 : 0;
    if (rf && !a.useTab) return fail(c, DIPGPU_ERR_STATE, "in-block reduced costs: %d blocks", c->nLocal);
    if (a.useTab) {
        const std::vector<int32_t> &capH = c->haveFix ? c->hCapEff : c->hCapacity;
        for (int lb = 0; lb <= c->nLocal; ++lb) {
            const int b = c->firstBlock + lb;
            a.tab[lb][0] = c->hBlockColStart[(size_t)b];
            a.tab[lb][1] = lb < c->nLocal ? capH[(size_t)b] : 0;
            a.tab[lb][2] = rf ? rf->hBlkEnt[b] : 0;
        }
    }
    if (g.fuse && (++c->fusedLaunches & 0xffffffull) == 0) {
        // a publication word that a launch does not rewrite (a vector beyond its nVec) must never meet its own
        // 26-bit epoch again: all words are cleared long before the epoch wraps
        DIPGPU_CUDA(c, cudaMemsetAsync(c->dPubWord, 0, sizeof(unsigned long long) * (size_t)std::max(c->nLocal, 1) * kPubWordStride, st));
        if (c->dPubWordBatch)
            DIPGPU_CUDA(c, cudaMemsetAsync(c->dPubWordBatch, 0, sizeof(unsigned long long) * (size_t)c->batchCap * (size_t)std::max(c->nLocal, 1) * kPubWordStride, st));
    }
    a.profitMode = c->profitMode;
    a.scale = batch ? c->dScaleB : c->dScale;
    a.shortcut = batch ? c->dShortcutB : c->dShortcut;
    a.zShort = batch ? c->dZShortB : c->dZShort;
    fillBackArgs(c, dRedCost, dAlpha, &a.back);
    a.filt = makeFilterArgs(c, redCostEps);
    if (batch) {
        // same layout, based at the batch buffer (vector 0); the kernel adds v*resStride
        const ptrdiff_t d = resBase - static_cast<char *>(c->dResult);
        auto mv = [d](auto *&ptr) { ptr = reinterpret_cast<std::remove_reference_t<decltype(ptr)>>(reinterpret_cast<char *>(const_cast<void *>(static_cast<const void *>(ptr))) + d); };
        mv(a.back.blockRedCost); mv(a.back.blockOrigCost); mv(a.back.blockNnz);
        mv(a.filt.hdr); mv(a.filt.blockRedCost); mv(a.filt.blockMostNeg); mv(a.filt.blockNnz);
        mv(a.filt.keptStart); mv(a.filt.keptBlock); mv(a.filt.keptInd);
        a.back.sel = c->profitMode == DIPGPU_PROFIT_PISINGER_INT ? c->dSelB : nullptr;
        a.back.shortcut = c->profitMode == DIPGPU_PROFIT_PISINGER_INT ? c->dShortcutB : nullptr;
    }
    a.filt.fuseCopy = c->nLocal <= kFuseFilterMaxBlocks ? 2 : 1;
    // host mirror of the packed result (the synchronous host entry points ask for it; option "result_zero_copy")
    a.hostOff = 0;
    if (g.fuse && (c->wantMirror || c->optMirrorAlways) && c->optZeroCopy) {
        if (!batch && c->mirrorBase) a.hostOff = (long long)(static_cast<char *>(c->mirrorBase) - static_cast<char *>(c->dResult));
        else if (!batch && c->dResultHost) a.hostOff = (long long)(static_cast<char *>(c->dResultHost) - static_cast<char *>(c->dResult));
        if (batch && c->dResultBatchHost)  // same per-vector stride on both sides
            a.hostOff = (long long)(static_cast<char *>(c->dResultBatchHost) - static_cast<char *>(c->dResultBatch));
    }
    c->resultMirrored = a.hostOff != 0;
    a.doneFlag = nullptr;
    a.doneVal = 0u;
    c->flagArmed = false;
    if (c->wantFlag && c->optFlagWait && a.hostOff != 0 && !batch && !c->mirrorBase && !c->stageTiming) {
        if (!c->hDoneFlag) {
            void *hp = nullptr, *dp = nullptr;
            DIPGPU_CUDA(c, cudaHostAlloc(&hp, 64, cudaHostAllocMapped | cudaHostAllocPortable));
            memset(hp, 0, 64);
            DIPGPU_CUDA(c, cudaHostGetDevicePointer(&dp, hp, 0));
            c->hDoneFlag = static_cast<unsigned int *>(hp);
            c->dDoneFlagHost = static_cast<unsigned int *>(dp);
        }
        if (++c->doneSerial == 0u) ++c->doneSerial;
        a.doneFlag = c->dDoneFlagHost;
        a.doneVal = c->doneSerial;
        c->flagArmed = true;
    }
    // Option "dp_pdl" (fused fp64 shapes): programmatic dependent launch -- the CTAs may become resident while
    // the preceding kernel of the stream (the SpMV) still runs; the kernel waits (griddepcontrol.wait) before it
    // reads the reduced costs.  Measured SLOWER on GAP-E (39.5 vs 35.9 us per round: the early CTAs take SM
    // resources from the 6 us SpMV they wait for), so it is off unless asked for.  (Pisinger mode reads the
    // prep kernel's output at its very top: always a plain launch.)
    a.pdl = (g.fuse && c->profitMode == DIPGPU_PROFIT_FP64 && c->optPdl && !rf) ? 1 : 0;
    if (c->hostTiming)
        c->hostUs[3] = 1e-3 * (double)(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() - c->hostT0);
    if (a.pdl || coop) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c->nLocal, nVec, 1);
        cfg.blockDim = dim3(g.T, 1, 1);
        cfg.dynamicSmemBytes = g.smemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        if (coop) {
            attr[0].id = cudaLaunchAttributeCooperative;
            attr[0].val.cooperative = 1;
        } else {
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
        }
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        DIPGPU_CUDA(c, cudaLaunchKernelEx(&cfg, fn, a));
    } else {
        fn<<<dim3(c->nLocal, nVec, 1), g.T, g.smemBytes, st>>>(a);
    }
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

int launchBacktrackEmit(Ctx *c, const double *dRedCost, const double *dAlpha, double redCostEps, bool fuseFilter,
                        cudaStream_t st)
{
    if (c->nLocal == 0) return DIPGPU_OK;
    BackArgs a;
    fillBackArgs(c, dRedCost, dAlpha, &a);
    a.doneCounter = fuseFilter ? c->dDoneCounter : nullptr;
    FilterArgs f = makeFilterArgs(c, redCostEps);
    f.fuseCopy = c->nLocal <= kFuseFilterMaxBlocks ? 2 : 1;
    const int warpsPerCta = kBackThreads / 32;
    const int grid = (c->nLocal + warpsPerCta - 1) / warpsPerCta;
    knap_backtrack_kernel<<<grid, kBackThreads, 0, st>>>(a, f);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

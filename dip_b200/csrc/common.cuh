// common.cuh -- shared declarations of libdipgpu (sm_100a only).
//
// The library implements DIP's pricing round (DecompAlgo::generateVars,
// Dip/src/DecompAlgo.cpp:4622-5260) as CUDA stages; this header holds the
// context, the packed result layout and the launcher prototypes.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/dipgpu.h"

namespace dipgpu {

// ---------------------------------------------------------------- packed result
//
// One contiguous buffer, identical on device and host, so that a shard's result
// can travel as raw bytes (D2H copy, NCCL gather) and be decoded anywhere.
//
//   ResultHeader                       64 B
//   double  blockRedCost[m]            sum rc - alpha          (all blocks)
//   double  blockMostNeg[m]            min(0, .)
//   double  blockOrigCost[m]
//   double  blockObj[m]                knapsack optimum
//   int64   keptStart[m + 1]           offsets into keptInd of the kept columns
//   int32   blockNnz[m]
//   int32   keptBlock[m]               LOCAL block ids of the kept columns
//   (pad to 16 B)
//   int32   keptInd[indCap]            x-space indices, ascending per column
struct ResultHeader {
    uint32_t magic;      // 'DIPR'
    int32_t m;           // blocks in this shard
    int32_t firstBlock;  // global id of local block 0
    int32_t nKept;
    int64_t nInd;
    int64_t cells;
    int64_t indCap;
    int64_t reserved[3];  // [0]: DP cells actually evaluated (<= cells: the fused kernel's reduction drops items)
                          // [1]: 1 = the index region holds (index, multiplicity) PAIRS (integer knapsack blocks): nInd,
                          //      keptStart and nnz count int32 words, two per entry
};
static_assert(sizeof(ResultHeader) == 64, "header is 64 bytes");
constexpr uint32_t kResultMagic = 0x52504944u;

struct ResultLayout {
    size_t offRedCost, offMostNeg, offOrigCost, offObj, offKeptStart, offNnz, offKeptBlock, offInd;
    size_t bytes;  // total with indCap indices
    __host__ __device__ static inline size_t alignUp(size_t x, size_t a) { return (x + a - 1) / a * a; }
    __host__ __device__ static inline ResultLayout make(int m, int64_t indCap)
    {
        ResultLayout L;
        size_t o = sizeof(ResultHeader);
        L.offRedCost = o;  o += sizeof(double) * (size_t)m;
        L.offMostNeg = o;  o += sizeof(double) * (size_t)m;
        L.offOrigCost = o; o += sizeof(double) * (size_t)m;
        L.offObj = o;      o += sizeof(double) * (size_t)m;
        L.offKeptStart = o; o += sizeof(int64_t) * ((size_t)m + 1);
        L.offNnz = o;      o += sizeof(int32_t) * (size_t)m;
        L.offKeptBlock = o; o += sizeof(int32_t) * (size_t)m;
        o = alignUp(o, 16);
        L.offInd = o;      o += sizeof(int32_t) * (size_t)indCap;
        L.bytes = alignUp(o, 16);
        return L;
    }
};

// Host side: the kept entries of a packed result into the caller's arrays.  wordsPerEntry == 2: (index, multiplicity)
// pairs (integer knapsack blocks); els (may be NULL) receives the multiplicities, 1.0 for plain index lists.
inline void unpackEntries(const int32_t *words, int64_t nWords, int wordsPerEntry, int32_t *ind, double *els)
{
    const int64_t n = nWords / wordsPerEntry;
    if (wordsPerEntry == 1) {
        if (ind && n > 0) memcpy(ind, words, sizeof(int32_t) * (size_t)n);
        if (els) for (int64_t q = 0; q < n; ++q) els[q] = 1.0;
        return;
    }
    for (int64_t q = 0; q < n; ++q) {
        if (ind) ind[q] = words[2 * q];
        if (els) els[q] = (double)words[2 * q + 1];
    }
}

constexpr int kFuseFilterMaxBlocks = 256;
constexpr int kPubWordStride = 16;     // fused DP kernel: one publication word per 128-byte line

// ---------------------------------------------------------------- SpMV tiling
constexpr int kSpmvChunk = 384;        // default entry budget of a warp tile (runtime knob "spmv_chunk")
constexpr int kSpmvChunkMax = 4096;    // a single column longer than the budget -> lane-cooperative kernel
constexpr int kSpmvColsPerLane = 4;    // a warp tile spans at most 32*4 consecutive columns (fewer when the
                                       // entry budget is reached first)
constexpr int kSpmvMaxStages = 8;
// A tile of the stream kernel is self-contained in the blob (one bulk copy).  Per slice of 32
// columns, lane l owns the slice's l-th LONGEST column (host-side stable sort inside a window of
// consecutive columns), so the columns that still have an entry at step i are a prefix of the lanes:
//   SpmvTileHeader                     16 B
//   uint16 len[nSlices*32]             entries of lane l's column inside this tile's step range
//   uint16 col[nSlices*32]             column of lane l, relative to firstCol (0xffff = no column)
//   double c[nSlices*32]               objective entry of lane l's column
//   uint16 stepOff[nSlices*stepStride] per slice: position of step i's first entry (padded to 16 B)
//   double val[nEnt (+1 to even)]      JAGGED: per slice, step 0 of every lane, then step 1 of the
//   RowT   row[nEnt (padded to 16 B)]  lanes whose column is longer than 1, ...
// strideFlags = stepStride | first << 16 | last << 17: a slice with more entries than the tile budget is
// cut by step range into consecutive single-slice tiles; `first` zeroes the lanes' sums, `last` stores.
struct SpmvTileHeader {
    int32_t firstCol, nSlices, nEnt, strideFlags;
};
constexpr int kSpmvWindow = 1024;      // columns sorted by length together (runtime knob "spmv_window")
// ------------------------------------------------- packed master columns (expand.cu)
//   MColsHeader                64 B
//   int64  colStart[m + 1]     offsets into entries
//   uint64 hash[m]             fingerprint of (block, x-space indices) for duplicate detection
//   MColEntry entries[cap]     (master row ascending per column, value)
struct MColsHeader {
    uint32_t magic;  // 'DIPM'
    uint32_t flags;  // bit 0: a column touched more rows than fit the staging area; bit 1: capacity
    int32_t nCols;
    int32_t pad;
    int64_t nNnz;
    int64_t reserved[5];
};
static_assert(sizeof(MColsHeader) == 64, "header is 64 bytes");
constexpr uint32_t kMColsMagic = 0x4d504944u;
struct MColEntry {
    int32_t row;
    int32_t pad;
    double val;
};
static_assert(sizeof(MColEntry) == 16, "entry is 16 bytes");

// ------------------------------------------------------------------ DP geometry
//
// A knapsack's DP row of (capacity+1) cells is owned by T threads, S consecutive
// cells each: thread t holds cells [t*S, t*S+S) in registers.
struct DpGeom {
    int T;         // threads per knapsack (CTA size, multiple of 32)
    int S;         // cells per thread (odd)
    int items;     // items advanced per barrier (1 or 2)
    int rowCells;  // T*S >= maxCapacity+1
    int chunkCols; // columns staged per TMA chunk
    int guard;     // -inf cells in front of each row buffer (0 = predicated reads)
    int fuse;      // 1: backtrack + filter run in the DP kernel's tail, decision bits stay in shared memory
    int big;       // 1: the row does not fit one CTA (> 12 800 cells): bits [item][cell/32] (bigdp.cu)
    int cluster;   // big: CTAs of the thread-block cluster that holds the row in distributed shared memory (S cells per thread); 0 = rows in global memory
    size_t smemBytes;
};

// ---------------------------------------------------------------------- context
struct Multi;  // multi.cu

// rf != NULL (fused shapes): the kernel forms the reduced costs of its blocks' columns itself -- stage (a) without a
// launch of its own -- from a CSC (start, idx, val) whose idx index uVec (vector v at uVec + v*uStride), writes
// them to dRedCost and reads alpha_b = uVec[alphaIdx[b]]
struct RcFuse {
    const int32_t *start, *hBlkEnt, *idx;  // hBlkEnt[b] (HOST array, m + 1): first stored entry of GLOBAL block b
    const double *val, *uVec;
    const int32_t *alphaIdx;
    int64_t uStride;
    int phase1;
    // co-resident grids: the kernel fetches u[support] from host memory itself (KnapArgs::pullSrc)
    const double *pullSrc = nullptr;
    const int32_t *pullIdx = nullptr;
    int64_t pullStride = 0;
    int pullN = 0;
};


struct Ctx {
    // non-NULL: this object is only the FRONT of a multi-device context (dipgpu_create_multi): it owns no device
    // memory, every entry point is routed to the per-device contexts behind it (multi.cu)
    Multi *multi = nullptr;
    int device = -1;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // core matrix (host copy in CSR for appends; device copy transposed)
    int R = 0, N = 0, nBaseRows = 0, numConvex = 0;
    std::vector<int64_t> hRowStart;
    std::vector<int32_t> hColInd;
    std::vector<double> hVal;
    bool haveCore = false;
    int64_t nnz = 0;
    int64_t *dColStart = nullptr;   // N+1, CSC of A''
    int32_t *dRowMaster = nullptr;  // nnz, MASTER dual index of each entry's row
    double *dVal = nullptr;         // nnz
    int spmvLanes = 0;              // 0 = auto (stream kernel when every column fits a tile)
    int32_t *dColStart32 = nullptr; // N+1 (only when nnz < 2^31)
    uint32_t *dTileRec = nullptr;   // nSpmvTiles x {byte offset / 16 into dSpmvBlob, bytes} (uint2)
    unsigned char *dSpmvBlob = nullptr;  // self-contained jagged tiles (SpmvTileHeader)
    int nSpmvTiles = 0;             // 0 = stream kernel not applicable
    bool tilesDirty = false;        // the tiles are rebuilt by the next sweep (ensureSpmvTiles)
    // columns the sweep covers: all of them, or (a block shard of a multi-device context) the shard's columns
    bool shardCols = false;
    int tileColFirst = 0, tileColEnd = -1;  // -1 = N
    int spmvStageBytes = 0;         // largest tile, bytes
    int32_t *dSpmvWarpStart = nullptr;  // launch warp -> first tile of its contiguous run (W + 1)
    int spmvMode = 0, spmvNst = 2, spmvWarps = 8, spmvGrid = 1, spmvUBytes = 0, optSpmvWindow = 0;  // launch shape (planSpmvLaunch)
    size_t spmvSmem = 0;
    int optSpmvWarps = 0, optSpmvUSmem = -1, optSpmvChunk = 0, optSpmvStages = 0, optSpmvMaxCols = 0;  // diagnostics knobs
    std::vector<double> hObj;       // host copy of the objective (tiles carry their columns' entries)
    void *spmvFn = nullptr;                   // kernel whose shared-memory limit is already set
    size_t spmvFnSmem = 0;
    int spmvClusters = 0;                     // resident CTA pairs of the cluster variant
    uint16_t *dRowMaster16 = nullptr;         // rowMaster as uint16 when R + numConvex <= 65536

    // Rows appended since the transposed structures (CSC, tiles) were last built -- dipgpu_append_core_rows on a large
    // model (DecompConstraintSet::appendRow, Dip/src/DecompConstraintSet.h:143-163): rows [0, baseR) are in the CSC and
    // the tiles, rows [baseR, R) in a small CSC of their own.  They are the LAST rows, i.e. the FIRST terms of every
    // column's sum in the reference's scatter order (rows R-1..0): redcost_delta_kernel forms those leading partial
    // sums (dSpmvInit; +0 for the untouched columns) and the stream kernel starts from them -- the same additions in
    // the same order as after a full rebuild.  mergeDelta() folds them in (any re-tiling, or the delta growing large).
    int baseR = 0;
    int64_t baseNnz = 0;
    int nDeltaCols = 0;               // columns with an entry in an appended row
    int64_t deltaNnz = 0;
    int32_t *dDeltaCol = nullptr;     // nDeltaCols: the columns
    int32_t *dDeltaStart = nullptr;   // nDeltaCols + 1
    int32_t *dDeltaRow = nullptr;     // master dual index of the entries, descending rows within a column
    double *dDeltaVal = nullptr;
    size_t deltaCapCols = 0, deltaCapNnz = 0;
    double *dSpmvInit = nullptr;      // initVecs x N leading partial sums (zero outside the delta's columns)
    int spmvInitVecs = 0;
    int spmvULen = 0;                 // dual entries the tiles' rows index (baseR + numConvex): what the u-in-smem mode stages
    bool spmvRow16 = false;           // the tiles store uint16 row ids
    int64_t optDeltaMinNnz = (int64_t)1 << 20;  // option "append_delta_min_nnz": smaller models are simply rebuilt

    // row-major copy of A'' (expand.cu re-walks touched rows in the reference's storage order); allocated with slack so
    // that appended rows are copied behind the old ones
    int32_t *dCsrRowStart = nullptr;  // R+1 (only when nnz < 2^31)
    int32_t *dCsrColInd = nullptr;
    double *dCsrVal = nullptr;
    size_t csrCapRows = 0, csrCapNnz = 0;
    int mcolsM = 0;                   // block count the packed master-column buffer was laid out for
    // packed master columns of the last expansion
    void *dMCols = nullptr, *hMCols = nullptr;
    void *dMColsHost = nullptr;        // device address of hMCols (mapped): the expansion kernel mirrors its output there
    unsigned int *hExpFlag = nullptr, *dExpFlagHost = nullptr;   // the expansion's done word (as Ctx::hDoneFlag for the round)
    unsigned int expSerialFlag = 0;
    bool expandFlagArmed = false, wantExpandMirror = false;
    size_t mcolsBytes = 0, mcolsOffColStart = 0, mcolsOffHash = 0, mcolsOffEntries = 0;
    int64_t mcolsCapEntries = 0;
    unsigned long long *dExpPub = nullptr;
    size_t expSmem = 0;
    std::vector<int64_t> hColStart;  // host copy of the CSC column starts
    int expMaxTouched = 0;           // bound on the rows one column can touch
    bool lastRoundOnHost = false;  // hResult holds the header of the last round
    // every enqueued round (and every change of the "last round": dipgpu_select_batch_result) gets a serial;
    // an expansion remembers the round it belongs to, so that stale master columns are never appended to the pool
    uint64_t roundSerial = 0, expandSerial = ~0ull;

    double *dObj = nullptr;  // c[N]
    size_t objLen = 0;       // entries of dObj
    bool haveObj = false;

    // blocks
    int m = 0;              // all blocks
    int nBlockCols = 0;     // blockColStart[m]
    bool haveBlocks = false;
    int profitMode = 0;
    // block kind: 0 = 0-1 knapsack (GAP, CSP), 1 = multiple-choice knapsack (MMKP: groups of columns),
    // 2 = integer (unbounded) knapsack with an optional use column (cutting-stock pricing, intknap.cu)
    int blockKind = 0;
    std::vector<int32_t> hUseCol;
    int32_t *dUseCol = nullptr;         // m (blockKind 2; NULL: no block has a use column)
    int32_t *dCandStart = nullptr;      // m + 1 (blockKind 2): first word of block b's candidate pairs in dCandInd
    size_t ikSmem = 0;
    std::vector<int32_t> hBlockGroupStart, hGroupColStart;
    int32_t *dBlockGroupStart = nullptr, *dGroupColStart = nullptr;
    size_t mckpSmem = 0;
    int mckpMaxWeight = 0;              // largest column weight of the multiple-choice blocks (guard cells of the DP rows)
    std::vector<int32_t> hBlockColStart, hCapacity;
    int32_t *dBlockColStart = nullptr;  // m+1
    int32_t *dWeight = nullptr;         // nBlockCols (+pad)
    int32_t *dCapacity = nullptr;       // m
    int64_t *dBitsOff = nullptr;        // m+1 word offsets into dBits (shard-local)
    uint32_t *dBits = nullptr;          // decision bits [word][cell]
    int64_t *dBigRowOff = nullptr;      // big rows (geom.big): local block -> first cell of its two global-memory rows
    double *dBigRows = nullptr, *dBigItemP = nullptr;
    size_t bitsWords = 0;
    int32_t *dItemW = nullptr;          // per column slot: weight of compacted item k of block b
    int32_t *dItemCol = nullptr;        // per column slot: local column of compacted item
    int32_t *dNItems = nullptr;         // m
    int32_t *dScale = nullptr, *dShortcut = nullptr, *dSel = nullptr;  // Pisinger mode, per local block
    double *dZShort = nullptr;
    int32_t *dCandInd = nullptr;        // per column slot: candidate column indices (ascending)
    int32_t *dScratch = nullptr;        // per column slot: descending taken list
    // node bounds of the columns (DecompAlgo::setSubProbBounds): dipgpu_set_col_bounds
    bool haveFix = false;
    uint8_t *dFix = nullptr;            // per column: bit 0 = fixed to 1, bit 1 = fixed to 0
    int32_t *dCapEff = nullptr;         // per block: capacity minus the fixed-to-1 weights (< 0: they do not fit)
    double *dRcEff = nullptr;           // large-shard path: reduced costs with the fixed columns masked to +0
    std::vector<int32_t> hWeight;
    std::vector<int32_t> hCapEff;       // host copy of dCapEff
    int maxCapacity = 0;       // of the shard
    int maxBlockCols = 0;      // of the shard
    int maxUsableWeight = 0;   // of the shard: max weight <= its block's capacity
    std::vector<int32_t> hMaxUsableW;  // per block
    int optThreads = 0, optCellsPerThread = 0, optItemsPerStep = 0, optNoGuard = 0, optNoPrune = 0, optPdl = 0;
    void *dpFnConfigured = nullptr;
    void *clusterFnConfigured = nullptr;
    bool clusterUnavailable = false;  // the device cannot place the planned cluster: rows in global memory instead
    int optClusterDebug = 0;       // option "dp_cluster_debug" (timing experiments, wrong results): see ClusterDpArgs::dbg
    int optClusterMaxCells = 0;    // option "dp_cluster_max_cells": cap on the cells per thread (smaller CTAs, larger clusters, two CTAs per SM)
    int optClusterThreads = 1024;  // option "dp_cluster_threads": 1024 (<= 13 cells per thread) or 512 (<= 26)
    int optNoCluster = 0;   // option "dp_no_cluster": rows beyond one CTA in global memory even when a cluster would hold them
    size_t dpFnSmem = 0;
    DpGeom geom{};

    // shard
    int firstBlock = 0, nLocal = 0;
    unsigned long long *dDbgTimes = nullptr;  // diagnostics (option dp_debug_times): 8 stamps per CTA
    unsigned long long *dPubWord = nullptr, *dPubCells = nullptr;  // fused DP kernel: per-block publication
    // fused DP kernel: [0] ticket counter, [1] launch epoch, [2] CTAs finished (KnapArgs::ticket / epochCtr / doneCtr);
    // [4..5]: the expansion kernel's 64-bit (epoch << 32 | ticket)
    unsigned int *dSyncWords = nullptr;
    unsigned long long *dExpTicket = nullptr;
    int dpCoResident = 0;                   // CTAs of the configured DP kernel that fit the GPU at once
    unsigned int *dPullCtr = nullptr;       // 64 per-vector arrival counters of the in-kernel dual upload
    unsigned int *dPullCtrBatch = nullptr;  // batchCap of them for batched launches (ensureBatch)
    int dpResidentBatch = 0;                // CTAs of the batch build (tickets) resident at once
    bool optBatchPull = true;               // option "batch_pull": ticketed batch launches fetch u[support] themselves
    int optBatchPullMin = 24;               // ... from this many vectors per launch on (16: 245 vs 237 us, 64: 725 vs 890 us per call)
    RcFuse batchPull;                       // (feedBatchDuals -> enqueueBatch: where the batch's kernel pulls its duals from)
    bool optPull = true;                    // option "pull_duals": co-resident fused launches fetch u[support] from host memory themselves
    // reduced costs formed inside the fused DP kernel (option "fuse_redcost", default on): the CSC with int32 row ids
    // is dColStart32 / dRowMaster / dVal; with a dual support, a second CSC holding only the entries of rows that can
    // carry a dual, indexed by SLOT in the compacted dual vector u[support] (dUCompact; slot ns = a zero)
    bool optFuseRc = true;
    bool optNoCoop = false;  // diagnostics (option "dp_coop" = 0): tickets even when the grid is co-resident
    int32_t *dCStart = nullptr, *dCIdx = nullptr, *dCRow = nullptr;  // (dCRow: the same entries' master rows, for dense vectors)
    double *dCVal = nullptr;
    int32_t *dAlphaIdxDense = nullptr, *dAlphaIdxCompact = nullptr;  // per block: where its convexity dual sits
    int32_t *dBlkEntFull = nullptr, *dBlkEntCompact = nullptr;       // per block (m + 1): its first stored entry in either CSC
    double *dUCompact = nullptr;      // (ns + 16) doubles, single rounds
    double *dUBatchCompact = nullptr; // batchCap x cStride doubles
    size_t cStride = 0;
    bool compactValid = false;        // the compact CSC matches the current core / support / blocks
    std::vector<int32_t> hBlkEntFull, hBlkEntCompact;  // host copies (which blocks fit the kernel's staging areas)
    bool dualsCompact = false;        // dUCompact holds the duals of the last call (dU may not)
    uint64_t fusedLaunches = 0;             // the publication words are cleared every 2^24 fused launches
    unsigned int *dDoneCounter = nullptr;  // ticket of the fused backtrack+filter kernel
    // -1 auto (2 when nLocal <= kFuseFilterMaxBlocks, else 0); 0 three kernels; 1 filter in the
    // backtrack kernel's tail; 2 backtrack + filter in the DP kernel's tail
    int optFuseFilter = -1;

    // duals / reduced costs
    double *dU = nullptr;        // uLen capacity
    size_t uCap = 0;
    double *dRedCost = nullptr;  // N (+pad)
    double *dAlpha = nullptr;    // m

    // results
    void *dResult = nullptr;
    void *hResult = nullptr;  // pinned mirror (mapped: dResultHost is its device address)
    void *dResultHost = nullptr;
    bool wantMirror = false;      // the caller of enqueueSolve will read the result on the host right after
    bool resultMirrored = false;  // the last fused launch wrote hResult itself
    bool optZeroCopy = true;
    size_t lastMColsBytes = 0;    // bytes of packed master columns the last expansion produced (sizes the next speculative D2H copy)
    bool optMirrorAlways = false;
    // non-NULL: the fused kernel mirrors the packed result THERE (device-visible: a peer device's gather buffer)
    // instead of into hResult
    void *mirrorBase = nullptr;
    // Completion of a mirrored single round without cudaStreamSynchronize: the last CTA of the fused kernel to finish
    // stores the round's serial into this word of mapped host memory (behind a system-scope fence: every CTA's part
    // of the mirrored result is visible first) and the host entry point spins on it -- the kernel's teardown and the
    // driver's completion handshake are off the caller's critical path (option "flag_wait", default on)
    unsigned int *hDoneFlag = nullptr, *dDoneFlagHost = nullptr;
    unsigned int doneSerial = 0;
    bool flagArmed = false, wantFlag = false, optFlagWait = true;
    long long flagWaits = 0, flagTimeouts = 0;   // rounds completed through the done word / waits that fell back to the synchronise
    bool hostTiming = false;      // option "host_timing": wall-clock stamps of the last dipgpu_price_round (diagnostics)
    double hostUs[4] = {0, 0, 0, 0};  // enqueue done, result seen, unpacked, DP launch call begins (us after entry)
    long long hostT0 = 0;             // entry time (steady clock, ns)
    bool optGatherPinned = true;  // sparse dual upload: read u[support] out of the caller's buffer when it is device-mapped
    size_t resultBytes = 0;
    size_t lastResultBytes = 0;  // bytes of the last packed result (sizes the next speculative D2H copy)
    int64_t indCap = 0;
    ResultLayout layout{};

    // batched dual vectors (dipgpu_price_rounds_batch / dipgpu_price_round_stab*): per-vector copies of
    // the dual vector, the reduced costs, the packed result and the per-block publication / Pisinger arrays
    int batchCap = 0;            // vectors the buffers below hold
    size_t uStride = 0;          // doubles between dual vectors (even, >= uLen)
    size_t rcStride = 0;         // doubles between reduced-cost vectors (even, >= N + 16)
    size_t resStride = 0;        // bytes between packed results (multiple of 16, >= resultBytes)
    double *dUBatch = nullptr, *dRedCostBatch = nullptr;
    void *dResultBatch = nullptr, *hResultBatch = nullptr, *dResultBatchHost = nullptr;  // (device address of the mapped host mirror)
    unsigned long long *dPubWordBatch = nullptr;
    int32_t *dScaleB = nullptr, *dShortcutB = nullptr, *dSelB = nullptr;
    double *dZShortB = nullptr;
    double *dAlphaLadder = nullptr;  // smoothing parameters of a ladder (device)
    // dual stabilisation (DecompAlgoPC::adjustMasterDualSolution): m_dual, the stability centre
    double *dCenter = nullptr;
    size_t centerLen = 0;
    bool haveCenter = false, haveST = false;  // haveST: dUBatch[0] holds the last smoothed vector
    bool dualsResident = false;               // dU holds the dual vector of the last host-buffer call
    // rows of the master dual vector that can be nonzero (dipgpu_set_dual_support): only those cross PCIe
    std::vector<int32_t> hSupport;
    int32_t *dSupport = nullptr;
    uint32_t *dSupportBits = nullptr;  // one bit per master row (pool.cu)
    size_t poolFnSmem[2] = {0, 0};
    size_t poolFnSmem2[4] = {0, 0, 0, 0};
    int optPoolChunk = 256, optPoolWarps = 24;   // options "pool_chunk" / "pool_warps": window size and warps per CTA of the pass over the compacted copy
    int optPoolKernel = 1;             // option "pool_kernel": 1 persistent (default), 2 persistent without the in-SM gather, 0 one CTA per 16 tiles
    uint16_t *dSupportRank = nullptr;  // per bitmap word: support rows below it (supports of < 65 536 rows)
    double *hSupVals = nullptr;  // mapped pinned staging the scatter kernel reads: supVecCap vectors of hSupport.size()
    int supVecCap = 0;
    bool dUClean = false, dUBatchClean = false;       // the device vectors are zero outside the support
    // waiting-column pool (DecompVarPool): master columns in device memory
    int64_t poolCols = 0, poolNnz = 0, poolColCap = 0, poolNnzCap = 0;
    int64_t *dPoolStart = nullptr;   // poolCols + 1
    int32_t *dPoolRow = nullptr;
    double *dPoolVal = nullptr, *dPoolOrigCost = nullptr, *dPoolRedCost = nullptr;
    int32_t *dPoolFlag = nullptr;    // [0]: some redCost <= -1e-10
    // With a dual support declared the re-pricing pass reads a COMPACTED copy of the pool that holds only the entries
    // of rows that can carry a dual (GAP master columns: one third of the entries).  The others contribute val * 0 =
    // +-0 to their column's sum, and a running sum that starts at +0 is never -0 (x + (-x) rounds to +0), so adding
    // +-0 is the identity bit for bit: leaving those entries out changes nothing.  Rebuilt (count, scan, fill) when the
    // pool or the support has changed since the last pass (option "pool_compact", default on).
    uint64_t poolSerial = 1, supportSerial = 1, poolCompactPool = 0, poolCompactSupport = 0;
    int64_t *dPoolCStart = nullptr, *dPoolCCount = nullptr;
    int32_t *dPoolCRow = nullptr;
    double *dPoolCVal = nullptr;
    int64_t poolCNnz = 0, poolCColCap = 0, poolCNnzCap = 0;
    void *dPoolScanTmp = nullptr;
    size_t poolScanTmpBytes = 0;
    bool optPoolCompact = true;

    // pinned staging for pageable inputs
    double *hStage = nullptr;
    size_t hStageBytes = 0;

    // counters / timing
    int64_t launches = 0, h2d = 0, d2h = 0;
    bool stageTiming = false;
    cudaEvent_t ev[8] = {};
    // batches priced in chunks: the gathers of the chunks' duals run on a second stream, chunk k's DP launch waits
    // for event k (batchPipelined in dipgpu_api.cu); option "batch_pipeline" (default 0: measured, gains <= 5 % on
    // batches of 64+ vectors and loses on smaller ones -- the gathers get no faster beside a DP launch)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t evChunk[9] = {};
    bool optBatchPipeline = false;
    float stageMs[6] = {0, 0, 0, 0, 0, 0};
};

// Where a context takes the master dual vectors of a call from (feedDuals, dipgpu_api.cu).
struct DualFeed {
    const double *host = nullptr;    // the caller's host vectors, vector v at host + v*uLen (pageable or pinned)
    const double *staged = nullptr;  // device-visible u[support], vector v at staged + v*stagedStride (needs a dual support)
    int64_t stagedStride = 0;
    const double *full = nullptr;    // device-visible whole vectors (mapped host memory, a peer device), v at full + v*fullStride
    int64_t fullStride = 0;
};
int feedDuals(Ctx *c, const DualFeed &f, int32_t uLen, int nVec, double *dDst, size_t dstStride, bool *clean, cudaStream_t st);
const double *deviceVisible(const double *u);
void gatherSupport(const std::vector<int32_t> &support, const double *u, int32_t uLen, int nVec, double *dst);

int fail(Ctx *c, int code, const char *fmt, ...);
// Opt-in dynamic shared memory of a kernel: the limit is a property of (function, device), shared by every
// context of the process -- it is raised ONCE to what the device allows (227 KiB minus the kernel's static
// shared memory) and never lowered, so contexts with different shapes cannot undercut each other.
int ensureSmemLimit(Ctx *c, const void *fn, size_t bytes);
int cudaFail(Ctx *c, cudaError_t e, const char *what);

#define DIPGPU_CUDA(c, call)                                            \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::dipgpu::cudaFail((c), e__, #call); \
    } while (0)

// ------------------------------------------------ shared host-side pieces (dipgpu_api.cu)
int bind(Ctx *c);
int setBlockRange(Ctx *c, int first, int count);
int checkRoundReady(Ctx *c, const char *who, int32_t uLen, int32_t phase);
int ensureDualBuffer(Ctx *c, int32_t uLen);
int ensureBatch(Ctx *c, int nVec);
int enqueueRound(Ctx *c, const DualFeed *feed, int32_t uLen, int32_t phase, double eps, bool mirror, cudaStream_t st, bool needAllRedCosts = false);
int enqueueRoundDev(Ctx *c, const double *uDev, int32_t uLen, int32_t phase, double eps, bool mirror, cudaStream_t st);
bool rcFuseEligible(const Ctx *c);
int rcFuseStageCap(const Ctx *c);  // knapsack.cu: stored entries of a block the fused kernel can stage
int enqueueBatch(Ctx *c, int nVec, int phase, double eps, cudaStream_t st, bool compact);
int fetchResult(Ctx *c, cudaStream_t st);
int fetchBatchPacked(Ctx *c, int nVec, cudaStream_t st);
int fetchBatch(Ctx *c, int nVec, dipgpu_cols *outs, cudaStream_t st);
void collectStageMs(Ctx *c);
int unpack(Ctx *c, const void *packed, size_t bytes, dipgpu_cols *out);
int batchCore(Ctx *c, int nVec, const DualFeed &f, int32_t uLen, int32_t phase, double eps, dipgpu_cols *outs);
int ladderCore(Ctx *c, const DualFeed &f, int32_t uLen, const double *al, int nSteps, int32_t phase, double eps, dipgpu_cols *outs);

// ------------------------------------------------ multi-device context (multi.cu)
inline bool isMulti(const dipgpu_ctx *ctx) { return ctx && reinterpret_cast<const Ctx *>(ctx)->multi != nullptr; }
int multiSetShardRounds(dipgpu_ctx *ctx, int64_t value);
dipgpu_ctx *multiPrimary(dipgpu_ctx *ctx);  // the primary device's whole-model context (serves what is not sharded)
dipgpu_ctx *multiForward(dipgpu_ctx *ctx);  // same, and "the last round" now lives there
int multiDestroy(dipgpu_ctx *ctx);
// fn on the whole-model context of every device (and, alsoShards, on its block-shard context): the model calls
int multiForAll(dipgpu_ctx *ctx, int (*fn)(dipgpu_ctx *sub, void *arg), void *arg, bool alsoShards, bool modelChanged);
void multiClearSupport(dipgpu_ctx *ctx);
int multiSolveBlocks(dipgpu_ctx *ctx, const double *redCostX, const double *alpha, double eps, dipgpu_cols *out);
int multiEnqueueSolve(Ctx *c, double eps, cudaStream_t st, bool timed);  // dipgpu_api.cu: stages (b)+(c) on dRedCost / dAlpha, mirrored
int multiSetBlocks(dipgpu_ctx *ctx);  // after the blocks are known everywhere: cut the block shards
int multiSetDualSupport(dipgpu_ctx *ctx, int32_t nIdx, const int32_t *idx);
int multiPriceRound(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double eps, dipgpu_cols *out, double *redCostXOpt);
int multiPriceRoundsBatch(dipgpu_ctx *ctx, int32_t nVec, const double *U, int32_t uLen, int32_t phase, double eps, dipgpu_cols *outs);
int multiStabLadder(dipgpu_ctx *ctx, const double *uRM, int32_t uLen, const double *al, int nSteps, int32_t phase, double eps, dipgpu_cols *outs);
int multiStabAccept(dipgpu_ctx *ctx, int32_t step);
int multiGetSmoothedDuals(dipgpu_ctx *ctx, int32_t step, double *dualST, int32_t uLen);
int multiSelectBatchResult(dipgpu_ctx *ctx, int32_t which);
int multiCalcRedCost(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double *redCostX);
int multiPriceRoundDev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase, double eps);
int multiResultDev(dipgpu_ctx *ctx, void **devPtr, size_t *bytes);
int multiGetCounters(const dipgpu_ctx *ctx, int64_t *launches, int64_t *h2d, int64_t *d2h);
int multiLastStageMs(const dipgpu_ctx *ctx, float *ms6);
int multiExpandColumns(dipgpu_ctx *ctx, double tolZero, dipgpu_mcols *out);

// ------------------------------------------------ Pisinger-mode pre-processing
struct PrepArgs {
    const double *redCost;
    const int32_t *weight;
    const int32_t *blockColStart;
    const int32_t *capacity;
    const uint8_t *fix; // node bounds (else NULL): per column bit 0 = fixed to 1, bit 1 = fixed to 0
    int firstBlock;
    long long rcStride; // batched dual vectors: vector v = blockIdx.y reads redCost + v*rcStride, writes slot v*gridDim.x + lb
    int32_t *scale;     // local block -> calcScaleFactor result
    int32_t *shortcut;  // local block -> 0 solve by DP, 1 take every active item, 2 explicit set (sel)
    int32_t *sel;       // 2 per local block: chosen local columns ascending, -1 = none
    double *zShort;     // local block -> optimum of a shortcut block (scaled-integer units)
};

// ------------------------------------------------------------------- launchers
// redcost.cu -- stage (a)
int launchRedCost(Ctx *c, const double *dU, int phase, cudaStream_t st);
// nVec dual vectors at once: vector v reads dU + v*uStride and writes dOut + v*outStride
int launchRedCostBatch(Ctx *c, const double *dU, int64_t uStride, double *dOut, int64_t outStride, int nVec, int phase,
                       cudaStream_t st);
int buildSpmvTiles(Ctx *c, const std::vector<int64_t> &colStart, const std::vector<int32_t> &rowMaster,
                   const std::vector<double> &val);
int ensureSpmvTiles(Ctx *c);  // dipgpu_api.cu
int mergeDelta(Ctx *c);       // dipgpu_api.cu: folds the rows appended since the last full build into the CSC / tiles
// pisinger.cu -- scale factor + trivial cases of GAP_KnapPisinger::solve (profit mode 1 only)
int launchKnapsackPrep(Ctx *c, const double *dRedCost, cudaStream_t st);
int launchKnapsackPrepBatch(Ctx *c, const double *dRedCost, int64_t rcStride, int nVec, cudaStream_t st);
// knapsack.cu -- stage (b)
int chooseDpGeom(Ctx *c, DpGeom *g, bool wantFuse);
int prepareKnapsackDp(Ctx *c);  // loads the chosen kernel, sets its shared-memory limit
int launchKnapsackDp(Ctx *c, const double *dRedCost, const double *dAlpha, double redCostEps, cudaStream_t st, const RcFuse *rf = nullptr);
int launchKnapsackDpBatch(Ctx *c, const double *dRedCost, int64_t rcStride, const double *dAlpha, int64_t alphaStride,
                          double redCostEps, int nVec, cudaStream_t st, const RcFuse *rf = nullptr, int vecBase = 0);
int launchBacktrackEmit(Ctx *c, const double *dRedCost, const double *dAlpha, double redCostEps, bool fuseFilter,
                        cudaStream_t st);
// bigdp.cu -- stage (b) for rows that do not fit one CTA
int launchBigDp(Ctx *c, const double *dRedCost, cudaStream_t st);
void planClusterDp(int64_t cells, int threads, int maxCells, int *clusterSize, int *S);  // bigdp.cu
// mckp.cu -- stage (b) for multiple-choice knapsack blocks (DP + backtrack; the filter kernels follow)
int launchMckp(Ctx *c, const double *dRedCost, const double *dAlpha, cudaStream_t st);
// intknap.cu -- stage (b) for integer knapsack blocks (table + counts + column; the filter kernels follow)
int launchIntKnap(Ctx *c, const double *dRedCost, const double *dAlpha, cudaStream_t st);
// filter.cu -- stage (c)
int launchFilter(Ctx *c, double redCostEps, cudaStream_t st);
// expand.cu -- master columns of the kept variables (the step after the round)
int launchExpandColumns(Ctx *c, double tolZero, int nKept, cudaStream_t st);
constexpr int kPoolPad = 512;  // entries allocated past the pool's capacity (aligned 512-entry windows are streamed whole)
// pool.cu -- waiting-column pool re-pricing (DecompVarPool::setReducedCosts) and Wentges smoothing
int launchPoolRedCost(Ctx *c, const double *dU, int feasible, cudaStream_t st);
int launchScatterDuals(Ctx *c, int n, const int32_t *dIdx, const double *dVals, int64_t valStride, bool gather, double *dDst,
                       int64_t dstStride, int nVec, cudaStream_t st);
int launchGatherDuals(Ctx *c, int n, const int32_t *dIdx, const double *dVals, int64_t valStride, double *dDst, int64_t dstStride,
                      int nVec, cudaStream_t st);
int launchSmoothDuals(Ctx *c, const double *dURM, const double *dCenter, const double *dAlphas, int nVec, int uLen,
                      double *dOut, int64_t outStride, cudaStream_t st);
// microbench.cu -- roofline denominators measured on the device
int measureSmemGbs(Ctx *c, double *gbs);

}  // namespace dipgpu

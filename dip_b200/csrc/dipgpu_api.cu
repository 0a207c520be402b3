// dipgpu_api.cu -- the C ABI of libdipgpu (include/dipgpu.h): context lifetime,
// model upload, and the orchestration of one pricing round
// (DecompAlgo::generateVars, Dip/src/DecompAlgo.cpp:4622-5260) as
//   H2D(u) -> reduced costs -> knapsack DP -> backtrack+emit -> filter -> D2H.
// No CPU fallback: every compute entry point needs a live sm_100 device.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>

#include "common.cuh"

namespace dipgpu {

int fail(Ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

int cudaFail(Ctx *c, cudaError_t e, const char *what)
{
    return fail(c, e == cudaErrorMemoryAllocation ? DIPGPU_ERR_NOMEM : DIPGPU_ERR_CUDA, "CUDA error %d (%s) at %s",
                (int)e, cudaGetErrorString(e), what);
}

int ensureSmemLimit(Ctx *c, const void *fn, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> limit;
    std::lock_guard<std::mutex> lk(mu);
    const std::pair<int, const void *> key(c->device, fn);
    auto it = limit.find(key);
    if (it == limit.end()) {
        cudaFuncAttributes at;
        DIPGPU_CUDA(c, cudaFuncGetAttributes(&at, fn));
        int optin = 0;
        DIPGPU_CUDA(c, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        const size_t maxDyn = (size_t)optin > at.sharedSizeBytes ? (size_t)optin - at.sharedSizeBytes : 0;
        if (maxDyn > 48 * 1024) DIPGPU_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxDyn));
        it = limit.emplace(key, std::max<size_t>(maxDyn, 48 * 1024 - std::min<size_t>(at.sharedSizeBytes, 48 * 1024))).first;
    }
    if (bytes > it->second)
        return fail(c, DIPGPU_ERR_UNSUPPORTED, "kernel needs %zu bytes of dynamic shared memory, the device allows %zu", bytes, it->second);
    return DIPGPU_OK;
}

template <typename T>
static int devAlloc(Ctx *c, T **p, size_t count)
{
    if (*p) {
        cudaFree(*p);
        *p = nullptr;
    }
    if (count == 0) count = 1;
    DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T)));
    return DIPGPU_OK;
}

template <typename T>
static void devFree(T **p)
{
    if (*p) cudaFree(*p);
    *p = nullptr;
}

static void freeBatch(Ctx *c);

int bind(Ctx *c)
{
    DIPGPU_CUDA(c, cudaSetDevice(c->device));
    return DIPGPU_OK;
}

// Transpose the host CSR of A'' into CSC, folding the master-dual row mapping (rows >= nBaseRows sit
// numConvex further down, DecompAlgo.cpp:4595-4597).  Rows [0, baseR): the rows appended since the last full
// build live in the delta (Ctx::baseR).
void transposeCore(const Ctx *c, std::vector<int64_t> &colStart, std::vector<int32_t> &rowMaster, std::vector<double> &val)
{
    const int R = c->baseR, N = c->N;
    const int64_t nnz = c->hRowStart[R];
    colStart.assign((size_t)N + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) colStart[(size_t)c->hColInd[k] + 1]++;
    for (int j = 0; j < N; ++j) colStart[j + 1] += colStart[j];
    std::vector<int64_t> fill(colStart.begin(), colStart.end() - 1);
    rowMaster.resize((size_t)std::max<int64_t>(nnz, 1));
    val.resize((size_t)std::max<int64_t>(nnz, 1));
    // rows descending: each column's entries are stored in the order the reference's scatter
    // adds them (CoinPackedMatrix::transposeTimes on a row-ordered matrix, rows R-1..0)
    for (int i = R - 1; i >= 0; --i) {
        const int32_t rm = i < c->nBaseRows ? i : i + c->numConvex;
        for (int64_t k = c->hRowStart[i]; k < c->hRowStart[i + 1]; ++k) {
            const int64_t dst = fill[c->hColInd[k]]++;
            rowMaster[dst] = rm;
            val[dst] = c->hVal[k];
        }
    }
}

// The warp tiles of the stream kernel (redcost.cu) are built when a sweep first needs them: the model calls
// (set_core_csr, set_objective, append_core_rows, a column range) each only mark them stale.
static int uploadCore(Ctx *c);

int ensureSpmvTiles(Ctx *c)
{
    if (!c->tilesDirty) return DIPGPU_OK;
    if (c->baseR != c->R) {  // re-tiling anyway: the appended rows go into the tiles
        const int rc = uploadCore(c);
        if (rc) return rc;
    }
    std::vector<int64_t> colStart;
    std::vector<int32_t> rowMaster;
    std::vector<double> val;
    transposeCore(c, colStart, rowMaster, val);
    const int rc = buildSpmvTiles(c, colStart, rowMaster, val);
    if (rc == DIPGPU_OK) c->tilesDirty = false;
    return rc;
}

// Full (re)build of the device copies of A'' from the host CSR: every row, also the ones appended since the last
// build (the delta is folded in and dropped).
static int uploadCore(Ctx *c)
{
    const int R = c->R, N = c->N;
    const int64_t nnz = c->hRowStart[R];
    c->baseR = R;
    c->baseNnz = nnz;
    c->nDeltaCols = 0;
    c->deltaNnz = 0;
    devFree(&c->dSpmvInit);  // (stale leading sums must not survive the merge)
    c->spmvInitVecs = 0;
    std::vector<int64_t> colStart;
    std::vector<int32_t> rowMaster;
    std::vector<double> val;
    transposeCore(c, colStart, rowMaster, val);
    c->nnz = nnz;
    c->hColStart = colStart;
    int rc;
    c->tilesDirty = true;
    c->nSpmvTiles = 0;
    devFree(&c->dColStart32);
    devFree(&c->dRowMaster16);
    if (nnz < (int64_t)1 << 31 && N > 0) {
        std::vector<int32_t> cs32((size_t)N + 1);
        for (int j = 0; j <= N; ++j) cs32[j] = (int32_t)colStart[j];
        if ((rc = devAlloc(c, &c->dColStart32, (size_t)N + 1 + 8))) return rc;  // (+8: bulk copies read whole 16-byte pieces)
        DIPGPU_CUDA(c, cudaMemset(c->dColStart32, 0, sizeof(int32_t) * ((size_t)N + 1 + 8)));
        DIPGPU_CUDA(c, cudaMemcpy(c->dColStart32, cs32.data(), sizeof(int32_t) * cs32.size(), cudaMemcpyHostToDevice));
        if ((int64_t)R + c->numConvex <= 65536) {
            std::vector<uint16_t> r16((size_t)nnz + 16, 0);
            for (int64_t k = 0; k < nnz; ++k) r16[k] = (uint16_t)rowMaster[k];
            if ((rc = devAlloc(c, &c->dRowMaster16, r16.size()))) return rc;
            DIPGPU_CUDA(c, cudaMemcpy(c->dRowMaster16, r16.data(), sizeof(uint16_t) * r16.size(), cudaMemcpyHostToDevice));
        }
    }
    // row-major copy for expand.cu
    devFree(&c->dCsrRowStart); devFree(&c->dCsrColInd); devFree(&c->dCsrVal);
    c->csrCapRows = c->csrCapNnz = 0;
    if (nnz < (int64_t)1 << 31) {
        std::vector<int32_t> rs32((size_t)R + 1);
        for (int i = 0; i <= R; ++i) rs32[i] = (int32_t)c->hRowStart[i];
        // (slack: rows appended later are copied behind these, appendCsrRows)
        c->csrCapRows = (size_t)R + (size_t)R / 16 + 64;
        c->csrCapNnz = (size_t)nnz + (size_t)nnz / 16 + 1024;
        if ((rc = devAlloc(c, &c->dCsrRowStart, c->csrCapRows + 1))) return rc;
        if ((rc = devAlloc(c, &c->dCsrColInd, c->csrCapNnz + 1))) return rc;
        if ((rc = devAlloc(c, &c->dCsrVal, c->csrCapNnz + 1))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dCsrRowStart, rs32.data(), sizeof(int32_t) * rs32.size(), cudaMemcpyHostToDevice));
        if (nnz > 0) {
            DIPGPU_CUDA(c, cudaMemcpy(c->dCsrColInd, c->hColInd.data(), sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(c->dCsrVal, c->hVal.data(), sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
        }
    }
    if ((rc = devAlloc(c, &c->dColStart, (size_t)N + 1))) return rc;
    if ((rc = devAlloc(c, &c->dRowMaster, (size_t)nnz + 8))) return rc;
    if ((rc = devAlloc(c, &c->dVal, (size_t)nnz + 8))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dColStart, colStart.data(), sizeof(int64_t) * ((size_t)N + 1), cudaMemcpyHostToDevice));
    if (nnz > 0) {
        DIPGPU_CUDA(c, cudaMemcpy(c->dRowMaster, rowMaster.data(), sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice));
        DIPGPU_CUDA(c, cudaMemcpy(c->dVal, val.data(), sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice));
    }
    return DIPGPU_OK;
}

int mergeDelta(Ctx *c)
{
    if (c->baseR == c->R) return DIPGPU_OK;
    return uploadCore(c);
}

// The objective buffer covers at least n columns: zero-filled when no objective was set (phase 1 / solve_blocks
// on reduced costs only), grown -- old entries kept, zeros behind them -- when the model gained columns without
// a new set_objective.
static int ensureObjective(Ctx *c, int n)
{
    if (c->dObj && c->objLen >= (size_t)n) return DIPGPU_OK;
    double *nw = nullptr;
    int rc;
    if ((rc = devAlloc(c, &nw, (size_t)n + 16))) return rc;
    DIPGPU_CUDA(c, cudaMemset(nw, 0, sizeof(double) * ((size_t)n + 16)));
    if (c->dObj && c->objLen > 0) DIPGPU_CUDA(c, cudaMemcpy(nw, c->dObj, sizeof(double) * c->objLen, cudaMemcpyDeviceToDevice));
    devFree(&c->dObj);
    c->dObj = nw;
    c->objLen = (size_t)n;
    return DIPGPU_OK;
}

int setBlockRange(Ctx *c, int first, int count)
{
    if (!c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "set_knapsack_blocks has not been called");
    if (first < 0 || count < 0 || first + count > c->m)
        return fail(c, DIPGPU_ERR_INVALID, "block range [%d,%d) outside [0,%d)", first, first + count, c->m);
    c->firstBlock = first;
    c->nLocal = count;
    freeBatch(c);  // per-vector buffers are sized by the shard
    int maxCap = 0, maxCols = 0, maxW = 0;
    for (int b = first; b < first + count; ++b) {
        maxCap = std::max(maxCap, c->hCapacity[b]);
        maxCols = std::max(maxCols, c->hBlockColStart[b + 1] - c->hBlockColStart[b]);
        maxW = std::max(maxW, c->hMaxUsableW[b]);
    }
    c->maxCapacity = maxCap;
    c->maxBlockCols = maxCols;
    c->maxUsableWeight = maxW;
    int rc;
    std::vector<int64_t> off((size_t)count + 1, 0);
    if (c->blockKind == 1) {
        // multiple-choice blocks: no DP geometry; dBits holds the uint16 argmin of every (group, cell),
        // dBitsOff the first entry of each block's table
        c->geom = DpGeom{};
        for (int lb = 0; lb < count; ++lb) {
            const int G = c->hBlockGroupStart[first + lb + 1] - c->hBlockGroupStart[first + lb];
            off[lb + 1] = off[lb] + (int64_t)G * ((int64_t)c->hCapacity[first + lb] + 1);
        }
        c->bitsWords = (size_t)((off[count] + 1) / 2 + 1);  // uint16 entries in uint32 words
    } else if (c->blockKind == 2) {
        // integer knapsack blocks: no DP geometry; dBits is the global scratch of the tables (F, item, src: 16 bytes per
        // cell) for shards whose rows do not fit shared memory beside the items (intknap.cu: launchIntKnap)
        c->geom = DpGeom{};
        c->bitsWords = (size_t)std::max(count, 1) * ((size_t)maxCap + 1) * 4 + 16;
    } else {
    const bool wantFuse = count > 0 && (c->optFuseFilter < 0 ? count <= kFuseFilterMaxBlocks : c->optFuseFilter == 2);
    if ((rc = chooseDpGeom(c, &c->geom, wantFuse))) return rc;
    devFree(&c->dBigRowOff); devFree(&c->dBigRows); devFree(&c->dBigItemP);
    if (c->geom.big) {
        // rows in global memory (bigdp.cu): per block two rows of (capacity + 1) cells, decisions [item][cell / 32]
        std::vector<int64_t> rowOff((size_t)count + 1, 0);
        for (int lb = 0; lb < count; ++lb) {
            const int n = c->hBlockColStart[first + lb + 1] - c->hBlockColStart[first + lb];
            const int64_t cells = (int64_t)c->hCapacity[first + lb] + 1;
            off[lb + 1] = off[lb] + (int64_t)n * ((cells + 31) >> 5);
            rowOff[lb + 1] = rowOff[lb] + 2 * ((cells + 3) & ~(int64_t)3);
        }
        c->bitsWords = (size_t)off[count];
        if ((rc = devAlloc(c, &c->dBigRowOff, (size_t)count + 1))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dBigRowOff, rowOff.data(), sizeof(int64_t) * ((size_t)count + 1), cudaMemcpyHostToDevice));
        if ((rc = devAlloc(c, &c->dBigRows, (size_t)rowOff[count] + 8))) return rc;
        if ((rc = devAlloc(c, &c->dBigItemP, (size_t)c->nBlockCols + 16))) return rc;
    } else {
    if (count > 0 && (rc = prepareKnapsackDp(c))) return rc;
    // decision bits: ceil(n_b/32) words per cell, rowCells cells per block
    for (int lb = 0; lb < count; ++lb) {
        const int n = c->hBlockColStart[first + lb + 1] - c->hBlockColStart[first + lb];
        off[lb + 1] = off[lb] + (int64_t)((n + 31) / 32) * c->geom.rowCells;
    }
    c->bitsWords = (size_t)off[count];
    }
    }
    if ((rc = devAlloc(c, &c->dBitsOff, (size_t)count + 1))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dBitsOff, off.data(), sizeof(int64_t) * ((size_t)count + 1), cudaMemcpyHostToDevice));
    if ((rc = devAlloc(c, &c->dBits, c->bitsWords))) return rc;
    if ((rc = devAlloc(c, &c->dNItems, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dPubWord, (size_t)std::max(count, 1) * kPubWordStride))) return rc;
    if ((rc = devAlloc(c, &c->dPubCells, (size_t)count))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dPubWord, 0, sizeof(unsigned long long) * std::max<size_t>((size_t)count, 1) * kPubWordStride));
    if ((rc = devAlloc(c, &c->dScale, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dShortcut, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dSel, 2 * (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dZShort, (size_t)count))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dShortcut, 0, sizeof(int32_t) * std::max<size_t>((size_t)count, 1)));
    // packed result
    c->indCap = count > 0 ? (int64_t)(c->hBlockColStart[first + count] - c->hBlockColStart[first]) : 0;
    if (c->blockKind == 2) c->indCap = 2 * (c->indCap + count);  // (index, multiplicity) pairs, one more per block for the use column
    c->layout = ResultLayout::make(count, c->indCap);
    c->resultBytes = c->layout.bytes;
    c->lastResultBytes = 0;
    if (c->dResult) cudaFree(c->dResult);
    c->dResult = nullptr;
    DIPGPU_CUDA(c, cudaMalloc(&c->dResult, c->resultBytes));
    DIPGPU_CUDA(c, cudaMemset(c->dResult, 0, c->resultBytes));
    if (c->hResult) cudaFreeHost(c->hResult);
    c->hResult = c->dResultHost = nullptr;
    DIPGPU_CUDA(c, cudaHostAlloc(&c->hResult, c->resultBytes, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(c->hResult, 0, c->resultBytes);
    DIPGPU_CUDA(c, cudaHostGetDevicePointer(&c->dResultHost, c->hResult, 0));
    return DIPGPU_OK;
}

static int ensureSupportStaging(Ctx *c, int nVec);

// ---------------------------------------------------------------- reduced costs inside the fused DP kernel
//
// Small shards (the GAP rounds) run stages (a)+(b)+(c) as ONE kernel: every block CTA forms the reduced costs of its
// own columns from the CSC of A'' before it compacts its items (knapsack.cu, KnapArgs::rcIdx).  With a dual support
// the kernel reads a second CSC that holds only the entries of rows that can carry a dual, indexed by slot in the
// compacted dual vector u[support] -- on GAP-E 130 k of 384 k entries and a 34 KB vector instead of 2 MB.
bool rcFuseEligible(const Ctx *c)
{
    return c->optFuseRc && c->geom.fuse && c->blockKind == 0 && c->profitMode == DIPGPU_PROFIT_FP64 && c->haveCore &&
           c->dColStart32 != nullptr && c->nLocal > 0 && c->nBlockCols <= c->N && c->numConvex == c->m;
}

static int ensureRcFuse(Ctx *c)
{
    int rc;
    const int m = c->m;
    if (!c->dAlphaIdxDense || !c->compactValid) {
        std::vector<int32_t> ai((size_t)std::max(m, 1));
        for (int b = 0; b < m; ++b) ai[(size_t)b] = c->nBaseRows + b;  // alpha_b = u[nBaseRows + b]  (DecompAlgo.cpp:4820)
        if ((rc = devAlloc(c, &c->dAlphaIdxDense, ai.size()))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dAlphaIdxDense, ai.data(), sizeof(int32_t) * ai.size(), cudaMemcpyHostToDevice));
        std::vector<int32_t> be((size_t)m + 1);
        for (int b = 0; b <= m; ++b) be[(size_t)b] = (int32_t)c->hColStart[(size_t)c->hBlockColStart[(size_t)b]];
        if ((rc = devAlloc(c, &c->dBlkEntFull, be.size()))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dBlkEntFull, be.data(), sizeof(int32_t) * be.size(), cudaMemcpyHostToDevice));
        c->hBlkEntFull = be;
    }
    if (c->compactValid) return DIPGPU_OK;
    const size_t ns = c->hSupport.size();
    if (ns > 0) {
        const int uLen = c->R + c->numConvex;
        std::vector<int32_t> slot((size_t)uLen, -1);
        for (size_t k = 0; k < ns; ++k) slot[(size_t)c->hSupport[k]] = (int32_t)k;
        std::vector<int64_t> colStart;
        std::vector<int32_t> rowMaster;
        std::vector<double> val;
        transposeCore(c, colStart, rowMaster, val);
        const int N = c->N;
        std::vector<int32_t> cs((size_t)N + 1, 0), ci, cr;
        std::vector<double> cv;
        ci.reserve((size_t)colStart[N] / 2 + 16);
        cv.reserve((size_t)colStart[N] / 2 + 16);
        for (int j = 0; j < N; ++j) {
            for (int64_t e = colStart[j]; e < colStart[j + 1]; ++e) {
                const int32_t sl = slot[(size_t)rowMaster[(size_t)e]];
                if (sl >= 0) {  // (order kept: descending rows, the reference's scatter order)
                    ci.push_back(sl);
                    cr.push_back(rowMaster[(size_t)e]);
                    cv.push_back(val[(size_t)e]);
                }
            }
            cs[(size_t)j + 1] = (int32_t)ci.size();
        }
        std::vector<int32_t> ai((size_t)std::max(m, 1));
        for (int b = 0; b < m; ++b) {
            const int32_t sl = slot[(size_t)(c->nBaseRows + b)];
            ai[(size_t)b] = sl >= 0 ? sl : (int32_t)ns;  // slot ns holds 0.0
        }
        if ((rc = devAlloc(c, &c->dCStart, cs.size() + 8))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dCStart, 0, sizeof(int32_t) * (cs.size() + 8)));
        if ((rc = devAlloc(c, &c->dCIdx, ci.size() + 8))) return rc;
        if ((rc = devAlloc(c, &c->dCRow, cr.size() + 8))) return rc;
        if ((rc = devAlloc(c, &c->dCVal, cv.size() + 8))) return rc;
        if ((rc = devAlloc(c, &c->dAlphaIdxCompact, ai.size()))) return rc;
        {
            std::vector<int32_t> be((size_t)m + 1);
            for (int b = 0; b <= m; ++b) be[(size_t)b] = cs[(size_t)c->hBlockColStart[(size_t)b]];
            if ((rc = devAlloc(c, &c->dBlkEntCompact, be.size()))) return rc;
            DIPGPU_CUDA(c, cudaMemcpy(c->dBlkEntCompact, be.data(), sizeof(int32_t) * be.size(), cudaMemcpyHostToDevice));
            c->hBlkEntCompact = be;
        }
        if ((rc = devAlloc(c, &c->dUCompact, ns + 16))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dCStart, cs.data(), sizeof(int32_t) * cs.size(), cudaMemcpyHostToDevice));
        if (!ci.empty()) {
            DIPGPU_CUDA(c, cudaMemcpy(c->dCIdx, ci.data(), sizeof(int32_t) * ci.size(), cudaMemcpyHostToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(c->dCRow, cr.data(), sizeof(int32_t) * cr.size(), cudaMemcpyHostToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(c->dCVal, cv.data(), sizeof(double) * cv.size(), cudaMemcpyHostToDevice));
        }
        DIPGPU_CUDA(c, cudaMemcpy(c->dAlphaIdxCompact, ai.data(), sizeof(int32_t) * ai.size(), cudaMemcpyHostToDevice));
        DIPGPU_CUDA(c, cudaMemset(c->dUCompact, 0, sizeof(double) * (ns + 16)));
        devFree(&c->dUBatchCompact);  // sized by the support
        c->cStride = 0;
    }
    c->dualsCompact = false;
    c->compactValid = true;
    return DIPGPU_OK;
}

// a dense dual vector: all stored entries, or -- with a dual support, outside of which the caller guarantees zeros --
// only the entries of the rows that can carry a dual (the same sums: the others would add u*a = 0)
static RcFuse rcFuseDense(const Ctx *c, const double *uVec, int64_t uStride, int phase)
{
    if (!c->hSupport.empty() && c->dCRow)
        return RcFuse{c->dCStart, c->hBlkEntCompact.data(), c->dCRow, c->dCVal, uVec, c->dAlphaIdxDense, uStride, phase == DIPGPU_PHASE_PRICE1 ? 1 : 0};
    return RcFuse{c->dColStart32, c->hBlkEntFull.data(), c->dRowMaster, c->dVal, uVec, c->dAlphaIdxDense, uStride, phase == DIPGPU_PHASE_PRICE1 ? 1 : 0};
}
// The whole round as one kernel?  Eligible shape, structures built, and every block of the shard fits the kernel's
// staging areas in the CSC the call would read (support-restricted when a dual support is declared).
static bool useRcFuse(Ctx *c, int *rcOut)
{
    *rcOut = DIPGPU_OK;
    if (!rcFuseEligible(c) || c->nLocal > kFuseFilterMaxBlocks) return false;
    if (c->baseR != c->R) return false;  // appended rows not folded in yet: the CSC the kernel reads lacks them
    if ((*rcOut = ensureRcFuse(c))) return false;
    const std::vector<int32_t> &be = c->hSupport.empty() ? c->hBlkEntFull : c->hBlkEntCompact;
    if ((int)be.size() != c->m + 1) return false;
    const int cap = rcFuseStageCap(c);
    for (int b = c->firstBlock; b < c->firstBlock + c->nLocal; ++b)
        if (be[(size_t)b + 1] - be[(size_t)b] > cap) return false;
    return true;
}

static RcFuse rcFuseCompact(const Ctx *c, const double *uVec, int64_t uStride, int phase)
{
    return RcFuse{c->dCStart, c->hBlkEntCompact.data(), c->dCIdx, c->dCVal, uVec, c->dAlphaIdxCompact, uStride, phase == DIPGPU_PHASE_PRICE1 ? 1 : 0};
}

// Would the fused DP kernel for nVec vectors be launched cooperatively (knapsack.cu: launchDp)?
static bool dpWouldCoop(const Ctx *c, int nVec)
{
    return c->geom.fuse && !c->optPdl && !c->optNoCoop && c->dpCoResident > 0 && (long long)c->nLocal * nVec <= c->dpCoResident;
}

// Co-resident launches fetch u[support] from host memory themselves: where from.  A page-locked `u` is read in
// place (the kernel picks u[support[k]]), a pageable one is gathered into mapped pinned staging first, staging
// somebody else filled (a multi-device context) is read as it is.
static int pullSource(Ctx *c, const DualFeed &f, int32_t uLen, int nVec, RcFuse *rf)
{
    const size_t ns = c->hSupport.size();
    rf->pullN = (int)ns;
    if (f.staged) {
        rf->pullSrc = f.staged;
        rf->pullStride = f.stagedStride;
        return DIPGPU_OK;
    }
    if (c->optGatherPinned) {
        const double *dv = deviceVisible(f.host);
        if (dv) {
            rf->pullSrc = dv;
            rf->pullIdx = c->dSupport;
            rf->pullStride = uLen;
            return DIPGPU_OK;
        }
    }
    int rc;
    if ((rc = ensureSupportStaging(c, nVec))) return rc;
    gatherSupport(c->hSupport, f.host, uLen, nVec, c->hSupVals);
    rf->pullSrc = c->hSupVals;  // (mapped + portable: one address on the host and on every device)
    rf->pullStride = (int64_t)ns;
    return DIPGPU_OK;
}

// u[support] of nVec dual vectors into the compact device vector(s) (vector v at dDst + v*dstStride): gathered on
// the host into pinned staging -- or taken from staging somebody else filled -- and copied in one piece.
static int feedCompact(Ctx *c, const DualFeed &f, int32_t uLen, int nVec, double *dDst, size_t dstStride, cudaStream_t st)
{
    const size_t ns = c->hSupport.size();
    const double *src = f.staged;
    size_t srcStride = (size_t)f.stagedStride;
    if (!src && c->optGatherPinned) {
        // page-locked dual vectors: the device picks u[support] out of them itself -- no pass over the (cold) vectors
        // on the host (12 us per 2 MB vector against 5 us of kernel)
        const double *dv = deviceVisible(f.host);
        if (dv) {
            c->h2d += (int64_t)sizeof(double) * (int64_t)ns * nVec;
            return launchGatherDuals(c, (int)ns, c->dSupport, dv, (int64_t)uLen, dDst, (int64_t)dstStride, nVec, st);
        }
    }
    if (!src) {
        int rc;
        if ((rc = ensureSupportStaging(c, nVec))) return rc;
        gatherSupport(c->hSupport, f.host, uLen, nVec, c->hSupVals);
        src = c->hSupVals;
        srcStride = ns;
    }
    if (nVec == 1) DIPGPU_CUDA(c, cudaMemcpyAsync(dDst, src, sizeof(double) * ns, cudaMemcpyHostToDevice, st));
    else DIPGPU_CUDA(c, cudaMemcpy2DAsync(dDst, sizeof(double) * dstStride, src, sizeof(double) * srcStride, sizeof(double) * ns,
                                          (size_t)nVec, cudaMemcpyHostToDevice, st));
    c->h2d += (int64_t)sizeof(double) * (int64_t)ns * nVec;
    return DIPGPU_OK;
}

// The dense dual vector dU, for the passes that read it (pool re-pricing, smoothing): after a round that only filled
// the compact vector it is formed on the device.
static int ensureDenseDuals(Ctx *c, cudaStream_t st)
{
    if (c->dualsResident || !c->dualsCompact) return DIPGPU_OK;
    const int uLen = c->R + c->numConvex;
    int rc;
    if ((rc = ensureDualBuffer(c, uLen))) return rc;
    if (!c->dUClean) {
        DIPGPU_CUDA(c, cudaMemsetAsync(c->dU, 0, sizeof(double) * (size_t)uLen, st));
        c->dUClean = true;
    }
    const int ns = (int)c->hSupport.size();
    if ((rc = launchScatterDuals(c, ns, c->dSupport, c->dUCompact, ns, false, c->dU, 0, 1, st))) return rc;
    c->dualsResident = true;
    return DIPGPU_OK;
}

// Enqueue stages (b) and (c) on `st`; dRedCost/dAlpha are device pointers
// (alpha indexed by global block id).
static int enqueueSolve(Ctx *c, const double *dRedCost, const double *dAlpha, double eps, cudaStream_t st,
                        bool timed, const RcFuse *rf = nullptr)
{
    int rc;
    c->resultMirrored = false;
    ++c->roundSerial;
    c->lastRoundOnHost = false;  // (the host entry points set it once the result has arrived)
    if (c->blockKind == 1) {
        // multiple-choice knapsack blocks: DP + backtrack (mckp.cu), then the stand-alone filter
        if ((rc = launchMckp(c, dRedCost, dAlpha, st))) return rc;
        if (timed) { cudaEventRecord(c->ev[3], st); cudaEventRecord(c->ev[4], st); }
        if ((rc = launchFilter(c, eps, st))) return rc;
        if (timed) cudaEventRecord(c->ev[5], st);
        return DIPGPU_OK;
    }
    if (c->blockKind == 2) {
        // integer knapsack blocks: table + counts + column (intknap.cu), then the stand-alone filter
        if ((rc = launchIntKnap(c, dRedCost, dAlpha, st))) return rc;
        if (timed) { cudaEventRecord(c->ev[3], st); cudaEventRecord(c->ev[4], st); }
        if ((rc = launchFilter(c, eps, st))) return rc;
        if (timed) cudaEventRecord(c->ev[5], st);
        return DIPGPU_OK;
    }
    int mode = c->geom.fuse ? 2 : c->optFuseFilter < 0 ? (c->nLocal <= kFuseFilterMaxBlocks ? 1 : 0)
                                                       : (c->optFuseFilter == 1 ? 1 : 0);
    if (c->nLocal == 0) mode = 0;
    if (!rf && (rc = launchKnapsackPrep(c, dRedCost, st))) return rc;
    if (c->geom.big) {
        if ((rc = launchBigDp(c, dRedCost, st))) return rc;
    } else if ((rc = launchKnapsackDp(c, dRedCost, dAlpha, eps, st, rf))) {
        return rc;
    }
    if (timed) cudaEventRecord(c->ev[3], st);
    if (mode != 2 && (rc = launchBacktrackEmit(c, dRedCost, dAlpha, eps, mode == 1, st))) return rc;
    if (timed) cudaEventRecord(c->ev[4], st);
    if (mode == 0 && (rc = launchFilter(c, eps, st))) return rc;
    if (timed) cudaEventRecord(c->ev[5], st);
    return DIPGPU_OK;
}

int unpack(Ctx *c, const void *packed, size_t bytes, dipgpu_cols *out)
{
    if (!packed || !out || bytes < sizeof(ResultHeader)) return fail(c, DIPGPU_ERR_INVALID, "bad packed result");
    const char *p = static_cast<const char *>(packed);
    const ResultHeader *h = reinterpret_cast<const ResultHeader *>(p);
    if (h->magic != kResultMagic) return fail(c, DIPGPU_ERR_INVALID, "packed result: bad magic");
    const int m = h->m;
    const ResultLayout L = ResultLayout::make(m, h->indCap);
    if (bytes < L.offInd + sizeof(int32_t) * (size_t)h->nInd)
        return fail(c, DIPGPU_ERR_INVALID, "packed result truncated");
    const double *rc = reinterpret_cast<const double *>(p + L.offRedCost);
    const double *mn = reinterpret_cast<const double *>(p + L.offMostNeg);
    const double *oc = reinterpret_cast<const double *>(p + L.offOrigCost);
    const double *ob = reinterpret_cast<const double *>(p + L.offObj);
    const int64_t *ks = reinterpret_cast<const int64_t *>(p + L.offKeptStart);
    const int32_t *nz = reinterpret_cast<const int32_t *>(p + L.offNnz);
    const int32_t *kb = reinterpret_cast<const int32_t *>(p + L.offKeptBlock);
    const int32_t *ki = reinterpret_cast<const int32_t *>(p + L.offInd);
    // integer knapsack blocks: the index region holds (index, multiplicity) pairs and every count is in words
    const int wpe = h->reserved[1] == 1 ? 2 : 1;
    out->nCols = h->nKept;
    out->nInd = h->nInd / wpe;
    out->cells = h->cells;
    out->cellsEvaluated = h->reserved[0];
    const bool wantBlocks = out->blockRedCost || out->blockMostNegRC || out->blockOrigCost || out->blockObj || out->blockNnz;
    if (wantBlocks && out->capBlocks < h->firstBlock + m)
        return fail(c, DIPGPU_ERR_OVERFLOW, "capBlocks %d < %d", out->capBlocks, h->firstBlock + m);
    double sum = 0.0;
    for (int lb = 0; lb < m; ++lb) {
        const int b = h->firstBlock + lb;
        if (out->blockRedCost) out->blockRedCost[b] = rc[lb];
        if (out->blockMostNegRC) out->blockMostNegRC[b] = mn[lb];
        if (out->blockOrigCost) out->blockOrigCost[b] = oc[lb];
        if (out->blockObj) out->blockObj[b] = ob[lb];
        if (out->blockNnz) out->blockNnz[b] = nz[lb] / wpe;
        sum += mn[lb];  // block order, DecompAlgo.cpp:5229-5232
    }
    out->mostNegReducedCost = sum;
    if (h->nKept > out->capCols || h->nInd / wpe > out->capInd)
        return fail(c, DIPGPU_ERR_OVERFLOW, "kept columns need capCols >= %d, capInd >= %lld", h->nKept,
                    (long long)(h->nInd / wpe));
    for (int k = 0; k < h->nKept; ++k) {
        const int lb = kb[k];
        if (out->colBlock) out->colBlock[k] = h->firstBlock + lb;
        if (out->colRedCost) out->colRedCost[k] = rc[lb];
        if (out->colOrigCost) out->colOrigCost[k] = oc[lb];
        if (out->colStart) out->colStart[k] = ks[k] / wpe;
    }
    if (out->colStart) out->colStart[h->nKept] = h->nInd / wpe;
    unpackEntries(ki, h->nInd, wpe, out->colInd, out->colEl);
    return DIPGPU_OK;
}

// D2H of the packed result: one speculative copy (header + per-block arrays +
// the first 64 KiB of indices), a second one only if more indices were kept.
int fetchResult(Ctx *c, cudaStream_t st)
{
    if (c->resultMirrored) {
        // the fused kernel stored the result into hResult itself (mapped pinned memory): nothing to copy
        if (c->stageTiming) cudaEventRecord(c->ev[6], st);
        bool seen = false;
        if (c->flagArmed) {
            // ... and it announces the complete result through the done word: no stream synchronisation on the
            // caller's critical path (a kernel that never gets there -- a fault -- ends the wait after 20 ms and the
            // synchronisation below reports it)
            const volatile unsigned int *flag = c->hDoneFlag;
            const unsigned int want = c->doneSerial;
            const auto t0 = std::chrono::steady_clock::now();
            for (unsigned int spins = 1;; ++spins) {
                if (*flag == want) { seen = true; break; }
                if ((spins & 0xfffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) break;
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            c->flagArmed = false;
            ++(seen ? c->flagWaits : c->flagTimeouts);
        }
        if (!seen) DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        const ResultHeader *h = static_cast<const ResultHeader *>(c->hResult);
        c->lastResultBytes = c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd;
        c->d2h += (int64_t)c->lastResultBytes;
        return DIPGPU_OK;
    }
    // speculative size: what the previous round needed plus a quarter (rounds of one node keep similar
    // columns), at least 8 KiB of indices
    const size_t guess = std::max<size_t>(c->lastResultBytes + c->lastResultBytes / 4, c->layout.offInd + (size_t)8192);
    const size_t spec = std::min(c->resultBytes, ResultLayout::alignUp(guess, 256));
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->hResult, c->dResult, spec, cudaMemcpyDeviceToHost, st));
    if (c->stageTiming) cudaEventRecord(c->ev[6], st);
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    c->d2h += (int64_t)spec;
    const ResultHeader *h = static_cast<const ResultHeader *>(c->hResult);
    const size_t need = c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd;
    c->lastResultBytes = need;
    if (need > spec) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(static_cast<char *>(c->hResult) + spec, static_cast<char *>(c->dResult) + spec,
                                       need - spec, cudaMemcpyDeviceToHost, st));
        DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        c->d2h += (int64_t)(need - spec);
    }
    return DIPGPU_OK;
}

void collectStageMs(Ctx *c)
{
    if (!c->stageTiming) return;
    for (int i = 0; i < 6; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]) != cudaSuccess) ms = 0.f;
        c->stageMs[i] = ms;
    }
    (void)cudaGetLastError();
}


// ------------------------------------------------------------ batched dual vectors

int launchPoolCopy(Ctx *c, int n, const void *dDesc, const void *srcEntries, const int32_t *srcRow, const double *srcVal,
                   int32_t *dstRow, double *dstVal, cudaStream_t st);  // pool.cu

static void freeBatch(Ctx *c)
{
    devFree(&c->dUBatch); devFree(&c->dRedCostBatch); devFree(&c->dPubWordBatch);
    devFree(&c->dScaleB); devFree(&c->dShortcutB); devFree(&c->dSelB); devFree(&c->dZShortB); devFree(&c->dAlphaLadder); devFree(&c->dPullCtrBatch);
    if (c->dResultBatch) cudaFree(c->dResultBatch);
    if (c->hResultBatch) cudaFreeHost(c->hResultBatch);
    c->dResultBatch = c->hResultBatch = c->dResultBatchHost = nullptr;
    c->batchCap = 0;
    c->haveST = false;
}

// Buffers for nVec dual vectors priced in one submission.
int ensureBatch(Ctx *c, int nVec)
{
    const size_t uLen = (size_t)c->R + c->numConvex;
    const size_t uStride = (uLen + 17) & ~(size_t)1;
    const size_t rcStride = ((size_t)c->N + 17) & ~(size_t)1;
    const size_t resStride = ResultLayout::alignUp(c->resultBytes, 16);
    if (c->batchCap >= nVec && c->uStride == uStride && c->rcStride == rcStride && c->resStride == resStride) return DIPGPU_OK;
    freeBatch(c);
    devFree(&c->dUBatchCompact);
    c->cStride = 0;
    int rc;
    const size_t V = (size_t)nVec, nl = (size_t)std::max(c->nLocal, 1);
    if ((rc = devAlloc(c, &c->dUBatch, V * uStride))) return rc;
    if ((rc = devAlloc(c, &c->dRedCostBatch, V * rcStride))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dUBatch, 0, sizeof(double) * V * uStride));
    DIPGPU_CUDA(c, cudaMemset(c->dRedCostBatch, 0, sizeof(double) * V * rcStride));
    if ((rc = devAlloc(c, &c->dPubWordBatch, V * nl * kPubWordStride))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dPubWordBatch, 0, sizeof(unsigned long long) * V * nl * kPubWordStride));
    if ((rc = devAlloc(c, &c->dScaleB, V * nl))) return rc;
    if ((rc = devAlloc(c, &c->dShortcutB, V * nl))) return rc;
    if ((rc = devAlloc(c, &c->dSelB, 2 * V * nl))) return rc;
    if ((rc = devAlloc(c, &c->dZShortB, V * nl))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dShortcutB, 0, sizeof(int32_t) * V * nl));
    if ((rc = devAlloc(c, &c->dAlphaLadder, V))) return rc;
    if ((rc = devAlloc(c, &c->dPullCtrBatch, V + 64))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dPullCtrBatch, 0, sizeof(unsigned int) * (V + 64)));
    DIPGPU_CUDA(c, cudaMalloc(&c->dResultBatch, V * resStride));
    DIPGPU_CUDA(c, cudaMemset(c->dResultBatch, 0, V * resStride));
    DIPGPU_CUDA(c, cudaHostAlloc(&c->hResultBatch, V * resStride, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(c->hResultBatch, 0, V * resStride);
    DIPGPU_CUDA(c, cudaHostGetDevicePointer(&c->dResultBatchHost, c->hResultBatch, 0));
    c->batchCap = nVec;
    c->dUBatchClean = true;  // freshly zeroed
    c->uStride = uStride;
    c->rcStride = rcStride;
    c->resStride = resStride;
    return DIPGPU_OK;
}

// Stages (a)-(c) for the nVec dual vectors held in dUBatch: three launches in all when the shard
// runs the fused DP kernel (gridDim.y = vectors), else vector after vector through the large-shard path.
int enqueueBatch(Ctx *c, int nVec, int phase, double eps, cudaStream_t st, bool compact)
{
    int rc;
    const bool timed = c->stageTiming;
    c->resultMirrored = false;
    ++c->roundSerial;
    c->lastRoundOnHost = false;
    if (timed) cudaEventRecord(c->ev[1], st);
    if (c->geom.fuse || c->nLocal == 0) {
        const bool fused = useRcFuse(c, &rc);
        if (rc) return rc;
        if (fused) {
            // ONE kernel for the whole batch: CTA (block, vector) forms its columns' reduced costs for its vector
            RcFuse rf = compact ? rcFuseCompact(c, c->dUBatchCompact, (int64_t)c->cStride, phase)
                                : rcFuseDense(c, c->dUBatch, (int64_t)c->uStride, phase);
            if (compact && c->batchPull.pullSrc) {
                rf.pullSrc = c->batchPull.pullSrc;
                rf.pullIdx = c->batchPull.pullIdx;
                rf.pullStride = c->batchPull.pullStride;
                rf.pullN = c->batchPull.pullN;
            }
            c->batchPull = RcFuse();
            if (timed) cudaEventRecord(c->ev[2], st);
            c->wantMirror = true;
            rc = launchKnapsackDpBatch(c, c->dRedCostBatch, (int64_t)c->rcStride, nullptr, 0, eps, nVec, st, &rf);
            c->wantMirror = false;
            if (timed) { cudaEventRecord(c->ev[3], st); cudaEventRecord(c->ev[4], st); cudaEventRecord(c->ev[5], st); }
            return rc;
        }
        if ((rc = launchRedCostBatch(c, c->dUBatch, (int64_t)c->uStride, c->dRedCostBatch, (int64_t)c->rcStride, nVec, phase, st))) return rc;
        if (timed) cudaEventRecord(c->ev[2], st);
        if ((rc = launchKnapsackPrepBatch(c, c->dRedCostBatch, (int64_t)c->rcStride, nVec, st))) return rc;
        c->wantMirror = true;  // every caller reads the results on the host right after
        rc = launchKnapsackDpBatch(c, c->dRedCostBatch, (int64_t)c->rcStride, c->dUBatch + c->nBaseRows, (int64_t)c->uStride, eps, nVec, st);
        c->wantMirror = false;
        if (timed) { cudaEventRecord(c->ev[3], st); cudaEventRecord(c->ev[4], st); cudaEventRecord(c->ev[5], st); }
        return rc;
    }
    if (timed) { cudaEventRecord(c->ev[2], st); }
    for (int v = 0; v < nVec; ++v) {
        const double *u = c->dUBatch + (size_t)v * c->uStride;
        if ((rc = launchRedCost(c, u, phase, st))) return rc;
        if ((rc = enqueueSolve(c, c->dRedCost, u + c->nBaseRows, eps, st, false))) return rc;
        DIPGPU_CUDA(c, cudaMemcpyAsync(static_cast<char *>(c->dResultBatch) + (size_t)v * c->resStride, c->dResult, c->resultBytes, cudaMemcpyDeviceToDevice, st));
    }
    if (timed) { cudaEventRecord(c->ev[3], st); cudaEventRecord(c->ev[4], st); cudaEventRecord(c->ev[5], st); }
    return DIPGPU_OK;
}

// The dual vectors of a batch: into dUBatch (dense), or -- fused shapes with a dual support, host vectors --
// u[support] only into dUBatchCompact.  *compact tells enqueueBatch which.
int feedBatchDuals(Ctx *c, const DualFeed &f, int32_t uLen, int nVec, cudaStream_t st, bool *compact)
{
    int rc;
    *compact = false;
    const size_t ns = c->hSupport.size();
    const bool fused = useRcFuse(c, &rc);
    if (rc) return rc;
    if (fused && ns > 0 && (f.host || f.staged)) {
        const size_t stride = (ns + 17) & ~(size_t)1;
        if (!c->dUBatchCompact || c->cStride != stride) {
            if ((rc = devAlloc(c, &c->dUBatchCompact, (size_t)std::max(c->batchCap, nVec) * stride))) return rc;
            DIPGPU_CUDA(c, cudaMemset(c->dUBatchCompact, 0, sizeof(double) * (size_t)std::max(c->batchCap, nVec) * stride));
            c->cStride = stride;
        }
        *compact = true;
        c->batchPull = RcFuse();
        // the kernel fetches u[support] itself: co-resident grids, and ticketed ones when a whole vector's CTAs are
        // resident at once (knapsack.cu: launchDp) -- the gathers of the later vectors then run under the DP of the
        // earlier ones instead of in a kernel of their own in front of it
        const bool ticketPull = c->optBatchPull && c->geom.fuse && c->dpResidentBatch >= c->nLocal && c->dPullCtrBatch != nullptr &&
                                !dpWouldCoop(c, nVec) && !c->optPdl && nVec >= c->optBatchPullMin;
        if (c->optPull && (dpWouldCoop(c, nVec) || ticketPull)) {
            c->h2d += (int64_t)sizeof(double) * (int64_t)ns * nVec;
            return pullSource(c, f, uLen, nVec, &c->batchPull);
        }
        return feedCompact(c, f, uLen, nVec, c->dUBatchCompact, stride, st);
    }
    return feedDuals(c, f, uLen, nVec, c->dUBatch, c->uStride, &c->dUBatchClean, st);
}

// D2H of nVec packed results into hResultBatch (one strided copy of the speculative prefix, a second copy per
// result that kept more indices).
int fetchBatchPacked(Ctx *c, int nVec, cudaStream_t st)
{
    if (c->resultMirrored) {
        // the fused kernel stored every vector's result into hResultBatch itself (mapped pinned memory)
        if (c->stageTiming) cudaEventRecord(c->ev[6], st);
        DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        collectStageMs(c);
        for (int v = 0; v < nVec; ++v) {
            const ResultHeader *h = reinterpret_cast<const ResultHeader *>(static_cast<char *>(c->hResultBatch) + (size_t)v * c->resStride);
            if (h->magic != kResultMagic) return fail(c, DIPGPU_ERR_CUDA, "batched round: result %d has no header", v);
            c->d2h += (int64_t)(c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd);
        }
        return DIPGPU_OK;
    }
    const size_t spec = std::min(c->resultBytes, c->layout.offInd + (size_t)65536);
    DIPGPU_CUDA(c, cudaMemcpy2DAsync(c->hResultBatch, c->resStride, c->dResultBatch, c->resStride, spec, (size_t)nVec, cudaMemcpyDeviceToHost, st));
    if (c->stageTiming) cudaEventRecord(c->ev[6], st);
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    collectStageMs(c);
    c->d2h += (int64_t)spec * nVec;
    bool more = false;
    for (int v = 0; v < nVec; ++v) {
        char *hp = static_cast<char *>(c->hResultBatch) + (size_t)v * c->resStride;
        const ResultHeader *h = reinterpret_cast<const ResultHeader *>(hp);
        if (h->magic != kResultMagic) return fail(c, DIPGPU_ERR_CUDA, "batched round: result %d has no header", v);
        const size_t need = c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd;
        if (need > spec) {
            DIPGPU_CUDA(c, cudaMemcpyAsync(hp + spec, static_cast<char *>(c->dResultBatch) + (size_t)v * c->resStride + spec, need - spec, cudaMemcpyDeviceToHost, st));
            c->d2h += (int64_t)(need - spec);
            more = true;
        }
    }
    if (more) DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    return DIPGPU_OK;
}

// ... and decoding.
int fetchBatch(Ctx *c, int nVec, dipgpu_cols *outs, cudaStream_t st)
{
    int first = fetchBatchPacked(c, nVec, st);
    if (first) return first;
    for (int v = 0; v < nVec; ++v) {
        const int rc = unpack(c, static_cast<char *>(c->hResultBatch) + (size_t)v * c->resStride, c->resultBytes, &outs[v]);
        if (rc && !first) first = rc;
    }
    return first;
}

int checkRoundReady(Ctx *c, const char *who, int32_t uLen, int32_t phase)
{
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "%s needs set_core_csr and set_knapsack_blocks", who);
    if (phase != DIPGPU_PHASE_PRICE1 && !c->haveObj) return fail(c, DIPGPU_ERR_STATE, "%s (phase 2) before set_objective", who);
    if (uLen != c->R + c->numConvex) return fail(c, DIPGPU_ERR_INVALID, "%s: uLen=%d, expected R+numConvex=%d", who, uLen, c->R + c->numConvex);
    if (c->numConvex != c->m) return fail(c, DIPGPU_ERR_INVALID, "%s: numConvex=%d but %d blocks", who, c->numConvex, c->m);
    return ensureObjective(c, c->N);
}

int ensureDualBuffer(Ctx *c, int32_t uLen)
{
    if ((size_t)uLen > c->uCap) {
        int rc;
        if ((rc = devAlloc(c, &c->dU, (size_t)uLen + 16))) return rc;
        c->uCap = (size_t)uLen;
        c->dualsResident = false;
        c->dUClean = false;
    }
    return DIPGPU_OK;
}

static int ensureSupportStaging(Ctx *c, int nVec)
{
    if (nVec <= c->supVecCap) return DIPGPU_OK;
    if (c->hSupVals) cudaFreeHost(c->hSupVals);
    c->hSupVals = nullptr;
    c->supVecCap = 0;
    const size_t n = std::max<size_t>(c->hSupport.size(), 1) * (size_t)nVec;
    DIPGPU_CUDA(c, cudaHostAlloc(reinterpret_cast<void **>(&c->hSupVals), sizeof(double) * n, cudaHostAllocMapped | cudaHostAllocPortable));
    c->supVecCap = nVec;
    return DIPGPU_OK;
}

// The nVec master dual vectors of a call (vector v at dDst + v*dstStride afterwards), from wherever the
// DualFeed says they are: the caller's host vectors (row v at host + v*uLen) -- the whole vectors, or, with a
// dual support, only the entries that can be nonzero, gathered into pinned staging on the host and scattered
// on the device (the rest of dDst is kept at zero) --, or a device-visible copy somebody else prepared
// (a multi-device context gathers once for all its devices; a peer device holds the vector).
int feedDuals(Ctx *c, const DualFeed &f, int32_t uLen, int nVec, double *dDst, size_t dstStride, bool *clean, cudaStream_t st)
{
    const size_t ns = c->hSupport.size();
    if (ns == 0) {
        const double *src = f.host ? f.host : f.full;
        const cudaMemcpyKind kind = f.host ? cudaMemcpyHostToDevice : cudaMemcpyDefault;
        if (!src) return fail(c, DIPGPU_ERR_INVALID, "no dual vector");
        const size_t srcStride = f.host ? (size_t)uLen : (size_t)f.fullStride;
        if (nVec == 1) DIPGPU_CUDA(c, cudaMemcpyAsync(dDst, src, sizeof(double) * (size_t)uLen, kind, st));
        else DIPGPU_CUDA(c, cudaMemcpy2DAsync(dDst, sizeof(double) * dstStride, src, sizeof(double) * srcStride,
                                              sizeof(double) * (size_t)uLen, (size_t)nVec, kind, st));
        c->h2d += (int64_t)sizeof(double) * uLen * nVec;
        *clean = false;
        return DIPGPU_OK;
    }
    int rc;
    if (!*clean) {
        DIPGPU_CUDA(c, cudaMemsetAsync(dDst, 0, sizeof(double) * (nVec == 1 ? (size_t)uLen : dstStride * (size_t)nVec), st));
        *clean = true;
    }
    c->h2d += (int64_t)sizeof(double) * (int64_t)ns * nVec;
    if (f.staged)  // u[support] of every vector, already gathered into device-visible memory
        return launchScatterDuals(c, (int)ns, c->dSupport, f.staged, f.stagedStride, false, dDst, (int64_t)dstStride, nVec, st);
    if (f.full)    // the whole vectors in device-visible memory: the kernel picks u[support] itself
        return launchScatterDuals(c, (int)ns, c->dSupport, f.full, f.fullStride, true, dDst, (int64_t)dstStride, nVec, st);
    const double *u = f.host;
    if (!u) return fail(c, DIPGPU_ERR_INVALID, "no dual vector");
    if (c->optGatherPinned) {
        // u in pinned (device-mapped) or managed host memory: the kernel picks u[support] out of the caller's
        // buffer itself -- no pass over the vector on the host (a gather of cold cache lines: ~9 us for 4 240
        // entries of a 2 MB vector)
        const double *dv = deviceVisible(u);
        if (dv) return launchScatterDuals(c, (int)ns, c->dSupport, dv, (int64_t)uLen, true, dDst, (int64_t)dstStride, nVec, st);
    }
    if ((rc = ensureSupportStaging(c, nVec))) return rc;
    gatherSupport(c->hSupport, u, uLen, nVec, c->hSupVals);
    // the scatter kernel reads the pinned staging directly (mapped host memory): the few KB cross PCIe inside
    // the kernel, without a separate copy and its latency
    return launchScatterDuals(c, (int)ns, c->dSupport, c->hSupVals, (int64_t)ns, false, dDst, (int64_t)dstStride, nVec, st);
}

// device address of a host pointer in page-locked (mapped) or managed memory, else NULL
const double *deviceVisible(const double *u)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, u) == cudaSuccess && (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged) &&
        at.devicePointer)
        return static_cast<const double *>(at.devicePointer);
    (void)cudaGetLastError();
    return nullptr;
}

// dst[v*ns + k] = u[v*uLen + idx[k]]
void gatherSupport(const std::vector<int32_t> &support, const double *u, int32_t uLen, int nVec, double *dst)
{
    const size_t ns = support.size();
    const int32_t *idx = support.data();
    for (int v = 0; v < nVec; ++v) {
        const double *uv = u + (size_t)v * (size_t)uLen;
        double *d = dst + (size_t)v * ns;
        for (size_t k = 0; k < ns; ++k) {
            if (k + 32 < ns) __builtin_prefetch(uv + idx[k + 32]);
            d[k] = uv[idx[k]];
        }
    }
}

static int uploadDuals(Ctx *c, const double *u, int32_t uLen, int nVec, double *dDst, size_t dstStride, bool *clean,
                       cudaStream_t st)
{
    DualFeed f;
    f.host = u;
    return feedDuals(c, f, uLen, nVec, dDst, dstStride, clean, st);
}

// ------------------------------------------------------------------ waiting-column pool

static int poolReserve(Ctx *c, int64_t cols, int64_t nnz)
{
    int rc;
    if (!c->dPoolFlag && (rc = devAlloc(c, &c->dPoolFlag, 4))) return rc;
    if (cols > c->poolColCap) {
        const int64_t cap = std::max<int64_t>(cols, std::max<int64_t>(1024, 2 * c->poolColCap));
        int64_t *ns = nullptr; double *no = nullptr, *nr = nullptr;
        if ((rc = devAlloc(c, &ns, (size_t)cap + 1))) return rc;
        if ((rc = devAlloc(c, &no, (size_t)cap))) return rc;
        if ((rc = devAlloc(c, &nr, (size_t)cap))) return rc;
        if (c->dPoolStart) DIPGPU_CUDA(c, cudaMemcpy(ns, c->dPoolStart, sizeof(int64_t) * ((size_t)c->poolCols + 1), cudaMemcpyDeviceToDevice));
        else DIPGPU_CUDA(c, cudaMemset(ns, 0, sizeof(int64_t)));
        if (c->poolCols > 0) DIPGPU_CUDA(c, cudaMemcpy(no, c->dPoolOrigCost, sizeof(double) * (size_t)c->poolCols, cudaMemcpyDeviceToDevice));
        devFree(&c->dPoolStart); devFree(&c->dPoolOrigCost); devFree(&c->dPoolRedCost);
        c->dPoolStart = ns; c->dPoolOrigCost = no; c->dPoolRedCost = nr;
        c->poolColCap = cap;
    }
    if (nnz > c->poolNnzCap) {
        const int64_t cap = std::max<int64_t>(nnz, std::max<int64_t>(65536, 2 * c->poolNnzCap));
        int32_t *nr = nullptr; double *nv = nullptr;
        // (+ kPoolPad: the re-pricing kernel streams whole aligned 512-entry windows of the two arrays)
        if ((rc = devAlloc(c, &nr, (size_t)cap + kPoolPad))) return rc;
        if ((rc = devAlloc(c, &nv, (size_t)cap + kPoolPad))) return rc;
        if (c->poolNnz > 0) {
            DIPGPU_CUDA(c, cudaMemcpy(nr, c->dPoolRow, sizeof(int32_t) * (size_t)c->poolNnz, cudaMemcpyDeviceToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(nv, c->dPoolVal, sizeof(double) * (size_t)c->poolNnz, cudaMemcpyDeviceToDevice));
        }
        devFree(&c->dPoolRow); devFree(&c->dPoolVal);
        c->dPoolRow = nr; c->dPoolVal = nv;
        c->poolNnzCap = cap;
    }
    return DIPGPU_OK;
}

int multiEnqueueSolve(Ctx *c, double eps, cudaStream_t st, bool timed)
{
    c->wantMirror = true;
    const int rc = enqueueSolve(c, c->dRedCost, c->dAlpha, eps, st, timed);
    c->wantMirror = false;
    return rc;
}

// One pricing round, enqueued on `st`: the duals (feed == NULL: the vector already in dU), the reduced costs of
// this context's columns, the block subproblems of its block range and the filter.  mirror: the fused kernel
// also stores the packed result into the context's mirror (mapped host memory, or what mirrorBase names).
int enqueueRound(Ctx *c, const DualFeed *feed, int32_t uLen, int32_t phase, double eps, bool mirror, cudaStream_t st, bool needAllRedCosts)
{
    int rc;
    if ((rc = ensureDualBuffer(c, uLen))) return rc;
    const bool timed = c->stageTiming;
    if (timed) cudaEventRecord(c->ev[0], st);
    const bool fused = !needAllRedCosts && useRcFuse(c, &rc);
    if (rc) return rc;
    if (fused) {
        // ONE kernel: the block CTAs form their columns' reduced costs themselves
        const size_t ns = c->hSupport.size();
        RcFuse rf;
        if (feed && ns > 0 && (feed->host || feed->staged)) {
            rf = rcFuseCompact(c, c->dUCompact, 0, phase);
            if (c->optPull && dpWouldCoop(c, 1)) {
                if ((rc = pullSource(c, *feed, uLen, 1, &rf))) return rc;
                c->h2d += (int64_t)sizeof(double) * (int64_t)ns;
            } else if ((rc = feedCompact(c, *feed, uLen, 1, c->dUCompact, 0, st))) {
                return rc;
            }
            c->dualsCompact = true;
            c->dualsResident = false;
        } else if (feed && feed->full) {
            rf = rcFuseDense(c, feed->full, 0, phase);  // a dense vector in device memory (this device or a peer): gathered from where it is
        } else if (feed) {
            if ((rc = feedDuals(c, *feed, uLen, 1, c->dU, 0, &c->dUClean, st))) return rc;
            c->dualsResident = true;
            c->dualsCompact = false;
            rf = rcFuseDense(c, c->dU, 0, phase);
        } else {
            rf = c->dualsCompact && !c->dualsResident ? rcFuseCompact(c, c->dUCompact, 0, phase) : rcFuseDense(c, c->dU, 0, phase);
        }
        if (timed) { cudaEventRecord(c->ev[1], st); cudaEventRecord(c->ev[2], st); }
        c->wantMirror = mirror;
        rc = enqueueSolve(c, c->dRedCost, nullptr, eps, st, timed, &rf);
        c->wantMirror = false;
        return rc;
    }
    const double *uUse = c->dU;
    if (feed) {
        cudaPointerAttributes at;
        if (feed->full && cudaPointerGetAttributes(&at, feed->full) == cudaSuccess &&
            at.type == cudaMemoryTypeDevice && at.device == c->device) {
            uUse = feed->full;  // already a dense vector in this device's memory
        } else {
            (void)cudaGetLastError();
            if ((rc = feedDuals(c, *feed, uLen, 1, c->dU, 0, &c->dUClean, st))) return rc;
            c->dualsResident = true;
            c->dualsCompact = false;
        }
    } else if ((rc = ensureDenseDuals(c, st))) {
        return rc;
    }
    if (timed) cudaEventRecord(c->ev[1], st);
    if ((rc = launchRedCost(c, uUse, phase, st))) return rc;
    if (timed) cudaEventRecord(c->ev[2], st);
    // alpha_b = u[nBaseRows + b]  (DecompAlgo.cpp:4820)
    c->wantMirror = mirror;
    rc = enqueueSolve(c, c->dRedCost, uUse + c->nBaseRows, eps, st, timed);
    c->wantMirror = false;
    return rc;
}

// The round on a dense dual vector that already lies in device memory -- this device's or (multi-device contexts)
// a peer's: fused shapes gather the duals straight from there, the others read a local vector directly and take a
// peer's through a copy of what they need.
int enqueueRoundDev(Ctx *c, const double *uDev, int32_t uLen, int32_t phase, double eps, bool mirror, cudaStream_t st)
{
    DualFeed f;
    f.full = uDev;
    f.fullStride = uLen;
    return enqueueRound(c, &f, uLen, phase, eps, mirror, st, false);
}

// A large batch of page-locked dual vectors on a fused shape with a dual support: what bounds it is not the DP but the
// gather of u[support] out of host memory -- thousands of small PCIe reads per vector, a request-rate limit of the
// link (5-9 us per GAP-E vector).  The batch is therefore priced in chunks: the gathers run one after the other on a
// second stream, chunk k's fused launch (vectors [k0, k0 + n) of the same per-vector buffers, KnapArgs::vecBase) waits
// only for ITS gather, so the DP of a chunk runs under the gathers of the next ones.  Returns 1 when it took the batch.
static int batchPipelined(Ctx *c, int nVec, const DualFeed &f, int32_t uLen, int32_t phase, double eps, cudaStream_t st, int *took)
{
    *took = 0;
    int rc;
    const size_t ns = c->hSupport.size();
    if (!c->optBatchPipeline || nVec < 8 || ns == 0 || !f.host || f.staged || !c->optGatherPinned) return DIPGPU_OK;
    const bool fused = useRcFuse(c, &rc);
    if (rc) return rc;
    if (!fused) return DIPGPU_OK;
    if (c->optPull && nVec <= 64 && dpWouldCoop(c, nVec)) return DIPGPU_OK;  // (the kernel fetches the duals itself)
    const double *dv = deviceVisible(f.host);
    if (!dv) return DIPGPU_OK;
    if (!c->stream2) {
        // highest priority: a gather's few CTAs must get the SM slots that free up while a fused launch still has
        // CTAs waiting, or the gathers would only run in the tails of the DP launches
        int prLo = 0, prHi = 0;
        DIPGPU_CUDA(c, cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
        DIPGPU_CUDA(c, cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prHi));
        for (auto &e : c->evChunk) DIPGPU_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const size_t stride = (ns + 17) & ~(size_t)1;
    if (!c->dUBatchCompact || c->cStride != stride) {
        if ((rc = devAlloc(c, &c->dUBatchCompact, (size_t)std::max(c->batchCap, nVec) * stride))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dUBatchCompact, 0, sizeof(double) * (size_t)std::max(c->batchCap, nVec) * stride));
        c->cStride = stride;
    }
    // chunks: a quarter of the batch, at least 4 vectors (a fused launch below ~300 CTAs is all latency), at most 8 chunks
    int chunk = std::max(4, (nVec + 3) / 4);
    while ((nVec + chunk - 1) / chunk > 8) ++chunk;
    const int nChunks = (nVec + chunk - 1) / chunk;
    const bool timed = c->stageTiming;
    c->resultMirrored = false;
    ++c->roundSerial;
    c->lastRoundOnHost = false;
    c->batchPull = RcFuse();
    DIPGPU_CUDA(c, cudaEventRecord(c->evChunk[8], st));            // (whatever the stream still holds comes first)
    DIPGPU_CUDA(c, cudaStreamWaitEvent(c->stream2, c->evChunk[8], 0));
    for (int k = 0; k < nChunks; ++k) {
        const int k0 = k * chunk, n = std::min(chunk, nVec - k0);
        if ((rc = launchGatherDuals(c, (int)ns, c->dSupport, dv + (size_t)k0 * (size_t)uLen, (int64_t)uLen,
                                    c->dUBatchCompact + (size_t)k0 * stride, (int64_t)stride, n, c->stream2))) return rc;
        DIPGPU_CUDA(c, cudaEventRecord(c->evChunk[k], c->stream2));
    }
    c->h2d += (int64_t)sizeof(double) * (int64_t)ns * nVec;
    if (timed) cudaEventRecord(c->ev[1], st);
    const RcFuse rf = rcFuseCompact(c, c->dUBatchCompact, (int64_t)stride, phase);
    for (int k = 0; k < nChunks; ++k) {
        const int k0 = k * chunk, n = std::min(chunk, nVec - k0);
        DIPGPU_CUDA(c, cudaStreamWaitEvent(st, c->evChunk[k], 0));
        if (timed && k == 0) cudaEventRecord(c->ev[2], st);       // "h2d" + "redcost" = the wait for the first gather
        c->wantMirror = true;
        rc = launchKnapsackDpBatch(c, c->dRedCostBatch, (int64_t)c->rcStride, nullptr, 0, eps, n, st, &rf, k0);
        c->wantMirror = false;
        if (rc) return rc;
    }
    if (timed) { cudaEventRecord(c->ev[3], st); cudaEventRecord(c->ev[4], st); cudaEventRecord(c->ev[5], st); }
    *took = 1;
    return DIPGPU_OK;
}

// nVec dual vectors priced in one submission and decoded into outs[0..nVec).
int batchCore(Ctx *c, int nVec, const DualFeed &f, int32_t uLen, int32_t phase, double eps, dipgpu_cols *outs)
{
    int rc;
    if ((rc = checkRoundReady(c, "price_rounds_batch", uLen, phase))) return rc;
    if (nVec == 0) return DIPGPU_OK;
    if ((rc = ensureBatch(c, nVec))) return rc;
    cudaStream_t st = c->stream;
    if (c->stageTiming) cudaEventRecord(c->ev[0], st);
    int took = 0;
    if ((rc = batchPipelined(c, nVec, f, uLen, phase, eps, st, &took))) return rc;
    if (took) {
        c->haveST = false;
        return fetchBatch(c, nVec, outs, st);
    }
    bool compact = false;
    if ((rc = feedBatchDuals(c, f, uLen, nVec, st, &compact))) return rc;
    c->haveST = false;
    if ((rc = enqueueBatch(c, nVec, phase, eps, st, compact))) return rc;
    c->lastRoundOnHost = false;
    return fetchBatch(c, nVec, outs, st);
}

// The rounds of the smoothing parameters al[0..nSteps) from ONE upload of dualRM (dipgpu_price_round_stab_ladder).
// nSteps == 0 (a device of a multi-device context without a step of its own): the upload and the first-call
// initialisation of the stability centre only, so that every device holds the same centre and dualRM.
int ladderCore(Ctx *c, const DualFeed &f, int32_t uLen, const double *al, int nSteps, int32_t phase, double eps, dipgpu_cols *outs)
{
    int rc;
    if ((rc = checkRoundReady(c, "price_round_stab_ladder", uLen, phase))) return rc;
    if (nSteps > 0 && (rc = ensureBatch(c, nSteps))) return rc;
    if ((rc = ensureDualBuffer(c, uLen))) return rc;
    cudaStream_t st = c->stream;
    if (c->stageTiming) cudaEventRecord(c->ev[0], st);
    if ((rc = feedDuals(c, f, uLen, 1, c->dU, 0, &c->dUClean, st))) return rc;
    c->dualsResident = true;
    c->dualsCompact = false;
    if (!c->haveCenter || c->centerLen != (size_t)uLen) {
        // first price call: m_dual := m_dualRM (DecompAlgoPC.cpp:163-173)
        if (c->centerLen != (size_t)uLen) {
            if ((rc = devAlloc(c, &c->dCenter, (size_t)uLen + 16))) return rc;
            c->centerLen = (size_t)uLen;
        }
        DIPGPU_CUDA(c, cudaMemcpyAsync(c->dCenter, c->dU, sizeof(double) * (size_t)uLen, cudaMemcpyDeviceToDevice, st));
        c->haveCenter = true;
    }
    if (nSteps == 0) {
        DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        return DIPGPU_OK;
    }
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dAlphaLadder, al, sizeof(double) * (size_t)nSteps, cudaMemcpyHostToDevice, st));
    c->h2d += (int64_t)sizeof(double) * nSteps;
    c->dUBatchClean = false;  // the smoothing writes every entry of the batch vectors
    if ((rc = launchSmoothDuals(c, c->dU, c->dCenter, c->dAlphaLadder, nSteps, uLen, c->dUBatch, (int64_t)c->uStride, st))) return rc;
    c->haveST = true;
    if ((rc = enqueueBatch(c, nSteps, phase, eps, st, false))) return rc;
    c->lastRoundOnHost = false;
    return fetchBatch(c, nSteps, outs, st);
}

}  // namespace dipgpu

using namespace dipgpu;

#define CTX_OR_FAIL(ctx)                        \
    Ctx *c = reinterpret_cast<Ctx *>(ctx);      \
    if (!c) return DIPGPU_ERR_INVALID;          \
    {                                           \
        int rc__ = bind(c);                     \
        if (rc__) return rc__;                  \
    }

extern "C" {

int dipgpu_abi_version(void) { return DIPGPU_ABI_VERSION; }

int dipgpu_create(dipgpu_ctx **ctx, int device)
{
    if (!ctx) return DIPGPU_ERR_INVALID;
    *ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        (void)cudaGetLastError();
        return DIPGPU_ERR_CUDA;  // no CPU fallback by design
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DIPGPU_ERR_CUDA;
    if (prop.major != 10) return DIPGPU_ERR_CUDA;  // kernels are built for sm_100a only
    Ctx *c = new (std::nothrow) Ctx();
    if (!c) return DIPGPU_ERR_NOMEM;
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return DIPGPU_ERR_CUDA;
    }
    for (auto &e : c->ev) cudaEventCreate(&e);
    unsigned int sync0[8 + 64] = {0u, 1u, 0u, 0u, 0u, 1u, 0u, 0u};  // epoch 1, no ticket drawn; expansion: (1 << 32) | 0; 64 pull counters
    if (cudaMalloc(reinterpret_cast<void **>(&c->dDoneCounter), sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(c->dDoneCounter, 0, sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&c->dSyncWords), sizeof(sync0)) != cudaSuccess ||
        cudaMemcpy(c->dSyncWords, sync0, sizeof(sync0), cudaMemcpyHostToDevice) != cudaSuccess) {
        if (c->dDoneCounter) cudaFree(c->dDoneCounter);
        if (c->dSyncWords) cudaFree(c->dSyncWords);
        delete c;
        return DIPGPU_ERR_NOMEM;
    }
    c->dExpTicket = reinterpret_cast<unsigned long long *>(c->dSyncWords + 4);
    c->dPullCtr = c->dSyncWords + 8;
    *ctx = reinterpret_cast<dipgpu_ctx *>(c);
    return DIPGPU_OK;
}

int dipgpu_destroy(dipgpu_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return DIPGPU_OK;
    if (c->multi) return multiDestroy(ctx);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    devFree(&c->dDoneCounter); devFree(&c->dSyncWords); devFree(&c->dCsrRowStart); devFree(&c->dDeltaCol); devFree(&c->dDeltaStart); devFree(&c->dDeltaRow); devFree(&c->dDeltaVal); devFree(&c->dSpmvInit);
    devFree(&c->dCStart); devFree(&c->dCIdx); devFree(&c->dCRow); devFree(&c->dCVal); devFree(&c->dAlphaIdxDense); devFree(&c->dAlphaIdxCompact); devFree(&c->dBlkEntFull); devFree(&c->dBlkEntCompact);
    devFree(&c->dUCompact); devFree(&c->dUBatchCompact); devFree(&c->dCsrColInd); devFree(&c->dCsrVal); devFree(&c->dExpPub);
    if (c->dMCols) cudaFree(c->dMCols);
    if (c->hDoneFlag) cudaFreeHost(c->hDoneFlag);
    if (c->hExpFlag) cudaFreeHost(c->hExpFlag);
    if (c->hMCols) cudaFreeHost(c->hMCols); devFree(&c->dPubWord); devFree(&c->dPubCells); devFree(&c->dDbgTimes); devFree(&c->dColStart); devFree(&c->dColStart32); devFree(&c->dTileRec); devFree(&c->dRowMaster16); devFree(&c->dSpmvBlob); devFree(&c->dSpmvWarpStart); devFree(&c->dRowMaster); devFree(&c->dVal); devFree(&c->dObj);
    devFree(&c->dBlockColStart); devFree(&c->dWeight); devFree(&c->dCapacity); devFree(&c->dBitsOff);
    devFree(&c->dBits); devFree(&c->dBigRowOff); devFree(&c->dBigRows); devFree(&c->dBigItemP); devFree(&c->dItemW); devFree(&c->dItemCol); devFree(&c->dNItems); devFree(&c->dScale); devFree(&c->dShortcut); devFree(&c->dSel); devFree(&c->dZShort);
    devFree(&c->dCandInd); devFree(&c->dScratch); devFree(&c->dU); devFree(&c->dRedCost); devFree(&c->dAlpha);
    freeBatch(c);
    devFree(&c->dFix); devFree(&c->dCapEff); devFree(&c->dRcEff);
    devFree(&c->dBlockGroupStart); devFree(&c->dGroupColStart);
    devFree(&c->dSupport); devFree(&c->dSupportBits); devFree(&c->dSupportRank);
    if (c->hSupVals) cudaFreeHost(c->hSupVals);
    devFree(&c->dCenter); devFree(&c->dPoolStart); devFree(&c->dPoolRow); devFree(&c->dPoolVal);
    devFree(&c->dPoolCStart); devFree(&c->dPoolCCount); devFree(&c->dPoolCRow); devFree(&c->dPoolCVal);
    if (c->dPoolScanTmp) cudaFree(c->dPoolScanTmp);
    devFree(&c->dPoolOrigCost); devFree(&c->dPoolRedCost); devFree(&c->dPoolFlag);
    if (c->dResult) cudaFree(c->dResult);
    if (c->hResult) cudaFreeHost(c->hResult);
    if (c->hStage) cudaFreeHost(c->hStage);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) {
        cudaStreamDestroy(c->stream2);
        for (auto &e : c->evChunk) if (e) cudaEventDestroy(e);
    }
    delete c;
    return DIPGPU_OK;
}

int dipgpu_last_error(const dipgpu_ctx *ctx, char *buf, size_t bufLen)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!buf || bufLen == 0) return DIPGPU_ERR_INVALID;
    // (a call a multi-device context handed to its primary device left its message there)
    if (c && c->multi && c->err.empty()) c = reinterpret_cast<const Ctx *>(multiPrimary(const_cast<dipgpu_ctx *>(ctx)));
    const char *s = c ? c->err.c_str() : "null context";
    snprintf(buf, bufLen, "%s", s);
    return DIPGPU_OK;
}

int dipgpu_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return DIPGPU_ERR_INVALID;
    // portable + mapped: every device of a multi-device context can read the dual vectors out of it
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? DIPGPU_ERR_NOMEM : DIPGPU_ERR_CUDA;
    }
    return DIPGPU_OK;
}

int dipgpu_host_free(void *ptr)
{
    if (!ptr) return DIPGPU_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? DIPGPU_OK : DIPGPU_ERR_CUDA;
}

int dipgpu_set_core_csr(dipgpu_ctx *ctx, int32_t R, int32_t N, const int64_t *rowStart, const int32_t *colInd,
                        const double *val, int32_t nBaseRows, int32_t numConvex)
{
    if (isMulti(ctx)) {
        struct A { int32_t R, N; const int64_t *rs; const int32_t *ci; const double *v; int32_t nb, nc; } a{R, N, rowStart, colInd, val, nBaseRows, numConvex};
        multiClearSupport(ctx);
        return multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_core_csr(sub, q->R, q->N, q->rs, q->ci, q->v, q->nb, q->nc);
        }, &a, true, true);
    }
    CTX_OR_FAIL(ctx);
    if (R < 0 || N < 0 || !rowStart || nBaseRows < 0 || nBaseRows > R || numConvex < 0)
        return fail(c, DIPGPU_ERR_INVALID, "set_core_csr: bad dimensions R=%d N=%d nBaseRows=%d numConvex=%d", R, N, nBaseRows, numConvex);
    if (rowStart[0] != 0) return fail(c, DIPGPU_ERR_INVALID, "set_core_csr: rowStart[0] != 0");
    for (int i = 0; i < R; ++i)
        if (rowStart[i + 1] < rowStart[i]) return fail(c, DIPGPU_ERR_INVALID, "set_core_csr: rowStart not monotone at row %d", i);
    const int64_t nnz = rowStart[R];
    if (nnz > 0 && (!colInd || !val)) return fail(c, DIPGPU_ERR_INVALID, "set_core_csr: NULL colInd/val");
    for (int64_t k = 0; k < nnz; ++k)
        if (colInd[k] < 0 || colInd[k] >= N) return fail(c, DIPGPU_ERR_INVALID, "set_core_csr: column index %d out of range at %lld", colInd[k], (long long)k);
    if (c->haveBlocks && c->nBlockCols > N) return fail(c, DIPGPU_ERR_INVALID, "set_core_csr: N=%d < block columns %d", N, c->nBlockCols);
    if (c->N != N) { devFree(&c->dObj); c->haveObj = false; c->hObj.clear(); c->objLen = 0; }
    if (c->N != N) { devFree(&c->dRedCost); devFree(&c->dRcEff); }
    c->R = R; c->N = N; c->nBaseRows = nBaseRows; c->numConvex = numConvex;
    c->dualsResident = false; c->haveCenter = false; c->haveST = false;
    c->dualsCompact = false; c->compactValid = false;
    c->hSupport.clear();
    ++c->supportSerial;
    // (with room for cut rows: an append must not copy the whole host CSR because a vector ran out of capacity)
    c->hRowStart.clear(); c->hColInd.clear(); c->hVal.clear();
    c->hRowStart.reserve((size_t)R + (size_t)R / 16 + 1024);
    c->hColInd.reserve((size_t)nnz + (size_t)nnz / 16 + 4096);
    c->hVal.reserve((size_t)nnz + (size_t)nnz / 16 + 4096);
    c->hRowStart.assign(rowStart, rowStart + R + 1);
    c->hColInd.assign(colInd, colInd + nnz);
    c->hVal.assign(val, val + nnz);
    int rc;
    if ((rc = uploadCore(c))) return rc;
    if (!c->dRedCost) {
        if ((rc = devAlloc(c, &c->dRedCost, (size_t)N + 16))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dRedCost, 0, sizeof(double) * ((size_t)N + 16)));
    }
    c->haveCore = true;
    return DIPGPU_OK;
}

// Rows [oldR, R) were appended to the host CSR: copy them behind the device's row-major copy and rebuild the (small)
// CSC of all rows appended since the last full build, [baseR, R).
static int appendDelta(Ctx *c, int oldR)
{
    int rc;
    const int R = c->R;
    const int64_t nnzOld = c->hRowStart[oldR], nnzNew = c->hRowStart[R];
    // ---- row-major copy (expand.cu): grow by half when the slack is used up (a device-to-device copy)
    if ((size_t)R > c->csrCapRows || (size_t)nnzNew > c->csrCapNnz) {
        const size_t capRows = std::max((size_t)R + (size_t)R / 2 + 64, c->csrCapRows);
        const size_t capNnz = std::max((size_t)nnzNew + (size_t)nnzNew / 2 + 1024, c->csrCapNnz);
        int32_t *rs = nullptr, *ci = nullptr;
        double *v = nullptr;
        if ((rc = devAlloc(c, &rs, capRows + 1))) return rc;
        if ((rc = devAlloc(c, &ci, capNnz + 1))) return rc;
        if ((rc = devAlloc(c, &v, capNnz + 1))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(rs, c->dCsrRowStart, sizeof(int32_t) * ((size_t)oldR + 1), cudaMemcpyDeviceToDevice));
        if (nnzOld > 0) {
            DIPGPU_CUDA(c, cudaMemcpy(ci, c->dCsrColInd, sizeof(int32_t) * (size_t)nnzOld, cudaMemcpyDeviceToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(v, c->dCsrVal, sizeof(double) * (size_t)nnzOld, cudaMemcpyDeviceToDevice));
        }
        devFree(&c->dCsrRowStart); devFree(&c->dCsrColInd); devFree(&c->dCsrVal);
        c->dCsrRowStart = rs; c->dCsrColInd = ci; c->dCsrVal = v;
        c->csrCapRows = capRows; c->csrCapNnz = capNnz;
    }
    {
        std::vector<int32_t> rs32((size_t)(R - oldR));
        for (int i = oldR + 1; i <= R; ++i) rs32[(size_t)(i - oldR - 1)] = (int32_t)c->hRowStart[(size_t)i];
        DIPGPU_CUDA(c, cudaMemcpy(c->dCsrRowStart + oldR + 1, rs32.data(), sizeof(int32_t) * rs32.size(), cudaMemcpyHostToDevice));
        if (nnzNew > nnzOld) {
            DIPGPU_CUDA(c, cudaMemcpy(c->dCsrColInd + nnzOld, c->hColInd.data() + nnzOld, sizeof(int32_t) * (size_t)(nnzNew - nnzOld), cudaMemcpyHostToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(c->dCsrVal + nnzOld, c->hVal.data() + nnzOld, sizeof(double) * (size_t)(nnzNew - nnzOld), cudaMemcpyHostToDevice));
        }
    }
    // ---- the delta's CSC: entries of a column in descending row order, the order the reference's scatter adds them
    struct E { int32_t col, row; double val; };
    std::vector<E> ent;
    ent.reserve((size_t)(nnzNew - c->baseNnz));
    for (int i = R - 1; i >= c->baseR; --i) {
        const int32_t rm = i < c->nBaseRows ? i : i + c->numConvex;
        for (int64_t k = c->hRowStart[(size_t)i]; k < c->hRowStart[(size_t)i + 1]; ++k) ent.push_back(E{c->hColInd[(size_t)k], rm, c->hVal[(size_t)k]});
    }
    std::stable_sort(ent.begin(), ent.end(), [](const E &x, const E &y) { return x.col < y.col; });
    std::vector<int32_t> cols, start, rows(ent.size());
    std::vector<double> vals(ent.size());
    for (size_t k = 0; k < ent.size(); ++k) {
        if (k == 0 || ent[k].col != ent[k - 1].col) { cols.push_back(ent[k].col); start.push_back((int32_t)k); }
        rows[k] = ent[k].row;
        vals[k] = ent[k].val;
    }
    start.push_back((int32_t)ent.size());
    if (cols.size() + 1 > c->deltaCapCols) {
        c->deltaCapCols = 2 * cols.size() + 64;
        if ((rc = devAlloc(c, &c->dDeltaCol, c->deltaCapCols))) return rc;
        if ((rc = devAlloc(c, &c->dDeltaStart, c->deltaCapCols + 1))) return rc;
    }
    if (ent.size() > c->deltaCapNnz) {
        c->deltaCapNnz = 2 * ent.size() + 64;
        if ((rc = devAlloc(c, &c->dDeltaRow, c->deltaCapNnz))) return rc;
        if ((rc = devAlloc(c, &c->dDeltaVal, c->deltaCapNnz))) return rc;
    }
    // (synchronous copies: a sweep enqueued earlier has finished with the old arrays)
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!cols.empty()) {
        DIPGPU_CUDA(c, cudaMemcpy(c->dDeltaCol, cols.data(), sizeof(int32_t) * cols.size(), cudaMemcpyHostToDevice));
        DIPGPU_CUDA(c, cudaMemcpy(c->dDeltaRow, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice));
        DIPGPU_CUDA(c, cudaMemcpy(c->dDeltaVal, vals.data(), sizeof(double) * vals.size(), cudaMemcpyHostToDevice));
    }
    DIPGPU_CUDA(c, cudaMemcpy(c->dDeltaStart, start.data(), sizeof(int32_t) * start.size(), cudaMemcpyHostToDevice));
    c->nDeltaCols = (int)cols.size();
    c->deltaNnz = (int64_t)ent.size();
    c->nnz = nnzNew;
    return DIPGPU_OK;
}

int dipgpu_append_core_rows(dipgpu_ctx *ctx, int32_t nNew, const int64_t *rowStart, const int32_t *colInd, const double *val)
{
    if (isMulti(ctx)) {
        struct A { int32_t n; const int64_t *rs; const int32_t *ci; const double *v; } a{nNew, rowStart, colInd, val};
        multiClearSupport(ctx);
        return multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_append_core_rows(sub, q->n, q->rs, q->ci, q->v);
        }, &a, true, true);
    }
    CTX_OR_FAIL(ctx);
    if (!c->haveCore) return fail(c, DIPGPU_ERR_STATE, "append_core_rows before set_core_csr");
    if (nNew < 0 || !rowStart || rowStart[0] != 0) return fail(c, DIPGPU_ERR_INVALID, "append_core_rows: bad arguments");
    for (int i = 0; i < nNew; ++i)
        if (rowStart[i + 1] < rowStart[i]) return fail(c, DIPGPU_ERR_INVALID, "append_core_rows: rowStart not monotone");
    const int64_t add = rowStart[nNew];
    for (int64_t k = 0; k < add; ++k)
        if (colInd[k] < 0 || colInd[k] >= c->N) return fail(c, DIPGPU_ERR_INVALID, "append_core_rows: column index out of range");
    const int64_t base = c->hRowStart[c->R];
    for (int i = 1; i <= nNew; ++i) c->hRowStart.push_back(base + rowStart[i]);
    c->hColInd.insert(c->hColInd.end(), colInd, colInd + add);
    c->hVal.insert(c->hVal.end(), val, val + add);
    c->R += nNew;
    c->dualsResident = false; c->haveST = false;  // the master dual vector grew
    c->dualsCompact = false; c->compactValid = false;
    c->hSupport.clear();
    ++c->supportSerial;
    if (c->haveCenter) {
        // DecompAlgoPC::adjustMasterDualSolution resizes m_dual to the new row count (Dip/src/DecompAlgoPC.cpp:135-140):
        // the stability centre is kept and zero-extended -- the cut rows sit at the end of the master dual vector
        const size_t newLen = (size_t)c->R + (size_t)c->numConvex;
        if (newLen > c->centerLen) {
            double *nw = nullptr;
            int rc;
            if ((rc = devAlloc(c, &nw, newLen + 16))) return rc;
            DIPGPU_CUDA(c, cudaMemset(nw, 0, sizeof(double) * (newLen + 16)));
            DIPGPU_CUDA(c, cudaMemcpy(nw, c->dCenter, sizeof(double) * c->centerLen, cudaMemcpyDeviceToDevice));
            devFree(&c->dCenter);
            c->dCenter = nw;
            c->centerLen = newLen;
        }
    }
    // Large model whose tiles are in use: the new rows go into the delta (O(new entries)), nothing is re-transposed
    // or re-tiled.  Otherwise (small model, tiles not built yet or about to be rebuilt, the lane-cooperative kernel,
    // the delta grown past an eighth of the matrix) everything is rebuilt with the new rows in it.
    const int64_t total = c->hRowStart[c->R];
    const int64_t deltaAfter = total - c->baseNnz;
    const bool useDelta = nNew > 0 && c->baseNnz >= c->optDeltaMinNnz && !c->tilesDirty && c->nSpmvTiles > 0 && c->spmvLanes == 0 &&
                          c->dCsrRowStart != nullptr && total < ((int64_t)1 << 31) && deltaAfter * 8 <= c->baseNnz;
    if (!useDelta) return uploadCore(c);
    return appendDelta(c, c->R - nNew);
}

int dipgpu_set_objective(dipgpu_ctx *ctx, const double *obj, int32_t N)
{
    if (isMulti(ctx)) {
        struct A { const double *c; int32_t N; } a{obj, N};
        return multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_objective(sub, q->c, q->N);
        }, &a, true, true);
    }
    CTX_OR_FAIL(ctx);
    if (!obj || N < 0) return fail(c, DIPGPU_ERR_INVALID, "set_objective: bad arguments");
    if (c->haveCore && N != c->N) return fail(c, DIPGPU_ERR_INVALID, "set_objective: N=%d but core has %d columns", N, c->N);
    if (!c->haveCore && c->haveBlocks && N < c->nBlockCols) return fail(c, DIPGPU_ERR_INVALID, "set_objective: N=%d < block columns %d", N, c->nBlockCols);
    int rc;
    devFree(&c->dObj);
    if ((rc = devAlloc(c, &c->dObj, (size_t)N + 16))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dObj, 0, sizeof(double) * ((size_t)N + 16)));
    DIPGPU_CUDA(c, cudaMemcpy(c->dObj, obj, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice));
    c->objLen = (size_t)N;
    if (!c->haveCore) c->N = std::max(c->N, (int)N);
    c->haveObj = true;
    c->hObj.assign(obj, obj + N);
    c->tilesDirty = true;  // the SpMV tiles carry their columns' objective entries
    return DIPGPU_OK;
}

int dipgpu_set_knapsack_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockColStart, const int32_t *weight,
                               const int32_t *capacity, int32_t profitMode)
{
    if (isMulti(ctx)) {
        struct A { int32_t m; const int32_t *bcs, *w, *cap; int32_t pm; } a{m, blockColStart, weight, capacity, profitMode};
        const int rc = multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_knapsack_blocks(sub, q->m, q->bcs, q->w, q->cap, q->pm);
        }, &a, true, true);
        return rc ? rc : multiSetBlocks(ctx);
    }
    CTX_OR_FAIL(ctx);
    if (m < 0 || !blockColStart || (m > 0 && !capacity)) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: bad arguments");
    if (profitMode != DIPGPU_PROFIT_FP64 && profitMode != DIPGPU_PROFIT_PISINGER_INT)
        return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: unknown profitMode %d", profitMode);
    if (blockColStart[0] < 0) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: negative column start");
    for (int b = 0; b < m; ++b) {
        if (blockColStart[b + 1] < blockColStart[b]) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: blockColStart not monotone at block %d", b);
        if (capacity[b] < 0) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: negative capacity at block %d", b);
    }
    const int nCols = blockColStart[m];
    if (nCols > 0 && !weight) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: NULL weight");
    for (int j = blockColStart[0]; j < nCols; ++j)
        if (weight[j] < 0) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: negative weight at column %d", j);
    if (c->haveCore && nCols > c->N) return fail(c, DIPGPU_ERR_INVALID, "set_knapsack_blocks: %d block columns > N=%d", nCols, c->N);
    c->m = m;
    c->nBlockCols = nCols;
    c->profitMode = profitMode;
    c->blockKind = 0;
    c->hBlockColStart.assign(blockColStart, blockColStart + m + 1);
    c->hCapacity.assign(capacity, capacity + m);
    c->hWeight.assign(weight, weight + nCols);
    c->haveFix = false;  // node bounds refer to the previous block structure
    c->hMaxUsableW.assign((size_t)m, 0);
    for (int b = 0; b < m; ++b) {
        int mw = 0;
        for (int j = blockColStart[b]; j < blockColStart[b + 1]; ++j)
            if (weight[j] <= capacity[b]) mw = std::max(mw, weight[j]);
        c->hMaxUsableW[b] = mw;
    }
    int rc;
    devFree(&c->dFix); devFree(&c->dCapEff);
    if ((rc = devAlloc(c, &c->dBlockColStart, (size_t)m + 1))) return rc;
    if ((rc = devAlloc(c, &c->dCapacity, (size_t)m))) return rc;
    if ((rc = devAlloc(c, &c->dWeight, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dItemW, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dItemCol, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dCandInd, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dScratch, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dAlpha, (size_t)m))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dBlockColStart, blockColStart, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice));
    if (m > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dCapacity, capacity, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemset(c->dWeight, 0, sizeof(int32_t) * ((size_t)nCols + 16)));
    if (nCols > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dWeight, weight, sizeof(int32_t) * (size_t)nCols, cudaMemcpyHostToDevice));
    if (!c->haveCore) {
        c->N = std::max(c->N, nCols);
        if ((rc = devAlloc(c, &c->dRedCost, (size_t)c->N + 16))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dRedCost, 0, sizeof(double) * ((size_t)c->N + 16)));
    }
    c->haveBlocks = true;
    c->compactValid = false;  // (the convexity duals' positions are per block)
    c->lastRoundOnHost = false;
    return setBlockRange(c, 0, m);
}

int dipgpu_set_mckp_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockGroupStart, const int32_t *groupColStart,
                           const int32_t *weight, const int32_t *capacity)
{
    if (isMulti(ctx)) {
        struct A { int32_t m; const int32_t *bgs, *gcs, *w, *cap; } a{m, blockGroupStart, groupColStart, weight, capacity};
        const int rc = multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_mckp_blocks(sub, q->m, q->bgs, q->gcs, q->w, q->cap);
        }, &a, true, true);
        return rc ? rc : multiSetBlocks(ctx);
    }
    CTX_OR_FAIL(ctx);
    if (m < 0 || !blockGroupStart || !groupColStart || (m > 0 && !capacity) || blockGroupStart[0] != 0 || groupColStart[0] < 0)
        return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: bad arguments");
    for (int b = 0; b < m; ++b) {
        if (blockGroupStart[b + 1] < blockGroupStart[b]) return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: blockGroupStart not monotone at block %d", b);
        if (capacity[b] < 0) return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: negative capacity at block %d", b);
    }
    const int nGroups = blockGroupStart[m];
    for (int g = 0; g < nGroups; ++g) {
        const int n = groupColStart[g + 1] - groupColStart[g];
        if (n < 1 || n >= 0xffff) return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: group %d has %d columns (1..65534)", g, n);
    }
    const int nCols = groupColStart[nGroups];
    if (nCols > 0 && !weight) return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: NULL weight");
    for (int j = groupColStart[0]; j < nCols; ++j)
        if (weight[j] < 0) return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: negative weight at column %d", j);
    if (c->haveCore && nCols > c->N) return fail(c, DIPGPU_ERR_INVALID, "set_mckp_blocks: %d block columns > N=%d", nCols, c->N);
    c->m = m;
    c->nBlockCols = nCols;
    c->profitMode = DIPGPU_PROFIT_FP64;
    c->blockKind = 1;
    c->haveFix = false;
    c->hBlockGroupStart.assign(blockGroupStart, blockGroupStart + m + 1);
    c->hGroupColStart.assign(groupColStart, groupColStart + nGroups + 1);
    c->hCapacity.assign(capacity, capacity + m);
    c->hWeight.assign(weight, weight + nCols);
    c->mckpMaxWeight = 0;
    for (int j = groupColStart[0]; j < nCols; ++j) c->mckpMaxWeight = std::max(c->mckpMaxWeight, weight[j]);
    c->hBlockColStart.assign((size_t)m + 1, 0);
    for (int b = 0; b <= m; ++b) c->hBlockColStart[b] = groupColStart[blockGroupStart[b]];
    c->hMaxUsableW.assign((size_t)m, 0);
    int rc;
    devFree(&c->dFix); devFree(&c->dCapEff);
    if ((rc = devAlloc(c, &c->dBlockColStart, (size_t)m + 1))) return rc;
    if ((rc = devAlloc(c, &c->dBlockGroupStart, (size_t)m + 1))) return rc;
    if ((rc = devAlloc(c, &c->dGroupColStart, (size_t)nGroups + 1))) return rc;
    if ((rc = devAlloc(c, &c->dCapacity, (size_t)m))) return rc;
    if ((rc = devAlloc(c, &c->dWeight, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dItemW, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dItemCol, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dCandInd, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dScratch, (size_t)nCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dAlpha, (size_t)m))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dBlockColStart, c->hBlockColStart.data(), sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dBlockGroupStart, blockGroupStart, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dGroupColStart, groupColStart, sizeof(int32_t) * ((size_t)nGroups + 1), cudaMemcpyHostToDevice));
    if (m > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dCapacity, capacity, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemset(c->dWeight, 0, sizeof(int32_t) * ((size_t)nCols + 16)));
    if (nCols > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dWeight, weight, sizeof(int32_t) * (size_t)nCols, cudaMemcpyHostToDevice));
    if (!c->haveCore) {
        c->N = std::max(c->N, nCols);
        if ((rc = devAlloc(c, &c->dRedCost, (size_t)c->N + 16))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dRedCost, 0, sizeof(double) * ((size_t)c->N + 16)));
    }
    c->haveBlocks = true;
    c->compactValid = false;  // (the convexity duals' positions are per block)
    c->lastRoundOnHost = false;
    return setBlockRange(c, 0, m);
}

int dipgpu_set_intknap_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockColStart, const int32_t *weight,
                              const int32_t *capacity, const int32_t *useCol)
{
    if (isMulti(ctx)) {
        struct A { int32_t m; const int32_t *bcs, *w, *cap, *use; } a{m, blockColStart, weight, capacity, useCol};
        const int rc = multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_intknap_blocks(sub, q->m, q->bcs, q->w, q->cap, q->use);
        }, &a, true, true);
        return rc ? rc : multiSetBlocks(ctx);
    }
    CTX_OR_FAIL(ctx);
    if (m < 0 || !blockColStart || (m > 0 && !capacity) || blockColStart[0] < 0)
        return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: bad arguments");
    for (int b = 0; b < m; ++b) {
        if (blockColStart[b + 1] < blockColStart[b]) return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: blockColStart not monotone at block %d", b);
        if (capacity[b] < 0) return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: negative capacity at block %d", b);
    }
    const int nItemCols = blockColStart[m];
    if (nItemCols > 0 && !weight) return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: NULL weight");
    // kp recurses on capacity - weight (test_cutting_stock.py:171): a weight below 1 never terminates there
    for (int j = blockColStart[0]; j < nItemCols; ++j)
        if (weight[j] < 1) return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: weight %d at column %d (must be >= 1)", weight[j], j);
    int nCols = nItemCols;
    if (useCol) {
        for (int b = 0; b < m; ++b) {
            const int u = useCol[b];
            if (u < -1) return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: useCol[%d] = %d", b, u);
            if (u >= blockColStart[b] && u < blockColStart[b + 1])
                return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: the use column of block %d is one of its item columns", b);
            nCols = std::max(nCols, u + 1);
        }
    }
    if (c->haveCore && nCols > c->N) return fail(c, DIPGPU_ERR_INVALID, "set_intknap_blocks: column %d beyond N=%d", nCols - 1, c->N);
    c->m = m;
    c->nBlockCols = nCols;  // every column the blocks read (the use columns may lie behind the item columns)
    c->profitMode = DIPGPU_PROFIT_FP64;
    c->blockKind = 2;
    c->haveFix = false;
    c->hBlockColStart.assign(blockColStart, blockColStart + m + 1);
    c->hCapacity.assign(capacity, capacity + m);
    c->hWeight.assign(weight, weight + nItemCols);
    c->hUseCol.clear();
    if (useCol) c->hUseCol.assign(useCol, useCol + m);
    c->hMaxUsableW.assign((size_t)m, 0);
    std::vector<int32_t> candStart((size_t)m + 1);
    for (int b = 0; b <= m; ++b) candStart[(size_t)b] = 2 * (blockColStart[b] - blockColStart[0] + b);
    int rc;
    devFree(&c->dFix); devFree(&c->dCapEff); devFree(&c->dUseCol);
    if ((rc = devAlloc(c, &c->dBlockColStart, (size_t)m + 1))) return rc;
    if ((rc = devAlloc(c, &c->dCandStart, (size_t)m + 1))) return rc;
    if ((rc = devAlloc(c, &c->dCapacity, (size_t)m))) return rc;
    if ((rc = devAlloc(c, &c->dWeight, (size_t)nItemCols + 16))) return rc;
    if ((rc = devAlloc(c, &c->dCandInd, 2 * ((size_t)nItemCols + (size_t)m) + 16))) return rc;
    if ((rc = devAlloc(c, &c->dAlpha, (size_t)m))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dBlockColStart, blockColStart, sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dCandStart, candStart.data(), sizeof(int32_t) * ((size_t)m + 1), cudaMemcpyHostToDevice));
    if (m > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dCapacity, capacity, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemset(c->dWeight, 0, sizeof(int32_t) * ((size_t)nItemCols + 16)));
    if (nItemCols > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dWeight, weight, sizeof(int32_t) * (size_t)nItemCols, cudaMemcpyHostToDevice));
    if (useCol && m > 0) {
        if ((rc = devAlloc(c, &c->dUseCol, (size_t)m))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dUseCol, useCol, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice));
    }
    if (!c->haveCore) {
        c->N = std::max(c->N, nCols);
        if ((rc = devAlloc(c, &c->dRedCost, (size_t)c->N + 16))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dRedCost, 0, sizeof(double) * ((size_t)c->N + 16)));
    }
    c->haveBlocks = true;
    c->compactValid = false;
    c->lastRoundOnHost = false;
    return setBlockRange(c, 0, m);
}

int dipgpu_set_col_bounds(dipgpu_ctx *ctx, const double *colLB, const double *colUB, int32_t N)
{
    if (isMulti(ctx)) {
        struct A { const double *lb, *ub; int32_t N; } a{colLB, colUB, N};
        return multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_col_bounds(sub, q->lb, q->ub, q->N);
        }, &a, true, true);
    }
    CTX_OR_FAIL(ctx);
    if (!c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "set_col_bounds before set_knapsack_blocks");
    c->lastRoundOnHost = false;
    if (!colLB && !colUB) {
        c->haveFix = false;
        return DIPGPU_OK;
    }
    if (!colLB || !colUB || N < c->nBlockCols) return fail(c, DIPGPU_ERR_INVALID, "set_col_bounds: need both bound arrays of length >= %d", c->nBlockCols);
    if (c->blockKind != 0)
        return fail(c, DIPGPU_ERR_UNSUPPORTED, "set_col_bounds: node bounds are enforced in 0-1 knapsack blocks only");
    const int n = c->nBlockCols;
    std::vector<uint8_t> fix((size_t)n + 16, 0);
    std::vector<int32_t> capEff((size_t)std::max(c->m, 1), 0);
    for (int b = 0; b < c->m; ++b) {
        int64_t left = c->hCapacity[b];
        bool clash = false;
        for (int j = c->hBlockColStart[b]; j < c->hBlockColStart[b + 1]; ++j) {
            const bool one = colLB[j] > 0.5, zero = colUB[j] < 0.5;  // binary columns: bounds are 0 or 1
            fix[j] = (uint8_t)((one ? 1 : 0) | (zero ? 2 : 0));
            if (one) left -= c->hWeight[j];
            if (one && zero) clash = true;
        }
        capEff[b] = (clash || left < 0) ? -1 : (int32_t)left;
    }
    int rc;
    if (!c->dFix && (rc = devAlloc(c, &c->dFix, (size_t)n + 16))) return rc;
    if (!c->dCapEff && (rc = devAlloc(c, &c->dCapEff, (size_t)std::max(c->m, 1)))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dFix, fix.data(), fix.size(), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dCapEff, capEff.data(), sizeof(int32_t) * capEff.size(), cudaMemcpyHostToDevice));
    c->hCapEff = capEff;
    c->h2d += (int64_t)fix.size() + 4 * c->m;
    c->haveFix = true;
    return DIPGPU_OK;
}

int dipgpu_set_block_range(dipgpu_ctx *ctx, int32_t first, int32_t count)
{
    if (isMulti(ctx)) return fail(reinterpret_cast<Ctx *>(ctx), DIPGPU_ERR_UNSUPPORTED, "set_block_range: a multi-device context cuts its own block shards");
    CTX_OR_FAIL(ctx);
    return setBlockRange(c, first, count);
}

int dipgpu_calc_redcost(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double *redCostX)
{
    if (isMulti(ctx)) return multiCalcRedCost(ctx, u, uLen, phase, redCostX);
    CTX_OR_FAIL(ctx);
    if (!c->haveCore) return fail(c, DIPGPU_ERR_STATE, "calc_redcost before set_core_csr");
    if (phase != DIPGPU_PHASE_PRICE1 && !c->haveObj) return fail(c, DIPGPU_ERR_STATE, "calc_redcost (phase 2) before set_objective");
    if (!u || !redCostX || uLen != c->R + c->numConvex)
        return fail(c, DIPGPU_ERR_INVALID, "calc_redcost: uLen=%d, expected R+numConvex=%d", uLen, c->R + c->numConvex);
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    if ((size_t)uLen > c->uCap) {
        if ((rc = devAlloc(c, &c->dU, (size_t)uLen + 16))) return rc;
        c->uCap = (size_t)uLen;
    }
    if ((rc = uploadDuals(c, u, uLen, 1, c->dU, 0, &c->dUClean, c->stream))) return rc;
    c->dualsResident = true;
    c->dualsCompact = false;
    if ((rc = launchRedCost(c, c->dU, phase, c->stream))) return rc;
    DIPGPU_CUDA(c, cudaMemcpyAsync(redCostX, c->dRedCost, sizeof(double) * (size_t)c->N, cudaMemcpyDeviceToHost, c->stream));
    c->d2h += (int64_t)sizeof(double) * c->N;
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return DIPGPU_OK;
}

int dipgpu_solve_blocks(dipgpu_ctx *ctx, const double *redCostX, const double *alpha, double redCostEps, dipgpu_cols *out)
{
    if (isMulti(ctx)) return multiSolveBlocks(ctx, redCostX, alpha, redCostEps, out);
    CTX_OR_FAIL(ctx);
    if (!c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "solve_blocks before set_knapsack_blocks");
    if (!redCostX || !alpha || !out) return fail(c, DIPGPU_ERR_INVALID, "solve_blocks: NULL argument");
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    const bool timed = c->stageTiming;
    cudaStream_t st = c->stream;
    if (timed) cudaEventRecord(c->ev[0], st);
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dRedCost, redCostX, sizeof(double) * (size_t)c->nBlockCols, cudaMemcpyHostToDevice, st));
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dAlpha, alpha, sizeof(double) * (size_t)c->m, cudaMemcpyHostToDevice, st));
    c->h2d += (int64_t)sizeof(double) * ((int64_t)c->nBlockCols + c->m);
    if (timed) { cudaEventRecord(c->ev[1], st); cudaEventRecord(c->ev[2], st); }
    c->wantMirror = true;
    rc = enqueueSolve(c, c->dRedCost, c->dAlpha, redCostEps, st, timed);
    c->wantMirror = false;
    if (rc) return rc;
    if ((rc = fetchResult(c, st))) return rc;
    c->lastRoundOnHost = true;
    collectStageMs(c);
    return unpack(c, c->hResult, c->resultBytes, out);
}

int dipgpu_price_round(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double redCostEps,
                       dipgpu_cols *out, double *redCostXOpt)
{
    if (isMulti(ctx)) return multiPriceRound(ctx, u, uLen, phase, redCostEps, out, redCostXOpt);
    CTX_OR_FAIL(ctx);
    if (!out) return fail(c, DIPGPU_ERR_INVALID, "price_round: NULL output");
    int rc;
    if ((rc = checkRoundReady(c, "price_round", uLen, phase))) return rc;
    if (!u && !c->dualsResident && !c->dualsCompact) return fail(c, DIPGPU_ERR_STATE, "price_round: u == NULL but no dual vector is resident");
    cudaStream_t st = c->stream;
    DualFeed f;
    f.host = u;
    const auto tEntry = std::chrono::steady_clock::now();
    c->hostT0 = std::chrono::duration_cast<std::chrono::nanoseconds>(tEntry.time_since_epoch()).count();
    auto usSince = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tEntry).count(); };
    c->wantFlag = redCostXOpt == nullptr;  // (the copy of the reduced costs needs the stream synchronised anyway)
    rc = enqueueRound(c, u ? &f : nullptr, uLen, phase, redCostEps, true, st, redCostXOpt != nullptr);
    c->wantFlag = false;
    if (rc) return rc;
    if (c->hostTiming) c->hostUs[0] = usSince();
    if (redCostXOpt) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(redCostXOpt, c->dRedCost, sizeof(double) * (size_t)c->N, cudaMemcpyDeviceToHost, st));
        c->d2h += (int64_t)sizeof(double) * c->N;
    }
    if ((rc = fetchResult(c, st))) return rc;
    if (c->hostTiming) c->hostUs[1] = usSince();
    c->lastRoundOnHost = true;
    collectStageMs(c);
    rc = unpack(c, c->hResult, c->resultBytes, out);
    if (c->hostTiming) c->hostUs[2] = usSince();
    return rc;
}

// ------------------------------------------------- batched / stabilised rounds

int dipgpu_price_rounds_batch(dipgpu_ctx *ctx, int32_t nVec, const double *U, int32_t uLen, int32_t phase,
                              double redCostEps, dipgpu_cols *outs)
{
    if (isMulti(ctx)) return multiPriceRoundsBatch(ctx, nVec, U, uLen, phase, redCostEps, outs);
    CTX_OR_FAIL(ctx);
    if (nVec < 0 || (nVec > 0 && (!U || !outs))) return fail(c, DIPGPU_ERR_INVALID, "price_rounds_batch: bad arguments");
    DualFeed f;
    f.host = U;
    return batchCore(c, nVec, f, uLen, phase, redCostEps, outs);
}

// Block subset with per-block user duals: every distinct dual vector is priced as one vector of
// a batch (all blocks of the shard -- they run side by side, a subset costs the same), then the listed blocks
// are picked out of the packed results of "their" vector on the host.
int dipgpu_price_round_which(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double redCostEps,
                             int32_t nWhich, const int32_t *whichBlocks, const double *const *userDuals,
                             dipgpu_cols *out, int32_t *solvedAll)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (!u || !out || nWhich < 0 || (nWhich > 0 && !whichBlocks)) return fail(c, DIPGPU_ERR_INVALID, "price_round_which: bad arguments");
    int rc;
    if ((rc = checkRoundReady(c, "price_round_which", uLen, phase))) return rc;
    const int m = c->nLocal, first = c->firstBlock;
    std::vector<int> vecOf((size_t)std::max(m, 1), -1);  // local block -> vector that prices it (-1: not listed)
    std::vector<const double *> vecs(1, u);
    for (int i = 0; i < nWhich; ++i) {
        const int lb = whichBlocks[i] - first;
        if (lb < 0 || lb >= m) return fail(c, DIPGPU_ERR_INVALID, "price_round_which: block %d outside this context's range [%d, %d)", whichBlocks[i], first, first + m);
        if (vecOf[lb] >= 0) return fail(c, DIPGPU_ERR_INVALID, "price_round_which: block %d listed twice", whichBlocks[i]);
        const double *ub = userDuals ? userDuals[i] : nullptr;
        int v = 0;
        if (ub && ub != u) {
            v = (int)(std::find(vecs.begin(), vecs.end(), ub) - vecs.begin());
            if (v == (int)vecs.size()) vecs.push_back(ub);
        }
        vecOf[lb] = v;
    }
    const int nVec = (int)vecs.size();
    std::vector<double> U;
    const double *Uptr = u;
    if (nVec > 1) {
        U.resize((size_t)nVec * (size_t)uLen);
        for (int v = 0; v < nVec; ++v) memcpy(U.data() + (size_t)v * uLen, vecs[v], sizeof(double) * (size_t)uLen);
        Uptr = U.data();
    }
    if ((rc = ensureBatch(c, nVec))) return rc;
    cudaStream_t st = c->stream;
    if (c->stageTiming) cudaEventRecord(c->ev[0], st);
    bool compact = false;
    {
        DualFeed f;
        f.host = Uptr;
        if ((rc = feedBatchDuals(c, f, uLen, nVec, st, &compact))) return rc;
    }
    c->haveST = false;
    if ((rc = enqueueBatch(c, nVec, phase, redCostEps, st, compact))) return rc;
    c->lastRoundOnHost = false;
    if ((rc = fetchBatchPacked(c, nVec, st))) return rc;

    const ResultLayout &L = c->layout;
    auto packed = [&](int v) { return static_cast<const char *>(c->hResultBatch) + (size_t)v * c->resStride; };
    auto dbl = [](const char *p, size_t off) { return reinterpret_cast<const double *>(p + off); };
    bool found = false;
    for (int lb = 0; lb < m && !found; ++lb)
        if (vecOf[lb] >= 0 && dbl(packed(vecOf[lb]), L.offRedCost)[lb] < -redCostEps) found = true;  // strict (:5216)
    char *dst = static_cast<char *>(c->hResult);
    const ResultHeader *h0 = reinterpret_cast<const ResultHeader *>(packed(0));
    long long cells = 0, evals = 0;
    for (int v = 0; v < nVec; ++v) {
        const ResultHeader *hv = reinterpret_cast<const ResultHeader *>(packed(v));
        cells += hv->cells;
        evals += hv->reserved[0];
    }
    double *dRc = reinterpret_cast<double *>(dst + L.offRedCost), *dMn = reinterpret_cast<double *>(dst + L.offMostNeg);
    double *dOc = reinterpret_cast<double *>(dst + L.offOrigCost), *dOb = reinterpret_cast<double *>(dst + L.offObj);
    int64_t *dKs = reinterpret_cast<int64_t *>(dst + L.offKeptStart);
    int32_t *dNz = reinterpret_cast<int32_t *>(dst + L.offNnz), *dKb = reinterpret_cast<int32_t *>(dst + L.offKeptBlock);
    int32_t *dKi = reinterpret_cast<int32_t *>(dst + L.offInd);
    ResultHeader hd = *h0;
    if (!found) {
        // no listed block prices out: all blocks at the standard duals (DecompAlgo.cpp:5076-5148) -- that round is
        // vector 0 of the batch.  The listed blocks' own candidates stay in the list the filter walks, so a
        // block priced with user duals contributes min(both reduced costs, 0) to mostNegReducedCost (:5212-5214).
        memcpy(dst, packed(0), L.offInd + sizeof(int32_t) * (size_t)h0->nInd);
        for (int lb = 0; lb < m; ++lb)
            if (vecOf[lb] > 0) dMn[lb] = std::min(dMn[lb], std::min(0.0, dbl(packed(vecOf[lb]), L.offRedCost)[lb]));
    } else {
        int nKept = 0;
        int64_t nInd = 0;
        for (int lb = 0; lb < m; ++lb) {
            const int v = vecOf[lb];
            if (v < 0) {  // not solved this round
                dRc[lb] = dMn[lb] = dOc[lb] = dOb[lb] = 0.0;
                dNz[lb] = 0;
                continue;
            }
            const char *src = packed(v);
            dRc[lb] = dbl(src, L.offRedCost)[lb];
            dMn[lb] = dbl(src, L.offMostNeg)[lb];
            dOc[lb] = dbl(src, L.offOrigCost)[lb];
            dOb[lb] = dbl(src, L.offObj)[lb];
            dNz[lb] = reinterpret_cast<const int32_t *>(src + L.offNnz)[lb];
            if (!(dRc[lb] < -redCostEps)) continue;
            const ResultHeader *hv = reinterpret_cast<const ResultHeader *>(src);
            const int32_t *kb = reinterpret_cast<const int32_t *>(src + L.offKeptBlock);
            const int64_t *ks = reinterpret_cast<const int64_t *>(src + L.offKeptStart);
            const int32_t *pos = std::lower_bound(kb, kb + hv->nKept, lb);
            if (pos == kb + hv->nKept || *pos != lb) return fail(c, DIPGPU_ERR_CUDA, "price_round_which: block %d priced out but is not among the kept columns", first + lb);
            const int k = (int)(pos - kb);
            const int64_t len = ks[k + 1] - ks[k];
            dKb[nKept] = lb;
            dKs[nKept] = nInd;
            memcpy(dKi + nInd, reinterpret_cast<const int32_t *>(src + L.offInd) + ks[k], sizeof(int32_t) * (size_t)len);
            nInd += len;
            ++nKept;
        }
        dKs[nKept] = nInd;
        hd.nKept = nKept;
        hd.nInd = nInd;
    }
    hd.cells = cells;
    hd.reserved[0] = evals;
    memcpy(dst, &hd, sizeof(hd));
    if (solvedAll) *solvedAll = found ? 0 : 1;
    // the composed round becomes the context's "last round" (dipgpu_expand_columns reads the device copy)
    const size_t need = L.offInd + sizeof(int32_t) * (size_t)hd.nInd;
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dResult, c->hResult, need, cudaMemcpyHostToDevice, st));
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    c->lastResultBytes = need;
    c->lastRoundOnHost = true;
    return unpack(c, c->hResult, c->resultBytes, out);
}

int dipgpu_set_dual_support(dipgpu_ctx *ctx, int32_t nIdx, const int32_t *idx)
{
    if (isMulti(ctx)) {
        if (nIdx < 0 || (nIdx > 0 && !idx)) return fail(reinterpret_cast<Ctx *>(ctx), DIPGPU_ERR_INVALID, "set_dual_support: bad arguments");
        return multiSetDualSupport(ctx, nIdx, idx);
    }
    CTX_OR_FAIL(ctx);
    if (!c->haveCore) return fail(c, DIPGPU_ERR_STATE, "set_dual_support before set_core_csr");
    if (nIdx < 0 || (nIdx > 0 && !idx)) return fail(c, DIPGPU_ERR_INVALID, "set_dual_support: bad arguments");
    const int uLen = c->R + c->numConvex;
    for (int k = 0; k < nIdx; ++k)
        if (idx[k] < 0 || idx[k] >= uLen || (k > 0 && idx[k] <= idx[k - 1]))
            return fail(c, DIPGPU_ERR_INVALID, "set_dual_support: indices must be ascending, unique and below %d", uLen);
    c->hSupport.assign(idx, idx + nIdx);
    ++c->supportSerial;
    c->supVecCap = 0;          // staging is sized by the support
    c->dUClean = false;        // whatever the device vectors hold now may be nonzero outside the new support
    c->dUBatchClean = false;
    c->dualsResident = false;
    c->dualsCompact = false; c->compactValid = false;
    int rc;
    if ((rc = devAlloc(c, &c->dSupport, (size_t)std::max(nIdx, 1)))) return rc;
    if (nIdx > 0) DIPGPU_CUDA(c, cudaMemcpy(c->dSupport, idx, sizeof(int32_t) * (size_t)nIdx, cudaMemcpyHostToDevice));
    std::vector<uint32_t> bits(((size_t)(uLen + 31) / 32 + 8) & ~(size_t)7, 0u);  // whole 16-byte pieces (pool kernel)
    for (int k = 0; k < nIdx; ++k) bits[(size_t)idx[k] >> 5] |= 1u << (idx[k] & 31);
    if ((rc = devAlloc(c, &c->dSupportBits, bits.size()))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dSupportBits, bits.data(), sizeof(uint32_t) * bits.size(), cudaMemcpyHostToDevice));
    // rank table of the bitmap (pool.cu, MODE 2): support rows below each word
    devFree(&c->dSupportRank);
    if (nIdx > 0 && nIdx < 65536) {
        std::vector<uint16_t> pre(bits.size(), 0);
        unsigned acc = 0;
        for (size_t w = 0; w < bits.size(); ++w) {
            pre[w] = (uint16_t)acc;
            acc += (unsigned)__builtin_popcount(bits[w]);
        }
        if ((rc = devAlloc(c, &c->dSupportRank, pre.size()))) return rc;
        DIPGPU_CUDA(c, cudaMemcpy(c->dSupportRank, pre.data(), sizeof(uint16_t) * pre.size(), cudaMemcpyHostToDevice));
    }
    return DIPGPU_OK;
}

int dipgpu_stab_set_center(dipgpu_ctx *ctx, const double *dual, int32_t uLen)
{
    if (isMulti(ctx)) {
        struct A { const double *d; int32_t n; } a{dual, uLen};
        return multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_stab_set_center(sub, q->d, q->n);
        }, &a, false, false);
    }
    CTX_OR_FAIL(ctx);
    if (!dual || uLen <= 0) return fail(c, DIPGPU_ERR_INVALID, "stab_set_center: bad arguments");
    int rc;
    if ((size_t)uLen != c->centerLen) {
        if ((rc = devAlloc(c, &c->dCenter, (size_t)uLen + 16))) return rc;
        c->centerLen = (size_t)uLen;
    }
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dCenter, dual, sizeof(double) * (size_t)uLen, cudaMemcpyHostToDevice, c->stream));
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += (int64_t)sizeof(double) * uLen;
    c->haveCenter = true;
    return DIPGPU_OK;
}

int dipgpu_stab_accept(dipgpu_ctx *ctx, int32_t step)
{
    if (isMulti(ctx)) return multiStabAccept(ctx, step);
    CTX_OR_FAIL(ctx);
    if (!c->haveST || step < 0 || step >= c->batchCap) return fail(c, DIPGPU_ERR_STATE, "stab_accept: no smoothed dual vector %d", step);
    if (!c->haveCenter) return fail(c, DIPGPU_ERR_STATE, "stab_accept: no stability centre");
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dCenter, c->dUBatch + (size_t)step * c->uStride, sizeof(double) * c->centerLen,
                                   cudaMemcpyDeviceToDevice, c->stream));
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return DIPGPU_OK;
}

int dipgpu_get_smoothed_duals(dipgpu_ctx *ctx, int32_t step, double *dualST, int32_t uLen)
{
    if (isMulti(ctx)) return multiGetSmoothedDuals(ctx, step, dualST, uLen);
    CTX_OR_FAIL(ctx);
    if (!dualST) return fail(c, DIPGPU_ERR_INVALID, "get_smoothed_duals: NULL output");
    if (!c->haveST || step < 0 || step >= c->batchCap || uLen != c->R + c->numConvex)
        return fail(c, DIPGPU_ERR_STATE, "get_smoothed_duals: no smoothed dual vector %d of length %d", step, uLen);
    DIPGPU_CUDA(c, cudaMemcpyAsync(dualST, c->dUBatch + (size_t)step * c->uStride, sizeof(double) * (size_t)uLen,
                                   cudaMemcpyDeviceToHost, c->stream));
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    c->d2h += (int64_t)sizeof(double) * uLen;
    return DIPGPU_OK;
}

int dipgpu_price_round_stab_ladder(dipgpu_ctx *ctx, const double *uRM, int32_t uLen, double alpha, double shrink,
                                   int32_t nSteps, int32_t phase, double redCostEps, dipgpu_cols *outs, double *alphasOut)
{
    Ctx *c0 = reinterpret_cast<Ctx *>(ctx);
    if (!c0) return DIPGPU_ERR_INVALID;
    if (!uRM || !outs || nSteps <= 0 || nSteps > 64) return fail(c0, DIPGPU_ERR_INVALID, "price_round_stab_ladder: bad arguments");
    // the ladder DualStabAlpha *= 0.90 (DecompAlgo.cpp:5646), by repeated multiplication as the reference
    double al[64];
    al[0] = alpha;
    for (int k = 1; k < nSteps; ++k) al[k] = al[k - 1] * shrink;
    if (alphasOut) memcpy(alphasOut, al, sizeof(double) * (size_t)nSteps);
    if (isMulti(ctx)) return multiStabLadder(ctx, uRM, uLen, al, nSteps, phase, redCostEps, outs);
    CTX_OR_FAIL(ctx);
    DualFeed f;
    f.host = uRM;
    return ladderCore(c, f, uLen, al, nSteps, phase, redCostEps, outs);
}

int dipgpu_price_round_stab(dipgpu_ctx *ctx, const double *uRM, int32_t uLen, double alpha, int32_t phase,
                            double redCostEps, dipgpu_cols *out, double *dualSTOpt)
{
    int rc = dipgpu_price_round_stab_ladder(ctx, uRM, uLen, alpha, 1.0, 1, phase, redCostEps, out, nullptr);
    if (rc) return rc;
    if ((rc = dipgpu_select_batch_result(ctx, 0))) return rc;
    if (dualSTOpt) return dipgpu_get_smoothed_duals(ctx, 0, dualSTOpt, uLen);
    return DIPGPU_OK;
}

int dipgpu_select_batch_result(dipgpu_ctx *ctx, int32_t which)
{
    if (isMulti(ctx)) return multiSelectBatchResult(ctx, which);
    CTX_OR_FAIL(ctx);
    if (which < 0 || which >= c->batchCap || !c->dResultBatch) return fail(c, DIPGPU_ERR_STATE, "select_batch_result: no batched result %d", which);
    const char *hp = static_cast<const char *>(c->hResultBatch) + (size_t)which * c->resStride;
    const ResultHeader *h = reinterpret_cast<const ResultHeader *>(hp);
    if (h->magic != kResultMagic) return fail(c, DIPGPU_ERR_STATE, "select_batch_result: result %d was never fetched", which);
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dResult, static_cast<char *>(c->dResultBatch) + (size_t)which * c->resStride, c->resultBytes,
                                   cudaMemcpyDeviceToDevice, c->stream));
    memcpy(c->hResult, hp, c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd);
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    ++c->roundSerial;  // a different round is "the last round" now
    c->lastRoundOnHost = true;
    return DIPGPU_OK;
}

// ------------------------------------------------------- waiting-column pool

int dipgpu_pool_clear(dipgpu_ctx *ctx)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    c->poolCols = c->poolNnz = 0;
    ++c->poolSerial;
    if (c->dPoolStart) DIPGPU_CUDA(c, cudaMemset(c->dPoolStart, 0, sizeof(int64_t)));
    return DIPGPU_OK;
}

int dipgpu_pool_size(const dipgpu_ctx *ctx, int64_t *nCols, int64_t *nNnz)
{
    if (isMulti(ctx)) ctx = multiPrimary(const_cast<dipgpu_ctx *>(ctx));
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c) return DIPGPU_ERR_INVALID;
    if (nCols) *nCols = c->poolCols;
    if (nNnz) *nNnz = c->poolNnz;
    return DIPGPU_OK;
}

int dipgpu_pool_append(dipgpu_ctx *ctx, int32_t nCols, const int64_t *colStart, const int32_t *rowInd, const double *val,
                       const double *origCost)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (nCols < 0 || (nCols > 0 && (!colStart || !origCost)) || (nCols > 0 && colStart[0] != 0))
        return fail(c, DIPGPU_ERR_INVALID, "pool_append: bad arguments");
    if (nCols == 0) return DIPGPU_OK;
    for (int k = 0; k < nCols; ++k)
        if (colStart[k + 1] < colStart[k]) return fail(c, DIPGPU_ERR_INVALID, "pool_append: colStart not monotone at %d", k);
    const int64_t add = colStart[nCols];
    if (add > 0 && (!rowInd || !val)) return fail(c, DIPGPU_ERR_INVALID, "pool_append: NULL rowInd/val");
    const int uLen = c->R + c->numConvex;
    if (c->haveCore)
        for (int64_t k = 0; k < add; ++k)
            if (rowInd[k] < 0 || rowInd[k] >= uLen) return fail(c, DIPGPU_ERR_INVALID, "pool_append: master row %d out of range", rowInd[k]);
    int rc;
    if ((rc = poolReserve(c, c->poolCols + nCols, c->poolNnz + add))) return rc;
    std::vector<int64_t> starts((size_t)nCols);
    for (int k = 0; k < nCols; ++k) starts[k] = c->poolNnz + colStart[k + 1];
    DIPGPU_CUDA(c, cudaMemcpy(c->dPoolStart + c->poolCols + 1, starts.data(), sizeof(int64_t) * (size_t)nCols, cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dPoolOrigCost + c->poolCols, origCost, sizeof(double) * (size_t)nCols, cudaMemcpyHostToDevice));
    if (add > 0) {
        DIPGPU_CUDA(c, cudaMemcpy(c->dPoolRow + c->poolNnz, rowInd, sizeof(int32_t) * (size_t)add, cudaMemcpyHostToDevice));
        DIPGPU_CUDA(c, cudaMemcpy(c->dPoolVal + c->poolNnz, val, sizeof(double) * (size_t)add, cudaMemcpyHostToDevice));
    }
    c->h2d += (int64_t)nCols * 16 + add * 12;
    c->poolCols += nCols;
    c->poolNnz += add;
    ++c->poolSerial;
    return DIPGPU_OK;
}

// descriptors of a device-side column copy (PoolCopyDesc in pool.cu: src, dst, len)
static int poolCopyColumns(Ctx *c, const std::vector<int64_t> &desc, const void *srcEntries, const int32_t *srcRow,
                           const double *srcVal, int32_t *dstRow, double *dstVal)
{
    const int n = (int)(desc.size() / 3);
    if (n == 0) return DIPGPU_OK;
    int64_t *dDesc = nullptr;
    int rc;
    if ((rc = devAlloc(c, &dDesc, desc.size()))) return rc;
    cudaError_t e = cudaMemcpyAsync(dDesc, desc.data(), sizeof(int64_t) * desc.size(), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        rc = launchPoolCopy(c, n, dDesc, srcEntries, srcRow, srcVal, dstRow, dstVal, c->stream);
        e = cudaStreamSynchronize(c->stream);
    }
    cudaFree(dDesc);
    if (e != cudaSuccess) return cudaFail(c, e, "pool column copy");
    return rc;
}

int dipgpu_pool_append_expanded(dipgpu_ctx *ctx, int32_t nWhich, const int32_t *which)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (nWhich < 0 || (nWhich > 0 && !which)) return fail(c, DIPGPU_ERR_INVALID, "pool_append_expanded: bad arguments");
    if (nWhich == 0) return DIPGPU_OK;
    if (!c->hMCols || !c->lastRoundOnHost) return fail(c, DIPGPU_ERR_STATE, "pool_append_expanded: no expanded columns");
    const char *hp = static_cast<const char *>(c->hMCols);
    const MColsHeader *mh = reinterpret_cast<const MColsHeader *>(hp);
    if (mh->magic != kMColsMagic) return fail(c, DIPGPU_ERR_STATE, "pool_append_expanded: no expanded columns");
    const int64_t *cs = reinterpret_cast<const int64_t *>(hp + c->mcolsOffColStart);
    // original cost of kept column k of the last round (packed result on the host)
    const char *rp = static_cast<const char *>(c->hResult);
    const ResultHeader *rh = reinterpret_cast<const ResultHeader *>(rp);
    const double *oc = reinterpret_cast<const double *>(rp + c->layout.offOrigCost);
    const int32_t *kb = reinterpret_cast<const int32_t *>(rp + c->layout.offKeptBlock);
    if (c->expandSerial != c->roundSerial || rh->nKept != mh->nCols)
        return fail(c, DIPGPU_ERR_STATE, "pool_append_expanded: expansion does not belong to the last round");
    int64_t add = 0;
    for (int k = 0; k < nWhich; ++k) {
        if (which[k] < 0 || which[k] >= mh->nCols) return fail(c, DIPGPU_ERR_INVALID, "pool_append_expanded: column %d out of range", which[k]);
        add += cs[which[k] + 1] - cs[which[k]];
    }
    int rc;
    if ((rc = poolReserve(c, c->poolCols + nWhich, c->poolNnz + add))) return rc;
    std::vector<int64_t> desc(3 * (size_t)nWhich), starts((size_t)nWhich);
    std::vector<double> cost((size_t)nWhich);
    int64_t dst = c->poolNnz;
    for (int k = 0; k < nWhich; ++k) {
        const int64_t len = cs[which[k] + 1] - cs[which[k]];
        desc[3 * k + 0] = cs[which[k]];
        desc[3 * k + 1] = dst;
        desc[3 * k + 2] = len;  // {int32 len, int32 pad}: little-endian, len < 2^31
        dst += len;
        starts[k] = dst;
        cost[k] = oc[kb[which[k]]];
    }
    DIPGPU_CUDA(c, cudaMemcpy(c->dPoolStart + c->poolCols + 1, starts.data(), sizeof(int64_t) * (size_t)nWhich, cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dPoolOrigCost + c->poolCols, cost.data(), sizeof(double) * (size_t)nWhich, cudaMemcpyHostToDevice));
    if ((rc = poolCopyColumns(c, desc, static_cast<char *>(c->dMCols) + c->mcolsOffEntries, nullptr, nullptr, c->dPoolRow, c->dPoolVal))) return rc;
    c->poolCols += nWhich;
    c->poolNnz += add;
    ++c->poolSerial;
    return DIPGPU_OK;
}

int dipgpu_pool_keep(dipgpu_ctx *ctx, int32_t nKeep, const int32_t *keepIdx)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (nKeep < 0 || (nKeep > 0 && !keepIdx)) return fail(c, DIPGPU_ERR_INVALID, "pool_keep: bad arguments");
    if (nKeep == 0) return dipgpu_pool_clear(ctx);
    std::vector<int64_t> hs((size_t)c->poolCols + 1);
    std::vector<double> hc((size_t)c->poolCols);
    DIPGPU_CUDA(c, cudaMemcpy(hs.data(), c->dPoolStart, sizeof(int64_t) * hs.size(), cudaMemcpyDeviceToHost));
    DIPGPU_CUDA(c, cudaMemcpy(hc.data(), c->dPoolOrigCost, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost));
    std::vector<int64_t> desc(3 * (size_t)nKeep), starts((size_t)nKeep + 1, 0);
    std::vector<double> cost((size_t)nKeep);
    int64_t dst = 0;
    for (int k = 0; k < nKeep; ++k) {
        const int j = keepIdx[k];
        if (j < 0 || j >= c->poolCols || (k > 0 && j <= keepIdx[k - 1]))
            return fail(c, DIPGPU_ERR_INVALID, "pool_keep: indices must be ascending and inside the pool");
        const int64_t len = hs[j + 1] - hs[j];
        desc[3 * k + 0] = hs[j];
        desc[3 * k + 1] = dst;
        desc[3 * k + 2] = len;
        dst += len;
        starts[k + 1] = dst;
        cost[k] = hc[j];
    }
    int rc;
    int32_t *nr = nullptr; double *nv = nullptr;
    if ((rc = devAlloc(c, &nr, (size_t)std::max<int64_t>(c->poolNnzCap, 1) + kPoolPad))) return rc;
    if ((rc = devAlloc(c, &nv, (size_t)std::max<int64_t>(c->poolNnzCap, 1) + kPoolPad))) { cudaFree(nr); return rc; }
    if ((rc = poolCopyColumns(c, desc, nullptr, c->dPoolRow, c->dPoolVal, nr, nv))) { cudaFree(nr); cudaFree(nv); return rc; }
    devFree(&c->dPoolRow); devFree(&c->dPoolVal);
    c->dPoolRow = nr; c->dPoolVal = nv;
    DIPGPU_CUDA(c, cudaMemcpy(c->dPoolStart, starts.data(), sizeof(int64_t) * starts.size(), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dPoolOrigCost, cost.data(), sizeof(double) * cost.size(), cudaMemcpyHostToDevice));
    c->poolCols = nKeep;
    c->poolNnz = dst;
    ++c->poolSerial;
    return DIPGPU_OK;
}

int dipgpu_pool_set_reduced_costs(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t feasible, double *redCost,
                                  int32_t *foundNegRC)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (!c->haveCore) return fail(c, DIPGPU_ERR_STATE, "pool_set_reduced_costs before set_core_csr");
    if (uLen != c->R + c->numConvex) return fail(c, DIPGPU_ERR_INVALID, "pool_set_reduced_costs: uLen=%d, expected R+numConvex=%d", uLen, c->R + c->numConvex);
    if (!u && !c->dualsResident && !c->dualsCompact) return fail(c, DIPGPU_ERR_STATE, "pool_set_reduced_costs: u == NULL but no dual vector is resident");
    int rc;
    if ((rc = ensureDualBuffer(c, uLen))) return rc;
    if ((rc = poolReserve(c, std::max<int64_t>(c->poolCols, 1), std::max<int64_t>(c->poolNnz, 1)))) return rc;
    cudaStream_t st = c->stream;
    if (u) {
        if ((rc = uploadDuals(c, u, uLen, 1, c->dU, 0, &c->dUClean, st))) return rc;
        c->dualsResident = true;
        c->dualsCompact = false;
    } else if ((rc = ensureDenseDuals(c, st))) {
        return rc;
    }
    if (c->stageTiming) cudaEventRecord(c->ev[1], st);
    if ((rc = launchPoolRedCost(c, c->dU, feasible ? 1 : 0, st))) return rc;
    if (c->stageTiming) cudaEventRecord(c->ev[2], st);
    int32_t flag = 0;
    if (redCost && c->poolCols > 0) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(redCost, c->dPoolRedCost, sizeof(double) * (size_t)c->poolCols, cudaMemcpyDeviceToHost, st));
        c->d2h += (int64_t)sizeof(double) * c->poolCols;
    }
    DIPGPU_CUDA(c, cudaMemcpyAsync(&flag, c->dPoolFlag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    if (c->stageTiming) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) != cudaSuccess) ms = 0.f;
        for (float &x : c->stageMs) x = 0.f;
        c->stageMs[1] = ms;  // slot [1]: the dual-vector pass (here the pool kernel incl. its flag memset)
    }
    c->d2h += 4;
    if (foundNegRC) *foundNegRC = flag;
    return DIPGPU_OK;
}

// Buffers and per-model constants of the expansion (allocated on first use, kept while the layout stays).
static int prepareExpand(Ctx *c)
{
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "expand_columns needs set_core_csr and set_knapsack_blocks");
    if (c->blockKind == 2) return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: columns of integer knapsack blocks carry multiplicities (expand them on the host)");
    if (!c->dCsrRowStart || !c->dColStart32) return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: more than 2^31 stored entries");
    int rc;
    // packed output buffer: worst case one entry per stored entry of A'' plus one per column (kept, with its slack,
    // while rows are appended: pinned host memory of this size is expensive to allocate)
    const int64_t needEntries = c->nnz + c->m + 16;
    if (!c->dMCols || c->mcolsM != c->m || needEntries > c->mcolsCapEntries) {
        const int64_t capEntries = needEntries + (c->baseR != c->R ? needEntries / 8 : 0);
        const size_t offColStart = sizeof(MColsHeader);
        const size_t offHash = offColStart + sizeof(int64_t) * ((size_t)c->m + 1);
        const size_t offEntries = ResultLayout::alignUp(offHash + sizeof(uint64_t) * (size_t)c->m, 16);
        const size_t bytes = offEntries + sizeof(MColEntry) * (size_t)capEntries;
        if (c->dMCols) cudaFree(c->dMCols);
        if (c->hMCols) cudaFreeHost(c->hMCols);
        c->dMCols = c->hMCols = nullptr;
        c->mcolsBytes = 0;
        DIPGPU_CUDA(c, cudaMalloc(&c->dMCols, bytes));
        DIPGPU_CUDA(c, cudaHostAlloc(&c->hMCols, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
        c->dMColsHost = nullptr;
        if (cudaHostGetDevicePointer(&c->dMColsHost, c->hMCols, 0) != cudaSuccess) {
            (void)cudaGetLastError();
            c->dMColsHost = nullptr;
        }
        memset(c->hMCols, 0, sizeof(MColsHeader));
        if ((rc = devAlloc(c, &c->dExpPub, (size_t)c->m))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dExpPub, 0, sizeof(unsigned long long) * (size_t)std::max(c->m, 1)));
        c->mcolsBytes = bytes;
        c->mcolsOffColStart = offColStart;
        c->mcolsOffHash = offHash;
        c->mcolsOffEntries = offEntries;
        c->mcolsCapEntries = capEntries;
        c->mcolsM = c->m;
        c->lastMColsBytes = 0;
    }
    // a column of block b touches at most as many rows as block b's x-columns have stored entries
    int64_t maxBlockNnz = 1;
    for (int b = c->firstBlock; b < c->firstBlock + c->nLocal; ++b) {
        const int s0 = c->hBlockColStart[b], s1 = c->hBlockColStart[b + 1];
        if (s1 <= c->N) maxBlockNnz = std::max<int64_t>(maxBlockNnz, c->hColStart[s1] - c->hColStart[s0]);
    }
    c->expMaxTouched = (int)std::min<int64_t>(maxBlockNnz + (c->R - c->baseR), c->R);  // (+ every appended row: expand.cu walks them all)
    return DIPGPU_OK;
}

// speculative size of the D2H copy of the packed master columns: header + starts + fingerprints + a first slab
// of entries sized by what the previous expansion needed plus a quarter (rounds of one node keep similar columns)
static size_t expandSpecBytes(const Ctx *c)
{
    return std::min(c->mcolsBytes, std::max(c->mcolsOffEntries + (size_t)65536,
                                            ResultLayout::alignUp(c->lastMColsBytes + c->lastMColsBytes / 4, 256)));
}

// The first `spec` bytes of the packed master columns are on the host (stream synchronised): checks, the rest of
// the entries if the speculative copy was short, decoding into the caller's arrays.
// The expansion mirrored its output into hMCols and rings c->hExpFlag: wait for the word (20 ms, then the stream).
static int waitExpandFlag(Ctx *c, cudaStream_t st)
{
    const volatile unsigned int *flag = c->hExpFlag;
    const unsigned int want = c->expSerialFlag;
    const auto t0 = std::chrono::steady_clock::now();
    bool seen = false;
    for (unsigned int spins = 1;; ++spins) {
        if (*flag == want) { seen = true; break; }
        if ((spins & 0xfffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) break;
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    c->expandFlagArmed = false;
    if (!seen) DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    return DIPGPU_OK;
}

static int finishExpand(Ctx *c, int nKept, size_t spec, dipgpu_mcols *out, cudaStream_t st)
{
    const char *hp = static_cast<const char *>(c->hMCols);
    const MColsHeader *mh = reinterpret_cast<const MColsHeader *>(hp);
    if (mh->magic != kMColsMagic || mh->nCols != nKept) return fail(c, DIPGPU_ERR_CUDA, "expand_columns: bad result header");
    if (mh->flags & 1u) return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: a column touches more rows than the shared-memory staging area holds");
    if (mh->flags & 2u) return fail(c, DIPGPU_ERR_NOMEM, "expand_columns: device output capacity exceeded");
    const size_t need = c->mcolsOffEntries + sizeof(MColEntry) * (size_t)mh->nNnz;
    c->lastMColsBytes = need;
    c->expandSerial = c->roundSerial;
    if (need > spec) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(static_cast<char *>(c->hMCols) + spec, static_cast<char *>(c->dMCols) + spec,
                                       need - spec, cudaMemcpyDeviceToHost, st));
        DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        c->d2h += (int64_t)(need - spec);
    }
    out->nCols = nKept;
    out->nNnz = mh->nNnz;
    if (nKept > out->capCols || mh->nNnz > out->capNnz)
        return fail(c, DIPGPU_ERR_OVERFLOW, "expand_columns needs capCols >= %d, capNnz >= %lld", nKept, (long long)mh->nNnz);
    const int64_t *cs = reinterpret_cast<const int64_t *>(hp + c->mcolsOffColStart);
    const uint64_t *hs = reinterpret_cast<const uint64_t *>(hp + c->mcolsOffHash);
    const MColEntry *en = reinterpret_cast<const MColEntry *>(hp + c->mcolsOffEntries);
    if (out->colStart) memcpy(out->colStart, cs, sizeof(int64_t) * ((size_t)nKept + 1));
    if (out->hash) memcpy(out->hash, hs, sizeof(uint64_t) * (size_t)nKept);
    for (int64_t q = 0; q < mh->nNnz; ++q) {
        if (out->rowInd) out->rowInd[q] = en[q].row;
        if (out->val) out->val[q] = en[q].val;
    }
    return DIPGPU_OK;
}

int dipgpu_expand_columns(dipgpu_ctx *ctx, double tolZero, dipgpu_mcols *out)
{
    if (isMulti(ctx)) return multiExpandColumns(ctx, tolZero, out);
    CTX_OR_FAIL(ctx);
    if (!out) return fail(c, DIPGPU_ERR_INVALID, "expand_columns: NULL output");
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "expand_columns needs set_core_csr and set_knapsack_blocks");
    if (c->blockKind == 2) return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: columns of integer knapsack blocks carry multiplicities (expand them on the host)");
    if (!c->lastRoundOnHost) return fail(c, DIPGPU_ERR_STATE, "expand_columns: no host-buffer pricing round since the model was set");
    const ResultHeader *rh = static_cast<const ResultHeader *>(c->hResult);
    const int nKept = rh->nKept;
    out->nCols = nKept;
    out->nNnz = 0;
    if (nKept == 0) {
        if (out->colStart && out->capCols >= 0) out->colStart[0] = 0;
        c->expandSerial = ~0ull;  // nothing to append
        return DIPGPU_OK;
    }
    int rc;
    if ((rc = prepareExpand(c))) return rc;
    cudaStream_t st = c->stream;
    c->wantExpandMirror = true;
    rc = launchExpandColumns(c, tolZero, nKept, st);
    c->wantExpandMirror = false;
    if (rc) return rc;
    if (c->expandFlagArmed) {
        // the kernel wrote the packed master columns into hMCols itself: no copy, no stream synchronisation
        if ((rc = waitExpandFlag(c, st))) return rc;
        const MColsHeader *mh = static_cast<const MColsHeader *>(c->hMCols);
        const size_t got = c->mcolsOffEntries + sizeof(MColEntry) * (size_t)std::max<int64_t>(mh->nNnz, 0);
        c->d2h += (int64_t)got;
        return finishExpand(c, nKept, c->mcolsBytes, out, st);
    }
    const size_t spec = expandSpecBytes(c);
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->hMCols, c->dMCols, spec, cudaMemcpyDeviceToHost, st));
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    c->d2h += (int64_t)spec;
    return finishExpand(c, nKept, spec, out, st);
}

// generateVars + the per-variable work of addVarsToPool in ONE submission: the expansion kernel is enqueued
// right behind the round (one CTA per block; the CTAs beyond the number of kept columns leave at once), the
// packed master columns follow in the same stream, and the host waits once.
int dipgpu_price_round_expand(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double redCostEps,
                              double tolZero, dipgpu_cols *out, dipgpu_mcols *mout)
{
    if (isMulti(ctx)) {
        // blocks partitioned over the devices: the round, then every shard expands the columns it kept
        if (!out || !mout) return fail(reinterpret_cast<Ctx *>(ctx), DIPGPU_ERR_INVALID, "price_round_expand: NULL output");
        const int rc = multiPriceRound(ctx, u, uLen, phase, redCostEps, out, nullptr);
        return rc ? rc : multiExpandColumns(ctx, tolZero, mout);
    }
    CTX_OR_FAIL(ctx);
    if (!out || !mout) return fail(c, DIPGPU_ERR_INVALID, "price_round_expand: NULL output");
    int rc;
    if ((rc = checkRoundReady(c, "price_round_expand", uLen, phase))) return rc;
    if (!u && !c->dualsResident && !c->dualsCompact) return fail(c, DIPGPU_ERR_STATE, "price_round_expand: u == NULL but no dual vector is resident");
    if ((rc = prepareExpand(c))) return rc;
    cudaStream_t st = c->stream;
    {
        DualFeed f;
        f.host = u;
        if ((rc = enqueueRound(c, u ? &f : nullptr, uLen, phase, redCostEps, true, st, false))) return rc;
    }
    size_t spec = 0;
    bool expMirrored = false;
    if (c->nLocal > 0) {
        c->wantExpandMirror = true;
        rc = launchExpandColumns(c, tolZero, c->nLocal, st);
        c->wantExpandMirror = false;
        if (rc) return rc;
        expMirrored = c->expandFlagArmed;
        if (!expMirrored) {
            spec = expandSpecBytes(c);
            DIPGPU_CUDA(c, cudaMemcpyAsync(c->hMCols, c->dMCols, spec, cudaMemcpyDeviceToHost, st));
            c->d2h += (int64_t)spec;
        }
    }
    if (expMirrored && !c->resultMirrored) {
        // (the round's result was not mirrored by its kernel: copy it, synchronise; the master columns are there by then)
        if ((rc = fetchResult(c, st))) return rc;
        c->expandFlagArmed = false;
        spec = c->mcolsBytes;
    } else if (expMirrored) {
        // the expansion rings its own word once its CTAs (which start behind the round's kernel) are done: the round's
        // result has landed by then too
        if ((rc = waitExpandFlag(c, st))) return rc;
        spec = c->mcolsBytes;
        c->flagArmed = false;
        const ResultHeader *h = static_cast<const ResultHeader *>(c->hResult);
        c->lastResultBytes = c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd;
        c->d2h += (int64_t)c->lastResultBytes;
        const MColsHeader *mh = static_cast<const MColsHeader *>(c->hMCols);
        c->d2h += (int64_t)(c->mcolsOffEntries + sizeof(MColEntry) * (size_t)std::max<int64_t>(mh->nNnz, 0));
    } else if ((rc = fetchResult(c, st))) return rc;  // (synchronises the stream: the master columns are on the host too)
    c->lastRoundOnHost = true;
    if ((rc = unpack(c, c->hResult, c->resultBytes, out))) return rc;
    mout->nCols = out->nCols;
    mout->nNnz = 0;
    if (c->nLocal == 0) {
        if (mout->colStart && mout->capCols >= 0) mout->colStart[0] = 0;
        return DIPGPU_OK;
    }
    return finishExpand(c, out->nCols, spec, mout, st);
}

int dipgpu_price_round_dev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase, double redCostEps, void *stream)
{
    if (isMulti(ctx)) {
        if (stream) return fail(reinterpret_cast<Ctx *>(ctx), DIPGPU_ERR_INVALID, "price_round_dev: a multi-device context runs on its own streams (pass NULL, then dipgpu_sync)");
        return multiPriceRoundDev(ctx, uDev, uLen, phase, redCostEps);
    }
    CTX_OR_FAIL(ctx);
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "price_round_dev needs set_core_csr and set_knapsack_blocks");
    if (!uDev || uLen != c->R + c->numConvex) return fail(c, DIPGPU_ERR_INVALID, "price_round_dev: bad dual vector");
    if (c->numConvex != c->m) return fail(c, DIPGPU_ERR_INVALID, "price_round_dev: numConvex=%d but %d blocks", c->numConvex, c->m);
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return enqueueRoundDev(c, uDev, uLen, phase, redCostEps, c->optMirrorAlways, st);
}

int dipgpu_calc_redcost_dev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase, void *stream)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (!c->haveCore) return fail(c, DIPGPU_ERR_STATE, "calc_redcost_dev before set_core_csr");
    if (phase != DIPGPU_PHASE_PRICE1 && !c->haveObj) return fail(c, DIPGPU_ERR_STATE, "calc_redcost_dev (phase 2) before set_objective");
    if (!uDev || uLen != c->R + c->numConvex) return fail(c, DIPGPU_ERR_INVALID, "calc_redcost_dev: bad dual vector");
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return launchRedCost(c, uDev, phase, st);
}

int dipgpu_measure_smem_gbs(dipgpu_ctx *ctx, double *gbs)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (!gbs) return fail(c, DIPGPU_ERR_INVALID, "measure_smem_gbs: NULL output");
    return measureSmemGbs(c, gbs);
}

int dipgpu_solve_blocks_dev(dipgpu_ctx *ctx, const double *redCostXDev, const double *alphaDev, double redCostEps, void *stream)
{
    if (isMulti(ctx)) ctx = multiForward(ctx);  // not partitioned: the primary device's whole-model context
    CTX_OR_FAIL(ctx);
    if (!c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "solve_blocks_dev before set_knapsack_blocks");
    if (!redCostXDev || !alphaDev) return fail(c, DIPGPU_ERR_INVALID, "solve_blocks_dev: NULL argument");
    if ((reinterpret_cast<uintptr_t>(redCostXDev) & 15u) != 0) return fail(c, DIPGPU_ERR_INVALID, "solve_blocks_dev: redCostXDev must be 16-byte aligned");
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    return enqueueSolve(c, redCostXDev, alphaDev, redCostEps, st, false);
}

int dipgpu_result_dev(dipgpu_ctx *ctx, void **devPtr, size_t *bytes)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !devPtr || !bytes) return DIPGPU_ERR_INVALID;
    if (c->multi) return multiResultDev(ctx, devPtr, bytes);
    if (!c->dResult) return fail(c, DIPGPU_ERR_STATE, "no result buffer yet");
    *devPtr = c->dResult;
    *bytes = c->resultBytes;
    return DIPGPU_OK;
}

int dipgpu_result_max_bytes(const dipgpu_ctx *ctx, size_t *bytes)
{
    if (isMulti(ctx)) ctx = multiPrimary(const_cast<dipgpu_ctx *>(ctx));
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !bytes) return DIPGPU_ERR_INVALID;
    *bytes = c->resultBytes;
    return DIPGPU_OK;
}

int dipgpu_unpack_result(const void *packedHost, size_t bytes, dipgpu_cols *out)
{
    return unpack(nullptr, packedHost, bytes, out);
}

int dipgpu_redcost_dev(dipgpu_ctx *ctx, double **devPtr)
{
    if (isMulti(ctx)) ctx = multiPrimary(ctx);
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !devPtr) return DIPGPU_ERR_INVALID;
    if (!c->dRedCost) return fail(c, DIPGPU_ERR_STATE, "no reduced-cost buffer yet");
    *devPtr = c->dRedCost;
    return DIPGPU_OK;
}

int dipgpu_get_counters(const dipgpu_ctx *ctx, int64_t *kernelLaunches, int64_t *h2dBytes, int64_t *d2hBytes)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c) return DIPGPU_ERR_INVALID;
    if (c->multi) return multiGetCounters(ctx, kernelLaunches, h2dBytes, d2hBytes);
    if (kernelLaunches) *kernelLaunches = c->launches;
    if (h2dBytes) *h2dBytes = c->h2d;
    if (d2hBytes) *d2hBytes = c->d2h;
    return DIPGPU_OK;
}

int dipgpu_last_stage_ms(const dipgpu_ctx *ctx, float *ms6)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !ms6) return DIPGPU_ERR_INVALID;
    for (int i = 0; i < 6; ++i) ms6[i] = c->stageMs[i];  // (the front of a multi-device context: per stage, the slowest device)
    return DIPGPU_OK;
}

static void geomOut(const DpGeom &g, int32_t out[7])
{
    out[0] = g.T; out[1] = (g.big && !g.cluster) ? 0 : g.S; out[2] = g.items; out[3] = g.guard; out[4] = g.chunkCols; out[5] = (int32_t)g.smemBytes;
    out[6] = g.fuse;
}

int dipgpu_plan_dp_geometry(int32_t smCount, int32_t nBlocks, int32_t maxCapacity, int32_t maxBlockCols,
                            int32_t maxUsableWeight, int32_t out[7])
{
    if (!out || smCount <= 0 || nBlocks < 0 || maxCapacity < 0 || maxBlockCols < 0 || maxUsableWeight < 0)
        return DIPGPU_ERR_INVALID;
    Ctx tmp;
    tmp.smCount = smCount;
    tmp.nLocal = nBlocks;
    tmp.maxCapacity = maxCapacity;
    tmp.maxBlockCols = maxBlockCols;
    tmp.maxUsableWeight = maxUsableWeight;
    DpGeom g{};
    int rc = chooseDpGeom(&tmp, &g, nBlocks > 0 && nBlocks <= kFuseFilterMaxBlocks);
    if (rc) return rc;
    geomOut(g, out);
    return DIPGPU_OK;
}

int dipgpu_get_dp_geometry(const dipgpu_ctx *ctx, int32_t out[7])
{
    if (isMulti(ctx)) ctx = multiPrimary(const_cast<dipgpu_ctx *>(ctx));
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !out) return DIPGPU_ERR_INVALID;
    if (!c->haveBlocks) return DIPGPU_ERR_STATE;
    geomOut(c->geom, out);
    return DIPGPU_OK;
}

int dipgpu_set_option(dipgpu_ctx *ctx, const char *name, int64_t value)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c || !name) return DIPGPU_ERR_INVALID;
    if (c->multi) {
        if (!strcmp(name, "dp_debug_times")) return dipgpu_set_option(multiPrimary(ctx), name, value);
        if (!strcmp(name, "shard_rounds")) return multiSetShardRounds(ctx, value);
        if (!strcmp(name, "round_timing")) {  // device-resident rounds: ONE event pair per device (dipgpu_last_device_ms)
            c->stageTiming = value != 0;
            return DIPGPU_OK;
        }
        if (!strcmp(name, "stage_timing")) c->stageTiming = value != 0;
        if (!strcmp(name, "gather_pinned_duals")) c->optGatherPinned = value != 0;
        struct A { const char *n; int64_t v; } a{name, value};
        const int rc = multiForAll(ctx, [](dipgpu_ctx *sub, void *p) {
            const A *q = static_cast<const A *>(p);
            return dipgpu_set_option(sub, q->n, q->v);
        }, &a, true, false);
        return rc;
    }
    if (!strcmp(name, "stage_timing") || !strcmp(name, "round_timing")) { c->stageTiming = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "dp_pdl")) { c->optPdl = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "fuse_redcost")) { c->optFuseRc = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "dp_coop")) { c->optNoCoop = value == 0; return DIPGPU_OK; }
    if (!strcmp(name, "pull_duals")) { c->optPull = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "batch_pipeline")) { c->optBatchPipeline = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "shard_rounds")) return DIPGPU_OK;  // (multi-device contexts only)
    if (!strcmp(name, "result_zero_copy")) {  // 2 (diagnostics): also from the device-resident entry points
        c->optZeroCopy = value != 0;
        c->optMirrorAlways = value == 2;
        return DIPGPU_OK;
    }
    if (!strcmp(name, "gather_pinned_duals")) { c->optGatherPinned = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "dp_no_cluster") || !strcmp(name, "dp_cluster_threads") || !strcmp(name, "dp_cluster_max_cells")) {
        if (!strcmp(name, "dp_no_cluster")) c->optNoCluster = value != 0;
        else if (!strcmp(name, "dp_cluster_max_cells")) c->optClusterMaxCells = (int)value;
        else if (value == 512 || value == 1024) c->optClusterThreads = (int)value;
        else return fail(c, DIPGPU_ERR_INVALID, "dp_cluster_threads: 512 or 1024");
        c->clusterFnConfigured = nullptr;
        c->clusterUnavailable = false;
        if (c->haveBlocks && c->blockKind == 0) return setBlockRange(c, c->firstBlock, c->nLocal);
        return DIPGPU_OK;
    }
    if (!strcmp(name, "batch_pull")) { c->optBatchPull = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "batch_pull_min")) { c->optBatchPullMin = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "dp_cluster_debug")) { c->optClusterDebug = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "pool_chunk")) { c->optPoolChunk = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "pool_warps")) { c->optPoolWarps = std::max(1, std::min(24, (int)value)); return DIPGPU_OK; }
    if (!strcmp(name, "pool_compact")) { c->optPoolCompact = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "pool_kernel")) { c->optPoolKernel = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "flag_wait")) { c->optFlagWait = value != 0; return DIPGPU_OK; }
    if (!strcmp(name, "flag_stats")) {  // diagnostics: value = host pointer receiving {rounds completed through the done word, timeouts}
        long long v[2] = {c->flagWaits, c->flagTimeouts};
        memcpy(reinterpret_cast<void *>(static_cast<uintptr_t>(value)), v, sizeof(v));
        return DIPGPU_OK;
    }
    if (!strcmp(name, "host_timing")) {  // diagnostics: 1 = on, 0 = off, else a host pointer receiving 4 doubles (us)
        if (value == 0 || value == 1) { c->hostTiming = value != 0; return DIPGPU_OK; }
        memcpy(reinterpret_cast<void *>(static_cast<uintptr_t>(value)), c->hostUs, sizeof(c->hostUs));
        return DIPGPU_OK;
    }
    if (!strcmp(name, "append_delta_min_nnz")) { c->optDeltaMinNnz = value; return DIPGPU_OK; }
    if (!strcmp(name, "merge_appended_rows")) { int rc = bind(c); return rc ? rc : mergeDelta(c); }
    if (!strcmp(name, "spmv_lanes")) { c->spmvLanes = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "dp_debug_times")) {
        // diagnostics: value = host pointer (as integer) receiving 8 x nLocal uint64 globaltimer stamps of
        // the LAST DP launch, or 1 to switch stamping on, 0 off
        int rc = bind(c);
        if (rc) return rc;
        if (value == 0) { devFree(&c->dDbgTimes); return DIPGPU_OK; }
        if (value == 1) {
            if ((rc = devAlloc(c, &c->dDbgTimes, 8 * (size_t)std::max(c->m, 1)))) return rc;
            DIPGPU_CUDA(c, cudaMemset(c->dDbgTimes, 0, 64 * (size_t)std::max(c->m, 1)));
            return DIPGPU_OK;
        }
        if (!c->dDbgTimes) return fail(c, DIPGPU_ERR_STATE, "dp_debug_times is off");
        DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
        DIPGPU_CUDA(c, cudaMemcpy(reinterpret_cast<void *>(static_cast<uintptr_t>(value)), c->dDbgTimes,
                                  64 * (size_t)std::max(c->nLocal, 1), cudaMemcpyDeviceToHost));
        return DIPGPU_OK;
    }
    if (!strcmp(name, "spmv_warps") || !strcmp(name, "spmv_stages") || !strcmp(name, "spmv_u_smem") ||
        !strcmp(name, "spmv_chunk") || !strcmp(name, "spmv_max_cols") || !strcmp(name, "spmv_window")) {
        // re-tiles the matrix: the tiles are dealt to the warps of the launch shape on the host
        int *opt = !strcmp(name, "spmv_warps") ? &c->optSpmvWarps : !strcmp(name, "spmv_stages") ? &c->optSpmvStages
                 : !strcmp(name, "spmv_u_smem") ? &c->optSpmvUSmem : !strcmp(name, "spmv_chunk") ? &c->optSpmvChunk
                 : !strcmp(name, "spmv_max_cols") ? &c->optSpmvMaxCols : &c->optSpmvWindow;
        if (*opt == (int)value) return DIPGPU_OK;
        *opt = (int)value;
        c->tilesDirty = true;
        return DIPGPU_OK;
    }
    if (!strcmp(name, "dp_threads") || !strcmp(name, "dp_cells_per_thread") || !strcmp(name, "dp_items_per_step") ||
        !strcmp(name, "dp_no_guard") || !strcmp(name, "fuse_filter") || !strcmp(name, "dp_no_prune")) {
        if (!strcmp(name, "dp_threads")) c->optThreads = (int)value;
        else if (!strcmp(name, "dp_no_prune")) c->optNoPrune = value != 0;
        else if (!strcmp(name, "dp_cells_per_thread")) c->optCellsPerThread = (int)value;
        else if (!strcmp(name, "dp_items_per_step")) c->optItemsPerStep = (int)value;
        else if (!strcmp(name, "fuse_filter")) c->optFuseFilter = (int)value;
        else c->optNoGuard = (int)value;
        if (c->haveBlocks) {
            int rc = bind(c);
            if (rc) return rc;
            return setBlockRange(c, c->firstBlock, c->nLocal);
        }
        return DIPGPU_OK;
    }
    return fail(c, DIPGPU_ERR_INVALID, "unknown option '%s'", name);
}

}  // extern "C"

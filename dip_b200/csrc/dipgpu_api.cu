// dipgpu_api.cu -- the C ABI of libdipgpu (include/dipgpu.h): context lifetime,
// model upload, and the orchestration of one pricing round
// (DecompAlgo::generateVars, Dip/src/DecompAlgo.cpp:4622-5260) as
//   H2D(u) -> reduced costs -> knapsack DP -> backtrack+emit -> filter -> D2H.
// No CPU fallback: every compute entry point needs a live sm_100 device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

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

static int bind(Ctx *c)
{
    DIPGPU_CUDA(c, cudaSetDevice(c->device));
    return DIPGPU_OK;
}

// Transpose the host CSR of A'' into the device CSC, folding the master-dual row
// mapping (rows >= nBaseRows sit numConvex further down, DecompAlgo.cpp:4595-4597).
static int uploadCore(Ctx *c)
{
    const int R = c->R, N = c->N;
    const int64_t nnz = c->hRowStart[R];
    std::vector<int64_t> colStart((size_t)N + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) colStart[(size_t)c->hColInd[k] + 1]++;
    for (int j = 0; j < N; ++j) colStart[j + 1] += colStart[j];
    std::vector<int64_t> fill(colStart.begin(), colStart.end() - 1);
    std::vector<int32_t> rowMaster((size_t)std::max<int64_t>(nnz, 1));
    std::vector<double> val((size_t)std::max<int64_t>(nnz, 1));
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
    c->nnz = nnz;
    c->hColStart = colStart;
    int rc;
    // warp tiles of the stream kernel: consecutive columns, <= kSpmvChunk stored entries (counted
    // from the 4-aligned start), <= 32*colsPerLane columns with colsPerLane fitted to the density
    c->nSpmvTiles = 0;
    if (nnz < (int64_t)1 << 31 && N > 0) {
        std::vector<int32_t> tileCol;
        std::vector<int32_t> cs32((size_t)N + 1);
        for (int j = 0; j <= N; ++j) cs32[j] = (int32_t)colStart[j];
        const bool row16 = (int64_t)R + c->numConvex <= 65536;
        // columns per warp tile: up to 32*kSpmvColsPerLane, fewer on small matrices so that every SM
        // still gets ~16 warp tiles (a GAP-E sweep is latency-bound: more, shorter tiles in flight)
        const int sms = c->smCount > 0 ? c->smCount : 148;
        int maxCols = (int)std::min<int64_t>(32 * kSpmvColsPerLane, (int64_t)N / (16 * sms)) & ~31;
        maxCols = std::max(32, maxCols);
        bool ok = true;
        int j = 0;
        tileCol.push_back(0);
        while (j < N) {
            int e = j;
            while (e < N && e - j < maxCols && colStart[e + 1] - colStart[j] <= kSpmvChunk) ++e;
            if (e == j) { ok = false; break; }  // a single column longer than a tile
            tileCol.push_back(e);
            j = e;
        }
        if (ok) {
            const size_t nt = tileCol.size() - 1;
            std::vector<int32_t> rec(4 * std::max<size_t>(nt, 1));
            std::vector<int64_t> blobOff(std::max<size_t>(nt, 1), 0);
            // tile-blocked blob: per tile its values then its rows, each padded to 16 bytes, so that ONE
            // bulk copy brings a tile into a shared-memory stage
            size_t blobBytes = 0;
            for (size_t k = 0; k < nt; ++k) {
                const int32_t first = cs32[tileCol[k]];
                const int32_t len = cs32[tileCol[k + 1]] - first;
                rec[4 * k + 0] = tileCol[k];
                rec[4 * k + 1] = tileCol[k + 1];
                rec[4 * k + 2] = first;
                rec[4 * k + 3] = len;
                blobOff[k] = (int64_t)blobBytes;
                blobBytes += (size_t)(((len + 1) & ~1) * 8);
                blobBytes += row16 ? (size_t)(((len + 7) & ~7) * 2) : (size_t)(((len + 3) & ~3) * 4);
            }
            std::vector<unsigned char> blob(blobBytes + 64, 0);
            for (size_t k = 0; k < nt; ++k) {
                const int32_t first = rec[4 * k + 2], len = rec[4 * k + 3];
                unsigned char *p = blob.data() + blobOff[k];
                memcpy(p, val.data() + first, sizeof(double) * (size_t)len);
                p += (size_t)(((len + 1) & ~1) * 8);
                if (row16) {
                    uint16_t *r16 = reinterpret_cast<uint16_t *>(p);
                    for (int q = 0; q < len; ++q) r16[q] = (uint16_t)rowMaster[(size_t)first + q];
                } else {
                    memcpy(p, rowMaster.data() + first, sizeof(int32_t) * (size_t)len);
                }
            }
            if ((rc = devAlloc(c, &c->dTileRec, rec.size()))) return rc;
            if ((rc = devAlloc(c, &c->dSpmvBlob, blob.size()))) return rc;
            if ((rc = devAlloc(c, &c->dSpmvBlobOff, blobOff.size()))) return rc;
            DIPGPU_CUDA(c, cudaMemcpy(c->dSpmvBlob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(c->dSpmvBlobOff, blobOff.data(), sizeof(int64_t) * blobOff.size(), cudaMemcpyHostToDevice));
            if ((rc = devAlloc(c, &c->dColStart32, (size_t)N + 1))) return rc;
            DIPGPU_CUDA(c, cudaMemcpy(c->dColStart32, cs32.data(), sizeof(int32_t) * cs32.size(), cudaMemcpyHostToDevice));
            DIPGPU_CUDA(c, cudaMemcpy(c->dTileRec, rec.data(), sizeof(int32_t) * rec.size(), cudaMemcpyHostToDevice));
            devFree(&c->dRowMaster16);
            if (row16) {
                std::vector<uint16_t> r16((size_t)nnz + 16, 0);
                for (int64_t k = 0; k < nnz; ++k) r16[k] = (uint16_t)rowMaster[k];
                if ((rc = devAlloc(c, &c->dRowMaster16, r16.size()))) return rc;
                DIPGPU_CUDA(c, cudaMemcpy(c->dRowMaster16, r16.data(), sizeof(uint16_t) * r16.size(), cudaMemcpyHostToDevice));
            }
            c->nSpmvTiles = (int)tileCol.size() - 1;
        }
    }
    // row-major copy for expand.cu
    devFree(&c->dCsrRowStart); devFree(&c->dCsrColInd); devFree(&c->dCsrVal);
    if (nnz < (int64_t)1 << 31) {
        std::vector<int32_t> rs32((size_t)R + 1);
        for (int i = 0; i <= R; ++i) rs32[i] = (int32_t)c->hRowStart[i];
        if ((rc = devAlloc(c, &c->dCsrRowStart, (size_t)R + 1))) return rc;
        if ((rc = devAlloc(c, &c->dCsrColInd, (size_t)nnz + 1))) return rc;
        if ((rc = devAlloc(c, &c->dCsrVal, (size_t)nnz + 1))) return rc;
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

static int ensureObjective(Ctx *c, int n)
{
    if (c->dObj) return DIPGPU_OK;
    int rc;
    if ((rc = devAlloc(c, &c->dObj, (size_t)n + 16))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dObj, 0, sizeof(double) * ((size_t)n + 16)));
    return DIPGPU_OK;
}

static int setBlockRange(Ctx *c, int first, int count)
{
    if (!c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "set_knapsack_blocks has not been called");
    if (first < 0 || count < 0 || first + count > c->m)
        return fail(c, DIPGPU_ERR_INVALID, "block range [%d,%d) outside [0,%d)", first, first + count, c->m);
    c->firstBlock = first;
    c->nLocal = count;
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
    const bool wantFuse = count > 0 && (c->optFuseFilter < 0 ? count <= kFuseFilterMaxBlocks : c->optFuseFilter == 2);
    if ((rc = chooseDpGeom(c, &c->geom, wantFuse))) return rc;
    if (count > 0 && (rc = prepareKnapsackDp(c))) return rc;
    // decision bits: ceil(n_b/32) words per cell, rowCells cells per block
    std::vector<int64_t> off((size_t)count + 1, 0);
    for (int lb = 0; lb < count; ++lb) {
        const int n = c->hBlockColStart[first + lb + 1] - c->hBlockColStart[first + lb];
        off[lb + 1] = off[lb] + (int64_t)((n + 31) / 32) * c->geom.rowCells;
    }
    c->bitsWords = (size_t)off[count];
    if ((rc = devAlloc(c, &c->dBitsOff, (size_t)count + 1))) return rc;
    DIPGPU_CUDA(c, cudaMemcpy(c->dBitsOff, off.data(), sizeof(int64_t) * ((size_t)count + 1), cudaMemcpyHostToDevice));
    if ((rc = devAlloc(c, &c->dBits, c->bitsWords))) return rc;
    if ((rc = devAlloc(c, &c->dNItems, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dPubWord, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dPubCells, (size_t)count))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dPubWord, 0, sizeof(unsigned long long) * std::max<size_t>((size_t)count, 1)));
    c->epoch = 0;
    if ((rc = devAlloc(c, &c->dScale, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dShortcut, (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dSel, 2 * (size_t)count))) return rc;
    if ((rc = devAlloc(c, &c->dZShort, (size_t)count))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dShortcut, 0, sizeof(int32_t) * std::max<size_t>((size_t)count, 1)));
    // packed result
    c->indCap = count > 0 ? (int64_t)(c->hBlockColStart[first + count] - c->hBlockColStart[first]) : 0;
    c->layout = ResultLayout::make(count, c->indCap);
    c->resultBytes = c->layout.bytes;
    if (c->dResult) cudaFree(c->dResult);
    c->dResult = nullptr;
    DIPGPU_CUDA(c, cudaMalloc(&c->dResult, c->resultBytes));
    DIPGPU_CUDA(c, cudaMemset(c->dResult, 0, c->resultBytes));
    if (c->hResult) cudaFreeHost(c->hResult);
    c->hResult = nullptr;
    DIPGPU_CUDA(c, cudaMallocHost(&c->hResult, c->resultBytes));
    return DIPGPU_OK;
}

// Enqueue stages (b) and (c) on `st`; dRedCost/dAlpha are device pointers
// (alpha indexed by global block id).
static int enqueueSolve(Ctx *c, const double *dRedCost, const double *dAlpha, double eps, cudaStream_t st,
                        bool timed)
{
    int rc;
    int mode = c->geom.fuse ? 2 : c->optFuseFilter < 0 ? (c->nLocal <= kFuseFilterMaxBlocks ? 1 : 0)
                                                       : (c->optFuseFilter == 1 ? 1 : 0);
    if (c->nLocal == 0) mode = 0;
    if ((rc = launchKnapsackPrep(c, dRedCost, st))) return rc;
    if ((rc = launchKnapsackDp(c, dRedCost, dAlpha, eps, st))) return rc;
    if (timed) cudaEventRecord(c->ev[3], st);
    if (mode != 2 && (rc = launchBacktrackEmit(c, dRedCost, dAlpha, eps, mode == 1, st))) return rc;
    if (timed) cudaEventRecord(c->ev[4], st);
    if (mode == 0 && (rc = launchFilter(c, eps, st))) return rc;
    if (timed) cudaEventRecord(c->ev[5], st);
    return DIPGPU_OK;
}

static int unpack(Ctx *c, const void *packed, size_t bytes, dipgpu_cols *out)
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
    out->nCols = h->nKept;
    out->nInd = h->nInd;
    out->cells = h->cells;
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
        if (out->blockNnz) out->blockNnz[b] = nz[lb];
        sum += mn[lb];  // block order, DecompAlgo.cpp:5229-5232
    }
    out->mostNegReducedCost = sum;
    if (h->nKept > out->capCols || h->nInd > out->capInd)
        return fail(c, DIPGPU_ERR_OVERFLOW, "kept columns need capCols >= %d, capInd >= %lld", h->nKept,
                    (long long)h->nInd);
    for (int k = 0; k < h->nKept; ++k) {
        const int lb = kb[k];
        if (out->colBlock) out->colBlock[k] = h->firstBlock + lb;
        if (out->colRedCost) out->colRedCost[k] = rc[lb];
        if (out->colOrigCost) out->colOrigCost[k] = oc[lb];
        if (out->colStart) out->colStart[k] = ks[k];
    }
    if (out->colStart) out->colStart[h->nKept] = h->nInd;
    if (out->colInd && h->nInd > 0) memcpy(out->colInd, ki, sizeof(int32_t) * (size_t)h->nInd);
    return DIPGPU_OK;
}

// D2H of the packed result: one speculative copy (header + per-block arrays +
// the first 64 KiB of indices), a second one only if more indices were kept.
static int fetchResult(Ctx *c, cudaStream_t st)
{
    const size_t spec = std::min(c->resultBytes, c->layout.offInd + (size_t)65536);
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->hResult, c->dResult, spec, cudaMemcpyDeviceToHost, st));
    if (c->stageTiming) cudaEventRecord(c->ev[6], st);
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    c->d2h += (int64_t)spec;
    const ResultHeader *h = static_cast<const ResultHeader *>(c->hResult);
    const size_t need = c->layout.offInd + sizeof(int32_t) * (size_t)h->nInd;
    if (need > spec) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(static_cast<char *>(c->hResult) + spec, static_cast<char *>(c->dResult) + spec,
                                       need - spec, cudaMemcpyDeviceToHost, st));
        DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        c->d2h += (int64_t)(need - spec);
    }
    return DIPGPU_OK;
}

static void collectStageMs(Ctx *c)
{
    if (!c->stageTiming) return;
    for (int i = 0; i < 6; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]) != cudaSuccess) ms = 0.f;
        c->stageMs[i] = ms;
    }
    (void)cudaGetLastError();
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
    if (cudaMalloc(reinterpret_cast<void **>(&c->dDoneCounter), sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(c->dDoneCounter, 0, sizeof(unsigned int)) != cudaSuccess) {
        delete c;
        return DIPGPU_ERR_NOMEM;
    }
    *ctx = reinterpret_cast<dipgpu_ctx *>(c);
    return DIPGPU_OK;
}

int dipgpu_destroy(dipgpu_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return DIPGPU_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    devFree(&c->dDoneCounter); devFree(&c->dCsrRowStart); devFree(&c->dCsrColInd); devFree(&c->dCsrVal); devFree(&c->dExpPub);
    if (c->dMCols) cudaFree(c->dMCols);
    if (c->hMCols) cudaFreeHost(c->hMCols); devFree(&c->dPubWord); devFree(&c->dPubCells); devFree(&c->dDbgTimes); devFree(&c->dColStart); devFree(&c->dColStart32); devFree(&c->dTileRec); devFree(&c->dRowMaster16); devFree(&c->dSpmvBlob); devFree(&c->dSpmvBlobOff); devFree(&c->dRowMaster); devFree(&c->dVal); devFree(&c->dObj);
    devFree(&c->dBlockColStart); devFree(&c->dWeight); devFree(&c->dCapacity); devFree(&c->dBitsOff);
    devFree(&c->dBits); devFree(&c->dItemW); devFree(&c->dItemCol); devFree(&c->dNItems); devFree(&c->dScale); devFree(&c->dShortcut); devFree(&c->dSel); devFree(&c->dZShort);
    devFree(&c->dCandInd); devFree(&c->dScratch); devFree(&c->dU); devFree(&c->dRedCost); devFree(&c->dAlpha);
    if (c->dResult) cudaFree(c->dResult);
    if (c->hResult) cudaFreeHost(c->hResult);
    if (c->hStage) cudaFreeHost(c->hStage);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return DIPGPU_OK;
}

int dipgpu_last_error(const dipgpu_ctx *ctx, char *buf, size_t bufLen)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!buf || bufLen == 0) return DIPGPU_ERR_INVALID;
    const char *s = c ? c->err.c_str() : "null context";
    snprintf(buf, bufLen, "%s", s);
    return DIPGPU_OK;
}

int dipgpu_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return DIPGPU_ERR_INVALID;
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 1);
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
    if (c->haveObj && c->N != N) { devFree(&c->dObj); c->haveObj = false; }
    if (c->N != N) devFree(&c->dRedCost);
    c->R = R; c->N = N; c->nBaseRows = nBaseRows; c->numConvex = numConvex;
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

int dipgpu_append_core_rows(dipgpu_ctx *ctx, int32_t nNew, const int64_t *rowStart, const int32_t *colInd, const double *val)
{
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
    return uploadCore(c);
}

int dipgpu_set_objective(dipgpu_ctx *ctx, const double *obj, int32_t N)
{
    CTX_OR_FAIL(ctx);
    if (!obj || N < 0) return fail(c, DIPGPU_ERR_INVALID, "set_objective: bad arguments");
    if (c->haveCore && N != c->N) return fail(c, DIPGPU_ERR_INVALID, "set_objective: N=%d but core has %d columns", N, c->N);
    if (!c->haveCore && c->haveBlocks && N < c->nBlockCols) return fail(c, DIPGPU_ERR_INVALID, "set_objective: N=%d < block columns %d", N, c->nBlockCols);
    int rc;
    devFree(&c->dObj);
    if ((rc = devAlloc(c, &c->dObj, (size_t)N + 16))) return rc;
    DIPGPU_CUDA(c, cudaMemset(c->dObj, 0, sizeof(double) * ((size_t)N + 16)));
    DIPGPU_CUDA(c, cudaMemcpy(c->dObj, obj, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice));
    if (!c->haveCore) c->N = std::max(c->N, (int)N);
    c->haveObj = true;
    return DIPGPU_OK;
}

int dipgpu_set_knapsack_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockColStart, const int32_t *weight,
                               const int32_t *capacity, int32_t profitMode)
{
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
    c->hBlockColStart.assign(blockColStart, blockColStart + m + 1);
    c->hCapacity.assign(capacity, capacity + m);
    c->hMaxUsableW.assign((size_t)m, 0);
    for (int b = 0; b < m; ++b) {
        int mw = 0;
        for (int j = blockColStart[b]; j < blockColStart[b + 1]; ++j)
            if (weight[j] <= capacity[b]) mw = std::max(mw, weight[j]);
        c->hMaxUsableW[b] = mw;
    }
    int rc;
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
    c->lastRoundOnHost = false;
    return setBlockRange(c, 0, m);
}

int dipgpu_set_block_range(dipgpu_ctx *ctx, int32_t first, int32_t count)
{
    CTX_OR_FAIL(ctx);
    return setBlockRange(c, first, count);
}

int dipgpu_calc_redcost(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double *redCostX)
{
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
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dU, u, sizeof(double) * (size_t)uLen, cudaMemcpyHostToDevice, c->stream));
    c->h2d += (int64_t)sizeof(double) * uLen;
    if ((rc = launchRedCost(c, c->dU, phase, c->stream))) return rc;
    DIPGPU_CUDA(c, cudaMemcpyAsync(redCostX, c->dRedCost, sizeof(double) * (size_t)c->N, cudaMemcpyDeviceToHost, c->stream));
    c->d2h += (int64_t)sizeof(double) * c->N;
    DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
    return DIPGPU_OK;
}

int dipgpu_solve_blocks(dipgpu_ctx *ctx, const double *redCostX, const double *alpha, double redCostEps, dipgpu_cols *out)
{
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
    if ((rc = enqueueSolve(c, c->dRedCost, c->dAlpha, redCostEps, st, timed))) return rc;
    if ((rc = fetchResult(c, st))) return rc;
    c->lastRoundOnHost = true;
    collectStageMs(c);
    return unpack(c, c->hResult, c->resultBytes, out);
}

int dipgpu_price_round(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double redCostEps,
                       dipgpu_cols *out, double *redCostXOpt)
{
    CTX_OR_FAIL(ctx);
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "price_round needs set_core_csr and set_knapsack_blocks");
    if (phase != DIPGPU_PHASE_PRICE1 && !c->haveObj) return fail(c, DIPGPU_ERR_STATE, "price_round (phase 2) before set_objective");
    if (!u || !out || uLen != c->R + c->numConvex)
        return fail(c, DIPGPU_ERR_INVALID, "price_round: uLen=%d, expected R+numConvex=%d", uLen, c->R + c->numConvex);
    if (c->numConvex != c->m) return fail(c, DIPGPU_ERR_INVALID, "price_round: numConvex=%d but %d blocks", c->numConvex, c->m);
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    if ((size_t)uLen > c->uCap) {
        if ((rc = devAlloc(c, &c->dU, (size_t)uLen + 16))) return rc;
        c->uCap = (size_t)uLen;
    }
    const bool timed = c->stageTiming;
    cudaStream_t st = c->stream;
    if (timed) cudaEventRecord(c->ev[0], st);
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->dU, u, sizeof(double) * (size_t)uLen, cudaMemcpyHostToDevice, st));
    c->h2d += (int64_t)sizeof(double) * uLen;
    if (timed) cudaEventRecord(c->ev[1], st);
    if ((rc = launchRedCost(c, c->dU, phase, st))) return rc;
    if (timed) cudaEventRecord(c->ev[2], st);
    // alpha_b = u[nBaseRows + b]  (DecompAlgo.cpp:4820)
    if ((rc = enqueueSolve(c, c->dRedCost, c->dU + c->nBaseRows, redCostEps, st, timed))) return rc;
    if (redCostXOpt) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(redCostXOpt, c->dRedCost, sizeof(double) * (size_t)c->N, cudaMemcpyDeviceToHost, st));
        c->d2h += (int64_t)sizeof(double) * c->N;
    }
    if ((rc = fetchResult(c, st))) return rc;
    c->lastRoundOnHost = true;
    collectStageMs(c);
    return unpack(c, c->hResult, c->resultBytes, out);
}

int dipgpu_expand_columns(dipgpu_ctx *ctx, double tolZero, dipgpu_mcols *out)
{
    CTX_OR_FAIL(ctx);
    if (!out) return fail(c, DIPGPU_ERR_INVALID, "expand_columns: NULL output");
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "expand_columns needs set_core_csr and set_knapsack_blocks");
    if (!c->lastRoundOnHost) return fail(c, DIPGPU_ERR_STATE, "expand_columns: no host-buffer pricing round since the model was set");
    if (!c->dCsrRowStart || !c->dColStart32) return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: more than 2^31 stored entries");
    const ResultHeader *rh = static_cast<const ResultHeader *>(c->hResult);
    const int nKept = rh->nKept;
    out->nCols = nKept;
    out->nNnz = 0;
    if (nKept == 0) {
        if (out->colStart && out->capCols >= 0) out->colStart[0] = 0;
        return DIPGPU_OK;
    }
    int rc;
    // packed output buffer: worst case one entry per stored entry of A'' plus one per column
    const int64_t capEntries = c->nnz + c->m + 16;
    const size_t offColStart = sizeof(MColsHeader);
    const size_t offHash = offColStart + sizeof(int64_t) * ((size_t)c->m + 1);
    const size_t offEntries = ResultLayout::alignUp(offHash + sizeof(uint64_t) * (size_t)c->m, 16);
    const size_t bytes = offEntries + sizeof(MColEntry) * (size_t)capEntries;
    if (bytes != c->mcolsBytes) {
        if (c->dMCols) cudaFree(c->dMCols);
        if (c->hMCols) cudaFreeHost(c->hMCols);
        c->dMCols = c->hMCols = nullptr;
        c->mcolsBytes = 0;
        DIPGPU_CUDA(c, cudaMalloc(&c->dMCols, bytes));
        DIPGPU_CUDA(c, cudaMallocHost(&c->hMCols, bytes));
        if ((rc = devAlloc(c, &c->dExpPub, (size_t)c->m))) return rc;
        DIPGPU_CUDA(c, cudaMemset(c->dExpPub, 0, sizeof(unsigned long long) * (size_t)std::max(c->m, 1)));
        c->expEpoch = 0;
        c->mcolsBytes = bytes;
        c->mcolsOffColStart = offColStart;
        c->mcolsOffHash = offHash;
        c->mcolsOffEntries = offEntries;
        c->mcolsCapEntries = capEntries;
    }
    // a column of block b touches at most as many rows as block b's x-columns have stored entries
    int64_t maxBlockNnz = 1;
    for (int b = c->firstBlock; b < c->firstBlock + c->nLocal; ++b) {
        const int s0 = c->hBlockColStart[b], s1 = c->hBlockColStart[b + 1];
        if (s1 <= c->N) maxBlockNnz = std::max<int64_t>(maxBlockNnz, c->hColStart[s1] - c->hColStart[s0]);
    }
    c->expMaxTouched = (int)std::min<int64_t>(maxBlockNnz, c->R);
    cudaStream_t st = c->stream;
    if ((rc = launchExpandColumns(c, tolZero, nKept, st))) return rc;
    // D2H: header + starts + fingerprints + a first slab of entries, the rest only if needed
    const size_t spec = std::min(bytes, offEntries + (size_t)131072);
    DIPGPU_CUDA(c, cudaMemcpyAsync(c->hMCols, c->dMCols, spec, cudaMemcpyDeviceToHost, st));
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    c->d2h += (int64_t)spec;
    const char *hp = static_cast<const char *>(c->hMCols);
    const MColsHeader *mh = reinterpret_cast<const MColsHeader *>(hp);
    if (mh->magic != kMColsMagic || mh->nCols != nKept) return fail(c, DIPGPU_ERR_CUDA, "expand_columns: bad result header");
    if (mh->flags & 1u) return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: a column touches more rows than the shared-memory staging area holds");
    if (mh->flags & 2u) return fail(c, DIPGPU_ERR_NOMEM, "expand_columns: device output capacity exceeded");
    const size_t need = offEntries + sizeof(MColEntry) * (size_t)mh->nNnz;
    if (need > spec) {
        DIPGPU_CUDA(c, cudaMemcpyAsync(static_cast<char *>(c->hMCols) + spec, static_cast<char *>(c->dMCols) + spec,
                                       need - spec, cudaMemcpyDeviceToHost, st));
        DIPGPU_CUDA(c, cudaStreamSynchronize(st));
        c->d2h += (int64_t)(need - spec);
    }
    out->nNnz = mh->nNnz;
    if (nKept > out->capCols || mh->nNnz > out->capNnz)
        return fail(c, DIPGPU_ERR_OVERFLOW, "expand_columns needs capCols >= %d, capNnz >= %lld", nKept, (long long)mh->nNnz);
    const int64_t *cs = reinterpret_cast<const int64_t *>(hp + offColStart);
    const uint64_t *hs = reinterpret_cast<const uint64_t *>(hp + offHash);
    const MColEntry *en = reinterpret_cast<const MColEntry *>(hp + offEntries);
    if (out->colStart) memcpy(out->colStart, cs, sizeof(int64_t) * ((size_t)nKept + 1));
    if (out->hash) memcpy(out->hash, hs, sizeof(uint64_t) * (size_t)nKept);
    for (int64_t q = 0; q < mh->nNnz; ++q) {
        if (out->rowInd) out->rowInd[q] = en[q].row;
        if (out->val) out->val[q] = en[q].val;
    }
    return DIPGPU_OK;
}

int dipgpu_price_round_dev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase, double redCostEps, void *stream)
{
    CTX_OR_FAIL(ctx);
    if (!c->haveCore || !c->haveBlocks) return fail(c, DIPGPU_ERR_STATE, "price_round_dev needs set_core_csr and set_knapsack_blocks");
    if (!uDev || uLen != c->R + c->numConvex) return fail(c, DIPGPU_ERR_INVALID, "price_round_dev: bad dual vector");
    if (c->numConvex != c->m) return fail(c, DIPGPU_ERR_INVALID, "price_round_dev: numConvex=%d but %d blocks", c->numConvex, c->m);
    int rc;
    if ((rc = ensureObjective(c, c->N))) return rc;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    if ((rc = launchRedCost(c, uDev, phase, st))) return rc;
    return enqueueSolve(c, c->dRedCost, uDev + c->nBaseRows, redCostEps, st, false);
}

int dipgpu_calc_redcost_dev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase, void *stream)
{
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
    CTX_OR_FAIL(ctx);
    if (!gbs) return fail(c, DIPGPU_ERR_INVALID, "measure_smem_gbs: NULL output");
    return measureSmemGbs(c, gbs);
}

int dipgpu_solve_blocks_dev(dipgpu_ctx *ctx, const double *redCostXDev, const double *alphaDev, double redCostEps, void *stream)
{
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
    if (!c->dResult) return fail(c, DIPGPU_ERR_STATE, "no result buffer yet");
    *devPtr = c->dResult;
    *bytes = c->resultBytes;
    return DIPGPU_OK;
}

int dipgpu_result_max_bytes(const dipgpu_ctx *ctx, size_t *bytes)
{
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
    if (kernelLaunches) *kernelLaunches = c->launches;
    if (h2dBytes) *h2dBytes = c->h2d;
    if (d2hBytes) *d2hBytes = c->d2h;
    return DIPGPU_OK;
}

int dipgpu_last_stage_ms(const dipgpu_ctx *ctx, float *ms6)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !ms6) return DIPGPU_ERR_INVALID;
    for (int i = 0; i < 6; ++i) ms6[i] = c->stageMs[i];
    return DIPGPU_OK;
}

static void geomOut(const DpGeom &g, int32_t out[7])
{
    out[0] = g.T; out[1] = g.S; out[2] = g.items; out[3] = g.guard; out[4] = g.chunkCols; out[5] = (int32_t)g.smemBytes;
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
    if (!strcmp(name, "stage_timing")) { c->stageTiming = value != 0; return DIPGPU_OK; }
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
    if (!strcmp(name, "spmv_warps")) { c->optSpmvWarps = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "spmv_u_smem")) { c->optSpmvUSmem = (int)value; return DIPGPU_OK; }
    if (!strcmp(name, "dp_threads") || !strcmp(name, "dp_cells_per_thread") || !strcmp(name, "dp_items_per_step") ||
        !strcmp(name, "dp_no_guard") || !strcmp(name, "fuse_filter")) {
        if (!strcmp(name, "dp_threads")) c->optThreads = (int)value;
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

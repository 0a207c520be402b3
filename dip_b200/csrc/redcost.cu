// redcost.cu -- stage (a) of the pricing round: redCostX = c - A''^T u.
//
// Replaces DecompAlgo::generateVarsAdjustDuals + generateVarsCalcRedCost
// (Dip/src/DecompAlgo.cpp:4550-4619, 4506-4547).  The reference scatters rows of
// the row-ordered A'' (CoinPackedMatrix::transposeTimes); here A'' is kept
// transposed on the device (CSC = CSR of A''^T) so each output column is a
// gather-and-reduce with no atomics, and the convexity-row gap of the master
// dual vector is folded into the stored row ids (rowMaster), so the "adjusted"
// dual vector of :4717-4722 is never materialised.
//
// HBM-bound: algorithmic bytes = 12*nnz + 20*N + 8*R.
//
// Main kernel (redcost_stream_kernel): the columns are cut on the host into
// self-contained tiles (<= 128 consecutive columns, <= "spmv_chunk" stored
// entries) whose entries are stored JAGGED per 32-column slice; a warp pulls a
// tile into shared memory with ONE TMA bulk copy (multi-stage ring per warp) and
// each lane walks its own column: conflict-free contiguous reads of the value
// and row streams, one gather of the dual, an unfused multiply and an add.
// Entries of a column come in DESCENDING row order, i.e. exactly the order and
// rounding of the reference's row-major scatter (rows R-1..0, y[col] += u[i]*a),
// so the reduced costs are bit-identical to the CPU restatement, not merely
// within 1e-12.  Matrices with a column longer than the largest tile fall back
// to the lane-cooperative kernel below (tolerance parity).
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace dipgpu {

__device__ __forceinline__ uint32_t rcSmemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void rcMbarWait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(rcSmemAddr(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ uint32_t ldsU32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ldsU16(uint32_t addr)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 ldsU64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double ldsF64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// byte sizes of a tile's parts inside the blob / a stage (16-byte multiples)
__host__ __device__ inline uint32_t tileValBytes(int len) { return (uint32_t)(((len + 1) & ~1) * 8); }
template <bool ROW16> __host__ __device__ inline uint32_t tileRowBytes(int len)
{
    return ROW16 ? (uint32_t)(((len + 7) & ~7) * 2) : (uint32_t)(((len + 3) & ~3) * 4);
}

struct StreamArgs {
    int nTiles;
    const uint2 *tileRec;       // per tile {byte offset / 16 into blob, bytes}
    const unsigned char *blob;  // self-contained tiles (layout: SpmvTileHeader in common.cuh)
    const double *u;
    int uLen;
    int phase1;
    double *out;
    int uSmemBytes;             // 0: gather the duals from global memory
    int stageBytes;             // bytes of one stage (>= the largest tile, multiple of 16)
    int nStages;                // stages per warp (2..kSpmvMaxStages)
    int64_t uStride, outStride; // batched dual vectors: vector v = blockIdx.y reads u + v*uStride
};

// Every WARP owns an nStages-deep TMA pipeline over its own tiles.  A tile is SELF-CONTAINED in
// the blob (header, column lengths / ids / objective entries, step offsets, values, rows: ONE bulk
// copy, no other global read on the warp's critical path).  Within a slice of 32 columns the host
// sorts the columns by DESCENDING length (lane l owns the l-th longest) and stores the entries
// JAGGED: first every column's entry 0 (lane order), then every column's entry 1 (only the columns
// that have one: a prefix of the lanes), ...  At step i lane l reads position stepOff[i] + l of the
// value / row streams: contiguous, conflict-free shared-memory reads, no ballot / popc, no second
// pass over stored products, and each lane adds its column's products SEQUENTIALLY in storage order
// (descending rows), unfused -- the reference's scatter order.  The only irregular access is the
// 8-byte gather of the dual (shared memory when the dual vector fits, else global/L2).
template <int UMODE, bool ROW16>
__global__ void __launch_bounds__(1024, 1) redcost_stream_kernel(const __grid_constant__ StreamArgs a)
{
    typedef typename std::conditional<ROW16, uint16_t, int32_t>::type RowT;
    extern __shared__ __align__(16) unsigned char sm[];
    const double *sU = reinterpret_cast<const double *>(sm);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const double *uVec = a.u + (int64_t)blockIdx.y * a.uStride;
    double *outVec = a.out + (int64_t)blockIdx.y * a.outStride;
    const int NST = a.nStages;
    const int barBytes = (NST * 8 + 15) & ~15;
    const size_t warpBytes = (size_t)barBytes + (size_t)NST * a.stageBytes;
    unsigned char *wbase = sm + a.uSmemBytes + (size_t)warp * warpBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
    unsigned char *stageBase = wbase + barBytes;
    const int warpsPerCta = blockDim.x >> 5;
    const int gw = blockIdx.x * warpsPerCta + warp;
    const int totalWarps = gridDim.x * warpsPerCta;

    auto issue = [&](const uint2 rec, int s) {
        uint64_t *bar = full + s;
        unsigned char *dst = stageBase + (size_t)s * a.stageBytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rcSmemAddr(bar)), "r"(rec.y)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                rcSmemAddr(dst)),
            "l"(a.blob + (size_t)rec.x * 16u), "r"(rec.y), "r"(rcSmemAddr(bar))
            : "memory");
    };

    if (lane == 0) {
        for (int s = 0; s < NST; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rcSmemAddr(full + s)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < NST; ++s) {
            const int tile = gw + s * totalWarps;
            if (tile < a.nTiles) issue(__ldg(a.tileRec + tile), s);
        }
    }
    if (UMODE != 0) {
        // the dual vector: one bulk copy (16-byte granules) + a plain tail, on its own mbarrier
        uint64_t *ubar = reinterpret_cast<uint64_t *>(sm + a.uSmemBytes - 16);
        double *w = const_cast<double *>(sU);
        const int cnt = a.uLen;
        const int bulk = (reinterpret_cast<uintptr_t>(uVec) & 15u) == 0 ? (cnt & ~1) : 0;
        if (t == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rcSmemAddr(ubar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (bulk > 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rcSmemAddr(ubar)),
                             "r"((uint32_t)bulk * 8u)
                             : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        rcSmemAddr(w)),
                    "l"(uVec), "r"((uint32_t)bulk * 8u), "r"(rcSmemAddr(ubar))
                    : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rcSmemAddr(ubar)) : "memory");
            }
        }
        for (int k = bulk + t; k < cnt; k += blockDim.x) w[k] = __ldg(uVec + k);
        __syncthreads();
        rcMbarWait(ubar, 0);
    } else {
        __syncwarp();
    }
    uint32_t sUaddr;  // computed ONCE (volatile: ptxas would otherwise rematerialise the conversion in the loop)
    asm volatile("{ .reg .u64 t64; cvta.to.shared.u64 t64, %1; cvt.u32.u64 %0, t64; }" : "=r"(sUaddr) : "l"(sm));

    int s = 0;
    uint32_t parity = 0;
    for (int tile = gw; tile < a.nTiles; tile += totalWarps) {
        // the record of the tile that will refill this stage: loaded now, used after the reduction
        const int tRefill = tile + NST * totalWarps;
        const uint2 recRefill = __ldg(a.tileRec + min(tRefill, a.nTiles - 1));
        const unsigned char *stage = stageBase + (size_t)s * a.stageBytes;
        rcMbarWait(full + s, parity);
        // shared-window addresses + ld.shared: the loads of a batch are issued back to back, in
        // the order written (rows, values, then the dependent gathers of the duals)
        const uint32_t stA = sUaddr + (uint32_t)(stage - sm);
        const int firstCol = (int)ldsU32(stA), nCols = (int)ldsU32(stA + 4), nEnt = (int)ldsU32(stA + 8);
        const uint32_t stepStride = ldsU32(stA + 12);
        const int nSlices = (nCols + 31) >> 5;
        const uint32_t lenA = stA + (uint32_t)sizeof(SpmvTileHeader);
        const uint32_t colA = lenA + (uint32_t)nSlices * 64u;
        const uint32_t cA = colA + (uint32_t)nSlices * 64u;
        const uint32_t offA = cA + (uint32_t)nSlices * 256u;
        const uint32_t valA = offA + (((uint32_t)nSlices * stepStride * 2u + 15u) & ~15u);
        const uint32_t rowA = valA + tileValBytes(nEnt);
        for (int q = 0; q < nSlices; ++q) {
            const int len = (int)ldsU16(lenA + (uint32_t)(q * 32 + lane) * 2u);
            const int colL = (int)ldsU16(colA + (uint32_t)(q * 32 + lane) * 2u);
            const double cj = ldsF64(cA + (uint32_t)(q * 32 + lane) * 8u);
            const int maxLen = (int)ldsU16(lenA + (uint32_t)q * 64u);  // lane 0 holds the slice's longest column
            const uint32_t offQ = offA + (uint32_t)q * stepStride * 2u;
            double sum = 0.0;
            for (int i0 = 0; i0 < maxLen; i0 += 4) {
                // first entry of steps i0..i0+3 (4 x uint16, warp-uniform); the lanes whose column is
                // longer than the step are a PREFIX of the warp (columns sorted by length on the host)
                const uint2 o = ldsU64(offQ + (uint32_t)i0 * 2u);
                uint32_t pos[4];
                pos[0] = (o.x & 0xffffu) + lane;
                pos[1] = (o.x >> 16) + lane;
                pos[2] = (o.y & 0xffffu) + lane;
                pos[3] = (o.y >> 16) + lane;
                uint32_t r[4];
                double v[4], g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    r[j] = 0u;
                    if (len > i0 + j) r[j] = ROW16 ? ldsU16(rowA + pos[j] * 2u) : ldsU32(rowA + pos[j] * 4u);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = 0.0;
                    if (len > i0 + j) v[j] = ldsF64(valA + pos[j] * 8u);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = 0.0;
                    if (len > i0 + j) g[j] = UMODE == 0 ? __ldg(uVec + r[j]) : ldsF64(sUaddr + r[j] * 8u);
                }
                // y += u*a, unfused.  A lane past its column's end adds 0*0 = +0, which leaves the sum
                // unchanged bit for bit (the sum starts at +0 and x + (-x) rounds to +0, so it is never -0).
#pragma unroll
                for (int j = 0; j < 4; ++j) sum = __dadd_rn(sum, __dmul_rn(g[j], v[j]));
            }
            if (q * 32 + lane < nCols) outVec[firstCol + colL] = a.phase1 ? -sum : cj - sum;  // DecompAlgo.cpp:4538-4545
        }
        // every lane is done reading the stage before the bulk copy (async proxy) refills it
        __syncwarp();
        if (lane == 0 && tRefill < a.nTiles) issue(recRefill, s);
        if (++s == NST) {
            s = 0;
            parity ^= 1u;
        }
    }
}

__device__ __forceinline__ double ldgStream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldgStream(const int32_t *p)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// L lanes cooperate on one column; columns are consecutive across the warp so a
// warp streams one contiguous span of (val, rowMaster).
template <int L>
__global__ void __launch_bounds__(256)
redcost_csc_kernel(int N, const int64_t *__restrict__ colStart,
                   const int32_t *__restrict__ rowMaster, const double *__restrict__ val,
                   const double *__restrict__ u, const double *__restrict__ c, int phase1,
                   double *__restrict__ out)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t col = gtid / L;
    const int sub = (int)(gtid % L);
    double sum = 0.0;
    if (col < N) {
        const int64_t beg = colStart[col], end = colStart[col + 1];
        for (int64_t k = beg + sub; k < end; k += L) {
            const double a = ldgStream(val + k);
            const int r = ldgStream(rowMaster + k);
            sum += __ldg(u + r) * a;
        }
    }
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (col < N && sub == 0) {
        // DecompAlgo.cpp:4538-4545
        out[col] = phase1 ? -sum : c[col] - sum;
    }
}


// Host side of the stream kernel: cut the columns into tiles and lay each tile out as the kernel
// reads it (SpmvTileHeader in common.cuh).  colStart/rowMaster/val: CSC of A'' with the entries
// of a column in the reference's scatter order (descending rows).
int buildSpmvTiles(Ctx *c, const std::vector<int64_t> &colStart, const std::vector<int32_t> &rowMaster,
                   const std::vector<double> &val)
{
    const int N = c->N;
    c->nSpmvTiles = 0;
    c->spmvStageBytes = 0;
    if (N <= 0 || colStart[N] >= (int64_t)1 << 31) return DIPGPU_OK;
    const bool row16 = (int64_t)c->R + c->numConvex <= 65536;
    const int sms = c->smCount > 0 ? c->smCount : 148;
    // columns per warp tile: up to 32*kSpmvColsPerLane, fewer on small matrices so that every SM
    // still gets ~16 warp tiles (a GAP-E sweep is latency-bound: more, shorter tiles in flight)
    int maxCols = (int)std::min<int64_t>(32 * kSpmvColsPerLane, (int64_t)N / (16 * sms)) & ~31;
    maxCols = std::max(32, maxCols);
    if (c->optSpmvMaxCols > 0) maxCols = std::max(32, std::min(c->optSpmvMaxCols & ~31, 32 * kSpmvColsPerLane));
    const int chunk = std::max(32, std::min(c->optSpmvChunk > 0 ? c->optSpmvChunk : kSpmvChunk, kSpmvChunkMax));
    std::vector<int32_t> tileCol(1, 0);
    for (int j = 0; j < N;) {
        int e = j;
        while (e < N && e - j < maxCols && colStart[e + 1] - colStart[j] <= chunk &&
               colStart[e + 1] - colStart[e] <= 65535)
            ++e;
        if (e == j) return DIPGPU_OK;  // a single column longer than a tile: lane-cooperative kernel
        tileCol.push_back(e);
        j = e;
    }
    const size_t nt = tileCol.size() - 1;
    std::vector<uint32_t> rec(2 * nt);
    std::vector<size_t> off(nt + 1, 0);
    std::vector<int32_t> stepStride(nt, 0);
    size_t maxBytes = 0;
    auto colLen = [&](int col) { return (int)(colStart[col + 1] - colStart[col]); };
    for (size_t k = 0; k < nt; ++k) {
        const int first = tileCol[k], nCols = tileCol[k + 1] - first;
        const int nEnt = (int)(colStart[tileCol[k + 1]] - colStart[first]);
        const int nSlices = (nCols + 31) / 32;
        int maxLen = 0;
        for (int col = first; col < first + nCols; ++col) maxLen = std::max(maxLen, colLen(col));
        stepStride[k] = (maxLen + 1 + 3) & ~3;  // step offsets of a slice: maxLen + 1 entries, read 4 at a time
        const size_t bytes = sizeof(SpmvTileHeader) + (size_t)nSlices * (64 + 64 + 256) +
                             (((size_t)nSlices * stepStride[k] * 2 + 15) & ~(size_t)15) + tileValBytes(nEnt) +
                             (row16 ? tileRowBytes<true>(nEnt) : tileRowBytes<false>(nEnt));
        off[k + 1] = off[k] + bytes;
        maxBytes = std::max(maxBytes, bytes);
        rec[2 * k + 1] = (uint32_t)bytes;
    }
    if (off[nt] / 16 >= ((size_t)1 << 32)) return DIPGPU_OK;
    std::vector<unsigned char> blob(off[nt] + 64, 0);
    const bool haveObj = (int)c->hObj.size() >= N;
    std::vector<int> order(32);
    for (size_t k = 0; k < nt; ++k) {
        rec[2 * k + 0] = (uint32_t)(off[k] / 16);
        const int first = tileCol[k], nCols = tileCol[k + 1] - first;
        const int nEnt = (int)(colStart[tileCol[k + 1]] - colStart[first]);
        const int nSlices = (nCols + 31) / 32;
        unsigned char *p = blob.data() + off[k];
        SpmvTileHeader *h = reinterpret_cast<SpmvTileHeader *>(p);
        h->firstCol = first;
        h->nCols = nCols;
        h->nEnt = nEnt;
        h->stepStride = stepStride[k];
        uint16_t *len = reinterpret_cast<uint16_t *>(p + sizeof(SpmvTileHeader));
        uint16_t *colL = len + (size_t)nSlices * 32;
        double *cc = reinterpret_cast<double *>(colL + (size_t)nSlices * 32);
        uint16_t *stepOff = reinterpret_cast<uint16_t *>(cc + (size_t)nSlices * 32);
        double *tv = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(stepOff) +
                                                (((size_t)nSlices * stepStride[k] * 2 + 15) & ~(size_t)15));
        unsigned char *tr = reinterpret_cast<unsigned char *>(tv) + tileValBytes(nEnt);
        int w = 0;
        for (int q = 0; q < nSlices; ++q) {
            const int c0 = first + 32 * q, n = std::min(first + nCols, c0 + 32) - c0;
            // lane l owns the l-th longest column of the slice (stable: equal lengths keep column order)
            for (int l = 0; l < n; ++l) order[l] = c0 + l;
            std::stable_sort(order.begin(), order.begin() + n, [&](int x, int y) { return colLen(x) > colLen(y); });
            for (int l = 0; l < n; ++l) {
                len[q * 32 + l] = (uint16_t)colLen(order[l]);
                colL[q * 32 + l] = (uint16_t)(order[l] - first);
                cc[q * 32 + l] = haveObj ? c->hObj[order[l]] : 0.0;
            }
            const int maxLen = n > 0 ? colLen(order[0]) : 0;
            uint16_t *so = stepOff + (size_t)q * stepStride[k];
            for (int i = 0; i < stepStride[k]; ++i) {
                so[i] = (uint16_t)w;
                if (i < maxLen)
                    for (int l = 0; l < n && colLen(order[l]) > i; ++l) {
                        const int64_t src = colStart[order[l]] + i;
                        tv[w] = val[src];
                        if (row16) reinterpret_cast<uint16_t *>(tr)[w] = (uint16_t)rowMaster[src];
                        else reinterpret_cast<int32_t *>(tr)[w] = rowMaster[src];
                        ++w;
                    }
            }
        }
    }
    int rc;
    if (c->dTileRec) { cudaFree(c->dTileRec); c->dTileRec = nullptr; }
    if (c->dSpmvBlob) { cudaFree(c->dSpmvBlob); c->dSpmvBlob = nullptr; }
    DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dTileRec), sizeof(uint32_t) * rec.size()));
    DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dSpmvBlob), blob.size()));
    DIPGPU_CUDA(c, cudaMemcpy(c->dSpmvBlob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dTileRec, rec.data(), sizeof(uint32_t) * rec.size(), cudaMemcpyHostToDevice));
    (void)rc;
    c->nSpmvTiles = (int)nt;
    c->spmvStageBytes = (int)((maxBytes + 15) & ~(size_t)15);
    return DIPGPU_OK;
}

int launchRedCost(Ctx *c, const double *dU, int phase, cudaStream_t st)
{
    return launchRedCostBatch(c, dU, 0, c->dRedCost, 0, 1, phase, st);
}

int launchRedCostBatch(Ctx *c, const double *dU, int64_t uStride, double *dOut, int64_t outStride, int nVec, int phase,
                       cudaStream_t st)
{
    if (c->N == 0 || nVec <= 0) return DIPGPU_OK;
    if (c->nSpmvTiles > 0 && c->spmvLanes == 0) {
        StreamArgs a;
        a.nTiles = c->nSpmvTiles;
        a.tileRec = reinterpret_cast<const uint2 *>(c->dTileRec);
        a.blob = c->dSpmvBlob;
        a.u = dU;
        a.uLen = c->R + c->numConvex;
        a.phase1 = phase == DIPGPU_PHASE_PRICE1 ? 1 : 0;
        a.out = dOut;
        a.uStride = uStride;
        a.outStride = outStride;
        a.stageBytes = c->spmvStageBytes;
        const int sms = c->smCount > 0 ? c->smCount : 148;
        const size_t maxSmem = 227 * 1024;
        const bool row16 = (int64_t)c->R + c->numConvex <= 65536;
        const size_t uBytes = (((size_t)a.uLen + 1) & ~(size_t)1) * 8 + 16;  // + the u mbarrier
        auto warpBytesFor = [&](int nst) { return (size_t)((nst * 8 + 15) & ~15) + (size_t)nst * a.stageBytes; };
        // u in shared memory pays when it is gathered irregularly and re-used: many more stored
        // entries than duals, and enough room left for the stages
        const bool reuse = c->nnz >= 8 * (int64_t)a.uLen;
        int mode = 0;
        if (reuse && uBytes + 8 * warpBytesFor(2) <= maxSmem && c->nSpmvTiles >= 24 * sms) mode = 1;
        if (c->optSpmvUSmem >= 0) mode = c->optSpmvUSmem != 0 ? 1 : 0;
        if (mode == 1 && uBytes + warpBytesFor(2) > maxSmem) mode = 0;
        typedef void (*Fn)(const StreamArgs);
        Fn fn;
        size_t bytes;
        int grid, threads;
        if (mode != 0) {
            // persistent, one CTA per SM: as many warps as fit beside the duals, each with a deep ring
            int nst = c->optSpmvStages > 0 ? std::min(c->optSpmvStages, kSpmvMaxStages) : 2;  // measured: 2 x 16+ warps beats 3 x 11
            nst = std::max(nst, 2);
            while (nst > 2 && uBytes + 8 * warpBytesFor(nst) > maxSmem) --nst;
            int warps = (int)std::min<size_t>(32, (maxSmem - uBytes) / warpBytesFor(nst));
            if (c->optSpmvWarps > 0) warps = std::min(warps, c->optSpmvWarps);
            warps = std::max(warps, 1);
            a.nStages = nst;
            a.uSmemBytes = (int)uBytes;
            bytes = uBytes + (size_t)warps * warpBytesFor(nst);
            fn = row16 ? redcost_stream_kernel<1, true> : redcost_stream_kernel<1, false>;
            grid = sms;
            threads = warps * 32;
        } else {
            int nst = c->optSpmvStages > 0 ? std::min(c->optSpmvStages, kSpmvMaxStages) : 2;
            nst = std::max(nst, 2);
            a.nStages = nst;
            a.uSmemBytes = 0;
            int warps = c->optSpmvWarps > 0 ? std::min(c->optSpmvWarps, 32) : 8;
            while (warps > 1 && (size_t)warps * warpBytesFor(nst) > maxSmem) --warps;
            bytes = (size_t)warps * warpBytesFor(nst);
            fn = row16 ? redcost_stream_kernel<0, true> : redcost_stream_kernel<0, false>;
            const int ctasPerSm = (int)std::max<size_t>(1, std::min<size_t>(4, maxSmem / std::max<size_t>(bytes, 1)));
            grid = std::max(1, std::min((c->nSpmvTiles + warps - 1) / warps, ctasPerSm * sms));
            threads = warps * 32;
        }
        if (bytes > maxSmem) return fail(c, DIPGPU_ERR_UNSUPPORTED, "reduced-cost tiles of %d bytes do not fit shared memory", a.stageBytes);
        if (c->spmvFn != (void *)fn || c->spmvFnSmem != bytes) {
            DIPGPU_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            c->spmvFn = (void *)fn;
            c->spmvFnSmem = bytes;
        }
        fn<<<dim3(grid, nVec, 1), threads, bytes, st>>>(a);
        c->launches++;
        DIPGPU_CUDA(c, cudaGetLastError());
        return DIPGPU_OK;
    }
    int L = c->spmvLanes;
    if (L == 0) {
        const double avg = c->N > 0 ? (double)c->nnz / (double)c->N : 0.0;
        L = avg <= 2.5 ? 1 : avg <= 6 ? 4 : avg <= 24 ? 8 : avg <= 96 ? 16 : 32;
    }
    const int phase1 = phase == DIPGPU_PHASE_PRICE1 ? 1 : 0;
    const int threads = 256;
    const int64_t total = (int64_t)c->N * L;
    const unsigned grid = (unsigned)((total + threads - 1) / threads);
#define LAUNCH(LL)                                                                              \
    redcost_csc_kernel<LL><<<grid, threads, 0, st>>>(c->N, c->dColStart, c->dRowMaster, c->dVal, \
                                                      dU + v * uStride, c->dObj, phase1, dOut + v * outStride)
    for (int64_t v = 0; v < nVec; ++v) {
        switch (L) {
            case 1: LAUNCH(1); break;
            case 2: LAUNCH(2); break;
            case 4: LAUNCH(4); break;
            case 8: LAUNCH(8); break;
            case 16: LAUNCH(16); break;
            default: LAUNCH(32); break;
        }
        c->launches++;
    }
#undef LAUNCH
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

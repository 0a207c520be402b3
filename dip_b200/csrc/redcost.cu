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
    const uint2 *tileRec;       // per tile {byte offset / 16 into blob, bytes}, grouped by owning warp
    const int32_t *warpStart;   // warp w of the launch owns tiles [warpStart[w], warpStart[w+1])
    const unsigned char *blob;  // self-contained tiles (layout: SpmvTileHeader in common.cuh)
    const double *u;
    int uLen;
    int phase1;
    double *out;
    int uSmemBytes;             // 0: gather the duals from global memory
    int stageBytes;             // bytes of one stage (>= the largest tile, multiple of 16)
    int nStages;                // stages per warp (2..kSpmvMaxStages)
    int64_t uStride, outStride; // batched dual vectors: vector v = blockIdx.y reads u + v*uStride
    // rows appended since the tiles were built (Ctx::baseR): init[col] = the column's sum over THOSE rows, the first
    // terms of the reference's scatter order (redcost_delta_kernel); NULL = every sum starts at +0
    const double *init;
    int64_t initStride;
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
    // programmatic dependent launch: the kernel that consumes the reduced costs may be scheduled now (it
    // waits, griddepcontrol.wait, before it reads them)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const double *sU = reinterpret_cast<const double *>(sm);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const double *uVec = a.u + (int64_t)blockIdx.y * a.uStride;
    double *outVec = a.out + (int64_t)blockIdx.y * a.outStride;
    const double *initVec = a.init != nullptr ? a.init + (int64_t)blockIdx.y * a.initStride : nullptr;
    const int NST = a.nStages;
    const int barBytes = (NST * 8 + 15) & ~15;
    const size_t warpBytes = (size_t)barBytes + (size_t)NST * a.stageBytes;
    unsigned char *wbase = sm + a.uSmemBytes + (size_t)warp * warpBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
    unsigned char *stageBase = wbase + barBytes;
    const int warpsPerCta = blockDim.x >> 5;
    const int gw = blockIdx.x * warpsPerCta + warp;
    // this warp's tiles: one contiguous run of the (host-ordered) tile list, so that the pieces of a
    // slice that was cut by step range follow each other on the same warp
    const int t0 = __ldg(a.warpStart + gw), t1 = __ldg(a.warpStart + gw + 1);

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
        for (int s = 0; s < NST; ++s)
            if (t0 + s < t1) issue(__ldg(a.tileRec + t0 + s), s);
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
    double sum = 0.0;  // lane's running column sum: carried across the pieces of a slice cut by step range
    for (int tile = t0; tile < t1; ++tile) {
        // the record of the tile that will refill this stage: loaded now, used after the reduction
        const int tRefill = tile + NST;
        const uint2 recRefill = __ldg(a.tileRec + min(tRefill, a.nTiles - 1));
        const unsigned char *stage = stageBase + (size_t)s * a.stageBytes;
        rcMbarWait(full + s, parity);
        // shared-window addresses + ld.shared: the loads of a batch are issued back to back, in
        // the order written (rows, values, then the dependent gathers of the duals)
        const uint32_t stA = sUaddr + (uint32_t)(stage - sm);
        const int firstCol = (int)ldsU32(stA), nSlices = (int)ldsU32(stA + 4), nEnt = (int)ldsU32(stA + 8);
        const uint32_t w3 = ldsU32(stA + 12);
        const uint32_t stepStride = w3 & 0xffffu;
        const bool pieceFirst = (w3 >> 16) & 1u, pieceLast = (w3 >> 17) & 1u;
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
            if (pieceFirst) sum = (initVec != nullptr && colL != 0xffff) ? __ldg(initVec + firstCol + colL) : 0.0;
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
            if (pieceLast && colL != 0xffff) outVec[firstCol + colL] = a.phase1 ? -sum : cj - sum;  // DecompAlgo.cpp:4538-4545
        }
        // every lane is done reading the stage before the bulk copy (async proxy) refills it
        __syncwarp();
        if (lane == 0 && tRefill < t1) issue(recRefill, s);
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

// The leading partial sums of the columns that have entries in appended rows (Ctx::baseR .. R): one thread per such
// column and dual vector adds its entries' products one after the other, from +0, in descending row order, unfused --
// what the reference's scatter has added to y[col] by the time it reaches row baseR - 1.
__global__ void __launch_bounds__(256)
redcost_delta_kernel(int nCols, const int32_t *__restrict__ cols, const int32_t *__restrict__ start,
                     const int32_t *__restrict__ rowMaster, const double *__restrict__ val, const double *__restrict__ u,
                     int64_t uStride, double *__restrict__ init, int64_t initStride)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nCols) return;
    const double *uVec = u + (int64_t)blockIdx.y * uStride;
    double sum = 0.0;
    for (int e = start[k]; e < start[k + 1]; ++e) sum = __dadd_rn(sum, __dmul_rn(__ldg(uVec + rowMaster[e]), val[e]));
    init[(int64_t)blockIdx.y * initStride + cols[k]] = sum;
}

// Launch shape of the stream kernel for this matrix (decided when the tiles are built, because the
// tiles are dealt to the launch's warps on the host).
static void planSpmvLaunch(Ctx *c, int nTiles)
{
    const int sms = c->smCount > 0 ? c->smCount : 148;
    const size_t maxSmem = 227 * 1024;
    const int uLen = c->baseR + c->numConvex;  // (rows appended later are not in the tiles)
    c->spmvULen = uLen;
    const size_t stageBytes = (size_t)c->spmvStageBytes;
    const size_t uBytes = (((size_t)uLen + 1) & ~(size_t)1) * 8 + 16;  // + the u mbarrier
    auto warpBytesFor = [&](int nst) { return (size_t)((nst * 8 + 15) & ~15) + (size_t)nst * stageBytes; };
    // u in shared memory pays when it is gathered irregularly and re-used: many more stored
    // entries than duals, and enough room left for the stages
    const bool reuse = c->nnz >= 8 * (int64_t)uLen;
    int mode = 0;
    if (reuse && uBytes + 8 * warpBytesFor(2) <= maxSmem && nTiles >= 24 * sms) mode = 1;
    if (c->optSpmvUSmem >= 0) mode = c->optSpmvUSmem != 0 ? 1 : 0;
    if (mode == 1 && uBytes + warpBytesFor(2) > maxSmem) mode = 0;
    int nst = c->optSpmvStages > 0 ? std::min(c->optSpmvStages, kSpmvMaxStages) : 2;  // measured: 2 x 16+ warps beats 3 x 11
    nst = std::max(nst, 2);
    int warps, grid;
    size_t bytes;
    if (mode != 0) {
        // persistent, one CTA per SM: as many warps as fit beside the duals
        while (nst > 2 && uBytes + 8 * warpBytesFor(nst) > maxSmem) --nst;
        warps = (int)std::min<size_t>(32, (maxSmem - uBytes) / warpBytesFor(nst));
        if (c->optSpmvWarps > 0) warps = std::min(warps, c->optSpmvWarps);
        warps = std::max(warps, 1);
        bytes = uBytes + (size_t)warps * warpBytesFor(nst);
        grid = sms;
    } else {
        warps = c->optSpmvWarps > 0 ? std::min(c->optSpmvWarps, 32) : 8;
        while (warps > 1 && (size_t)warps * warpBytesFor(nst) > maxSmem) --warps;
        bytes = (size_t)warps * warpBytesFor(nst);
        const int ctasPerSm = (int)std::max<size_t>(1, std::min<size_t>(4, maxSmem / std::max<size_t>(bytes, 1)));
        grid = std::max(1, std::min((nTiles + warps - 1) / warps, ctasPerSm * sms));
    }
    c->spmvMode = mode;
    c->spmvNst = nst;
    c->spmvWarps = warps;
    c->spmvGrid = grid;
    c->spmvSmem = bytes;
    c->spmvUBytes = mode != 0 ? (int)uBytes : 0;
}

// Host side of the stream kernel: cut the columns into tiles and lay each tile out as the kernel
// reads it (SpmvTileHeader in common.cuh).  colStart/rowMaster/val: CSC of A'' with the entries
// of a column in the reference's scatter order (descending rows).
//
// Columns are taken window by window ("spmv_window" consecutive columns); inside a window they are
// sorted by length and cut into slices of 32, so the lanes of a slice carry nearly equal work.  Short
// slices are packed several to a tile; a slice with more entries than the tile budget is cut by STEP
// RANGE into pieces that follow each other on one warp, the lanes' running sums carried from piece
// to piece -- any column length runs through the stream kernel, in the reference's order.
int buildSpmvTiles(Ctx *c, const std::vector<int64_t> &colStart, const std::vector<int32_t> &rowMaster,
                   const std::vector<double> &val)
{
    const int N = c->N;
    c->nSpmvTiles = 0;
    c->spmvStageBytes = 0;
    if (N <= 0 || colStart[N] >= (int64_t)1 << 31) return DIPGPU_OK;
    // a block shard of a multi-device context sweeps its own columns only (the SpMV is column-partitioned)
    const int colFirst = std::max(0, std::min(c->tileColFirst, N));
    const int colEnd = c->tileColEnd < 0 ? N : std::max(colFirst, std::min(c->tileColEnd, N));
    const int nSweep = colEnd - colFirst;
    const bool row16 = (int64_t)c->baseR + c->numConvex <= 65536;
    c->spmvRow16 = row16;
    const int sms = c->smCount > 0 ? c->smCount : 148;
    // slices per tile: up to kSpmvColsPerLane, fewer on small matrices so that every SM still gets
    // ~16 warp tiles (a GAP-E sweep is latency-bound: more, shorter tiles in flight)
    int maxCols = (int)std::min<int64_t>(32 * kSpmvColsPerLane, (int64_t)nSweep / (16 * sms)) & ~31;
    maxCols = std::max(32, maxCols);
    if (c->optSpmvMaxCols > 0) maxCols = std::max(32, std::min(c->optSpmvMaxCols & ~31, 32 * kSpmvColsPerLane));
    const int maxSlices = maxCols / 32;
    const int chunk = std::max(32, std::min(c->optSpmvChunk > 0 ? c->optSpmvChunk : kSpmvChunk, kSpmvChunkMax));
    int window = c->optSpmvWindow > 0 ? c->optSpmvWindow : kSpmvWindow;
    window = std::max(32, std::min(window, 32768)) & ~31;
    auto colLen = [&](int col) { return (int)(colStart[col + 1] - colStart[col]); };

    struct Slice { int lanes; int col[32]; int s0, s1; };       // columns (lane order) and step range
    struct Tile { int winFirst; int nSl; int firstSlice; int nEnt; int stepStride; bool first, last; size_t bytes; };
    std::vector<Slice> slices;
    std::vector<Tile> tiles;
    std::vector<std::pair<int, int>> units;  // {first tile, tiles}: pieces of one slice stay together
    std::vector<int> order((size_t)window);
    auto tileBytes = [&](int nSl, int stepStride, int nEnt) {
        return sizeof(SpmvTileHeader) + (size_t)nSl * (64 + 64 + 256) + (((size_t)nSl * stepStride * 2 + 15) & ~(size_t)15) +
               tileValBytes(nEnt) + (row16 ? tileRowBytes<true>(nEnt) : tileRowBytes<false>(nEnt));
    };
    auto entriesInSteps = [&](const Slice &sl, int s0, int s1) {
        int e = 0;
        for (int l = 0; l < sl.lanes; ++l) e += std::max(0, std::min(colLen(sl.col[l]), s1) - s0);
        return e;
    };
    for (int wf = colFirst; wf < colEnd; wf += window) {
        const int wn = std::min(window, colEnd - wf);
        for (int l = 0; l < wn; ++l) order[l] = wf + l;
        // lane l of a slice owns the slice's l-th longest column (stable: equal lengths keep column order)
        std::stable_sort(order.begin(), order.begin() + wn, [&](int x, int y) { return colLen(x) > colLen(y); });
        Tile pend{wf, 0, 0, 0, 0, true, true, 0};
        auto flush = [&]() {
            if (pend.nSl == 0) return;
            pend.bytes = tileBytes(pend.nSl, pend.stepStride, pend.nEnt);
            units.push_back({(int)tiles.size(), 1});
            tiles.push_back(pend);
            pend = Tile{wf, 0, 0, 0, 0, true, true, 0};
        };
        for (int q0 = 0; q0 < wn; q0 += 32) {
            Slice sl;
            sl.lanes = std::min(32, wn - q0);
            for (int l = 0; l < sl.lanes; ++l) sl.col[l] = order[q0 + l];
            const int maxLen = colLen(sl.col[0]);
            sl.s0 = 0;
            sl.s1 = maxLen;
            const int ent = entriesInSteps(sl, 0, maxLen);
            const int stride = (maxLen + 1 + 3) & ~3;  // step offsets: maxLen + 1 entries, read 4 at a time
            if (ent <= chunk) {
                if (pend.nSl > 0 && (pend.nEnt + ent > chunk || pend.nSl >= maxSlices)) flush();
                if (pend.nSl == 0) pend.firstSlice = (int)slices.size();
                slices.push_back(sl);
                pend.nSl++;
                pend.nEnt += ent;
                pend.stepStride = std::max(pend.stepStride, stride);
            } else {
                flush();
                // cut by step range: pieces of <= chunk entries (a step holds <= 32)
                const int firstTile = (int)tiles.size();
                for (int s0 = 0; s0 < maxLen;) {
                    int s1 = s0, e = 0;
                    while (s1 < maxLen) {
                        const int add = entriesInSteps(sl, s1, s1 + 1);
                        if (e + add > chunk && s1 > s0) break;
                        e += add;
                        ++s1;
                    }
                    Slice pc = sl;
                    pc.s0 = s0;
                    pc.s1 = s1;
                    Tile t{wf, 1, (int)slices.size(), e, (s1 - s0 + 1 + 3) & ~3, s0 == 0, s1 == maxLen, 0};
                    t.bytes = tileBytes(1, t.stepStride, e);
                    slices.push_back(pc);
                    tiles.push_back(t);
                    s0 = s1;
                }
                units.push_back({firstTile, (int)tiles.size() - firstTile});
            }
        }
        flush();
    }
    const size_t nt = tiles.size();
    size_t maxBytes = 0;
    for (const Tile &t : tiles) maxBytes = std::max(maxBytes, t.bytes);
    c->spmvStageBytes = (int)((maxBytes + 15) & ~(size_t)15);
    planSpmvLaunch(c, (int)nt);
    if (c->spmvSmem > 227 * 1024) { c->spmvStageBytes = 0; return DIPGPU_OK; }  // lane-cooperative kernel instead

    // deal the units to the launch's warps (unit k -> warp k mod W): each warp's tiles become one
    // contiguous run of the tile list
    const int W = c->spmvGrid * c->spmvWarps;
    std::vector<std::vector<int>> perWarp((size_t)W);
    for (size_t k = 0; k < units.size(); ++k)
        for (int q = 0; q < units[k].second; ++q) perWarp[k % (size_t)W].push_back(units[k].first + q);
    std::vector<int32_t> warpStart((size_t)W + 1, 0);
    std::vector<int> tileOrder;
    tileOrder.reserve(nt);
    for (int w = 0; w < W; ++w) {
        warpStart[(size_t)w] = (int32_t)tileOrder.size();
        tileOrder.insert(tileOrder.end(), perWarp[(size_t)w].begin(), perWarp[(size_t)w].end());
    }
    warpStart[(size_t)W] = (int32_t)tileOrder.size();

    std::vector<size_t> off(nt + 1, 0);
    for (size_t k = 0; k < nt; ++k) off[k + 1] = off[k] + tiles[(size_t)tileOrder[k]].bytes;
    if (off[nt] / 16 >= ((size_t)1 << 32)) { c->spmvStageBytes = 0; return DIPGPU_OK; }
    std::vector<uint32_t> rec(2 * std::max<size_t>(nt, 1), 0);
    std::vector<unsigned char> blob(off[nt] + 64, 0);
    const bool haveObj = (int)c->hObj.size() >= N;
    for (size_t k = 0; k < nt; ++k) {
        const Tile &t = tiles[(size_t)tileOrder[k]];
        rec[2 * k + 0] = (uint32_t)(off[k] / 16);
        rec[2 * k + 1] = (uint32_t)t.bytes;
        unsigned char *p = blob.data() + off[k];
        SpmvTileHeader *h = reinterpret_cast<SpmvTileHeader *>(p);
        h->firstCol = t.winFirst;
        h->nSlices = t.nSl;
        h->nEnt = t.nEnt;
        h->strideFlags = (int32_t)((uint32_t)t.stepStride | (t.first ? 1u << 16 : 0u) | (t.last ? 1u << 17 : 0u));
        uint16_t *len = reinterpret_cast<uint16_t *>(p + sizeof(SpmvTileHeader));
        uint16_t *colL = len + (size_t)t.nSl * 32;
        double *cc = reinterpret_cast<double *>(colL + (size_t)t.nSl * 32);
        uint16_t *stepOff = reinterpret_cast<uint16_t *>(cc + (size_t)t.nSl * 32);
        double *tv = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(stepOff) +
                                                (((size_t)t.nSl * t.stepStride * 2 + 15) & ~(size_t)15));
        unsigned char *tr = reinterpret_cast<unsigned char *>(tv) + tileValBytes(t.nEnt);
        int w = 0;
        for (int q = 0; q < t.nSl; ++q) {
            const Slice &sl = slices[(size_t)t.firstSlice + q];
            for (int l = 0; l < 32; ++l) {
                const bool on = l < sl.lanes;
                len[q * 32 + l] = (uint16_t)(on ? std::max(0, std::min(colLen(sl.col[l]), sl.s1) - sl.s0) : 0);
                colL[q * 32 + l] = (uint16_t)(on ? sl.col[l] - t.winFirst : 0xffff);
                cc[q * 32 + l] = (on && haveObj) ? c->hObj[sl.col[l]] : 0.0;
            }
            uint16_t *so = stepOff + (size_t)q * t.stepStride;
            for (int i = 0; i < t.stepStride; ++i) {
                so[i] = (uint16_t)w;
                const int step = sl.s0 + i;
                if (step < sl.s1)
                    for (int l = 0; l < sl.lanes && colLen(sl.col[l]) > step; ++l) {
                        const int64_t src = colStart[sl.col[l]] + step;
                        tv[w] = val[src];
                        if (row16) reinterpret_cast<uint16_t *>(tr)[w] = (uint16_t)rowMaster[src];
                        else reinterpret_cast<int32_t *>(tr)[w] = rowMaster[src];
                        ++w;
                    }
            }
        }
    }
    if (c->dTileRec) { cudaFree(c->dTileRec); c->dTileRec = nullptr; }
    if (c->dSpmvBlob) { cudaFree(c->dSpmvBlob); c->dSpmvBlob = nullptr; }
    if (c->dSpmvWarpStart) { cudaFree(c->dSpmvWarpStart); c->dSpmvWarpStart = nullptr; }
    DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dTileRec), sizeof(uint32_t) * rec.size()));
    DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dSpmvBlob), blob.size()));
    DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dSpmvWarpStart), sizeof(int32_t) * warpStart.size()));
    DIPGPU_CUDA(c, cudaMemcpy(c->dSpmvBlob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dTileRec, rec.data(), sizeof(uint32_t) * rec.size(), cudaMemcpyHostToDevice));
    DIPGPU_CUDA(c, cudaMemcpy(c->dSpmvWarpStart, warpStart.data(), sizeof(int32_t) * warpStart.size(), cudaMemcpyHostToDevice));
    c->nSpmvTiles = (int)nt;
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
    {
        const int rc = ensureSpmvTiles(c);
        if (rc) return rc;
    }
    const bool delta = c->baseR != c->R;
    if (delta && !(c->nSpmvTiles > 0 && c->spmvLanes == 0)) {
        // the lane-cooperative kernel reads the plain CSC: fold the appended rows in first
        int rc = mergeDelta(c);
        if (rc || (rc = ensureSpmvTiles(c))) return rc;
    }
    if (c->nSpmvTiles > 0 && c->spmvLanes == 0) {
        StreamArgs a;
        a.init = nullptr;
        a.initStride = 0;
        if (c->baseR != c->R) {
            if (c->spmvInitVecs < nVec) {
                if (c->dSpmvInit) { cudaFree(c->dSpmvInit); c->dSpmvInit = nullptr; }
                c->spmvInitVecs = 0;
                const size_t cnt = (size_t)nVec * (size_t)c->N + 16;
                DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dSpmvInit), sizeof(double) * cnt));
                DIPGPU_CUDA(c, cudaMemsetAsync(c->dSpmvInit, 0, sizeof(double) * cnt, st));
                c->spmvInitVecs = nVec;
            }
            if (c->nDeltaCols > 0) {
                redcost_delta_kernel<<<dim3((unsigned)((c->nDeltaCols + 255) / 256), (unsigned)nVec, 1), 256, 0, st>>>(
                    c->nDeltaCols, c->dDeltaCol, c->dDeltaStart, c->dDeltaRow, c->dDeltaVal, dU, uStride, c->dSpmvInit, (int64_t)c->N);
                c->launches++;
                DIPGPU_CUDA(c, cudaGetLastError());
            }
            a.init = c->dSpmvInit;
            a.initStride = (int64_t)c->N;
        }
        a.nTiles = c->nSpmvTiles;
        a.tileRec = reinterpret_cast<const uint2 *>(c->dTileRec);
        a.warpStart = c->dSpmvWarpStart;
        a.blob = c->dSpmvBlob;
        a.u = dU;
        a.uLen = c->spmvULen;
        a.phase1 = phase == DIPGPU_PHASE_PRICE1 ? 1 : 0;
        a.out = dOut;
        a.uStride = uStride;
        a.outStride = outStride;
        a.stageBytes = c->spmvStageBytes;
        a.nStages = c->spmvNst;
        a.uSmemBytes = c->spmvUBytes;
        const bool row16 = c->spmvRow16;
        typedef void (*Fn)(const StreamArgs);
        const Fn fn = c->spmvMode != 0 ? (row16 ? redcost_stream_kernel<1, true> : redcost_stream_kernel<1, false>)
                                       : (row16 ? redcost_stream_kernel<0, true> : redcost_stream_kernel<0, false>);
        const size_t bytes = c->spmvSmem;
        if (c->spmvFn != (void *)fn || c->spmvFnSmem != bytes) {
            const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(fn), bytes);
            if (rc) return rc;
            c->spmvFn = (void *)fn;
            c->spmvFnSmem = bytes;
        }
        fn<<<dim3(c->spmvGrid, nVec, 1), c->spmvWarps * 32, bytes, st>>>(a);
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

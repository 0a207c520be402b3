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
// tiles of at most kSpmvChunk stored entries; a warp pulls its tile's (value,
// row) stream into shared memory with two TMA bulk copies (double-buffered),
// all lanes multiply by the gathered duals in place, and then one lane per
// column adds that column's products SEQUENTIALLY in storage order.  Entries are stored per column in
// DESCENDING row order and the arithmetic is an unfused multiply followed by an
// add, i.e. exactly the order and rounding of the reference's row-major scatter
// (rows R-1..0, y[col] += u[i]*a), so the reduced costs are bit-identical to the
// CPU restatement, not merely within 1e-12.  Matrices with a column longer than
// one tile fall back to the lane-cooperative kernel below (tolerance parity).
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

// byte sizes of a tile's two parts inside the blob / a stage (16-byte multiples)
__host__ __device__ inline uint32_t tileValBytes(int len) { return (uint32_t)(((len + 1) & ~1) * 8); }
template <bool ROW16> __host__ __device__ inline uint32_t tileRowBytes(int len)
{
    return ROW16 ? (uint32_t)(((len + 7) & ~7) * 2) : (uint32_t)(((len + 3) & ~3) * 4);
}

struct StreamArgs {
    int nTiles;
    const int4 *tileRec;        // per tile {first column, end column, first stored entry, entries}
    const int32_t *colStart32;  // N+1
    const unsigned char *blob;  // tile-blocked copy of the streams: per tile [values | rows], 16-byte padded
    const int64_t *blobOff;     // per tile: byte offset of its block in `blob`
    const double *u;
    int uLen;
    const double *c;
    int phase1;
    double *out;
    int uSmemBytes;             // 0: gather the duals from global memory
    int uHalf;                  // UMODE 2: duals per CTA of the cluster pair (even)
};

// A stage holds up to kSpmvChunk entries counted from the aligned start (alignment: 16 bytes
// of the row stream = 4 int32 or 8 uint16 entries) plus that much slack.
template <bool ROW16> struct SpmvStage {
    static constexpr int kAlign = ROW16 ? 8 : 4;
    static constexpr int kValBytes = (kSpmvChunk + kAlign) * 8;
    static constexpr int kRowBytes = (kSpmvChunk + kAlign) * (ROW16 ? 2 : 4);
    static constexpr int kBytes = kValBytes + kRowBytes;  // multiple of 16
    static constexpr int kWarpBytes = 16 + 2 * kBytes;    // 2 mbarriers + 2 stages
};

struct ColMeta {  // one lane's columns of a tile: RAW loaded values, consumed one tile later
    int b[kSpmvColsPerLane], e[kSpmvColsPerLane];  // colStart32[col], colStart32[col+1]
    double cj[kSpmvColsPerLane];
};

// Only issues the loads: nothing here may depend on their results (a warp issues in order,
// so the first dependent instruction would stall everything behind it).
__device__ __forceinline__ void loadColMeta(const StreamArgs &a, const int4 rec, int lane, ColMeta &m)
{
#pragma unroll
    for (int q = 0; q < kSpmvColsPerLane; ++q) {
        const int col = rec.x + lane + 32 * q;
        m.b[q] = m.e[q] = 0;
        m.cj[q] = 0.0;
        if (col < rec.y) {
            m.b[q] = __ldg(a.colStart32 + col);
            m.e[q] = __ldg(a.colStart32 + col + 1);
            if (!a.phase1) m.cj[q] = __ldg(a.c + col);
        }
    }
}

// Every WARP owns a two-stage TMA pipeline over its own tiles (<= 32*kSpmvColsPerLane
// consecutive columns, <= kSpmvChunk stored entries): lane 0 issues the bulk copies of the
// (value, row) streams of tile k+2 while the warp reduces tile k out of shared memory, one
// lane per column, sequentially in storage order.  The per-tile record and the per-column
// extents / objective entries of tile k+1 are fetched into registers while tile k is being
// reduced, so no cold global load sits on the warp's critical path; there is no CTA-wide
// barrier in the loop.  Where the duals are gathered from (UMODE): 0 = global memory (from L1 a
// random 8-byte gather costs a wavefront per lane: fine for structured matrices, 71 us for 10^7
// random gathers); 1 = the whole dual vector in shared memory; 2 = the dual vector SPLIT over the
// two CTAs of a thread-block cluster, the other half read through distributed shared memory
// (mapa + ld.shared::cluster) -- half the footprint, which doubles the bytes the TMA stages can
// keep in flight per SM (the stream is in-flight-limited).
template <int UMODE, bool ROW16>
__global__ void __launch_bounds__(768) redcost_stream_kernel(const __grid_constant__ StreamArgs a)
{
    typedef SpmvStage<ROW16> St;
    typedef typename std::conditional<ROW16, uint16_t, int32_t>::type RowT;
    constexpr int kSpmvStageBytes = St::kBytes, kSpmvWarpBytes = St::kWarpBytes;
    extern __shared__ __align__(16) unsigned char sm[];
    const double *sU = reinterpret_cast<const double *>(sm);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned int crank = 0;
    if (UMODE == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int uHalf = a.uHalf;  // UMODE 2: duals [rank*uHalf, (rank+1)*uHalf) live in CTA `rank`
    const uint32_t sUaddr = rcSmemAddr(sm);
    auto gatherU = [&](int r) -> double {
        if (UMODE == 0) return __ldg(a.u + r);
        if (UMODE == 1) return sU[r];
        const unsigned int owner = r >= uHalf ? 1u : 0u;
        const uint32_t la = sUaddr + (uint32_t)(r - (owner ? uHalf : 0)) * 8u;
        uint32_t ra;
        double v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(owner));
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra));
        return v;
    };
    unsigned char *wbase = sm + a.uSmemBytes + (size_t)warp * kSpmvWarpBytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
    unsigned char *stageBase = wbase + 16;
    const int warpsPerCta = blockDim.x >> 5;
    const int gw = blockIdx.x * warpsPerCta + warp;
    const int totalWarps = gridDim.x * warpsPerCta;
    const int4 none = make_int4(0, 0, 0, 0);

    // one bulk copy per tile: its values and rows sit back to back in the tile-blocked blob
    auto issue = [&](const int4 rec, int64_t off, int s) {
        uint64_t *bar = full + s;
        const int len = rec.w;
        if (len > 0) {
            unsigned char *dst = stageBase + (size_t)s * kSpmvStageBytes;
            const uint32_t bytes = tileValBytes(len) + tileRowBytes<ROW16>(len);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rcSmemAddr(bar)), "r"(bytes)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    rcSmemAddr(dst)),
                "l"(a.blob + off), "r"(bytes), "r"(rcSmemAddr(bar))
                : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rcSmemAddr(bar)) : "memory");
        }
    };

    // tile records: cur = tile being reduced, nxt = the one after (already in flight),
    // nn = two ahead (issued when cur's stage is free)
    int4 cur = gw < a.nTiles ? __ldg(a.tileRec + gw) : none;
    int4 nxt = gw + totalWarps < a.nTiles ? __ldg(a.tileRec + gw + totalWarps) : none;
    const int64_t offCur = gw < a.nTiles ? __ldg(a.blobOff + gw) : 0;
    const int64_t offNxt = gw + totalWarps < a.nTiles ? __ldg(a.blobOff + gw + totalWarps) : 0;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rcSmemAddr(full)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rcSmemAddr(full + 1)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (gw < a.nTiles) issue(cur, offCur, 0);
        if (gw + totalWarps < a.nTiles) issue(nxt, offNxt, 1);
    }
    ColMeta mc;
    loadColMeta(a, cur, lane, mc);
    if (UMODE != 0) {
        // the dual vector (UMODE 2: this CTA's half): one bulk copy (16-byte granules) + a plain
        // tail, on its own mbarrier
        uint64_t *ubar = reinterpret_cast<uint64_t *>(sm + a.uSmemBytes - 16);
        double *w = const_cast<double *>(sU);
        const int lo = UMODE == 2 ? (int)crank * uHalf : 0;
        const int cnt = UMODE == 2 ? max(0, min(a.uLen - lo, uHalf)) : a.uLen;
        const double *src = a.u + lo;
        const int bulk = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 ? (cnt & ~1) : 0;
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
                    "l"(src), "r"((uint32_t)bulk * 8u), "r"(rcSmemAddr(ubar))
                    : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rcSmemAddr(ubar)) : "memory");
            }
        }
        for (int k = bulk + t; k < cnt; k += blockDim.x) w[k] = __ldg(src + k);
        __syncthreads();
        rcMbarWait(ubar, 0);
        if (UMODE == 2) {
            // both halves must be in place before anyone gathers across the cluster
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
    } else {
        __syncwarp();
    }

    int it = 0;
    for (int tile = gw; tile < a.nTiles; tile += totalWarps, ++it) {
        const int s = it & 1;
        const uint32_t parity = (uint32_t)((it >> 1) & 1);
        double *sVal = reinterpret_cast<double *>(stageBase + (size_t)s * kSpmvStageBytes);
        const RowT *sRow =
            reinterpret_cast<const RowT *>(stageBase + (size_t)s * kSpmvStageBytes + tileValBytes(cur.w));
        // prefetch (registers) what the next two tiles need; consumed after this tile is done
        const int tn = tile + totalWarps, tnn = tile + 2 * totalWarps;
        // unconditional (clamped) load: a select on the result would stall the warp right here
        const int4 nn = __ldg(a.tileRec + min(tnn, a.nTiles - 1));
        const int64_t offNn = __ldg(a.blobOff + min(tnn, a.nTiles - 1));
        ColMeta mn;
        loadColMeta(a, nxt, lane, mn);  // nxt is {0,0,0,0} past the end: no loads
        rcMbarWait(full + s, parity);
        // phase 1, all lanes, coalesced: products u[i]*a in place (unfused multiply, as the
        // reference's  y += u*a).  Entries in front of the tile's first column (alignment pad)
        // are real entries of the previous column, so their rows are valid.
        {
            const int len = cur.w;
            int k = lane;
            for (; k + 96 < len; k += 128) {
                int r[4];
                double v[4], g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = sRow[k + 32 * j];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = sVal[k + 32 * j];
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] = gatherU(r[j]);
#pragma unroll
                for (int j = 0; j < 4; ++j) sVal[k + 32 * j] = __dmul_rn(g[j], v[j]);
            }
            {
                // remainder (< 4 rounds): still one batch, so the gathers overlap
                int r[4];
                double v[4], g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = k + 32 * j;
                    r[j] = kk < len ? (int)sRow[kk] : 0;
                    v[j] = kk < len ? sVal[kk] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] = (k + 32 * j < len) ? gatherU(r[j]) : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k + 32 * j < len) sVal[k + 32 * j] = __dmul_rn(g[j], v[j]);
            }
        }
        __syncwarp();
        // phase 2, one lane per column: add the products sequentially in storage order
#pragma unroll
        for (int q = 0; q < kSpmvColsPerLane; ++q) {
            const int col = cur.x + lane + 32 * q;
            if (col < cur.y) {
                int k = mc.b[q] - cur.z;
                const int e = mc.e[q] - cur.z;
                double sum = 0.0;
                double p = sVal[k];  // the stage has 4 entries of slack: reading one past is safe
                while (k < e) {
                    const double pn = sVal[k + 1];
                    sum = __dadd_rn(sum, p);
                    p = pn;
                    ++k;
                }
                a.out[col] = a.phase1 ? -sum : mc.cj[q] - sum;  // DecompAlgo.cpp:4538-4545
            }
        }
        // the stage was written through the generic proxy; order that before the bulk copy
        // (async proxy) that refills it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && tnn < a.nTiles) issue(nn, offNn, s);
        (void)tn;
        cur = nxt;
        nxt = tnn < a.nTiles ? nn : none;
        mc = mn;
    }
    if (UMODE == 2) {
        // a CTA's half of the duals must outlive its partner's last gather
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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

int launchRedCost(Ctx *c, const double *dU, int phase, cudaStream_t st)
{
    if (c->N == 0) return DIPGPU_OK;
    if (c->nSpmvTiles > 0 && c->spmvLanes == 0) {
        StreamArgs a;
        a.nTiles = c->nSpmvTiles;
        a.tileRec = reinterpret_cast<const int4 *>(c->dTileRec);
        a.colStart32 = c->dColStart32;
        a.blob = c->dSpmvBlob;
        a.blobOff = c->dSpmvBlobOff;
        a.u = dU;
        a.uLen = c->R + c->numConvex;
        a.c = c->dObj;
        a.phase1 = phase == DIPGPU_PHASE_PRICE1 ? 1 : 0;
        a.out = c->dRedCost;
        const int sms = c->smCount > 0 ? c->smCount : 148;
        const size_t maxSmem = 227 * 1024;
        const bool row16 = c->dRowMaster16 != nullptr;
        const size_t warpBytes = row16 ? SpmvStage<true>::kWarpBytes : SpmvStage<false>::kWarpBytes;
        const size_t uBytes = (((size_t)a.uLen + 1) & ~(size_t)1) * 8 + 16;  // + the u mbarrier
        const int uHalf = (((a.uLen + 1) / 2) + 1) & ~1;
        const size_t uHalfBytes = (size_t)uHalf * 8 + 16;
        // u in shared memory pays when it is gathered irregularly and re-used: many more stored
        // entries than duals.  mode 2 (split over a 2-CTA cluster) leaves twice the room for stages.
        const bool reuse = c->nnz >= 8 * (int64_t)a.uLen;
        int mode = 0;
        // (mode 2 is never chosen automatically: on B200 the remote 8-byte gathers sustain only ~0.25
        // per cycle and SM, 78 us on the MILPBlock shape against 44 us for mode 1)
        if (reuse && uBytes + 6 * warpBytes <= maxSmem && c->nSpmvTiles >= 24 * sms) mode = 1;
        if (c->optSpmvUSmem >= 0) mode = c->optSpmvUSmem;
        if (mode == 1 && uBytes + warpBytes > maxSmem) mode = 0;
        if (mode == 2 && uHalfBytes + warpBytes > maxSmem) mode = 0;
        typedef void (*Fn)(const StreamArgs);
        Fn fn;
        size_t bytes;
        int grid, threads;
        a.uHalf = uHalf;
        if (mode != 0) {
            const size_t ub = mode == 2 ? uHalfBytes : uBytes;
            int warps = (int)std::min<size_t>(24, (maxSmem - ub) / warpBytes);
            if (c->optSpmvWarps > 0) warps = std::min(warps, c->optSpmvWarps);
            a.uSmemBytes = (int)ub;
            bytes = ub + (size_t)warps * warpBytes;
            fn = mode == 2 ? (row16 ? redcost_stream_kernel<2, true> : redcost_stream_kernel<2, false>)
                           : (row16 ? redcost_stream_kernel<1, true> : redcost_stream_kernel<1, false>);
            grid = sms & ~1;
            threads = warps * 32;
        } else {
            a.uSmemBytes = 0;
            const int warps = c->optSpmvWarps > 0 ? std::min(c->optSpmvWarps, 16) : 8;
            bytes = (size_t)warps * warpBytes;
            fn = row16 ? redcost_stream_kernel<0, true> : redcost_stream_kernel<0, false>;
            const int ctasPerSm = (int)std::max<size_t>(1, std::min<size_t>(4, maxSmem / bytes));
            grid = std::max(1, std::min((c->nSpmvTiles + warps - 1) / warps, ctasPerSm * sms));
            threads = warps * 32;
        }
        if (c->spmvFn != (void *)fn || c->spmvFnSmem != bytes) {
            DIPGPU_CUDA(c, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            c->spmvFn = (void *)fn;
            c->spmvFnSmem = bytes;
            c->spmvClusters = 0;
        }
        if (mode == 2) {
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(threads, 1, 1);
            cfg.dynamicSmemBytes = bytes;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            if (c->spmvClusters == 0) {
                // persistent kernel: no more CTA pairs than can be resident at once
                cfg.gridDim = dim3(grid, 1, 1);
                int nc = 0;
                DIPGPU_CUDA(c, cudaOccupancyMaxActiveClusters(&nc, fn, &cfg));
                c->spmvClusters = std::max(1, std::min(nc, grid / 2));
            }
            cfg.gridDim = dim3(2 * c->spmvClusters, 1, 1);
            DIPGPU_CUDA(c, cudaLaunchKernelEx(&cfg, fn, a));
        } else {
            fn<<<grid, threads, bytes, st>>>(a);
        }
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
                                                      dU, c->dObj, phase1, c->dRedCost)
    switch (L) {
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        default: LAUNCH(32); break;
    }
#undef LAUNCH
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

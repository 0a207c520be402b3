// pool.cu -- the two dual-vector passes that surround a pricing round in DecompAlgoPC:
//
//  * re-pricing of the waiting columns, DecompVarPool::setReducedCosts ->
//    DecompWaitingCol::setReducedCost (Dip/src/DecompVarPool.cpp:185-203, :22-40; called once per
//    price call, Dip/src/DecompAlgo.cpp:1890-1908):  redCost_s = origCost_s - col_s . u  (feasible
//    master) or -col_s . u (dual ray), col_s = the waiting column in MASTER row space;
//  * Wentges smoothing of the master duals, DecompAlgoPC::adjustMasterDualSolution
//    (Dip/src/DecompAlgoPC.cpp:126-212):  dualST = alpha*dual + (1-alpha)*dualRM, for one alpha or
//    for the whole ladder alpha, 0.9 alpha, ... that addVarsToPool walks when every new column is
//    a duplicate (Dip/src/DecompAlgo.cpp:5642-5655).
//
// Both are HBM streams: the pool pass reads 12 B per stored entry + 24 B per column, the
// smoothing reads 16 B and writes 8 B per row and vector.
#include <algorithm>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace dipgpu {

constexpr int kPoolWarpsMax = 16;
constexpr int kPoolChunk = 512;   // entries per window: 4 KiB of values + 2 KiB of rows (256 / 32 warps: 66 us, 1024 / 8 warps: 88 us, this: 55 us)
constexpr int kPoolStages = 2;    // windows in flight per warp
static_assert(kPoolChunk <= kPoolPad, "the pool arrays are padded by one window");
constexpr int kPoolPerLane = kPoolChunk / 32;

// One warp per 32 consecutive waiting columns.  Their stored entries are one contiguous span
// [B, E) of the pool.  The warp walks the aligned 512-entry windows that intersect the span from
// the END, two windows in flight: cp.async streams a window's rows and values into the warp's
// shared-memory ring (16-byte pieces, nothing held in registers while the bytes travel), all lanes
// then turn the values into products val*u[row] in place (unfused multiply), and lane l adds the
// products of ITS column, last entry first -- the order of CoinPackedVectorBase::dotProduct as the
// oracle restates it (dp = 0; for i = n-1..0: dp += els[i]*u[ind[i]]), so the values are bit-equal
// to the oracle.
//
// BITMAP: a dual support is declared (dipgpu_set_dual_support) and its bitmap (one bit per master row)
// fits shared memory.  Most entries of a master column sit in rows that cannot carry a dual (GAP: the
// free branch rows, 2/3 of every column): the gather of the dual -- one 32-byte sector per 8 useful
// bytes -- is skipped for them, a shared-memory bit test decides.  The product is still formed
// (val * 0), so the result is the same bit for bit.
template <bool BITMAP>
__global__ void __launch_bounds__(kPoolWarpsMax * 32, 1)
pool_redcost_kernel(long long nCols, const int64_t *__restrict__ colStart, const int32_t *__restrict__ row,
                    const double *__restrict__ val, const double *__restrict__ origCost,
                    const double *__restrict__ u, int feasible, double *__restrict__ out, int32_t *flag,
                    const uint32_t *__restrict__ supportBits, int bitWords)
{
    extern __shared__ __align__(16) unsigned char poolSm[];
    const int warpsPerCta = blockDim.x >> 5;
    constexpr int kStageBytes = kPoolChunk * 12;  // values, then rows
    const uint32_t *sBits = reinterpret_cast<const uint32_t *>(poolSm + (size_t)warpsPerCta * kPoolStages * kStageBytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (BITMAP) {
        // the bitmap travels as its own cp.async group (16-byte pieces; the array is padded to a multiple of 16
        // bytes), behind it the warp's first windows: nobody waits for the 32 KiB before the stream has started
        const uint32_t dstB = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(sBits));
        for (int k = threadIdx.x * 4; k < bitWords; k += blockDim.x * 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstB + 4u * k), "l"(supportBits + k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const long long c0 = ((long long)blockIdx.x * warpsPerCta + warp) * 32;
    const bool active = c0 < nCols;
    const long long col = c0 + lane;
    const long long cEnd = min(nCols, c0 + 32);
    int64_t B = 0, E = 0;
    if (active) {
        B = colStart[c0];
        E = colStart[cEnd];
    }
    int64_t b = 0, e = 0;
    double oc = 0.0;
    if (col < nCols) {
        b = colStart[col];
        e = colStart[col + 1];
        oc = origCost[col];
    }
    double sum = 0.0;
    unsigned char *ring = poolSm + (size_t)warp * kPoolStages * kStageBytes;
    const uint32_t ringA = (uint32_t)__cvta_generic_to_shared(ring);
    const int64_t kLast = E > B ? (E - 1) / kPoolChunk : -1, kFirst = E > B ? B / kPoolChunk : 0;  // no entries: no window
    // window k -> ring slot, 16-byte pieces
    auto issue = [&](int64_t k, int slot) {
        const uint32_t dst = ringA + (uint32_t)slot * kStageBytes;
        const char *sv = reinterpret_cast<const char *>(val + k * kPoolChunk);
        const char *sr = reinterpret_cast<const char *>(row + k * kPoolChunk);
#pragma unroll
        for (int j = 0; j < kPoolChunk / 64; ++j) {
            const int off = (lane + 32 * j) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(sv + off) : "memory");
        }
#pragma unroll
        for (int j = 0; j < kPoolChunk / 128; ++j) {
            const int off = (lane + 32 * j) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + kPoolChunk * 8 + off), "l"(sr + off) : "memory");
        }
    };
    int64_t kIssue = kLast;
#pragma unroll
    for (int s = 0; s < kPoolStages; ++s) {
        if (kIssue >= kFirst) issue(kIssue--, s);
        asm volatile("cp.async.commit_group;" ::: "memory");  // (possibly empty: the group count stays uniform)
    }
    if (BITMAP) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kPoolStages) : "memory");  // this thread's pieces of the bitmap
        __syncthreads();                                                       // ... and everybody else's
    }
    if (!active) return;
    {
        int slot = 0;
        for (int64_t k = kLast; k >= kFirst; --k) {
            asm volatile("cp.async.wait_group %0;" ::"n"(kPoolStages - 1) : "memory");
            __syncwarp();
            double *p = reinterpret_cast<double *>(ring + (size_t)slot * kStageBytes);
            const int32_t *rw = reinterpret_cast<const int32_t *>(p + kPoolChunk);
            const int64_t w0 = k * kPoolChunk;
            const int lo = (int)(max(B, w0) - w0), hi = (int)(min(E, w0 + kPoolChunk) - w0);  // this warp's part of the window
            // the lane's entries of the window at once: the rows, then their (few) dependent gathers of the dual
            // -- all in flight together, one round trip per window --, then the products in place
            {
                double g[kPoolPerLane];
#pragma unroll
                for (int j = 0; j < kPoolPerLane; ++j) {
                    const int i = lane + 32 * j;
                    bool on = i >= lo && i < hi;
                    int r = 0;
                    if (on) {
                        r = rw[i];
                        if (BITMAP) on = (sBits[r >> 5] >> (r & 31)) & 1u;
                    }
                    g[j] = on ? __ldg(u + r) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < kPoolPerLane; ++j) {
                    const int i = lane + 32 * j;
                    if (i >= lo && i < hi) p[i] = __dmul_rn(p[i], g[j]);
                }
            }
            __syncwarp();
            const int64_t kHi = min(w0 + hi, e), kLo = max(w0 + lo, b);
            if (kHi > kLo) {
                // four products are fetched while the previous four are added: the chain of dependent additions
                // (one lane, last entry first) never waits for shared memory
                int q = (int)(kHi - 1 - w0);
                const int qLo = (int)(kLo - w0);
                if (q - 3 >= qLo) {
                    double a0 = p[q], a1 = p[q - 1], a2 = p[q - 2], a3 = p[q - 3];
                    q -= 4;
                    while (q - 3 >= qLo) {
                        const double n0 = p[q], n1 = p[q - 1], n2 = p[q - 2], n3 = p[q - 3];
                        sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, a0), a1), a2), a3);
                        a0 = n0;
                        a1 = n1;
                        a2 = n2;
                        a3 = n3;
                        q -= 4;
                    }
                    sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, a0), a1), a2), a3);
                }
                for (; q >= qLo; --q) sum = __dadd_rn(sum, p[q]);
            }
            __syncwarp();
            if (kIssue >= kFirst) issue(kIssue--, slot);
            asm volatile("cp.async.commit_group;" ::: "memory");
            slot = slot + 1 == kPoolStages ? 0 : slot + 1;
        }
    }
    if (col < nCols) {
        const double rc = feasible ? oc - sum : -sum;  // DecompVarPool.cpp:29 / :36
        out[col] = rc;
        if (rc <= -0.0000000001) *flag = 1;            // :30, :37 -> found_negrc_var (:199)
    }
}

// ---- persistent variant: one CTA per SM, every warp walks the 32-column tiles  w, w + W, w + 2W, ...  (W = warps of
// the launch) with ONE window ring that keeps streaming across tile borders: no wave quantisation (391 CTAs of 16
// tiles on 148 SMs were 2.6 waves), the per-CTA set-up paid once per SM, and the next tile's first windows already in
// flight while the last ones of the current tile are added.
//   MODE 0: duals gathered from global memory (L2);
//   MODE 1: dual support declared: shared-memory bitmap skips the rows that cannot carry a dual;
//   MODE 2: as 1, and the gather itself stays on the SM: u[support] is compacted into shared memory once per CTA
//           (GAP-E: 4 240 doubles) and a row finds its slot by rank = prefix[word] + popc(bits below it in the
//           word); no L2 round trip between a window's arrival and its products.
// The additions and their order are those of the kernel above (per column: last entry first, unfused products).
struct PoolArgs {
    long long nCols;
    const int64_t *colStart;
    const int32_t *row;
    const double *val;
    const double *origCost;
    const double *u;
    int feasible;
    double *out;
    int32_t *flag;
    const uint32_t *supportBits;   // MODE >= 1
    int bitWords;                  // padded to a multiple of 4
    const uint16_t *rankPrefix;    // MODE 2: support rows below word w
    const int32_t *support;        // MODE 2
    int nSupport;
};

template <int MODE, int CH>
__global__ void __launch_bounds__(CH == 512 ? kPoolWarpsMax * 32 : 768, 1) pool_redcost_persist_kernel(const __grid_constant__ PoolArgs a)
{
    constexpr int kPerLane = CH / 32;
    extern __shared__ __align__(16) unsigned char poolSm[];
    const int warpsPerCta = blockDim.x >> 5;
    constexpr int kStageBytes = CH * 12;  // values, then rows
    unsigned char *extra = poolSm + (size_t)warpsPerCta * kPoolStages * kStageBytes;
    const uint32_t *sBits = reinterpret_cast<const uint32_t *>(extra);
    const uint16_t *sPre = reinterpret_cast<const uint16_t *>(extra + (size_t)a.bitWords * 4);
    double *sU = reinterpret_cast<double *>(extra + (size_t)a.bitWords * 4 + (((size_t)a.bitWords * 2 + 15) & ~(size_t)15));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (MODE >= 1) {
        const uint32_t dstB = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(sBits));
        for (int k = threadIdx.x * 4; k < a.bitWords; k += blockDim.x * 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstB + 4u * k), "l"(a.supportBits + k) : "memory");
        if (MODE == 2) {
            const uint32_t dstP = (uint32_t)__cvta_generic_to_shared(const_cast<uint16_t *>(sPre));
            for (int k = threadIdx.x * 8; k < a.bitWords; k += blockDim.x * 8)  // (bitWords % 8 == 0 in this mode)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstP + 2u * k), "l"(a.rankPrefix + k) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    const long long nTiles = (a.nCols + 31) / 32;
    const long long W = (long long)gridDim.x * warpsPerCta;
    // a tile as the warp sees it: its span [B, E) of the pool (uniform) and the lane's own column [b, e), cost oc
    struct Tile { long long id; int64_t B, E, kFirst, kLast, b, e; double oc; };
    auto loadTile = [&](long long id) {
        Tile t;
        t.id = id;
        t.B = t.E = t.b = t.e = 0;
        t.kFirst = 0;
        t.kLast = -1;
        t.oc = 0.0;
        if (id < nTiles) {
            const long long c0 = id * 32, col = c0 + lane;
            const int64_t cs = a.colStart[min(col, a.nCols)];
            t.E = a.colStart[min(c0 + 32, a.nCols)];
            t.B = __shfl_sync(0xffffffffu, cs, 0);
            const int64_t nx = __shfl_down_sync(0xffffffffu, cs, 1);
            t.b = cs;
            t.e = lane == 31 ? t.E : nx;
            t.oc = col < a.nCols ? a.origCost[col] : 0.0;
            if (t.E > t.B) {
                t.kLast = (t.E - 1) / CH;
                t.kFirst = t.B / CH;
            }
        }
        return t;
    };
    unsigned char *ring = poolSm + (size_t)warp * kPoolStages * kStageBytes;
    const uint32_t ringA = (uint32_t)__cvta_generic_to_shared(ring);
    auto issue = [&](int64_t k, int slot) {
        const uint32_t dst = ringA + (uint32_t)slot * kStageBytes;
        const char *sv = reinterpret_cast<const char *>(a.val + k * CH);
        const char *sr = reinterpret_cast<const char *>(a.row + k * CH);
#pragma unroll
        for (int j = 0; j < CH / 64; ++j) {
            const int off = (lane + 32 * j) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(sv + off) : "memory");
        }
#pragma unroll
        for (int j = 0; j < CH / 128; ++j) {
            const int off = (lane + 32 * j) * 16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + CH * 8 + off), "l"(sr + off) : "memory");
        }
    };

    Tile A = loadTile((long long)blockIdx.x * warpsPerCta + warp);
    Tile N = loadTile(A.id + W);
    unsigned pc = 0, cc = 0;     // windows issued / consumed by this warp (ring slot = count % kPoolStages)
    int prodInNext = 0;          // the producer has moved on to tile N
    int64_t pk = A.kLast;        // next window the producer issues (descending), in A or in N
    auto topUp = [&]() {
        while (pc - cc < (unsigned)kPoolStages) {
            if (!prodInNext && pk < A.kFirst) {
                if (N.kLast < N.kFirst) break;   // (no next tile, or an empty one: the producer waits for the rotation)
                prodInNext = 1;
                pk = N.kLast;
            }
            if (prodInNext && pk < N.kFirst) break;
            issue(pk--, (int)(pc % kPoolStages));
            asm volatile("cp.async.commit_group;" ::: "memory");
            ++pc;
        }
    };
    topUp();   // the stream starts before anybody waits for the bitmap / the compacted duals
    if (MODE == 2) {
        for (int k = threadIdx.x; k < a.nSupport; k += blockDim.x) sU[k] = __ldg(a.u + __ldg(a.support + k));
        if (threadIdx.x == 0) sU[a.nSupport] = 0.0;   // where the rows without a dual point
    }
    if (MODE >= 1) {
        // this thread's pieces of the bitmap: the group committed first, i.e. everything but the windows issued since
        if (pc == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else if (pc == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncthreads();
    }
    static_assert(kPoolStages == 2, "the group accounting below is written for two stages");

    int64_t ck = A.kLast;
    double sum = 0.0;
    while (A.id < nTiles) {
        if (ck >= A.kFirst) {
            // the group of window cc: (pc - cc - 1) younger groups may stay in flight
            if (pc - cc >= 2u) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            double *p = reinterpret_cast<double *>(ring + (size_t)(cc % kPoolStages) * kStageBytes);
            const int32_t *rw = reinterpret_cast<const int32_t *>(p + CH);
            const int64_t w0 = ck * CH;
            const int lo = (int)(max(A.B, w0) - w0), hi = (int)(min(A.E, w0 + CH) - w0);
            {
                double g[kPerLane];
#pragma unroll
                for (int j = 0; j < kPerLane; ++j) {
                    const int i = lane + 32 * j;
                    const bool in = i >= lo && i < hi;
                    const int r = in ? rw[i] : 0;
                    if (MODE == 2) {
                        const uint32_t bw = sBits[r >> 5];
                        const uint32_t bit = 1u << (r & 31);
                        const int slot = (int)sPre[r >> 5] + __popc(bw & (bit - 1u));
                        g[j] = sU[(in && (bw & bit)) ? slot : a.nSupport];
                    } else {
                        bool on = in;
                        if (MODE == 1 && on) on = (sBits[r >> 5] >> (r & 31)) & 1u;
                        g[j] = on ? __ldg(a.u + r) : 0.0;
                    }
                }
#pragma unroll
                for (int j = 0; j < kPerLane; ++j) {
                    const int i = lane + 32 * j;
                    if (i >= lo && i < hi) p[i] = __dmul_rn(p[i], g[j]);
                }
            }
            __syncwarp();
            const int64_t kHi = min(w0 + hi, A.e), kLo = max(w0 + lo, A.b);
            if (kHi > kLo) {
                int q = (int)(kHi - 1 - w0);
                const int qLo = (int)(kLo - w0);
                if (q - 3 >= qLo) {
                    double a0 = p[q], a1 = p[q - 1], a2 = p[q - 2], a3 = p[q - 3];
                    q -= 4;
                    while (q - 3 >= qLo) {
                        const double n0 = p[q], n1 = p[q - 1], n2 = p[q - 2], n3 = p[q - 3];
                        sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, a0), a1), a2), a3);
                        a0 = n0;
                        a1 = n1;
                        a2 = n2;
                        a3 = n3;
                        q -= 4;
                    }
                    sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, a0), a1), a2), a3);
                }
                for (; q >= qLo; --q) sum = __dadd_rn(sum, p[q]);
            }
            __syncwarp();
            --ck;
            ++cc;
            topUp();
        }
        if (ck < A.kFirst) {
            // the tile is done: its 32 reduced costs, then the next tile becomes the current one
            const long long col = A.id * 32 + lane;
            if (col < a.nCols) {
                const double rc = a.feasible ? A.oc - sum : -sum;  // DecompVarPool.cpp:29 / :36
                a.out[col] = rc;
                if (rc <= -0.0000000001) *a.flag = 1;              // :30, :37 -> found_negrc_var (:199)
            }
            A = N;
            sum = 0.0;
            if (prodInNext) prodInNext = 0;   // (pk keeps counting down inside what is now A)
            else pk = A.kLast;
            ck = A.kLast;
            N = loadTile(A.id + W);
            topUp();
        }
    }
}

// ---- the compacted copy of the pool (Ctx::dPoolC*): entries of rows that can carry a dual, in their column's order.
// One warp per column; count pass, exclusive scan (cub), fill pass.
__global__ void __launch_bounds__(256)
pool_compact_kernel(long long nCols, const int64_t *__restrict__ colStart, const int32_t *__restrict__ row,
                    const double *__restrict__ val, const uint32_t *__restrict__ supportBits,
                    int64_t *__restrict__ count, const int64_t *__restrict__ cStart, int32_t *__restrict__ cRow,
                    double *__restrict__ cVal)
{
    const long long col = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (col >= nCols) return;
    const int64_t b = colStart[col], e = colStart[col + 1];
    int64_t out = cStart != nullptr ? cStart[col] : 0;
    int64_t n = 0;
    for (int64_t k0 = b; k0 < e; k0 += 32) {
        const int64_t k = k0 + lane;
        int r = 0;
        bool on = false;
        if (k < e) {
            r = row[k];
            on = (supportBits[r >> 5] >> (r & 31)) & 1u;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, on);
        if (cStart != nullptr && on) {
            const int64_t dst = out + __popc(bal & ((1u << lane) - 1u));
            cRow[dst] = r;
            cVal[dst] = val[k];
        }
        out += __popc(bal);
        n += __popc(bal);
    }
    if (cStart == nullptr && lane == 0) count[col] = n;
}

// Builds the compacted copy when the pool or the support changed since it was last built.
static int ensurePoolCompact(Ctx *c, cudaStream_t st)
{
    if (c->poolCompactPool == c->poolSerial && c->poolCompactSupport == c->supportSerial) return DIPGPU_OK;
    const long long n = c->poolCols;
    if (n + 1 > c->poolCColCap) {
        if (c->dPoolCStart) cudaFree(c->dPoolCStart);
        if (c->dPoolCCount) cudaFree(c->dPoolCCount);
        c->dPoolCStart = c->dPoolCCount = nullptr;
        c->poolCColCap = n + n / 4 + 64;
        DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dPoolCStart), sizeof(int64_t) * (size_t)c->poolCColCap));
        DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dPoolCCount), sizeof(int64_t) * (size_t)c->poolCColCap));
    }
    const unsigned grid = (unsigned)((n * 32 + 255) / 256);
    DIPGPU_CUDA(c, cudaMemsetAsync(c->dPoolCCount + n, 0, sizeof(int64_t), st));   // the scan's last input: the total comes out at [n]
    pool_compact_kernel<<<grid, 256, 0, st>>>(n, c->dPoolStart, c->dPoolRow, c->dPoolVal, c->dSupportBits, c->dPoolCCount, nullptr, nullptr, nullptr);
    c->launches++;
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->dPoolCCount, c->dPoolCStart, (int)(n + 1), st);
    if (tmp > c->poolScanTmpBytes) {
        if (c->dPoolScanTmp) cudaFree(c->dPoolScanTmp);
        c->dPoolScanTmp = nullptr;
        DIPGPU_CUDA(c, cudaMalloc(&c->dPoolScanTmp, tmp + 256));
        c->poolScanTmpBytes = tmp + 256;
    }
    tmp = c->poolScanTmpBytes;
    if (cub::DeviceScan::ExclusiveSum(c->dPoolScanTmp, tmp, c->dPoolCCount, c->dPoolCStart, (int)(n + 1), st) != cudaSuccess)
        return fail(c, DIPGPU_ERR_CUDA, "pool compaction: scan failed");
    c->launches++;
    int64_t total = 0;
    DIPGPU_CUDA(c, cudaMemcpyAsync(&total, c->dPoolCStart + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    DIPGPU_CUDA(c, cudaStreamSynchronize(st));
    if (total + kPoolPad > c->poolCNnzCap) {
        if (c->dPoolCRow) cudaFree(c->dPoolCRow);
        if (c->dPoolCVal) cudaFree(c->dPoolCVal);
        c->dPoolCRow = nullptr;
        c->dPoolCVal = nullptr;
        c->poolCNnzCap = total + total / 4 + kPoolPad + 64;
        DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dPoolCRow), sizeof(int32_t) * (size_t)c->poolCNnzCap));
        DIPGPU_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&c->dPoolCVal), sizeof(double) * (size_t)c->poolCNnzCap));
        DIPGPU_CUDA(c, cudaMemsetAsync(c->dPoolCRow, 0, sizeof(int32_t) * (size_t)c->poolCNnzCap, st));   // (whole windows are streamed: defined bytes behind the end)
        DIPGPU_CUDA(c, cudaMemsetAsync(c->dPoolCVal, 0, sizeof(double) * (size_t)c->poolCNnzCap, st));
    }
    pool_compact_kernel<<<grid, 256, 0, st>>>(n, c->dPoolStart, c->dPoolRow, c->dPoolVal, c->dSupportBits, c->dPoolCCount, c->dPoolCStart, c->dPoolCRow, c->dPoolCVal);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    c->poolCNnz = total;
    c->poolCompactPool = c->poolSerial;
    c->poolCompactSupport = c->supportSerial;
    return DIPGPU_OK;
}

int launchPoolRedCost(Ctx *c, const double *dU, int feasible, cudaStream_t st)
{
    DIPGPU_CUDA(c, cudaMemsetAsync(c->dPoolFlag, 0, sizeof(int32_t), st));
    if (c->poolCols == 0) return DIPGPU_OK;
    const long long warps = (c->poolCols + 31) / 32;
    // the bitmap variants pay when the pool is large (the bitmap is loaded once per CTA) and the support small
    const int bitWordsRaw = (c->R + c->numConvex + 31) / 32;
    const bool bitmap = c->dSupportBits != nullptr && !c->hSupport.empty() && bitWordsRaw * 4 <= 64 * 1024 &&
                        c->poolCols >= 32 * 1024 && (int64_t)c->hSupport.size() * 4 <= (int64_t)c->R + c->numConvex;
    const size_t perWarp = (size_t)kPoolStages * kPoolChunk * 12;
    if (c->optPoolKernel != 0) {
        // persistent kernel (default).  MODE 2 (option "pool_kernel" = 2) needs the rank table and room for at least
        // 8 warps' rings beside bitmap + prefixes + compacted duals; measured SLOWER than the L2 gathers of MODE 1 on
        // the GAP-E pool (66 vs 55 us: 82 KB of per-CTA set-up and 12 instead of 16 warps), so it is not the default
        const int bitWords = (bitWordsRaw + 7) & ~7;   // (the arrays are allocated with this padding)
        const size_t ns = c->hSupport.size();
        const size_t bitsB = sizeof(uint32_t) * (size_t)bitWords, preB = ((size_t)bitWords * 2 + 15) & ~(size_t)15, uB = sizeof(double) * (ns + 2);
        int mode = bitmap ? 1 : 0;
        if (bitmap && c->dSupportRank != nullptr && 227 * 1024 >= bitsB + preB + uB + 8 * perWarp && c->optPoolKernel == 2) mode = 2;
        // the compacted copy (only support-row entries): every entry left has a dual to gather, no bitmap in the kernel
        const bool compact = bitmap && c->optPoolCompact && mode == 1;
        if (compact) {
            const int rc = ensurePoolCompact(c, st);
            if (rc) return rc;
            mode = 0;
        }
        const size_t extra = mode == 2 ? bitsB + preB + uB : mode == 1 ? bitsB : 0;
        // the compacted copy has short columns (GAP: 17 entries): 256-entry windows and more warps per SM
        const bool small = compact && c->optPoolChunk != 512;
        const size_t perWarpK = small ? perWarp / 2 : perWarp;
        int wpc = (int)std::min<size_t>(small ? (size_t)c->optPoolWarps : (size_t)kPoolWarpsMax, (227 * 1024 - extra) / perWarpK);
        const int sms = std::max(c->smCount, 1);
        while (wpc > 2 && (warps + wpc - 1) / wpc < sms) wpc /= 2;
        const unsigned grid = (unsigned)std::min<long long>(sms, (warps + wpc - 1) / wpc);
        const size_t smem = perWarpK * (size_t)wpc + extra;
        auto fn = small ? pool_redcost_persist_kernel<0, 256>
                        : mode == 2 ? pool_redcost_persist_kernel<2, 512> : mode == 1 ? pool_redcost_persist_kernel<1, 512> : pool_redcost_persist_kernel<0, 512>;
        const int slot = small ? 3 : mode;
        if (c->poolFnSmem2[slot] < smem) {
            const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(fn), smem);
            if (rc) return rc;
            c->poolFnSmem2[slot] = smem;
        }
        PoolArgs a;
        a.nCols = c->poolCols;
        a.colStart = compact ? c->dPoolCStart : c->dPoolStart;
        a.row = compact ? c->dPoolCRow : c->dPoolRow;
        a.val = compact ? c->dPoolCVal : c->dPoolVal;
        a.origCost = c->dPoolOrigCost;
        a.u = dU;
        a.feasible = feasible;
        a.out = c->dPoolRedCost;
        a.flag = c->dPoolFlag;
        a.supportBits = c->dSupportBits;
        a.bitWords = bitWords;
        a.rankPrefix = c->dSupportRank;
        a.support = c->dSupport;
        a.nSupport = (int)ns;
        fn<<<grid, wpc * 32, smem, st>>>(a);
        c->launches++;
        DIPGPU_CUDA(c, cudaGetLastError());
        return DIPGPU_OK;
    }
    const int bitWords = bitWordsRaw;
    const size_t bits = bitmap ? sizeof(uint32_t) * (((size_t)bitWords + 3) & ~(size_t)3) : 0;
    // one CTA per SM, as many warps as the ring buffers leave room for (12 KiB each); small pools: fewer warps per
    // CTA so that the columns spread over the SMs
    int wpc = (int)std::min<size_t>(kPoolWarpsMax, (227 * 1024 - bits) / perWarp);
    while (wpc > 2 && (warps + wpc - 1) / wpc < c->smCount) wpc /= 2;
    const size_t smem = perWarp * (size_t)wpc + bits;
    const unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
    auto fn = bitmap ? pool_redcost_kernel<true> : pool_redcost_kernel<false>;
    if (c->poolFnSmem[bitmap ? 1 : 0] < smem) {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(fn), smem);
        if (rc) return rc;
        c->poolFnSmem[bitmap ? 1 : 0] = smem;
    }
    fn<<<grid, wpc * 32, smem, st>>>(c->poolCols, c->dPoolStart, c->dPoolRow, c->dPoolVal, c->dPoolOrigCost, dU, feasible,
                                    c->dPoolRedCost, c->dPoolFlag, c->dSupportBits, bitWords);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// dualST_v[r] = (alpha_v * dual[r]) + ((1 - alpha_v) * dualRM[r])   (DecompAlgoPC.cpp:153-154, :179-181)
// for nVec smoothing parameters at once; two unfused multiplies and one add, as the reference.
__global__ void __launch_bounds__(256)
smooth_duals_kernel(const double *__restrict__ uRM, const double *__restrict__ center,
                    const double *__restrict__ alphas, int nVec, int uLen, double *__restrict__ out, long long outStride)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= uLen) return;
    const double d = center[r], rm = uRM[r];
    for (int v = 0; v < nVec; ++v) {
        const double alpha = alphas[v];
        const double alpha1 = __dsub_rn(1.0, alpha);
        out[(long long)v * outStride + r] = __dadd_rn(__dmul_rn(alpha, d), __dmul_rn(alpha1, rm));
    }
}

int launchSmoothDuals(Ctx *c, const double *dURM, const double *dCenter, const double *dAlphas, int nVec, int uLen,
                      double *dOut, int64_t outStride, cudaStream_t st)
{
    if (uLen <= 0 || nVec <= 0) return DIPGPU_OK;
    smooth_duals_kernel<<<(uLen + 255) / 256, 256, 0, st>>>(dURM, dCenter, dAlphas, nVec, uLen, dOut, outStride);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// dst_v[idx[k]] = vals_v[k] (GATHER: vals_v[idx[k]] -- vals is the caller's whole dual vector in mapped host
// memory): the support of a sparse dual upload (dipgpu_set_dual_support)
template <bool GATHER>
__global__ void __launch_bounds__(256)
scatter_duals_kernel(int n, const int32_t *__restrict__ idx, const double *__restrict__ vals, long long valStride,
                     double *__restrict__ dst, long long dstStride)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int i = idx[k];
        dst[(long long)blockIdx.y * dstStride + i] = vals[(long long)blockIdx.y * valStride + (GATHER ? i : k)];
    }
}

// dst_v[k] = vals_v[idx[k]]: u[support] picked out of whole dual vectors in device-visible (page-locked host) memory
// into the compact vectors the fused kernel's in-block reduced costs read
__global__ void __launch_bounds__(256)
gather_duals_kernel(int n, const int32_t *__restrict__ idx, const double *__restrict__ vals, long long valStride,
                    double *__restrict__ dst, long long dstStride)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[(long long)blockIdx.y * dstStride + k] = vals[(long long)blockIdx.y * valStride + idx[k]];
}

int launchGatherDuals(Ctx *c, int n, const int32_t *dIdx, const double *dVals, int64_t valStride, double *dDst, int64_t dstStride,
                      int nVec, cudaStream_t st)
{
    if (n <= 0 || nVec <= 0) return DIPGPU_OK;
    gather_duals_kernel<<<dim3((n + 255) / 256, nVec, 1), 256, 0, st>>>(n, dIdx, dVals, valStride, dDst, dstStride);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

int launchScatterDuals(Ctx *c, int n, const int32_t *dIdx, const double *dVals, int64_t valStride, bool gather, double *dDst,
                       int64_t dstStride, int nVec, cudaStream_t st)
{
    if (n <= 0 || nVec <= 0) return DIPGPU_OK;
    const dim3 grid((n + 255) / 256, nVec, 1);
    if (gather) scatter_duals_kernel<true><<<grid, 256, 0, st>>>(n, dIdx, dVals, valStride, dDst, dstStride);
    else scatter_duals_kernel<false><<<grid, 256, 0, st>>>(n, dIdx, dVals, valStride, dDst, dstStride);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

// ------------------------------------------------------------------ pool upkeep
//
// Columns enter the pool either from host arrays (DecompVarPool::push_back of a waiting column
// the host built) or straight from the packed master columns of the last expansion, and leave it
// when addVarsFromPool moves them into the master (Dip/src/DecompAlgo.cpp:5678-5900).

struct PoolCopyDesc {
    int64_t src;  // first entry in the source
    int64_t dst;  // first entry in the pool
    int32_t len;
    int32_t pad;
};

// srcEntries != NULL: packed master columns (MColEntry); else (srcRow, srcVal).
__global__ void __launch_bounds__(256)
pool_copy_kernel(int n, const PoolCopyDesc *__restrict__ d, const MColEntry *__restrict__ srcEntries,
                 const int32_t *__restrict__ srcRow, const double *__restrict__ srcVal, int32_t *__restrict__ dstRow,
                 double *__restrict__ dstVal)
{
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= n) return;
    const PoolCopyDesc dk = d[k];
    for (int q = lane; q < dk.len; q += 32) {
        if (srcEntries != nullptr) {
            const MColEntry en = srcEntries[dk.src + q];
            dstRow[dk.dst + q] = en.row;
            dstVal[dk.dst + q] = en.val;
        } else {
            dstRow[dk.dst + q] = srcRow[dk.src + q];
            dstVal[dk.dst + q] = srcVal[dk.src + q];
        }
    }
}

int launchPoolCopy(Ctx *c, int n, const void *dDesc, const void *srcEntries, const int32_t *srcRow, const double *srcVal,
                   int32_t *dstRow, double *dstVal, cudaStream_t st)
{
    if (n <= 0) return DIPGPU_OK;
    pool_copy_kernel<<<(n + 7) / 8, 256, 0, st>>>(n, static_cast<const PoolCopyDesc *>(dDesc),
                                                  static_cast<const MColEntry *>(srcEntries), srcRow, srcVal, dstRow, dstVal);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

// expand.cu -- the step right after the pricing round: every kept column
// s is expanded into its master-LP column  [A''s ; e_block]  and fingerprinted for duplicate
// detection.  Replaces the per-column work of DecompAlgo::addVarsToPool
// (Dip/src/DecompAlgo.cpp:5452-5558): fillDenseArr (N) + M->times (nnz) + dense->sparse scan (R)
// per column on the CPU, and the string hash of Dip/src/UtilHash.cpp:52-68 that
// DecompVarPool::isDuplicate compares (Dip/src/DecompVarPool.cpp:150-181).
//
// One CTA per kept column, all out of the device-resident packed result of the last round:
//   1. the column's x-indices set a membership bitmap; the CSC entries of those x-columns mark the
//      touched core rows in a shared-memory bitmap;
//   2. the touched rows are enumerated in ascending order (popc + block scan);
//   3. one thread per touched row re-walks that row of the ROW-ORDERED matrix in storage order and
//      adds the coefficients of the member columns -- exactly the additions, in exactly the order,
//      of CoinPackedMatrix::times on the reference's row-major A'';
//   4. |v| > TolZero entries (UtilPackedVectorFromDense, Dip/src/UtilMacrosDecomp.cpp:46-60), with
//      the cut rows moved behind the convexity rows and the 1.0 of the block's convexity row
//      (:5532-5540), are compacted in order;
//   5. columns find their offset in the packed output by publishing their length and polling the
//      lower-numbered columns (same scheme as the fused DP kernel), so the output is tight and
//      needs no second launch.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace dipgpu {

constexpr int kExpThreads = 512;   // (256: 22.5 us, 512: 18.1 us, 1024: 19 us per GAP-E round, cold L2)

struct ExpandArgs {
    // last round's packed result (device)
    const ResultHeader *hdr;
    const int64_t *keptStart;
    const int32_t *keptBlock;
    const int32_t *keptInd;
    const int32_t *blockColStart;
    int firstBlock;
    // A'' column-major (for the touched rows) and row-major (for the sums)
    const int32_t *colStart32;
    const void *rowMaster;  // uint16 or int32
    int row16;
    const int32_t *csrRowStart;
    const int32_t *csrColInd;
    const double *csrVal;
    int R, nBaseRows, numConvex;
    int deltaRow0;  // rows [deltaRow0, R) were appended after the column-major copy was built (Ctx::baseR): not in it
    double tolZero;
    // output
    MColsHeader *outHdr;
    int64_t *outColStart;  // nKept+1
    uint64_t *outHash;     // nKept
    MColEntry *outEntries;
    int64_t capEntries;
    // synchronisation
    unsigned long long *pubWord;  // per column: valid (1) | epoch (23) | length (40)
    // (launch epoch << 32) | CTAs started: a CTA's column is the TICKET it draws on entry, so the columns it
    // waits for below belong to CTAs that started before it, whatever the dispatch order (see knapsack.cu)
    unsigned long long *ticket;
    // host mirror (0 = none): every store into the packed master columns is repeated hostOff bytes further on, in mapped
    // host memory; the last CTA to finish then stores doneVal into doneFlag (mapped host memory) behind a system-scope
    // fence -- the host reads the columns without a device-to-host copy and without synchronising the stream
    long long hostOff;
    unsigned int *doneCtr;   // device: CTAs finished (reset by the last one)
    unsigned int *doneFlag;
    unsigned int doneVal;
    int rowWords;   // ceil(R/32)
    int rowWordsPhys;  // with the skew: rowWords + rowWords/32 + 1
    int colWords;   // ceil(maxBlockCols/32)
    int capStage;   // staged rows per column
};

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, *total the sum
__device__ __forceinline__ int blockExclusiveScan(int v, int *sWarp, int *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += up;
    }
    __syncthreads();  // sWarp may still be read from a previous scan
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
    for (int w = 0; w < kExpThreads / 32; ++w) {
        const int c = sWarp[w];
        off += w < warp ? c : 0;
        tot += c;
    }
    *total = tot;
    return off + incl - v;
}

__global__ void __launch_bounds__(kExpThreads) expand_columns_kernel(const __grid_constant__ ExpandArgs a)
{
    extern __shared__ __align__(16) unsigned char smRaw[];
    // the row bitmap is stored skewed, word w at w + w/32: a thread scans a contiguous run of words (step 2), and
    // 32 such runs side by side would otherwise sit in one shared-memory bank
    uint32_t *rowBits = reinterpret_cast<uint32_t *>(smRaw);            // rowWordsPhys
    uint32_t *colBits = rowBits + a.rowWordsPhys;                        // colWords
#define RB(w) rowBits[(w) + ((w) >> 5)]
    int32_t *sRows = reinterpret_cast<int32_t *>(rowBits + ((a.rowWordsPhys + a.colWords + 3) & ~3));  // capStage: touched core rows
    double *sVals = reinterpret_cast<double *>(sRows + ((a.capStage + 1) & ~1));  // capStage
    __shared__ int sWarp[kExpThreads / 32];
    __shared__ unsigned long long sHashW[kExpThreads / 32];
    __shared__ unsigned long long sOffW[kExpThreads / 32];
    __shared__ int sOverflow;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __shared__ unsigned long long sTicket;
    if (t == 0) {
        const unsigned long long tk = atomicAdd(a.ticket, 1ull);
        if ((unsigned int)tk == gridDim.x - 1u) atomicAdd(a.ticket, (1ull << 32) - (unsigned long long)gridDim.x);
        sTicket = tk;
    }
    __syncthreads();
    const int k = (int)(unsigned int)sTicket;
    // a store into the packed master columns and, when asked for, into their host mirror
#define MC_PUT(lhs, val)                                                                                          \
    do {                                                                                                          \
        auto *p__ = &(lhs);                                                                                       \
        const auto v__ = (val);                                                                                   \
        *p__ = v__;                                                                                               \
        if (a.hostOff != 0) *reinterpret_cast<decltype(p__)>(reinterpret_cast<char *>(p__) + a.hostOff) = v__;    \
    } while (0)
    // every CTA of the launch reports here when its stores are out; the last one rings the host
    auto finish = [&]() {
        if (a.doneFlag == nullptr) return;
        __syncthreads();
        if (t == 0) {
            __threadfence_system();
            if (atomicAdd(a.doneCtr, 1u) == gridDim.x - 1u) {
                *a.doneCtr = 0u;
                __threadfence();
                // (the error flags are OR-ed into the device header by atomics; the host's copy gets their final value here)
                if (a.hostOff != 0)
                    *reinterpret_cast<volatile unsigned int *>(reinterpret_cast<char *>(&a.outHdr->flags) + a.hostOff) =
                        *reinterpret_cast<volatile unsigned int *>(&a.outHdr->flags);
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int *>(a.doneFlag) = a.doneVal;
            }
        }
    };
    const unsigned long long epochTag = (1ull << 23) | ((sTicket >> 32) & 0x7fffffull);  // valid bit + epoch
    const int nKept = a.hdr->nKept;
    if (k >= nKept) {
        // (launched with one CTA per block behind a round whose outcome the host does not know yet)
        if (nKept == 0 && k == 0 && t == 0) {
            MC_PUT(a.outColStart[0], (int64_t)0);
            MC_PUT(a.outHdr->magic, kMColsMagic);
            MC_PUT(a.outHdr->nCols, 0);
            MC_PUT(a.outHdr->nNnz, (int64_t)0);
        }
        finish();
        return;
    }
    const int lb = a.keptBlock[k];
    const int b = a.firstBlock + lb;
    const int cs = a.blockColStart[b];
    const int64_t i0 = a.keptStart[k], i1 = a.keptStart[k + 1];
    const uint16_t *row16 = static_cast<const uint16_t *>(a.rowMaster);
    const int32_t *row32 = static_cast<const int32_t *>(a.rowMaster);

    for (int w = t; w < a.rowWordsPhys + a.colWords; w += kExpThreads) rowBits[w] = 0u;  // both bitmaps
    if (t == 0) sOverflow = 0;
    __syncthreads();

    // ---- 1. membership bitmap, touched rows, fingerprint
    unsigned long long h = 0ull;
    for (int64_t idx = i0 + t; idx < i1; idx += kExpThreads) {
        const int j = a.keptInd[idx];
        atomicOr(&colBits[(j - cs) >> 5], 1u << ((j - cs) & 31));
        h += mix64(((uint64_t)(uint32_t)j << 32) | (uint64_t)(idx - i0));
        for (int e = a.colStart32[j]; e < a.colStart32[j + 1]; ++e) {
            const int rm = a.row16 ? (int)row16[e] : row32[e];
            const int i = rm < a.nBaseRows ? rm : rm - a.numConvex;  // master index -> core row
            atomicOr(&RB(i >> 5), 1u << (i & 31));
        }
    }
    // rows appended since the column-major copy was built: every one of them is walked (step 3 finds the ones the
    // column does not touch at exactly +0, which step 4 drops like any other zero)
    for (int i = a.deltaRow0 + t; i < a.R; i += kExpThreads) atomicOr(&RB(i >> 5), 1u << (i & 31));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) h += __shfl_xor_sync(0xffffffffu, h, off);
    if (lane == 0) sHashW[warp] = h;
    __syncthreads();

    // ---- 2. touched rows in ascending order: thread t owns a contiguous run of bitmap words
    const int wpt = (a.rowWords + kExpThreads - 1) / kExpThreads;
    const int w0 = min(a.rowWords, t * wpt), w1 = min(a.rowWords, w0 + wpt);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(RB(w));
    int nTouched;
    int pos = blockExclusiveScan(cnt, sWarp, &nTouched);
    if (nTouched > a.capStage) {
        if (t == 0) sOverflow = 1;
    } else {
        for (int w = w0; w < w1; ++w)
            for (uint32_t bits = RB(w); bits; bits &= bits - 1) sRows[pos++] = (w << 5) + (__ffs(bits) - 1);
    }
    __syncthreads();
    const bool overflow = sOverflow != 0;

    // ---- 3. the touched rows' dot products with s, in storage order.  Short rows (GAP: the branch rows, one or
    // two entries) take one thread each; a long row (an assignment row: one entry per block) takes a warp --
    // 32 membership tests at a time, the member values added in entry order -- instead of one thread walking
    // it alone while the others wait.
    constexpr int kShortRow = 8;
    if (!overflow) {
        auto member = [&](int e) -> bool {
            const int jl = a.csrColInd[e] - cs;
            return jl >= 0 && (jl >> 5) < a.colWords && ((colBits[jl >> 5] >> (jl & 31)) & 1u);
        };
        for (int p = t; p < nTouched; p += kExpThreads) {
            const int i = sRows[p];
            const int e0 = a.csrRowStart[i], e1 = a.csrRowStart[i + 1];
            if (e1 - e0 > kShortRow) continue;
            double y = 0.0;
            for (int e = e0; e < e1; ++e)
                if (member(e)) y = __dadd_rn(y, a.csrVal[e]);  // val * 1.0
            sVals[p] = y;
        }
        for (int p = warp; p < nTouched; p += kExpThreads / 32) {
            const int i = sRows[p];
            const int e0 = a.csrRowStart[i], e1 = a.csrRowStart[i + 1];
            if (e1 - e0 <= kShortRow) continue;
            double y = 0.0;
            // 128 entries per pass: the four column-index loads of a lane are independent (one round trip), the
            // member values follow (a second one), then the ballots in entry order
            for (int eb = e0; eb < e1; eb += 128) {
                int jl[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = eb + 32 * q + lane;
                    jl[q] = e < e1 ? a.csrColInd[e] - cs : -1;
                }
                bool mem[4];
                double v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    mem[q] = jl[q] >= 0 && (jl[q] >> 5) < a.colWords && ((colBits[jl[q] >> 5] >> (jl[q] & 31)) & 1u);
                    v[q] = mem[q] ? a.csrVal[eb + 32 * q + lane] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    unsigned m = __ballot_sync(0xffffffffu, mem[q]);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        y = __dadd_rn(y, __shfl_sync(0xffffffffu, v[q], src));
                    }
                }
            }
            if (lane == 0) sVals[p] = y;
        }
    }
    __syncthreads();

    // ---- 4. |v| > TolZero, ascending master row, the convexity 1.0 in its place
    // thread t owns the contiguous run [p0, p1) of touched rows
    const int ppt = (nTouched + kExpThreads - 1) / kExpThreads;
    const int p0 = overflow ? 0 : min(nTouched, t * ppt), p1 = overflow ? 0 : min(nTouched, p0 + ppt);
    int nk = 0, nkBelow = 0;
    for (int p = p0; p < p1; ++p) {
        const bool keep = fabs(sVals[p]) > a.tolZero;
        nk += keep ? 1 : 0;
        nkBelow += (keep && sRows[p] < a.nBaseRows) ? 1 : 0;
    }
    int totKeep, totBelow;
    const int posKeep = blockExclusiveScan(nk, sWarp, &totKeep);
    (void)blockExclusiveScan(nkBelow, sWarp, &totBelow);
    const unsigned long long len = overflow ? 0ull : (unsigned long long)totKeep + 1ull;

    // ---- 5. offset in the packed output: publish the length, poll the lower columns
    if (t == 0) {
        const unsigned long long word = (epochTag << 40) | len;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(a.pubWord + k), "l"(word) : "memory");
    }
    unsigned long long off = 0ull;
    for (int c = t; c < k; c += kExpThreads) {
        unsigned long long wv;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(wv) : "l"(a.pubWord + c) : "memory");
        } while ((wv >> 40) != epochTag);
        off += wv & 0xffffffffffull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
    __syncthreads();
    if (lane == 0) sOffW[warp] = off;
    __syncthreads();
    off = 0ull;
    for (int w = 0; w < kExpThreads / 32; ++w) off += sOffW[w];
    const bool fits = (long long)(off + len) <= a.capEntries;

    if (!overflow && fits) {
        int q = posKeep;
        for (int p = p0; p < p1; ++p) {
            const double v = sVals[p];
            if (fabs(v) > a.tolZero) {
                const int i = sRows[p];
                const bool cut = i >= a.nBaseRows;
                MColEntry e;
                e.row = cut ? i + a.numConvex : i;  // DecompAlgo.cpp:5532-5534
                e.pad = 0;
                e.val = v;
                MC_PUT(a.outEntries[off + q + (cut ? 1 : 0)], e);
                ++q;
            }
        }
        if (t == 0) {
            MColEntry e;
            e.row = a.nBaseRows + b;  // the block's convexity row, DecompAlgo.cpp:5540
            e.pad = 0;
            e.val = 1.0;
            MC_PUT(a.outEntries[off + totBelow], e);
        }
    }
    if (t == 0) {
        unsigned long long hh = 0ull;
        for (int w = 0; w < kExpThreads / 32; ++w) hh += sHashW[w];
        MC_PUT(a.outHash[k], mix64(hh ^ ((uint64_t)(uint32_t)b << 40) ^ (uint64_t)(i1 - i0)));
        MC_PUT(a.outColStart[k], (int64_t)off);
        if (overflow || !fits) {
            atomicOr(&a.outHdr->flags, overflow ? 1u : 2u);
        }
        if (k == nKept - 1) {
            MC_PUT(a.outColStart[nKept], (int64_t)(off + len));
            MC_PUT(a.outHdr->magic, kMColsMagic);
            MC_PUT(a.outHdr->nCols, nKept);
            MC_PUT(a.outHdr->nNnz, (int64_t)(off + len));
        }
    }
    finish();
#undef MC_PUT
}

int launchExpandColumns(Ctx *c, double tolZero, int nKept, cudaStream_t st)
{
    if (nKept <= 0) return DIPGPU_OK;
    char *res = static_cast<char *>(c->dResult);
    char *mc = static_cast<char *>(c->dMCols);
    ExpandArgs a;
    a.hdr = reinterpret_cast<const ResultHeader *>(res);
    a.keptStart = reinterpret_cast<const int64_t *>(res + c->layout.offKeptStart);
    a.keptBlock = reinterpret_cast<const int32_t *>(res + c->layout.offKeptBlock);
    a.keptInd = reinterpret_cast<const int32_t *>(res + c->layout.offInd);
    a.blockColStart = c->dBlockColStart;
    a.firstBlock = c->firstBlock;
    a.colStart32 = c->dColStart32;
    a.row16 = c->dRowMaster16 != nullptr ? 1 : 0;
    a.rowMaster = a.row16 ? static_cast<const void *>(c->dRowMaster16) : static_cast<const void *>(c->dRowMaster);
    a.csrRowStart = c->dCsrRowStart;
    a.csrColInd = c->dCsrColInd;
    a.csrVal = c->dCsrVal;
    a.R = c->R;
    a.deltaRow0 = c->baseR;
    a.nBaseRows = c->nBaseRows;
    a.numConvex = c->numConvex;
    a.tolZero = tolZero;
    a.outHdr = reinterpret_cast<MColsHeader *>(mc);
    a.outColStart = reinterpret_cast<int64_t *>(mc + c->mcolsOffColStart);
    a.outHash = reinterpret_cast<uint64_t *>(mc + c->mcolsOffHash);
    a.outEntries = reinterpret_cast<MColEntry *>(mc + c->mcolsOffEntries);
    a.capEntries = c->mcolsCapEntries;
    a.pubWord = c->dExpPub;
    a.ticket = c->dExpTicket;
    a.hostOff = 0;
    a.doneCtr = c->dSyncWords + 6;
    a.doneFlag = nullptr;
    a.doneVal = 0u;
    c->expandFlagArmed = false;
    if (c->optFlagWait && c->optZeroCopy && c->wantExpandMirror && !c->stageTiming && c->hMCols != nullptr && c->dMColsHost != nullptr) {
        if (!c->hExpFlag) {
            void *hp = nullptr, *dp = nullptr;
            DIPGPU_CUDA(c, cudaHostAlloc(&hp, 64, cudaHostAllocMapped | cudaHostAllocPortable));
            memset(hp, 0, 64);
            DIPGPU_CUDA(c, cudaHostGetDevicePointer(&dp, hp, 0));
            c->hExpFlag = static_cast<unsigned int *>(hp);
            c->dExpFlagHost = static_cast<unsigned int *>(dp);
        }
        if (++c->expSerialFlag == 0u) ++c->expSerialFlag;
        a.hostOff = (long long)(static_cast<char *>(c->dMColsHost) - static_cast<char *>(c->dMCols));
        a.doneFlag = c->dExpFlagHost;
        a.doneVal = c->expSerialFlag;
        c->expandFlagArmed = true;
    }
    a.rowWords = (c->R + 31) / 32;
    a.rowWordsPhys = a.rowWords + a.rowWords / 32 + 1;
    a.colWords = (c->maxBlockCols + 31) / 32;
    // shared memory: two bitmaps + staged (row, value) pairs
    const size_t maxSmem = 227 * 1024 - 1024;
    const size_t bitmapBytes = sizeof(uint32_t) * (((size_t)a.rowWordsPhys + (size_t)a.colWords + 3) & ~(size_t)3);
    if (bitmapBytes + 4096 > maxSmem)
        return fail(c, DIPGPU_ERR_UNSUPPORTED, "expand_columns: %d core rows do not fit the row bitmap", c->R);
    int cap = (int)((maxSmem - bitmapBytes) / 12) & ~1;
    cap = std::min(cap, std::max((c->expMaxTouched + 1) & ~1, 2));
    a.capStage = cap;
    const size_t bytes = bitmapBytes + sizeof(int32_t) * (size_t)((cap + 1) & ~1) + sizeof(double) * (size_t)cap;
    {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(expand_columns_kernel), bytes);
        if (rc) return rc;
    }
    DIPGPU_CUDA(c, cudaMemsetAsync(&a.outHdr->flags, 0, sizeof(uint32_t), st));
    expand_columns_kernel<<<nKept, kExpThreads, bytes, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

// multi.cu -- one libdipgpu context over several GPUs of one box (dipgpu_create_multi).
//
// DIP is ONE host process: DecompAlgo::generateVars (Dip/src/DecompAlgo.cpp:4622-5260) loops over the
// independent block subproblems -- in parallel under OpenMP when SubProbParallel is set (:4815-4848) -- and
// DecompAlgoPC prices whole ladders of smoothed dual vectors (DecompAlgo.cpp:5642-5655,
// DecompAlgoPC.cpp:126-212).  Both are what this file spreads over the GPUs, behind the same C ABI and the
// same caller thread:
//
//   * ONE dual vector (dipgpu_price_round): the BLOCKS are partitioned, device d prices the contiguous
//     block range shard[d] -- reduced costs of its own columns only (the SpMV is column-partitioned: no
//     reduction), its blocks' knapsacks, its part of the filter;
//   * SEVERAL dual vectors (dipgpu_price_rounds_batch, the stabilisation ladder): the VECTORS are
//     partitioned, device d prices all blocks for its vectors on full[d] and returns complete rounds.
//
// Exchange, per call: every device takes the duals it needs straight from the caller's buffer -- u[support]
// gathered once into mapped pinned staging (or picked out of a pinned `u` by the device itself): 34 KB on
// GAP-E, never the dense 2 MB -- and the fused DP kernel of every device stores its part of the packed
// result straight into mapped host memory.  No collective, no gather, no second hop through a "rank 0" GPU.
// The device-resident entry points (dipgpu_price_round_dev) keep the exchange on NVLink instead: the shards
// pull u[support] out of the primary device's HBM with peer loads and mirror their packed results into the
// primary's gather buffer with peer stores.
//
// One worker thread per additional device (the caller's thread drives device 0) keeps the launches of the
// devices side by side: a pricing round is tens of microseconds, three launches per device issued one after
// the other from a single thread would serialise the GPUs.  Workers spin briefly after a call, then sleep.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>

#include "common.cuh"

namespace dipgpu {

namespace {

inline void cpuRelax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> posted{0}, finished{0};
    std::atomic<bool> quit{false};
    std::function<int(int)> *job = nullptr;
    int rc = 0;
    int slot = 0;    // index of the device in the context
    int device = 0;  // CUDA ordinal

    void loop()
    {
        cudaSetDevice(device);
        uint64_t seen = 0;
        for (;;) {
            int spins = 0;
            while (posted.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_acquire)) {
                if (++spins < 40000) {
                    cpuRelax();
                } else {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait_for(lk, std::chrono::microseconds(500), [&] {
                        return posted.load(std::memory_order_acquire) != seen || quit.load(std::memory_order_acquire);
                    });
                }
            }
            if (quit.load(std::memory_order_acquire)) return;
            rc = (*job)(slot);
            ++seen;
            finished.store(seen, std::memory_order_release);
        }
    }
};

}  // namespace

struct Multi {
    Ctx *front = nullptr;
    int ndev = 0;
    std::vector<int> dev;
    std::vector<Ctx *> full;   // all blocks: vector-partitioned batches; full[0] serves every call that is not partitioned
    std::vector<Ctx *> shard;  // block range of device d: the single round (== full when ndev == 1)
    std::vector<Worker *> workers;  // [1, ndev)
    uint64_t posted = 0;
    std::vector<int> shardFirst, shardCount;
    // devices that take part in a SINGLE round (block partition).  Option "shard_rounds": -1 (default) = all of them
    // unless one GPU runs every block side by side anyway (a fused shape whose blocks are co-resident on the primary:
    // each block is one CTA-long latency chain, a partition cannot shorten it and pays the exchange) -- then 1;
    // k >= 1 forces min(k, ndev).  Batches of dual vectors always use every device.
    int nRound = 1;
    int optShardRounds = -1;
    bool peerOK = false;  // the primary device and every other device can address each other's memory
    // duals staged once for all devices (mapped, portable pinned memory): vector v at hStage + v * support.size()
    std::vector<int32_t> support;
    double *hStage = nullptr;
    size_t stageCap = 0;  // doubles
    // vector -> (device, local index) of the last batched / ladder call
    std::vector<int> vecDev, vecLocal;
    double lastAlphas[64] = {};
    int lastSteps = 0;
    // which context holds "the last round" (dipgpu_expand_columns): NULL = the shards of the last single round
    Ctx *lastCtx = nullptr;
    bool shardedRoundValid = false;
    // device-resident path: the shards' packed results, one after the other, in the primary device's memory
    void *dGather = nullptr;
    size_t gatherBytes = 0;
    std::vector<size_t> gatherOff;
    std::atomic<Ctx *> errFrom{nullptr};
    std::vector<float> devMs;  // device time of each device's part of the last timed call (option stage_timing)
};

static Multi *M_(dipgpu_ctx *ctx) { return reinterpret_cast<Ctx *>(ctx)->multi; }
static const Multi *M_(const dipgpu_ctx *ctx) { return reinterpret_cast<const Ctx *>(ctx)->multi; }
static dipgpu_ctx *H_(Ctx *c) { return reinterpret_cast<dipgpu_ctx *>(c); }

// fn(d) on every device: devices 1.. on their workers, device 0 on the calling thread.  First non-zero code wins;
// the message of the context that failed becomes the front's.
static int runAll(Multi *M, std::function<int(int)> fn, int nUse = -1)
{
    const int n = nUse < 0 ? M->ndev : std::min(nUse, M->ndev);
    M->errFrom.store(nullptr);
    ++M->posted;
    for (int d = 1; d < n; ++d) {
        Worker *w = M->workers[(size_t)d];
        w->job = &fn;
        w->posted.store(M->posted, std::memory_order_release);
        w->cv.notify_one();
    }
    int rc = 0;
    if (n > 0) {
        cudaSetDevice(M->dev[0]);
        rc = fn(0);
    }
    for (int d = 1; d < n; ++d) {
        Worker *w = M->workers[(size_t)d];
        int spins = 0;
        while (w->finished.load(std::memory_order_acquire) != M->posted) {
            if (++spins < 200000) cpuRelax();
            else std::this_thread::yield();
        }
        if (!rc && w->rc) rc = w->rc;
    }
    // workers that sat this call out still have to see the new sequence number
    for (int d = std::max(n, 1); d < M->ndev; ++d) {
        Worker *w = M->workers[(size_t)d];
        static std::function<int(int)> nop = [](int) { return 0; };
        w->job = &nop;
        w->posted.store(M->posted, std::memory_order_release);
        w->cv.notify_one();
    }
    for (int d = std::max(n, 1); d < M->ndev; ++d) {
        Worker *w = M->workers[(size_t)d];
        while (w->finished.load(std::memory_order_acquire) != M->posted) cpuRelax();
    }
    if (rc) {
        Ctx *e = M->errFrom.load();
        if (e && e != M->front) M->front->err = e->err;
    }
    return rc;
}

static int subFail(Multi *M, Ctx *c, int rc)
{
    if (rc) {
        Ctx *expect = nullptr;
        M->errFrom.compare_exchange_strong(expect, c);
    }
    return rc;
}

static void destroyAll(Multi *M)
{
    for (Worker *w : M->workers) {
        if (!w) continue;
        w->quit.store(true, std::memory_order_release);
        w->cv.notify_one();
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    for (int d = 0; d < (int)M->full.size(); ++d) {
        if (d < (int)M->shard.size() && M->shard[(size_t)d] && M->shard[(size_t)d] != M->full[(size_t)d]) dipgpu_destroy(H_(M->shard[(size_t)d]));
        if (M->full[(size_t)d]) dipgpu_destroy(H_(M->full[(size_t)d]));
    }
    if (M->hStage) cudaFreeHost(M->hStage);
    if (M->dGather && !M->dev.empty()) {
        cudaSetDevice(M->dev[0]);
        cudaFree(M->dGather);
    }
    delete M->front;
    delete M;
}

int multiDestroy(dipgpu_ctx *ctx)
{
    destroyAll(M_(ctx));
    return DIPGPU_OK;
}

dipgpu_ctx *multiPrimary(dipgpu_ctx *ctx) { return H_(M_(ctx)->full[0]); }

dipgpu_ctx *multiForward(dipgpu_ctx *ctx)
{
    Multi *M = M_(ctx);
    M->lastCtx = M->full[0];
    M->shardedRoundValid = false;
    M->front->err.clear();  // dipgpu_last_error then reports the primary's message
    return H_(M->full[0]);
}

int multiForAll(dipgpu_ctx *ctx, int (*fn)(dipgpu_ctx *sub, void *arg), void *arg, bool alsoShards, bool modelChanged)
{
    Multi *M = M_(ctx);
    if (modelChanged) {
        M->shardedRoundValid = false;
        M->lastCtx = nullptr;
    }
    return runAll(M, [&](int d) {
        int rc = subFail(M, M->full[(size_t)d], fn(H_(M->full[(size_t)d]), arg));
        if (rc || !alsoShards) return rc;
        if (M->shard[(size_t)d] != M->full[(size_t)d]) rc = subFail(M, M->shard[(size_t)d], fn(H_(M->shard[(size_t)d]), arg));
        return rc;
    });
}

void multiClearSupport(dipgpu_ctx *ctx) { M_(ctx)->support.clear(); }
int multiSetDualSupport(dipgpu_ctx *ctx, int32_t nIdx, const int32_t *idx);

// Contiguous block ranges balanced by the DP work n_b * (cap_b + 1), and the columns each shard's
// reduced-cost sweep covers: its blocks' columns, the first shard also everything in front of block 0, the
// last one everything behind the last block.
int multiSetBlocks(dipgpu_ctx *ctx)
{
    Multi *M = M_(ctx);
    Ctx *f0 = M->full[0];
    const int m = f0->m;
    int n = M->ndev;
    if (M->optShardRounds >= 1) n = std::min(M->optShardRounds, M->ndev);
    else if (f0->blockKind == 0 && f0->geom.fuse && f0->dpCoResident > 0 && m <= f0->dpCoResident) n = 1;
    M->nRound = n;
    M->shardFirst.assign((size_t)M->ndev, m);
    M->shardCount.assign((size_t)M->ndev, 0);
    std::vector<double> pre((size_t)m + 1, 0.0);
    for (int b = 0; b < m; ++b)
        pre[(size_t)b + 1] = pre[(size_t)b] + (double)std::max(f0->hBlockColStart[(size_t)b + 1] - f0->hBlockColStart[(size_t)b], 1) *
                                                  ((double)f0->hCapacity[(size_t)b] + 1.0);
    std::vector<int> cut((size_t)n + 1, 0);
    for (int r = 1; r < n; ++r) {
        const double want = pre[(size_t)m] * r / n;
        cut[(size_t)r] = (int)(std::lower_bound(pre.begin(), pre.end(), want) - pre.begin());
        cut[(size_t)r] = std::max(cut[(size_t)r], cut[(size_t)r - 1]);
        cut[(size_t)r] = std::min(cut[(size_t)r], m);
    }
    cut[(size_t)n] = m;
    for (int d = 0; d < n; ++d) {
        M->shardFirst[(size_t)d] = cut[(size_t)d];
        M->shardCount[(size_t)d] = cut[(size_t)d + 1] - cut[(size_t)d];
    }
    if (M->ndev == 1) return DIPGPU_OK;
    return runAll(M, [&](int d) {
        Ctx *s = M->shard[(size_t)d];
        int rc = bind(s);
        if (!rc) rc = setBlockRange(s, M->shardFirst[(size_t)d], M->shardCount[(size_t)d]);
        if (!rc) {
            s->shardCols = n > 1;
            s->tileColFirst = d == 0 ? 0 : s->hBlockColStart[(size_t)M->shardFirst[(size_t)d]];
            s->tileColEnd = d >= n - 1 ? -1 : s->hBlockColStart[(size_t)(M->shardFirst[(size_t)d] + M->shardCount[(size_t)d])];
            s->tilesDirty = true;
        }
        return subFail(M, s, rc);
    });
}

// option "shard_rounds" (see Multi::nRound): re-partitions the blocks when they are already set
int multiSetShardRounds(dipgpu_ctx *ctx, int64_t value)
{
    Multi *M = M_(ctx);
    M->optShardRounds = value < 1 ? -1 : (int)std::min<int64_t>(value, 64);
    M->lastCtx = nullptr;
    M->shardedRoundValid = false;
    if (!M->full[0]->haveBlocks) return DIPGPU_OK;
    const int rc = multiSetBlocks(ctx);
    if (rc) return rc;
    if (!M->support.empty()) {
        // (the shards' compact structures follow the block ranges)
        std::vector<int32_t> sup = M->support;
        return multiSetDualSupport(ctx, (int32_t)sup.size(), sup.data());
    }
    return DIPGPU_OK;
}

int multiSetDualSupport(dipgpu_ctx *ctx, int32_t nIdx, const int32_t *idx)
{
    Multi *M = M_(ctx);
    struct A { int32_t n; const int32_t *idx; } a{nIdx, idx};
    int rc = multiForAll(ctx, [](dipgpu_ctx *sub, void *arg) {
        const A *q = static_cast<const A *>(arg);
        return dipgpu_set_dual_support(sub, q->n, q->idx);
    }, &a, true, false);
    if (rc) return rc;
    M->support.assign(idx, idx + nIdx);
    return DIPGPU_OK;
}

static int ensureStage(Multi *M, size_t doubles)
{
    if (doubles <= M->stageCap) return DIPGPU_OK;
    if (M->hStage) cudaFreeHost(M->hStage);
    M->hStage = nullptr;
    M->stageCap = 0;
    const size_t cap = std::max<size_t>(doubles, 4096);
    if (cudaHostAlloc(reinterpret_cast<void **>(&M->hStage), sizeof(double) * cap, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(M->front, DIPGPU_ERR_NOMEM, "pinned staging for %zu dual entries", cap);
    }
    M->stageCap = cap;
    return DIPGPU_OK;
}

// The master duals of a call, prepared ONCE for all devices: a page-locked vector is left where it is (every
// device reads u[support] out of it), a pageable one is gathered into shared pinned staging (u[support], 34 KB
// per vector on GAP-E) that every device reads; without a dual support every device copies the whole vector.
// gatherHere: false = the caller of this function gathers (per device, in parallel) -- only the feed is formed.
static int stageDuals(Multi *M, const double *u, int32_t uLen, int nVec, bool gatherHere, DualFeed *f)
{
    *f = DualFeed();
    const size_t ns = M->support.size();
    if (ns == 0) {
        f->host = u;
        return DIPGPU_OK;
    }
    if (M->front->optGatherPinned && deviceVisible(u)) {
        f->host = u;  // page-locked: every device picks u[support] out of the caller's buffer itself
        return DIPGPU_OK;
    }
    const int rc = ensureStage(M, ns * (size_t)nVec);
    if (rc) return rc;
    if (gatherHere) gatherSupport(M->support, u, uLen, nVec, M->hStage);
    f->staged = M->hStage;  // (unified addressing: a mapped, portable allocation has one address everywhere)
    f->stagedStride = (int64_t)ns;
    return DIPGPU_OK;
}

// Decodes the shards' packed results (host copies, ascending block ranges) into ONE round: per-block arrays at
// their global positions, the kept columns one shard after the other (= block order), mostNegReducedCost added
// in block order like DecompAlgo.cpp:5229-5232.
static int unpackShards(Ctx *front, const void *const *packed, const size_t *bytes, int nShards, dipgpu_cols *out)
{
    int nKept = 0;
    int64_t nInd = 0, cells = 0, evals = 0;
    double sum = 0.0;
    int needBlocks = 0;
    int wpe = 1;  // words per kept entry: 2 = (index, multiplicity) pairs (integer knapsack blocks)
    struct Part { const char *p; const ResultHeader *h; ResultLayout L; };
    std::vector<Part> parts((size_t)nShards);
    for (int s = 0; s < nShards; ++s) {
        const char *p = static_cast<const char *>(packed[s]);
        const ResultHeader *h = reinterpret_cast<const ResultHeader *>(p);
        if (!p || bytes[s] < sizeof(ResultHeader) || h->magic != kResultMagic) return fail(front, DIPGPU_ERR_INVALID, "packed result of shard %d: bad header", s);
        parts[(size_t)s] = Part{p, h, ResultLayout::make(h->m, h->indCap)};
        if (bytes[s] < parts[(size_t)s].L.offInd + sizeof(int32_t) * (size_t)h->nInd) return fail(front, DIPGPU_ERR_INVALID, "packed result of shard %d truncated", s);
        if (h->reserved[1] == 1) wpe = 2;
        nKept += h->nKept;
        nInd += h->nInd;
        cells += h->cells;
        evals += h->reserved[0];
        needBlocks = std::max(needBlocks, h->firstBlock + h->m);
    }
    nInd /= wpe;
    out->nCols = nKept;
    out->nInd = nInd;
    out->cells = cells;
    out->cellsEvaluated = evals;
    const bool wantBlocks = out->blockRedCost || out->blockMostNegRC || out->blockOrigCost || out->blockObj || out->blockNnz;
    if (wantBlocks && out->capBlocks < needBlocks) return fail(front, DIPGPU_ERR_OVERFLOW, "capBlocks %d < %d", out->capBlocks, needBlocks);
    for (const Part &q : parts) {
        const double *rc = reinterpret_cast<const double *>(q.p + q.L.offRedCost);
        const double *mn = reinterpret_cast<const double *>(q.p + q.L.offMostNeg);
        const double *oc = reinterpret_cast<const double *>(q.p + q.L.offOrigCost);
        const double *ob = reinterpret_cast<const double *>(q.p + q.L.offObj);
        const int32_t *nz = reinterpret_cast<const int32_t *>(q.p + q.L.offNnz);
        for (int lb = 0; lb < q.h->m; ++lb) {
            const int b = q.h->firstBlock + lb;
            if (out->blockRedCost) out->blockRedCost[b] = rc[lb];
            if (out->blockMostNegRC) out->blockMostNegRC[b] = mn[lb];
            if (out->blockOrigCost) out->blockOrigCost[b] = oc[lb];
            if (out->blockObj) out->blockObj[b] = ob[lb];
            if (out->blockNnz) out->blockNnz[b] = nz[lb] / wpe;
            sum += mn[lb];
        }
    }
    out->mostNegReducedCost = sum;
    if (nKept > out->capCols || nInd > out->capInd)
        return fail(front, DIPGPU_ERR_OVERFLOW, "kept columns need capCols >= %d, capInd >= %lld", nKept, (long long)nInd);
    int k0 = 0;
    int64_t i0 = 0;
    for (const Part &q : parts) {
        const double *rc = reinterpret_cast<const double *>(q.p + q.L.offRedCost);
        const double *oc = reinterpret_cast<const double *>(q.p + q.L.offOrigCost);
        const int64_t *ks = reinterpret_cast<const int64_t *>(q.p + q.L.offKeptStart);
        const int32_t *kb = reinterpret_cast<const int32_t *>(q.p + q.L.offKeptBlock);
        const int32_t *ki = reinterpret_cast<const int32_t *>(q.p + q.L.offInd);
        for (int k = 0; k < q.h->nKept; ++k) {
            const int lb = kb[k];
            if (out->colBlock) out->colBlock[k0 + k] = q.h->firstBlock + lb;
            if (out->colRedCost) out->colRedCost[k0 + k] = rc[lb];
            if (out->colOrigCost) out->colOrigCost[k0 + k] = oc[lb];
            if (out->colStart) out->colStart[k0 + k] = i0 + ks[k] / wpe;
        }
        unpackEntries(ki, q.h->nInd, wpe, out->colInd ? out->colInd + i0 : nullptr, out->colEl ? out->colEl + i0 : nullptr);
        k0 += q.h->nKept;
        i0 += q.h->nInd / wpe;
    }
    if (out->colStart) out->colStart[nKept] = nInd;
    return DIPGPU_OK;
}

static void collectDevMs(Multi *M, const std::vector<Ctx *> &ctxs, int n)
{
    if (!M->front->stageTiming) return;
    for (int d = 0; d < M->ndev; ++d) M->devMs[(size_t)d] = 0.f;
    for (int k = 0; k < 6; ++k) M->front->stageMs[k] = 0.f;
    for (int d = 0; d < n; ++d) {
        float tot = 0.f;
        for (int k = 0; k < 6; ++k) {
            tot += ctxs[(size_t)d]->stageMs[k];
            M->front->stageMs[k] = std::max(M->front->stageMs[k], ctxs[(size_t)d]->stageMs[k]);  // per stage: the slowest device
        }
        M->devMs[(size_t)d] = tot;
    }
}

int multiPriceRound(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double eps, dipgpu_cols *out, double *redCostXOpt)
{
    Multi *M = M_(ctx);
    Ctx *front = M->front;
    if (!out) return fail(front, DIPGPU_ERR_INVALID, "price_round: NULL output");
    if (!u || M->nRound == 1) {
        // the dual vector already on the primary device (left there by the pool re-pricing, which is not partitioned)
        M->lastCtx = M->full[0];
        M->shardedRoundValid = false;
        const int rc = dipgpu_price_round(H_(M->full[0]), u, uLen, phase, eps, out, redCostXOpt);
        if (rc) front->err = M->full[0]->err;
        if (!rc && front->stageTiming) {
            std::vector<Ctx *> one(1, M->full[0]);
            collectDevMs(M, one, 1);
        }
        return rc;
    }
    DualFeed f;
    int rc = stageDuals(M, u, uLen, 1, true, &f);
    if (rc) return rc;
    rc = runAll(M, [&](int d) {
        Ctx *s = M->shard[(size_t)d];
        int r = bind(s);
        if (!r) r = checkRoundReady(s, "price_round", uLen, phase);
        if (r) return subFail(M, s, r);
        cudaStream_t st = s->stream;
        if ((r = enqueueRound(s, &f, uLen, phase, eps, true, st))) return subFail(M, s, r);
        if (redCostXOpt) {
            const int c0 = s->tileColFirst, c1 = s->tileColEnd < 0 ? s->N : s->tileColEnd;
            if (c1 > c0 && cudaMemcpyAsync(redCostXOpt + c0, s->dRedCost + c0, sizeof(double) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return subFail(M, s, fail(s, DIPGPU_ERR_CUDA, "price_round: copy of the reduced costs"));
            s->d2h += (int64_t)sizeof(double) * (c1 - c0);
        }
        if ((r = fetchResult(s, st))) return subFail(M, s, r);
        s->lastRoundOnHost = true;
        collectStageMs(s);
        return 0;
    }, M->nRound);
    if (rc) return rc;
    collectDevMs(M, M->shard, M->nRound);
    M->lastCtx = nullptr;
    M->shardedRoundValid = true;
    const void *packed[64];
    size_t bytes[64];
    for (int d = 0; d < M->nRound; ++d) {
        packed[d] = M->shard[(size_t)d]->hResult;
        bytes[d] = M->shard[(size_t)d]->resultBytes;
    }
    return unpackShards(front, packed, bytes, M->nRound, out);
}

// vectors [v0[d], v0[d+1]) go to device d: as even as possible over the first min(nVec, ndev) devices
static int splitVectors(Multi *M, int nVec, std::vector<int> &v0)
{
    const int nUse = std::max(1, std::min(nVec, M->ndev));
    v0.assign((size_t)nUse + 1, 0);
    for (int d = 0; d <= nUse; ++d) v0[(size_t)d] = (int)((int64_t)nVec * d / nUse);
    M->vecDev.assign((size_t)nVec, 0);
    M->vecLocal.assign((size_t)nVec, 0);
    for (int d = 0; d < nUse; ++d)
        for (int v = v0[(size_t)d]; v < v0[(size_t)d + 1]; ++v) {
            M->vecDev[(size_t)v] = d;
            M->vecLocal[(size_t)v] = v - v0[(size_t)d];
        }
    return nUse;
}

int multiPriceRoundsBatch(dipgpu_ctx *ctx, int32_t nVec, const double *U, int32_t uLen, int32_t phase, double eps, dipgpu_cols *outs)
{
    Multi *M = M_(ctx);
    Ctx *front = M->front;
    if (nVec < 0 || (nVec > 0 && (!U || !outs))) return fail(front, DIPGPU_ERR_INVALID, "price_rounds_batch: bad arguments");
    if (nVec == 0) return DIPGPU_OK;
    std::vector<int> v0;
    const int nUse = splitVectors(M, nVec, v0);
    DualFeed f;
    int rc = stageDuals(M, U, uLen, nVec, false, &f);
    if (rc) return rc;
    M->lastSteps = 0;
    M->lastCtx = nullptr;
    M->shardedRoundValid = false;
    const size_t ns = M->support.size();
    rc = runAll(M, [&](int d) {
        Ctx *c = M->full[(size_t)d];
        const int a = v0[(size_t)d], n = v0[(size_t)d + 1] - a;
        int r = bind(c);
        if (r) return subFail(M, c, r);
        DualFeed fd = f;  // this device's vectors
        if (fd.host) fd.host += (size_t)a * (size_t)uLen;
        if (fd.full) fd.full += (size_t)a * (size_t)fd.fullStride;
        if (fd.staged) {
            // pageable duals: every device gathers ITS vectors into its part of the shared staging, side by side
            double *mine = M->hStage + (size_t)a * ns;
            gatherSupport(M->support, U + (size_t)a * (size_t)uLen, uLen, n, mine);
            fd.staged = mine;
        }
        return subFail(M, c, batchCore(c, n, fd, uLen, phase, eps, outs + a));
    }, nUse);
    if (!rc) collectDevMs(M, M->full, nUse);
    return rc;
}

int multiStabLadder(dipgpu_ctx *ctx, const double *uRM, int32_t uLen, const double *al, int nSteps, int32_t phase, double eps, dipgpu_cols *outs)
{
    Multi *M = M_(ctx);
    std::vector<int> v0;
    const int nUse = splitVectors(M, nSteps, v0);
    DualFeed f;
    int rc = stageDuals(M, uRM, uLen, 1, true, &f);
    if (rc) return rc;
    memcpy(M->lastAlphas, al, sizeof(double) * (size_t)nSteps);
    M->lastSteps = nSteps;
    M->lastCtx = nullptr;
    M->shardedRoundValid = false;
    // EVERY device takes dualRM (and initialises its copy of the stability centre on the first call), also the
    // ones without a step of their own: dipgpu_stab_accept then moves the centre on all of them alike
    rc = runAll(M, [&](int d) {
        Ctx *c = M->full[(size_t)d];
        int r = bind(c);
        if (r) return subFail(M, c, r);
        const int a = d < nUse ? v0[(size_t)d] : 0, n = d < nUse ? v0[(size_t)d + 1] - a : 0;
        return subFail(M, c, ladderCore(c, f, uLen, al + a, n, phase, eps, outs + a));
    });
    if (!rc) collectDevMs(M, M->full, nUse);
    return rc;
}

// m_dual := m_dualST of step `step` (DecompAlgoPC::setObjBound, DecompAlgoPC.h:124-133) on EVERY device: each holds
// the centre and dualRM of the last ladder call and forms alpha*dual + (1-alpha)*dualRM itself, with the operations
// of the smoothing kernel -- the same bits the owner of the step priced with.
int multiStabAccept(dipgpu_ctx *ctx, int32_t step)
{
    Multi *M = M_(ctx);
    if (step < 0 || step >= M->lastSteps) return fail(M->front, DIPGPU_ERR_STATE, "stab_accept: no smoothed dual vector %d", step);
    const double alpha = M->lastAlphas[step];
    return runAll(M, [&](int d) {
        Ctx *c = M->full[(size_t)d];
        int r = bind(c);
        if (r) return subFail(M, c, r);
        if (!c->haveCenter || !c->dualsResident) return subFail(M, c, fail(c, DIPGPU_ERR_STATE, "stab_accept: no stability centre"));
        double *tmp = nullptr, *dAl = nullptr;
        if (cudaMalloc(reinterpret_cast<void **>(&tmp), sizeof(double) * (c->centerLen + 16)) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void **>(&dAl), sizeof(double)) != cudaSuccess) {
            if (tmp) cudaFree(tmp);
            (void)cudaGetLastError();
            return subFail(M, c, fail(c, DIPGPU_ERR_NOMEM, "stab_accept: device memory"));
        }
        cudaMemcpyAsync(dAl, &alpha, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        cudaMemsetAsync(tmp, 0, sizeof(double) * (c->centerLen + 16), c->stream);
        r = launchSmoothDuals(c, c->dU, c->dCenter, dAl, 1, (int)c->centerLen, tmp, 0, c->stream);
        const cudaError_t e = cudaStreamSynchronize(c->stream);
        cudaFree(dAl);
        if (r || e != cudaSuccess) {
            cudaFree(tmp);
            return subFail(M, c, r ? r : cudaFail(c, e, "stab_accept"));
        }
        cudaFree(c->dCenter);
        c->dCenter = tmp;
        return 0;
    });
}

int multiGetSmoothedDuals(dipgpu_ctx *ctx, int32_t step, double *dualST, int32_t uLen)
{
    Multi *M = M_(ctx);
    if (step < 0 || step >= M->lastSteps) return fail(M->front, DIPGPU_ERR_STATE, "get_smoothed_duals: no smoothed dual vector %d", step);
    Ctx *c = M->full[(size_t)M->vecDev[(size_t)step]];
    const int rc = dipgpu_get_smoothed_duals(H_(c), M->vecLocal[(size_t)step], dualST, uLen);
    if (rc) M->front->err = c->err;
    return rc;
}

int multiSelectBatchResult(dipgpu_ctx *ctx, int32_t which)
{
    Multi *M = M_(ctx);
    if (which < 0 || which >= (int)M->vecDev.size()) return fail(M->front, DIPGPU_ERR_STATE, "select_batch_result: no batched result %d", which);
    Ctx *c = M->full[(size_t)M->vecDev[(size_t)which]];
    const int rc = dipgpu_select_batch_result(H_(c), M->vecLocal[(size_t)which]);
    if (rc) M->front->err = c->err;
    else {
        M->lastCtx = c;
        M->shardedRoundValid = false;
    }
    return rc;
}

// Stages (b)+(c) for a caller-supplied reduced-cost vector, the blocks partitioned: every shard uploads the reduced
// costs of ITS blocks' columns (and the convexity duals) and returns its part of the round.
int multiSolveBlocks(dipgpu_ctx *ctx, const double *redCostX, const double *alpha, double eps, dipgpu_cols *out)
{
    Multi *M = M_(ctx);
    Ctx *front = M->front;
    if (!redCostX || !alpha || !out) return fail(front, DIPGPU_ERR_INVALID, "solve_blocks: NULL argument");
    if (M->nRound == 1 || M->shardCount.empty()) {
        M->lastCtx = M->full[0];
        const int rc = dipgpu_solve_blocks(H_(M->full[0]), redCostX, alpha, eps, out);
        if (rc) front->err = M->full[0]->err;
        return rc;
    }
    int rc = runAll(M, [&](int d) {
        Ctx *s = M->shard[(size_t)d];
        int r = bind(s);
        if (r) return subFail(M, s, r);
        if (!s->haveBlocks) return subFail(M, s, fail(s, DIPGPU_ERR_STATE, "solve_blocks before set_knapsack_blocks"));
        cudaStream_t st = s->stream;
        const bool timed = s->stageTiming;
        if (timed) cudaEventRecord(s->ev[0], st);
        const int c0 = s->hBlockColStart[(size_t)s->firstBlock], c1 = s->hBlockColStart[(size_t)(s->firstBlock + s->nLocal)];
        cudaError_t e = cudaSuccess;
        if (c1 > c0) e = cudaMemcpyAsync(s->dRedCost + c0, redCostX + c0, sizeof(double) * (size_t)(c1 - c0), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && s->m > 0) e = cudaMemcpyAsync(s->dAlpha, alpha, sizeof(double) * (size_t)s->m, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return subFail(M, s, cudaFail(s, e, "solve_blocks: upload"));
        s->h2d += (int64_t)sizeof(double) * ((int64_t)(c1 - c0) + s->m);
        if (timed) { cudaEventRecord(s->ev[1], st); cudaEventRecord(s->ev[2], st); }
        if ((r = multiEnqueueSolve(s, eps, st, timed))) return subFail(M, s, r);
        if ((r = fetchResult(s, st))) return subFail(M, s, r);
        s->lastRoundOnHost = true;
        collectStageMs(s);
        return 0;
    }, M->nRound);
    if (rc) return rc;
    collectDevMs(M, M->shard, M->nRound);
    M->lastCtx = nullptr;
    M->shardedRoundValid = true;
    const void *packed[64];
    size_t bytes[64];
    for (int d = 0; d < M->nRound; ++d) {
        packed[d] = M->shard[(size_t)d]->hResult;
        bytes[d] = M->shard[(size_t)d]->resultBytes;
    }
    return unpackShards(front, packed, bytes, M->nRound, out);
}

// Stage (a) alone, column-partitioned: every shard sweeps its own columns and copies them into the caller's vector.
int multiCalcRedCost(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double *redCostX)
{
    Multi *M = M_(ctx);
    if (!u || !redCostX) return fail(M->front, DIPGPU_ERR_INVALID, "calc_redcost: NULL argument");
    if (M->nRound == 1 || M->shardCount.empty()) {
        const int rc = dipgpu_calc_redcost(H_(M->full[0]), u, uLen, phase, redCostX);
        if (rc) M->front->err = M->full[0]->err;
        return rc;
    }
    DualFeed f;
    int rc = stageDuals(M, u, uLen, 1, true, &f);
    if (rc) return rc;
    return runAll(M, [&](int d) {
        Ctx *s = M->shard[(size_t)d];
        int r = bind(s);
        if (r) return subFail(M, s, r);
        if (!s->haveCore) return subFail(M, s, fail(s, DIPGPU_ERR_STATE, "calc_redcost before set_core_csr"));
        if (phase != DIPGPU_PHASE_PRICE1 && !s->haveObj) return subFail(M, s, fail(s, DIPGPU_ERR_STATE, "calc_redcost (phase 2) before set_objective"));
        if (uLen != s->R + s->numConvex) return subFail(M, s, fail(s, DIPGPU_ERR_INVALID, "calc_redcost: uLen=%d, expected R+numConvex=%d", uLen, s->R + s->numConvex));
        if ((r = ensureDualBuffer(s, uLen))) return subFail(M, s, r);
        cudaStream_t st = s->stream;
        if ((r = feedDuals(s, f, uLen, 1, s->dU, 0, &s->dUClean, st))) return subFail(M, s, r);
        s->dualsResident = true;
        if ((r = launchRedCost(s, s->dU, phase, st))) return subFail(M, s, r);
        const int c0 = s->tileColFirst, c1 = s->tileColEnd < 0 ? s->N : s->tileColEnd;
        cudaError_t e = cudaSuccess;
        if (c1 > c0) e = cudaMemcpyAsync(redCostX + c0, s->dRedCost + c0, sizeof(double) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        s->d2h += (int64_t)sizeof(double) * (c1 - c0);
        return e == cudaSuccess ? 0 : subFail(M, s, cudaFail(s, e, "calc_redcost"));
    }, M->nRound);
}

// ---------------------------------------------------------------- device-resident round over NVLink
//
// uDev: the master dual vector in the PRIMARY device's memory.  Every shard pulls what it needs out of it with
// peer loads (u[support]; the whole vector by a peer copy when no support is declared) and its fused DP kernel
// repeats every result store into its slice of the primary's gather buffer (peer stores): when the devices'
// streams have drained (dipgpu_sync), the primary holds the shards' packed results one after the other
// (dipgpu_result_dev / dipgpu_result_shards / dipgpu_unpack_results).  Only enqueues.
static int ensureGather(Multi *M)
{
    size_t tot = 0;
    std::vector<size_t> off((size_t)M->nRound + 1, 0);
    for (int d = 0; d < M->nRound; ++d) {
        off[(size_t)d] = tot;
        tot += ResultLayout::alignUp(M->shard[(size_t)d]->resultBytes, 256);
    }
    off[(size_t)M->nRound] = tot;
    if (M->dGather && tot == M->gatherBytes && off == M->gatherOff) return DIPGPU_OK;
    cudaSetDevice(M->dev[0]);
    if (M->dGather) cudaFree(M->dGather);
    M->dGather = nullptr;
    if (cudaMalloc(&M->dGather, std::max<size_t>(tot, 256)) != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(M->front, DIPGPU_ERR_NOMEM, "gather buffer of %zu bytes on the primary device", tot);
    }
    cudaMemset(M->dGather, 0, std::max<size_t>(tot, 256));
    M->gatherBytes = tot;
    M->gatherOff = off;
    return DIPGPU_OK;
}

int multiPriceRoundDev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase, double eps)
{
    Multi *M = M_(ctx);
    Ctx *front = M->front;
    if (!uDev) return fail(front, DIPGPU_ERR_INVALID, "price_round_dev: NULL dual vector");
    int rc = ensureGather(M);
    if (rc) return rc;
    DualFeed f;
    f.full = uDev;
    f.fullStride = uLen;
    M->lastCtx = nullptr;
    M->shardedRoundValid = false;
    rc = runAll(M, [&](int d) {
        Ctx *s = M->shard[(size_t)d];
        int r = bind(s);
        if (!r) r = checkRoundReady(s, "price_round_dev", uLen, phase);
        if (r) return subFail(M, s, r);
        cudaStream_t st = s->stream;
        char *slice = static_cast<char *>(M->dGather) + M->gatherOff[(size_t)d];
        const bool direct = s->geom.fuse && s->optZeroCopy && (d == 0 || M->peerOK);
        if (front->stageTiming) cudaEventRecord(s->ev[0], st);
        s->mirrorBase = direct ? slice : nullptr;
        r = enqueueRound(s, &f, uLen, phase, eps, direct, st);
        s->mirrorBase = nullptr;
        if (r) return subFail(M, s, r);
        if (!direct) {
            // large shards (stand-alone filter kernel) or no peer access: one copy of the packed result behind the round
            if (cudaMemcpyPeerAsync(slice, M->dev[0], s->dResult, s->device, s->resultBytes, st) != cudaSuccess)
                return subFail(M, s, fail(s, DIPGPU_ERR_CUDA, "price_round_dev: peer copy of the packed result"));
        }
        if (front->stageTiming) cudaEventRecord(s->ev[6], st);
        return 0;
    }, M->nRound);
    return rc;
}

int multiResultDev(dipgpu_ctx *ctx, void **devPtr, size_t *bytes)
{
    Multi *M = M_(ctx);
    if (!M->dGather) return fail(M->front, DIPGPU_ERR_STATE, "no result buffer yet");
    *devPtr = M->dGather;
    *bytes = M->gatherBytes;
    return DIPGPU_OK;
}

int multiGetCounters(const dipgpu_ctx *ctx, int64_t *launches, int64_t *h2d, int64_t *d2h)
{
    const Multi *M = M_(ctx);
    int64_t a = 0, b = 0, c = 0;
    for (int d = 0; d < M->ndev; ++d) {
        const Ctx *x[2] = {M->full[(size_t)d], M->shard[(size_t)d] != M->full[(size_t)d] ? M->shard[(size_t)d] : nullptr};
        for (const Ctx *q : x)
            if (q) {
                a += q->launches;
                b += q->h2d;
                c += q->d2h;
            }
    }
    if (launches) *launches = a;
    if (h2d) *h2d = b;
    if (d2h) *d2h = c;
    return DIPGPU_OK;
}

int multiLastStageMs(const dipgpu_ctx *ctx, float *ms6)
{
    const Multi *M = M_(ctx);
    for (int k = 0; k < 6; ++k) ms6[k] = M->front->stageMs[k];
    return DIPGPU_OK;
}

// Master columns of the kept columns of the last round: from the context that holds it, or -- after a round whose
// blocks were partitioned -- every shard expands ITS kept columns and the pieces are strung together in block order.
int multiExpandColumns(dipgpu_ctx *ctx, double tolZero, dipgpu_mcols *out)
{
    Multi *M = M_(ctx);
    Ctx *front = M->front;
    if (!out) return fail(front, DIPGPU_ERR_INVALID, "expand_columns: NULL output");
    if (M->lastCtx || M->nRound == 1) {
        Ctx *c = M->lastCtx ? M->lastCtx : M->full[0];
        const int rc = dipgpu_expand_columns(H_(c), tolZero, out);
        if (rc) front->err = c->err;
        return rc;
    }
    if (!M->shardedRoundValid) return fail(front, DIPGPU_ERR_STATE, "expand_columns: no host-buffer pricing round since the model was set");
    struct Piece {
        std::vector<int64_t> cs;
        std::vector<int32_t> ri;
        std::vector<double> va;
        std::vector<uint64_t> hs;
        dipgpu_mcols mc;
    };
    std::vector<Piece> pc((size_t)M->nRound);
    int rc = runAll(M, [&](int d) {
        Ctx *s = M->shard[(size_t)d];
        Piece &p = pc[(size_t)d];
        const int cap = std::max(s->nLocal, 1);
        p.cs.assign((size_t)cap + 1, 0);
        p.hs.assign((size_t)cap, 0);
        size_t capNnz = std::max<size_t>(s->lastMColsBytes / sizeof(MColEntry) + 1024, 4096);
        for (int attempt = 0; attempt < 2; ++attempt) {
            p.ri.resize(capNnz);
            p.va.resize(capNnz);
            memset(&p.mc, 0, sizeof(p.mc));
            p.mc.capCols = cap;
            p.mc.capNnz = (int64_t)capNnz;
            p.mc.colStart = p.cs.data();
            p.mc.rowInd = p.ri.data();
            p.mc.val = p.va.data();
            p.mc.hash = p.hs.data();
            const int r = dipgpu_expand_columns(H_(s), tolZero, &p.mc);
            if (r == DIPGPU_ERR_OVERFLOW && attempt == 0 && p.mc.nNnz > (int64_t)capNnz) {
                capNnz = (size_t)p.mc.nNnz;
                continue;
            }
            return subFail(M, s, r);
        }
        return 0;
    }, M->nRound);
    if (rc) return rc;
    int nCols = 0;
    int64_t nNnz = 0;
    for (const Piece &p : pc) {
        nCols += p.mc.nCols;
        nNnz += p.mc.nNnz;
    }
    out->nCols = nCols;
    out->nNnz = nNnz;
    if (nCols > out->capCols || nNnz > out->capNnz)
        return fail(front, DIPGPU_ERR_OVERFLOW, "expand_columns needs capCols >= %d, capNnz >= %lld", nCols, (long long)nNnz);
    int k0 = 0;
    int64_t e0 = 0;
    for (const Piece &p : pc) {
        for (int k = 0; k < p.mc.nCols; ++k) {
            if (out->colStart) out->colStart[k0 + k] = e0 + p.cs[(size_t)k];
            if (out->hash) out->hash[k0 + k] = p.hs[(size_t)k];
        }
        if (p.mc.nNnz > 0) {
            if (out->rowInd) memcpy(out->rowInd + e0, p.ri.data(), sizeof(int32_t) * (size_t)p.mc.nNnz);
            if (out->val) memcpy(out->val + e0, p.va.data(), sizeof(double) * (size_t)p.mc.nNnz);
        }
        k0 += p.mc.nCols;
        e0 += p.mc.nNnz;
    }
    if (out->colStart) out->colStart[nCols] = nNnz;
    return DIPGPU_OK;
}

}  // namespace dipgpu

using namespace dipgpu;

extern "C" {

int dipgpu_create_multi(dipgpu_ctx **ctx, const int32_t *devices, int32_t nDevices)
{
    if (!ctx) return DIPGPU_ERR_INVALID;
    *ctx = nullptr;
    if (!devices || nDevices < 1 || nDevices > 64) return DIPGPU_ERR_INVALID;
    // (an ordinal may be listed more than once: its shards then share that GPU -- how the partitioned paths are
    // tested on a one-GPU box)
    Multi *M = new (std::nothrow) Multi();
    Ctx *front = new (std::nothrow) Ctx();
    if (!M || !front) {
        delete M;
        delete front;
        return DIPGPU_ERR_NOMEM;
    }
    front->multi = M;
    front->device = devices[0];
    M->front = front;
    M->ndev = nDevices;
    M->dev.assign(devices, devices + nDevices);
    M->full.assign((size_t)nDevices, nullptr);
    M->shard.assign((size_t)nDevices, nullptr);
    M->workers.assign((size_t)nDevices, nullptr);
    M->devMs.assign((size_t)nDevices, 0.f);
    int rc = DIPGPU_OK;
    for (int d = 0; d < nDevices && !rc; ++d) {
        dipgpu_ctx *a = nullptr, *b = nullptr;
        rc = dipgpu_create(&a, devices[d]);
        M->full[(size_t)d] = reinterpret_cast<Ctx *>(a);
        if (!rc && nDevices > 1) {
            rc = dipgpu_create(&b, devices[d]);
            M->shard[(size_t)d] = reinterpret_cast<Ctx *>(b);
        } else {
            M->shard[(size_t)d] = M->full[(size_t)d];
        }
    }
    if (rc) {
        destroyAll(M);
        return rc;
    }
    front->smCount = M->full[0]->smCount;
    // peer access between the primary device and every other one (device-resident rounds over NVLink)
    M->peerOK = nDevices > 1;
    for (int d = 1; d < nDevices; ++d) {
        if (devices[d] == devices[0]) continue;  // the same GPU: plain device memory
        int ab = 0, ba = 0;
        cudaDeviceCanAccessPeer(&ab, devices[0], devices[d]);
        cudaDeviceCanAccessPeer(&ba, devices[d], devices[0]);
        if (!ab || !ba) {
            M->peerOK = false;
            continue;
        }
        cudaSetDevice(devices[0]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[d], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) M->peerOK = false;
        cudaSetDevice(devices[d]);
        e = cudaDeviceEnablePeerAccess(devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) M->peerOK = false;
        (void)cudaGetLastError();
    }
    cudaSetDevice(devices[0]);
    for (int d = 1; d < nDevices; ++d) {
        Worker *w = new (std::nothrow) Worker();
        if (!w) {
            destroyAll(M);
            return DIPGPU_ERR_NOMEM;
        }
        w->slot = d;
        w->device = devices[d];
        M->workers[(size_t)d] = w;
        w->th = std::thread([w] { w->loop(); });
    }
    *ctx = H_(front);
    return DIPGPU_OK;
}

int dipgpu_device_count(const dipgpu_ctx *ctx, int32_t *nDevices, int32_t *peerAccess)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c) return DIPGPU_ERR_INVALID;
    if (nDevices) *nDevices = c->multi ? c->multi->ndev : 1;
    if (peerAccess) *peerAccess = c->multi ? (c->multi->peerOK ? 1 : 0) : 0;
    return DIPGPU_OK;
}

int dipgpu_shard_ranges(const dipgpu_ctx *ctx, int32_t cap, int32_t *firstBlock, int32_t *blockCount)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || cap < 0) return DIPGPU_ERR_INVALID;
    if (!c->multi) {
        if (cap >= 1) {
            if (firstBlock) firstBlock[0] = c->firstBlock;
            if (blockCount) blockCount[0] = c->nLocal;
        }
        return DIPGPU_OK;
    }
    const Multi *M = c->multi;
    for (int d = 0; d < std::min<int>(cap, (int)M->shardFirst.size()); ++d) {
        if (firstBlock) firstBlock[d] = M->shardFirst[(size_t)d];
        if (blockCount) blockCount[d] = M->shardCount[(size_t)d];
    }
    return DIPGPU_OK;
}

int dipgpu_sync(dipgpu_ctx *ctx)
{
    Ctx *c = reinterpret_cast<Ctx *>(ctx);
    if (!c) return DIPGPU_ERR_INVALID;
    if (!c->multi) {
        DIPGPU_CUDA(c, cudaSetDevice(c->device));
        DIPGPU_CUDA(c, cudaStreamSynchronize(c->stream));
        return DIPGPU_OK;
    }
    Multi *M = c->multi;
    int rc = DIPGPU_OK;
    for (int d = 0; d < M->ndev; ++d) {
        for (Ctx *s : {M->shard[(size_t)d], M->full[(size_t)d]}) {
            cudaSetDevice(s->device);
            const cudaError_t e = cudaStreamSynchronize(s->stream);
            if (e != cudaSuccess && !rc) rc = cudaFail(c, e, "dipgpu_sync");
        }
    }
    if (!rc && c->stageTiming) {
        // device time of each device's part of the last device-resident round (events on its own stream)
        for (int k = 0; k < 6; ++k) c->stageMs[k] = 0.f;
        for (int d = 0; d < M->ndev; ++d) M->devMs[(size_t)d] = 0.f;
        for (int d = 0; d < M->nRound; ++d) {
            Ctx *s = M->shard[(size_t)d];
            float ms = 0.f;
            cudaSetDevice(s->device);
            if (cudaEventElapsedTime(&ms, s->ev[0], s->ev[6]) != cudaSuccess) ms = 0.f;
            (void)cudaGetLastError();
            M->devMs[(size_t)d] = ms;
        }
    }
    cudaSetDevice(M->dev[0]);
    return rc;
}

int dipgpu_last_device_ms(const dipgpu_ctx *ctx, int32_t cap, float *ms, float *maxMs)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c) return DIPGPU_ERR_INVALID;
    float mx = 0.f;
    if (c->multi) {
        const Multi *M = c->multi;
        for (int d = 0; d < M->ndev; ++d) {
            if (ms && d < cap) ms[d] = M->devMs[(size_t)d];
            mx = std::max(mx, M->devMs[(size_t)d]);
        }
    } else {
        float tot = 0.f;
        for (int k = 0; k < 6; ++k) tot += c->stageMs[k];
        if (ms && cap >= 1) ms[0] = tot;
        mx = tot;
    }
    if (maxMs) *maxMs = mx;
    return DIPGPU_OK;
}

int dipgpu_result_shards(const dipgpu_ctx *ctx, int32_t cap, int32_t *nShards, size_t *offsets, size_t *sizes)
{
    const Ctx *c = reinterpret_cast<const Ctx *>(ctx);
    if (!c || !nShards) return DIPGPU_ERR_INVALID;
    if (!c->multi) {
        *nShards = 1;
        if (cap >= 1) {
            if (offsets) offsets[0] = 0;
            if (sizes) sizes[0] = c->resultBytes;
        }
        return DIPGPU_OK;
    }
    const Multi *M = c->multi;
    *nShards = M->nRound;
    for (int d = 0; d < std::min<int>(cap, M->nRound); ++d) {
        if (offsets) offsets[d] = d < (int)M->gatherOff.size() ? M->gatherOff[(size_t)d] : 0;
        if (sizes) sizes[d] = M->shard[(size_t)d]->resultBytes;
    }
    return DIPGPU_OK;
}

int dipgpu_unpack_results(const void *packedHost, int32_t nShards, const size_t *offsets, const size_t *sizes, dipgpu_cols *out)
{
    if (!packedHost || nShards < 1 || nShards > 64 || !offsets || !sizes || !out) return DIPGPU_ERR_INVALID;
    const void *packed[64];
    size_t bytes[64];
    for (int s = 0; s < nShards; ++s) {
        packed[s] = static_cast<const char *>(packedHost) + offsets[s];
        bytes[s] = sizes[s];
    }
    return unpackShards(nullptr, packed, bytes, nShards, out);
}

}  // extern "C"

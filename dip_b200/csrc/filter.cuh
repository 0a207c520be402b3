// filter.cuh -- device code of stage (c) shared by filter.cu (stand-alone kernels for large
// shards) and knapsack.cu (fused into the tail of the backtrack kernel for small shards).
#pragma once

#include "common.cuh"

namespace dipgpu {

struct FilterArgs {
    ResultHeader *hdr;
    const double *blockRedCost;
    double *blockMostNeg;
    const int32_t *blockNnz;
    int64_t *keptStart;
    int32_t *keptBlock;
    int32_t *keptInd;
    const int32_t *candInd;
    const int32_t *blockColStart;
    const int32_t *capacity;
    const int32_t *nItems;
    int firstBlock, nLocal;
    int64_t indCap;
    double eps;
    int fuseCopy;
    int pairs;  // 1: candidate / kept entries are (index, multiplicity) pairs, counted in words (ResultHeader::reserved[1])
};


__device__ __forceinline__ void copyColumn(const FilterArgs &a, int k, int lane)
{
    const int lb = a.keptBlock[k];
    const int cs = a.blockColStart[a.firstBlock + lb];
    const int64_t dst = a.keptStart[k];
    const int n = __ldcg(a.blockNnz + lb);
    for (int q = lane; q < n; q += 32) a.keptInd[dst + q] = __ldcg(a.candInd + cs + q);
}

// The scan over a shard's blocks by ONE CTA of NT threads, NT a multiple of 32 (see filter.cu
// for what it restates).
__device__ __forceinline__ void filterScanBody(const FilterArgs &a, const int NT)
{
    __shared__ int sCnt[32];
    __shared__ long long sNnz[32];
    __shared__ long long sCells[32];
    __shared__ int sCarryCnt;
    __shared__ long long sCarryNnz;
    __shared__ long long sDst[kFuseFilterMaxBlocks];  // fuseCopy == 2: kept column -> first output index
    __shared__ int sSrc[kFuseFilterMaxBlocks];        //                 kept column -> first candidate index
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        sCarryCnt = 0;
        sCarryNnz = 0;
    }
    long long cells = 0;
    __syncthreads();
    for (int base = 0; base < a.nLocal; base += NT) {
        const int lb = base + t;
        bool keep = false;
        int nnz = 0;
        if (lb < a.nLocal) {
            const double rc = __ldcg(a.blockRedCost + lb);
            a.blockMostNeg[lb] = rc < 0.0 ? rc : 0.0;  // min(0, rc)
            keep = rc < -a.eps;                        // strict, :5216
            nnz = __ldcg(a.blockNnz + lb);
            cells += (long long)__ldcg(a.nItems + lb) * ((long long)a.capacity[a.firstBlock + lb] + 1);
        }
        // warp-shuffle compaction
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        long long v = keep ? nnz : 0;
        long long incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += up;
        }
        if (lane == 31) {
            sCnt[warp] = __popc(bal);
            sNnz[warp] = incl;
        }
        __syncthreads();
        int cntOff = sCarryCnt;
        long long nnzOff = sCarryNnz;
        int totCnt = 0;
        long long totNnz = 0;
        for (int wv = 0; wv < NT / 32; ++wv) {
            const int cw = sCnt[wv];
            const long long nw = sNnz[wv];
            if (wv < warp) {
                cntOff += cw;
                nnzOff += nw;
            }
            totCnt += cw;
            totNnz += nw;
        }
        if (keep) {
            const int pos = cntOff + wpre;
            a.keptBlock[pos] = lb;
            a.keptStart[pos] = nnzOff + (incl - v);
            if (a.fuseCopy == 2) {
                sDst[pos] = nnzOff + (incl - v);
                sSrc[pos] = a.blockColStart[a.firstBlock + lb];
            }
        }
        __syncthreads();
        if (t == 0) {
            sCarryCnt += totCnt;
            sCarryNnz += totNnz;
        }
        __syncthreads();
    }
    // total DP cells of the shard
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, off);
    if (lane == 0) sCells[warp] = cells;
    __syncthreads();
    if (t == 0) {
        long long tot = 0;
        for (int wv = 0; wv < NT / 32; ++wv) tot += sCells[wv];
        a.keptStart[sCarryCnt] = sCarryNnz;
        a.hdr->magic = kResultMagic;
        a.hdr->m = a.nLocal;
        a.hdr->firstBlock = a.firstBlock;
        a.hdr->nKept = sCarryCnt;
        a.hdr->nInd = sCarryNnz;
        a.hdr->cells = tot;
        a.hdr->reserved[0] = tot;  // the large-shard path evaluates every cell
        a.hdr->reserved[1] = a.pairs;
        a.hdr->indCap = a.indCap;
    }
    if (a.fuseCopy == 2) {
        // small shards (<= kFuseFilterMaxBlocks): flat copy, every index an independent load
        __syncthreads();
        const int nKept = sCarryCnt;
        const long long total = sCarryNnz;
        // batches of 8 independent loads per thread: the L2 round trips overlap instead of
        // serialising (a warp issues in order and would stall at each load's first use)
        for (long long base = t; base < total; base += 8LL * NT) {
            int v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const long long idx = base + (long long)q * NT;
                v[q] = 0;
                if (idx < total) {
                    int lo = 0, hi = nKept - 1;
                    while (lo < hi) {  // last kept column whose first output index is <= idx
                        const int mid = (lo + hi + 1) >> 1;
                        if (sDst[mid] <= idx) lo = mid; else hi = mid - 1;
                    }
                    v[q] = __ldcg(a.candInd + sSrc[lo] + (int)(idx - sDst[lo]));
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const long long idx = base + (long long)q * NT;
                if (idx < total) a.keptInd[idx] = v[q];
            }
        }
    } else if (a.fuseCopy) {
        __threadfence_block();
        __syncthreads();
        const int nKept = sCarryCnt;
        for (int k = warp; k < nKept; k += NT / 32) copyColumn(a, k, lane);
    }
}


FilterArgs makeFilterArgs(Ctx *c, double redCostEps);  // filter.cu

}  // namespace dipgpu

// intknap.cu -- integer (unbounded) knapsack blocks: the pricing problem of DIP's cutting-stock example.
//
// Reference: kp(obj, weights, capacity) and relaxed_solver of Dip/src/dippy/tests/test_cutting_stock.py:154-179 and
// :98-151 (the copy of kp in Dip/src/dippy/examples/csp/cutting_stock_func.py:92-118 compares `zbest > zyes` and always
// returns 0).  kp is a recursion on the capacity:
//     kp(0) = (0, zeros);  kp(c): best = kp(c-1);  for i = 0..n-1: if w_i <= c: zyes = kp(c-w_i).z + obj_i;
//                                                                   if zyes > zbest: best = (zyes, kp(c-w_i).sol + e_i)
// i.e. F[c] = max(F[c-1], max_i F[c-w_i] + obj_i), the FIRST candidate in the order (skip, item 0, item 1, ...) that
// reaches the maximum wins, and a cell's value is the chain sum "source cell + obj_i" -- so the table is evaluated
// here exactly like that: cells sequential in the capacity, the items of a cell in parallel under an ordered max.
//
// Shape: one CTA per block.  All cells of a WINDOW of wmin = (smallest usable weight) consecutive capacities read only
// cells below the window, except through the "skip" candidate F[c-1]:
//   phase A  one warp per cell of the window: M[c] = max_i F[c-w_i] + obj_i with the smallest i among the maxima
//            (lanes stride over the items in ascending order under a strict >, then an ordered shuffle reduction);
//   phase B  warp 0 closes the window with a prefix maximum: F[c] = max(F[c-1], M[c]), the item is taken iff
//            M[c] > F[c-1] (strict: skip wins ties), and src[c] = the last cell <= c where an item was taken -- the
//            backtrack then jumps from taken cell to taken cell instead of walking the skips.
// The tables (F, item, src) sit in shared memory when (capacity + 1) * 16 bytes fit beside the items, else in global
// scratch (same code through generic pointers).  Result per block: (x-space index, multiplicity) pairs ascending, the
// "use" column of the block (the roll, test_cutting_stock.py:139-141) merged in with multiplicity 1 when anything is
// cut; varRedCost = -z + redCost[use] (:129) and 0 for the empty column (:131), WITHOUT the convexity dual as on the
// C++ plugin surface (DecompAlgo::solveRelaxed subtracts it, Dip/src/DecompAlgo.cpp:6350-6356); origCost = sum of
// c_j * x_j in ascending column order.  The stand-alone filter kernels pack the round: a pair is two int32 words of
// the packed index region (ResultHeader::reserved[1] = 1 tells the decoder).
#include <algorithm>

#include "filter.cuh"

namespace dipgpu {

constexpr int kIkThreads = 256;

struct IntKnapArgs {
    const double *redCost;
    const double *obj;
    const double *alpha;           // indexed by GLOBAL block
    const int32_t *blockColStart;  // global block -> first item column
    const int32_t *weight;
    const int32_t *capacity;
    const int32_t *useCol;         // global block -> x-space index of its use column (-1: none); may be NULL
    const int32_t *candStart;      // global block -> first WORD of its candidate pairs in candInd
    int32_t *candInd;
    int32_t *nItems;
    double *blockObj, *blockRedCost, *blockOrigCost;
    int32_t *blockNnz;
    int firstBlock, nLocal;
    int maxItems;                  // items staged in shared memory per block (>= the largest block)
    int tablesInSmem;              // 1: F / item / src of (maxCapacity + 1) cells follow the items in shared memory
    int maxCells;                  // maxCapacity + 1
    double *gF;                    // global scratch when the tables do not fit: nLocal x maxCells each
    int32_t *gItem, *gSrc;
};

__global__ void __launch_bounds__(kIkThreads) intknap_kernel(const __grid_constant__ IntKnapArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    __shared__ int sWarpCnt[kIkThreads / 32];
    __shared__ int sBase, sN, sNRef, sWmin, sTotal, sEntries, sUsePos;
    __shared__ double sPartV[kIkThreads];   // phase A: best candidate of (item run, cell)
    __shared__ int sPartI[kIkThreads];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NW = kIkThreads / 32;
    const int lb = blockIdx.x;
    const int b = a.firstBlock + lb;
    const int cs = a.blockColStart[b], ce = a.blockColStart[b + 1];
    const int cap = a.capacity[b];
    const double negInf = __longlong_as_double(0xfff0000000000000LL);

    double *sP = reinterpret_cast<double *>(smemRaw);                   // profit of active item k
    int32_t *sW = reinterpret_cast<int32_t *>(sP + a.maxItems);         // weight
    int32_t *sCol = sW + a.maxItems;                                    // x-space column
    int32_t *sCnt = sCol + a.maxItems;                                  // multiplicity in the solution
    double *F;
    int32_t *item, *src;
    if (a.tablesInSmem) {
        F = reinterpret_cast<double *>(sCnt + ((a.maxItems + 1) & ~1));
        item = reinterpret_cast<int32_t *>(F + a.maxCells);
        src = item + a.maxCells;
    } else {
        F = a.gF + (size_t)lb * a.maxCells;
        item = a.gItem + (size_t)lb * a.maxCells;
        src = a.gSrc + (size_t)lb * a.maxCells;
    }

    // ---- items: columns with redCost < 0 (test_cutting_stock.py:103-104) in ascending order; an item heavier than the
    // capacity is never a candidate (:170) and is left out of the staged list (it still counts as an item of the
    // reference's table)
    if (t == 0) { sBase = 0; sNRef = 0; sWmin = 0x7fffffff; sTotal = 0; }
    __syncthreads();
    for (int j0 = cs; j0 < ce; j0 += kIkThreads) {
        const int j = j0 + t;
        double rc = 0.0;
        int w = 0;
        if (j < ce) { rc = a.redCost[j]; w = a.weight[j]; }
        const bool neg = j < ce && rc < 0.0;
        const bool act = neg && w <= cap;
        const unsigned bal = __ballot_sync(0xffffffffu, act);
        const unsigned balRef = __ballot_sync(0xffffffffu, neg);
        if (lane == 0) sWarpCnt[warp] = __popc(bal);
        if (lane == 0 && balRef) atomicAdd(&sNRef, __popc(balRef));
        __syncthreads();
        int pre = sBase;
        for (int v = 0; v < warp; ++v) pre += sWarpCnt[v];
        if (act) {
            const int k = pre + __popc(bal & ((1u << lane) - 1u));
            sP[k] = -rc;  // :106
            sW[k] = w;
            sCol[k] = j;
            sCnt[k] = 0;
            atomicMin(&sWmin, w);
        }
        __syncthreads();
        if (t == 0) {
            int tot = 0;
            for (int v = 0; v < NW; ++v) tot += sWarpCnt[v];
            sBase += tot;
        }
        __syncthreads();
    }
    if (t == 0) sN = sBase;
    __syncthreads();
    const int n = sN;
    const int wmin = sWmin;

    // ---- the table, window by window
    double z = 0.0;
    if (n > 0 && cap > 0) {
        if (t == 0) { F[0] = 0.0; item[0] = -1; src[0] = 0; }
        __syncthreads();
        for (int c0 = 1; c0 <= cap; c0 += wmin) {
            const int c1 = min(cap, c0 + wmin - 1);
            // phase A: candidates of the items.  The CTA's threads are laid over (cell of the window) x (group of
            // items): thread (ci, g) scans a CONTIGUOUS run of items, ascending, under
            // a strict >, so the first maximum of its run stays; the runs of a cell are then combined in run order,
            // again under a strict >: the first maximum over (item 0, item 1, ...) wins, as in the reference's loop.
            for (int cb = c0; cb <= c1; cb += kIkThreads) {
                const int wc = min(c1 - cb + 1, kIkThreads);            // cells of this pass
                const int G = max(1, min(kIkThreads / wc, (n + 3) / 4)); // item runs per cell (>= 4 items each)
                const int per = (n + G - 1) / G;
                const int ci = t % wc, g = t / wc;
                const int c = cb + ci;
                double best = negInf;
                int arg = 0x7fffffff;
                if (g < G) {
                    const int i1 = min(n, (g + 1) * per);
                    for (int i = g * per; i < i1; ++i) {
                        const int w = sW[i];
                        if (w <= c) {                                  // :170
                            const double zyes = F[c - w] + sP[i];      // :171-173
                            if (zyes > best) { best = zyes; arg = i; } // :175 (ascending i: the first maximum stays)
                        }
                    }
                }
                if (G == 1) {
                    if (g == 0) { F[c] = best; item[c] = arg; }        // (cells of the window are not read in phase A)
                } else {
                    if (g < G) { sPartV[g * wc + ci] = best; sPartI[g * wc + ci] = arg; }
                    __syncthreads();
                    if (g == 0) {
                        for (int q = 1; q < G; ++q) {
                            const double ov = sPartV[q * wc + ci];
                            if (ov > best) { best = ov; arg = sPartI[q * wc + ci]; }
                        }
                        F[c] = best;
                        item[c] = arg;
                    }
                    __syncthreads();   // the partials are free for the next pass
                }
            }
            __syncthreads();
            // phase B: F[c] = max(F[c-1], M[c]); taken iff M[c] > F[c-1]
            if (warp == 0) {
                double carry = F[c0 - 1];
                int carrySrc = src[c0 - 1];
                for (int q0 = c0; q0 <= c1; q0 += 32) {
                    const int c = q0 + lane;
                    const double mv = c <= c1 ? F[c] : negInf;
                    double inc = mv;  // inclusive prefix maximum of M over the chunk
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const double up = __shfl_up_sync(0xffffffffu, inc, off);
                        if (lane >= off && up > inc) inc = up;
                    }
                    double exc = __shfl_up_sync(0xffffffffu, inc, 1);  // F[c-1]: maximum over the cells before this one
                    if (lane == 0 || carry > exc) exc = carry;
                    const bool taken = c <= c1 && mv > exc;            // strict: skip wins a tie (:167, :175)
                    int s = taken ? c : 0;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const int up = __shfl_up_sync(0xffffffffu, s, off);
                        if (lane >= off && up > s) s = up;
                    }
                    if (s == 0) s = carrySrc;
                    if (c <= c1) {
                        F[c] = taken ? mv : exc;
                        src[c] = s;
                    }
                    const double fLast = taken ? mv : exc;
                    carry = __shfl_sync(0xffffffffu, fLast, 31);
                    carrySrc = __shfl_sync(0xffffffffu, s, 31);
                }
            }
            __syncthreads();
        }
        z = F[cap];
        // ---- the counts, from taken cell to taken cell (:174 solyes[i] += 1 along the chain)
        if (t == 0) {
            int c = src[cap], total = 0;
            while (c > 0) {
                const int i = item[c];
                sCnt[i] += 1;
                ++total;
                c = src[c - sW[i]];
            }
            sTotal = total;
        }
        __syncthreads();
    }

    // ---- the column: (index, multiplicity) pairs ascending, the use column merged in
    const int use = a.useCol != nullptr ? a.useCol[b] : -1;
    const int total = sTotal;
    int32_t *cand = a.candInd + a.candStart[b];
    if (t == 0) { sBase = 0; sUsePos = 0; }
    __syncthreads();
    const bool withUse = total > 0 && use >= 0;
    if (total > 0) {
        for (int k0 = 0; k0 < n; k0 += kIkThreads) {
            const int k = k0 + t;
            const bool in = k < n && sCnt[k] > 0;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (lane == 0) sWarpCnt[warp] = __popc(bal);
            __syncthreads();
            int pre = sBase;
            for (int v = 0; v < warp; ++v) pre += sWarpCnt[v];
            if (in) {
                const int col = sCol[k];
                const bool before = withUse && use < col;  // the use column sits in front of this entry
                const int pos = pre + __popc(bal & ((1u << lane) - 1u)) + (before ? 1 : 0);
                cand[2 * pos] = col;
                cand[2 * pos + 1] = sCnt[k];
                if (withUse && !before) atomicAdd(&sUsePos, 1);
            }
            __syncthreads();
            if (t == 0) {
                int tot = 0;
                for (int v = 0; v < NW; ++v) tot += sWarpCnt[v];
                sBase += tot;
            }
            __syncthreads();
        }
        if (t == 0) {
            int e = sBase;
            if (withUse) {
                cand[2 * sUsePos] = use;
                cand[2 * sUsePos + 1] = 1;
                ++e;
            }
            sEntries = e;
        }
        __syncthreads();
    }
    if (t == 0) {
        double rc = 0.0, oc = 0.0;
        int e = 0;
        if (total > 0) {
            e = sEntries;
            for (int q = 0; q < e; ++q) oc += a.obj[cand[2 * q]] * (double)cand[2 * q + 1];
            rc = -z;
            if (use >= 0) rc = rc + a.redCost[use];  // :129
        }
        a.blockObj[lb] = z;
        a.blockRedCost[lb] = rc - a.alpha[b];        // DecompAlgo.cpp:6350-6356
        a.blockOrigCost[lb] = oc;
        a.blockNnz[lb] = 2 * e;                      // words of the packed index region
        a.nItems[lb] = sNRef;
    }
}

int launchIntKnap(Ctx *c, const double *dRedCost, const double *dAlpha, cudaStream_t st)
{
    if (c->nLocal == 0) return DIPGPU_OK;
    char *res = static_cast<char *>(c->dResult);
    IntKnapArgs a;
    a.redCost = dRedCost;
    a.obj = c->dObj;
    a.alpha = dAlpha;
    a.blockColStart = c->dBlockColStart;
    a.weight = c->dWeight;
    a.capacity = c->dCapacity;
    a.useCol = c->dUseCol;
    a.candStart = c->dCandStart;
    a.candInd = c->dCandInd;
    a.nItems = c->dNItems;
    a.blockObj = reinterpret_cast<double *>(res + c->layout.offObj);
    a.blockRedCost = reinterpret_cast<double *>(res + c->layout.offRedCost);
    a.blockOrigCost = reinterpret_cast<double *>(res + c->layout.offOrigCost);
    a.blockNnz = reinterpret_cast<int32_t *>(res + c->layout.offNnz);
    a.firstBlock = c->firstBlock;
    a.nLocal = c->nLocal;
    a.maxItems = std::max(c->maxBlockCols, 1);
    a.maxCells = c->maxCapacity + 1;
    const size_t itemBytes = (size_t)a.maxItems * 8 + (size_t)a.maxItems * 8 + (size_t)((a.maxItems + 1) & ~1) * 4;
    const size_t tableBytes = (size_t)a.maxCells * 16;
    const size_t limit = 227 * 1024 - 1024;
    if (itemBytes > limit) return fail(c, DIPGPU_ERR_UNSUPPORTED, "integer knapsack block with %d columns: the items do not fit shared memory", c->maxBlockCols);
    a.tablesInSmem = itemBytes + tableBytes <= limit ? 1 : 0;
    a.gF = nullptr;
    a.gItem = a.gSrc = nullptr;
    if (!a.tablesInSmem) {
        // tables in global scratch (reuses the decision-bit buffer: sized for it by setBlockRange)
        char *base = reinterpret_cast<char *>(c->dBits);
        const size_t cells = (size_t)c->nLocal * (size_t)a.maxCells;
        a.gF = reinterpret_cast<double *>(base);
        a.gItem = reinterpret_cast<int32_t *>(base + cells * 8);
        a.gSrc = a.gItem + cells;
    }
    const size_t smem = itemBytes + (a.tablesInSmem ? tableBytes : 0) + 16;
    if (c->ikSmem < smem) {
        const int rc = ensureSmemLimit(c, reinterpret_cast<const void *>(intknap_kernel), smem);
        if (rc) return rc;
        c->ikSmem = smem;
    }
    intknap_kernel<<<c->nLocal, kIkThreads, smem, st>>>(a);
    c->launches++;
    DIPGPU_CUDA(c, cudaGetLastError());
    return DIPGPU_OK;
}

}  // namespace dipgpu

/*
 * dip_oracle.c -- CPU restatement of DIP's pricing round (see dip_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path and the "port" CPU
 * baseline of bench.py.  Never linked into libdipgpu.
 *
 * Every function cites the reference lines (relative to /root/reference) whose
 * arithmetic it restates.  Written from the behaviour of those lines; no
 * reference source is copied.
 */
#include "dip_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- duals -- */

void dip_oracle_adjust_duals(const double *uOld, int nMasterRows, int nBaseCoreRows,
                             int numConvex, double *uNew)
{
    /* DecompAlgo.cpp:4566 -- core (original + branch) rows keep their slot */
    memcpy(uNew, uOld, sizeof(double) * (size_t)nBaseCoreRows);
    /* DecompAlgo.cpp:4595-4597 -- cut rows slide down over the convexity rows */
    int nCuts = nMasterRows - nBaseCoreRows - numConvex;
    if (nCuts > 0)
        memcpy(uNew + nBaseCoreRows, uOld + nBaseCoreRows + numConvex,
               sizeof(double) * (size_t)nCuts);
}

/* ------------------------------------------------------------ red. cost -- */

void dip_oracle_calc_redcost(int R, int N, const int64_t *rowStart, const int32_t *colInd,
                             const double *val, const double *uAdj, const double *c,
                             int phase1, double *redCostX)
{
    /* DecompAlgo.cpp:4532: modelCore->M->transposeTimes(u, redCostX) on a
     * row-ordered matrix (DecompConstraintSet.cpp:41-43) = scatter of u[i]*row i.
     * CoinUtils walks the major vectors from the last to the first and skips
     * exact-zero multipliers (un-vendored; order is "parity unpinned"). */
    for (int j = 0; j < N; ++j) redCostX[j] = 0.0;
    for (int i = R - 1; i >= 0; --i) {
        const double ui = uAdj[i];
        if (ui == 0.0) continue;
        for (int64_t k = rowStart[i]; k < rowStart[i + 1]; ++k)
            redCostX[colInd[k]] += ui * val[k];
    }
    if (phase1) { /* DecompAlgo.cpp:4538-4541 */
        for (int j = 0; j < N; ++j) redCostX[j] = -redCostX[j];
    } else { /* DecompAlgo.cpp:4543-4545 */
        for (int j = 0; j < N; ++j) redCostX[j] = c[j] - redCostX[j];
    }
}

/* ------------------------------------------------------------- knapsack -- */

double dip_oracle_knapsack01(int n, const double *obj, const int32_t *weights, int capacity,
                             int32_t *solution, int *nSol, uint8_t *added)
{
    /* gap_func.py:139-141 */
    if (nSol) *nSol = 0;
    if (n <= 0) return 0.0;

    const size_t W = (size_t)capacity + 1;
    uint8_t *tab = added ? added : (uint8_t *)calloc((size_t)n * W, 1);
    if (added) memset(tab, 0, (size_t)n * W);
    double *row = (double *)calloc(W, sizeof(double)); /* c[i-1][.], row -1 == 0 */

    /* gap_func.py:160-170.  The two-row table is evaluated in place with the
     * capacity index descending, which reads exactly c[i-1][j - w]; cells with
     * w > j keep c[i-1][j].  The take test is the strict '>' of :166. */
    for (int i = 0; i < n; ++i) {
        const int w = weights[i];
        const double p = obj[i];
        uint8_t *t = tab + (size_t)i * W;
        if (w > capacity) continue;
        for (int j = capacity; j >= w; --j) {
            const double cand = p + row[j - w];
            if (cand > row[j]) {
                row[j] = cand;
                t[j] = 1;
            }
        }
    }
    const double z = row[capacity];

    /* gap_func.py:173-181: walk items from the last one down */
    int cnt = 0;
    int j = capacity;
    for (int i = n - 1; i >= 0 && j >= 0; --i) {
        if (tab[(size_t)i * W + (size_t)j]) {
            if (solution) solution[cnt] = i;
            ++cnt;
            j -= weights[i];
        }
    }
    if (nSol) *nSol = cnt;
    free(row);
    if (!added) free(tab);
    return z;
}

/* Dip/src/UtilMacros.h:521-531 */
static double util_frac_part(double x)
{
    const double fl = floor(x);
    const double flp = floor(x + 0.5);
    if (fabs(flp - x) < 1.0e-6 * (fabs(flp) + 1.0)) return 0.0;
    return x - fl;
}

/* Dip/src/UtilMacros.h:548-552 */
static int util_is_zero(double x) { return fabs(x) < 1.0e-8; }

int dip_oracle_calc_scale_factor(int n, const double *redCost, double *work)
{
    /* GAP_KnapPisinger.h:89-143, epsTol = 1e-4 */
    const double epsTol = 1.0e-4;
    const double oneOverEps = 1.0 / epsTol;
    int scale = 1, nFrac = 0;
    for (int i = 0; i < n; ++i) {
        if (redCost[i] >= 0) continue;
        double frac = util_frac_part(-redCost[i]);
        if (!util_is_zero(frac)) {
            frac *= oneOverEps;
            work[nFrac++] = (int)frac * (double)epsTol; /* truncate to 1e-4 */
        }
    }
    for (int i = 0; i < nFrac; ++i) {
        int tooBig = 0;
        work[i] *= scale;
        while (!util_is_zero(util_frac_part(work[i]))) {
            scale *= 10;
            if (scale >= oneOverEps) {
                tooBig = 1;
                break;
            }
            work[i] *= 10;
        }
        if (tooBig) break;
    }
    return scale;
}

/* ---- CPU-baseline variant (bench.py only): the reference's own Horowitz-Sahni branch and bound,
 * KnapsackOptimizeHS (Dip/src/UtilKnapsack.cpp:246-502), compiled from the reference sources into
 * oracle/_ref/libdip_ref_knap.so and handed in as a function pointer.  profitMode 2 prices a block with it:
 * items = columns with redCost < 0 that fit, sorted by non-increasing p/w as KnapsackSortRatioOut does
 * (UtilKnapsack.cpp:204-243; its own sort prints from inside the loop and is restated here), the solution mapped
 * back to ascending column order.  n == 1 is decided directly (the reference's n == 1 branch writes x[1], :314). */
typedef int (*dip_hs_fn)(const int, const double, double *, double *, int *, double *, int *);
static dip_hs_fn g_hs = 0;
void dip_oracle_set_hs(void *fn) { g_hs = (dip_hs_fn)fn; }

typedef struct { double r; int i; } ratio_t;
static int ratio_cmp(const void *a, const void *b)
{
    const ratio_t *x = (const ratio_t *)a, *y = (const ratio_t *)b;
    if (x->r > y->r) return -1;
    if (x->r < y->r) return 1;
    return x->i - y->i;
}

static double solve_hs(int nItems, const double *p, const int32_t *w, int capacity, uint8_t *x)
{
    ratio_t *ra = (ratio_t *)malloc(sizeof(ratio_t) * (size_t)nItems);
    int n = 0;
    for (int i = 0; i < nItems; ++i)
        if (w[i] <= capacity && w[i] > 0) { ra[n].r = p[i] / (double)w[i]; ra[n].i = i; ++n; }
    double z = 0.0;
    if (n == 1) {
        x[ra[0].i] = 1;
        z = p[ra[0].i];
    } else if (n >= 2) {
        qsort(ra, (size_t)n, sizeof(ratio_t), ratio_cmp);
        double *ps = (double *)malloc(sizeof(double) * (size_t)(n + 2));
        double *ws = (double *)malloc(sizeof(double) * (size_t)(n + 2));
        int *xs = (int *)calloc((size_t)(n + 2), sizeof(int));
        for (int k = 0; k < n; ++k) { ps[k] = p[ra[k].i]; ws[k] = (double)w[ra[k].i]; }
        ps[n] = 0.0; ws[n] = 1.0e20; ps[n + 1] = 0.0; ws[n + 1] = 1.0e20;
        int status = 0;
        g_hs(n, (double)capacity, ps, ws, xs, &z, &status);
        for (int k = 0; k < n; ++k) if (xs[k]) x[ra[k].i] = 1;
        free(ps); free(ws); free(xs);
    }
    free(ra);
    return z;
}

int dip_oracle_solve_block(int n, const double *redCost, const double *origCost,
                           const int32_t *weight, int capacity, int colOffset, int profitMode,
                           int32_t *ind, double *varRedCost, double *varOrigCost, double *zOut,
                           int64_t *cellsOut)
{
    int32_t *toOrig = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t *w = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    uint8_t *x = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    int nItems = 0;
    long weightSum = 0;
    double z = 0.0;

    if (profitMode == 0 || profitMode == 2) {
        /* gap_func.py:102-105: tasks with negative reduced cost, obj = -rc */
        for (int j = 0; j < n; ++j) {
            if (redCost[j] < 0) {
                toOrig[nItems] = j;
                w[nItems] = weight[j];
                p[nItems] = -redCost[j];
                ++nItems;
            }
        }
    } else {
        /* GAP_KnapPisinger.h:167-186 */
        double *work = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        const int scale = dip_oracle_calc_scale_factor(n, redCost, work);
        free(work);
        for (int j = 0; j < n; ++j) {
            if (redCost[j] < 0.0) {
                const double scaled = (-redCost[j]) * scale;
                toOrig[nItems] = j;
                w[nItems] = weight[j];
                p[nItems] = (double)(long)floor(scaled + 0.5);
                weightSum += weight[j];
                ++nItems;
            }
        }
    }

    int64_t cells = 0;
    int solved = 0;
    if (profitMode == 1) {
        /* GAP_KnapPisinger.h:203-238: trivial cases before combo */
        if (nItems == 1) {
            /* the reference asserts w <= capacity here (:204); an item that does not fit
             * is left out, as the DP would */
            if (w[0] <= capacity) {
                x[0] = 1;
                z = p[0];
            }
            solved = 1;
        } else if (nItems >= 1 && weightSum <= capacity) {
            for (int i = 0; i < nItems; ++i) {
                x[i] = 1;
                z += p[i];
            }
            solved = 1;
        } else if (nItems == 2) {
            double best = 0;
            if (w[1] <= capacity && p[1] > best) { best = p[1]; x[0] = 0; x[1] = 1; }
            if (w[0] <= capacity && p[0] > best) { best = p[0]; x[0] = 1; x[1] = 0; }
            if (w[0] + w[1] <= capacity && p[0] + p[1] > best) { best = p[0] + p[1]; x[0] = 1; x[1] = 1; }
            z = best;
            solved = 1;
        }
    }
    if (!solved && nItems > 0 && profitMode == 2 && g_hs) {
        z = solve_hs(nItems, p, w, capacity, x);
        solved = 1;
    }
    if (!solved && nItems > 0) {
        /* gap_func.py:107 / GAP_KnapPisinger.h:243-249 (combo stand-in) */
        int32_t *sol = (int32_t *)malloc(sizeof(int32_t) * (size_t)nItems);
        int nSol = 0;
        z = dip_oracle_knapsack01(nItems, p, w, capacity, sol, &nSol, NULL);
        for (int k = 0; k < nSol; ++k) x[sol[k]] = 1;
        free(sol);
        for (int i = 0; i < nItems; ++i)
            if (w[i] <= capacity) cells += (int64_t)capacity + 1;
    }

    /* GAP_KnapPisinger.h:263-272 + GAP_DecompApp.cpp:252-253: x-space indices,
     * element 1.0, sums of the TRUE reduced / original costs.  Emission order
     * here is ascending index (Appendix A.8 (iv) of SURVEY.md). */
    int nnz = 0;
    double rc = 0.0, oc = 0.0;
    for (int i = 0; i < nItems; ++i) {
        if (x[i]) {
            const int j = toOrig[i];
            rc += redCost[j];
            oc += origCost[j];
            if (ind) ind[nnz] = colOffset + j;
            ++nnz;
        }
    }
    *varRedCost = rc;
    *varOrigCost = oc;
    if (zOut) *zOut = z;
    if (cellsOut) *cellsOut = cells;
    free(toOrig);
    free(w);
    free(p);
    free(x);
    return nnz;
}

int dip_oracle_solve_blocks(int m, const int32_t *blockColStart, const int32_t *weight,
                            const int32_t *capacity, int profitMode, const double *redCostX,
                            const double *origCost, const double *alpha, double redCostEps,
                            int nThreads, int maxBlockCols, int32_t *colNnz,
                            double *colRedCost, double *colOrigCost, int32_t *colKeep,
                            double *mostNegRC, double *z, int32_t *colIndOut,
                            double *mostNegReducedCost, int64_t *cellsOut)
{
    int64_t cellsTotal = 0;
    (void)nThreads;
    /* DecompAlgo.cpp:4815-4848: one task per block, dynamic schedule */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nThreads > 0 ? nThreads : 1) reduction(+ : cellsTotal)
#endif
    for (int b = 0; b < m; ++b) {
        const int s = blockColStart[b];
        const int n = blockColStart[b + 1] - s;
        double rc = 0, oc = 0, zb = 0;
        int64_t cells = 0;
        int32_t *ind = colIndOut ? colIndOut + (size_t)b * (size_t)maxBlockCols : NULL;
        /* GAP_DecompApp.cpp:247-281 */
        const int nnz = dip_oracle_solve_block(n, redCostX + s, origCost + s, weight + s,
                                               capacity[b], s, profitMode, ind, &rc, &oc, &zb,
                                               &cells);
        colNnz[b] = nnz;
        /* DecompAlgo.cpp:6350-6356: the driver subtracts alpha for Point vars */
        colRedCost[b] = rc - alpha[b];
        colOrigCost[b] = oc;
        if (z) z[b] = zb;
        cellsTotal += cells;
    }
    /* DecompAlgo.cpp:4761, 5151-5232 */
    int kept = 0;
    double sum = 0.0;
    for (int b = 0; b < m; ++b) {
        double mn = 0.0;
        if (colRedCost[b] < mn) mn = colRedCost[b];
        mostNegRC[b] = mn;
        colKeep[b] = colRedCost[b] < -redCostEps ? 1 : 0;
        kept += colKeep[b];
    }
    for (int b = 0; b < m; ++b) sum += mostNegRC[b];
    if (mostNegReducedCost) *mostNegReducedCost = sum;
    if (cellsOut) *cellsOut = cellsTotal;
    return kept;
}

int dip_oracle_price_round(int R, int N, const int64_t *rowStart, const int32_t *colInd,
                           const double *val, const double *c, const double *u, int nBase,
                           int phase1, int m, const int32_t *blockColStart,
                           const int32_t *weight, const int32_t *capacity, int profitMode,
                           double redCostEps, int nThreads, int maxBlockCols,
                           double *redCostX, int32_t *colNnz, double *colRedCost,
                           double *colOrigCost, int32_t *colKeep, double *mostNegRC,
                           double *z, int32_t *colIndOut, double *mostNegReducedCost,
                           int64_t *cellsOut)
{
    /* DecompAlgo.cpp:4692-4726 */
    double *uAdj = (double *)malloc(sizeof(double) * (size_t)(R > 0 ? R : 1));
    dip_oracle_adjust_duals(u, R + m, nBase, m, uAdj);
    dip_oracle_calc_redcost(R, N, rowStart, colInd, val, uAdj, c, phase1, redCostX);
    free(uAdj);
    /* DecompAlgo.cpp:4820: alpha_b = u[nBaseCoreRows + b] */
    const int kept = dip_oracle_solve_blocks(m, blockColStart, weight, capacity, profitMode,
                                             redCostX, c, u + nBase, redCostEps, nThreads,
                                             maxBlockCols, colNnz, colRedCost, colOrigCost,
                                             colKeep, mostNegRC, z, colIndOut,
                                             mostNegReducedCost, cellsOut);
    return kept;
}

int dip_oracle_expand_column(int R, int N, const int64_t *rowStart, const int32_t *colInd,
                             const double *val, int nBase, int numConvex, int nInd,
                             const int32_t *ind, const double *els, int blockId, double tolZero,
                             int32_t *outRow, double *outVal)
{
    /* DecompVar::fillDenseArr, DecompVar.cpp:58-70 */
    double *x = (double *)calloc((size_t)(N > 0 ? N : 1), sizeof(double));
    for (int k = 0; k < nInd; ++k) x[ind[k]] = els ? els[k] : 1.0;
    /* modelCore->M->times(x, denseCol), DecompAlgo.cpp:5484 (row-major: one dot product per row) */
    const int nRows = R + numConvex;
    double *dense = (double *)calloc((size_t)(nRows > 0 ? nRows : 1), sizeof(double));
    for (int i = 0; i < R; ++i) {
        double y = 0.0;
        for (int64_t k = rowStart[i]; k < rowStart[i + 1]; ++k) y += val[k] * x[colInd[k]];
        dense[i] = y;
    }
    /* DecompAlgo.cpp:5532-5540: cuts slide behind the convexity rows, which are zeroed but for the block */
    const int nCuts = R - nBase;
    for (int r = nRows - 1; r >= nRows - nCuts; --r) dense[r] = dense[r - numConvex];
    for (int b = 0; b < numConvex; ++b) dense[nBase + b] = 0.0;
    dense[nBase + blockId] = 1.0;
    /* UtilPackedVectorFromDense, UtilMacrosDecomp.cpp:46-60 */
    int nnz = 0;
    for (int i = 0; i < nRows; ++i) {
        if (fabs(dense[i]) > tolZero) {
            outRow[nnz] = i;
            outVal[nnz] = dense[i];
            ++nnz;
        }
    }
    free(x);
    free(dense);
    return nnz;
}

int dip_oracle_string_hash(int len, const int32_t *ind, const double *els, char *buf,
                           int bufLen)
{
    /* UtilHash.cpp:52-68: "ind_el_" per non-zero element, default precision 6 */
    int pos = 0;
    if (bufLen < 1) return -1;
    buf[0] = '\0';
    for (int i = 0; i < len; ++i) {
        if (fabs(els[i]) < 1.0e-8) continue;
        int k = snprintf(buf + pos, (size_t)(bufLen - pos), "%d_%.6g_", ind[i], els[i]);
        if (k < 0 || k >= bufLen - pos) return -1;
        pos += k;
    }
    return pos;
}

/* ------------------------------------------------- var pool / smoothing -- */

int dip_oracle_pool_reduced_costs(int nCols, const int64_t *colStart, const int32_t *rowInd,
                                  const double *val, const double *origCost, const double *u,
                                  int feasible, double *redCost)
{
    /* DecompVarPool.cpp:185-203: every waiting column; found |= (rc <= -1e-10) */
    int found = 0;
    for (int s = 0; s < nCols; ++s) {
        /* CoinPackedVectorBase::dotProduct (CoinUtils, un-vendored): last entry first */
        double dp = 0.0;
        for (int64_t k = colStart[s + 1] - 1; k >= colStart[s]; --k) dp += val[k] * u[rowInd[k]];
        /* DecompVarPool.cpp:29 (feasible) / :36 (dual ray) */
        const double rc = feasible ? origCost[s] - dp : -dp;
        redCost[s] = rc;
        if (rc <= -0.0000000001) found = 1; /* :30, :37 */
    }
    return found;
}

void dip_oracle_smooth_duals(int nRows, double alpha, const double *dual, const double *dualRM,
                             double *dualST)
{
    const double alpha1 = 1.0 - alpha; /* DecompAlgoPC.cpp:154 */
    for (int r = 0; r < nRows; ++r)    /* :179-181 */
        dualST[r] = (alpha * dual[r]) + (alpha1 * dualRM[r]);
}

/* ---------------------------------------------- node bounds in the blocks -- */

int dip_oracle_solve_blocks_bounded(int m, const int32_t *blockColStart, const int32_t *weight,
                                    const int32_t *capacity, const double *redCostX,
                                    const double *origCost, const double *alpha, const double *colLB,
                                    const double *colUB, double redCostEps, int maxBlockCols,
                                    int32_t *colNnz, double *colRedCost, double *colOrigCost,
                                    int32_t *colKeep, double *mostNegRC, double *z, int32_t *colIndOut,
                                    double *mostNegReducedCost, int64_t *cellsOut)
{
    int64_t cellsTotal = 0;
    for (int b = 0; b < m; ++b) {
        const int s = blockColStart[b];
        const int n = blockColStart[b + 1] - s;
        int32_t *ind = colIndOut ? colIndOut + (size_t)b * (size_t)maxBlockCols : NULL;
        /* DecompAlgo.cpp:6403-6405: the node's bounds on the block's columns */
        long left = capacity[b];
        int clash = 0;
        for (int j = 0; j < n; ++j) {
            if (colLB[s + j] > 0.5) left -= weight[s + j];
            if (colLB[s + j] > 0.5 && colUB[s + j] < 0.5) clash = 1;
        }
        if (clash || left < 0) { /* infeasible subproblem: no variable comes back */
            colNnz[b] = 0;
            colRedCost[b] = INFINITY;
            colOrigCost[b] = 0.0;
            if (z) z[b] = -INFINITY;
            continue;
        }
        /* free columns with negative reduced cost: gap_func.py:102-105 */
        int32_t *toOrig = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        int32_t *w = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        double *p = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        uint8_t *x = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
        int nItems = 0;
        for (int j = 0; j < n; ++j) {
            if (colLB[s + j] > 0.5) { x[j] = 1; continue; }
            if (colUB[s + j] < 0.5) continue;
            if (redCostX[s + j] < 0) {
                toOrig[nItems] = j;
                w[nItems] = weight[s + j];
                p[nItems] = -redCostX[s + j];
                ++nItems;
            }
        }
        double zb = 0.0;
        if (nItems > 0) {
            int32_t *sol = (int32_t *)malloc(sizeof(int32_t) * (size_t)nItems);
            int nSol = 0;
            zb = dip_oracle_knapsack01(nItems, p, w, (int)left, sol, &nSol, NULL);
            for (int k = 0; k < nSol; ++k) x[toOrig[sol[k]]] = 1;
            free(sol);
            for (int i = 0; i < nItems; ++i)
                if (w[i] <= left) cellsTotal += (int64_t)left + 1;
        }
        int nnz = 0;
        double rc = 0.0, oc = 0.0;
        for (int j = 0; j < n; ++j)
            if (x[j]) {
                rc += redCostX[s + j];
                oc += origCost[s + j];
                if (ind) ind[nnz] = s + j;
                ++nnz;
            }
        colNnz[b] = nnz;
        colRedCost[b] = rc - alpha[b]; /* DecompAlgo.cpp:6350-6356 */
        colOrigCost[b] = oc;
        if (z) z[b] = zb;
        free(toOrig);
        free(w);
        free(p);
        free(x);
    }
    /* DecompAlgo.cpp:4761, 5151-5232 */
    int kept = 0;
    double sum = 0.0;
    for (int b = 0; b < m; ++b) {
        double mn = 0.0;
        if (colRedCost[b] < mn) mn = colRedCost[b];
        mostNegRC[b] = mn;
        colKeep[b] = colRedCost[b] < -redCostEps ? 1 : 0;
        kept += colKeep[b];
    }
    for (int b = 0; b < m; ++b) sum += mostNegRC[b];
    if (mostNegReducedCost) *mostNegReducedCost = sum;
    if (cellsOut) *cellsOut = cellsTotal;
    return kept;
}

/* ------------------------------------------- multiple-choice knapsack blocks -- */

int dip_oracle_solve_mckp_blocks(int m, const int32_t *blockGroupStart, const int32_t *groupColStart,
                                 const int32_t *weight, const int32_t *capacity, const double *redCostX,
                                 const double *origCost, const double *alpha, double redCostEps,
                                 int maxBlockCols, int32_t *colNnz, double *colRedCost,
                                 double *colOrigCost, int32_t *colKeep, double *mostNegRC, double *z,
                                 int32_t *colIndOut, double *mostNegReducedCost, int64_t *cellsOut)
{
    int64_t cellsTotal = 0;
    for (int b = 0; b < m; ++b) {
        const int g0 = blockGroupStart[b], g1 = blockGroupStart[b + 1];
        const int G = g1 - g0;
        const int cap = capacity[b];
        const size_t W = (size_t)cap + 1;
        int32_t *ind = colIndOut ? colIndOut + (size_t)b * (size_t)maxBlockCols : NULL;
        double *prev = (double *)calloc(W, sizeof(double)); /* c[-1][.] = 0 */
        double *cur = (double *)malloc(W * sizeof(double));
        int32_t *choice = (int32_t *)malloc(sizeof(int32_t) * W * (size_t)(G > 0 ? G : 1));
        for (int g = 0; g < G; ++g) {
            const int cs = groupColStart[g0 + g], ce = groupColStart[g0 + g + 1];
            cellsTotal += (int64_t)(ce - cs) * (int64_t)W;
            for (int j = 0; j <= cap; ++j) {
                double best = INFINITY;
                int arg = -1;
                for (int k = cs; k < ce; ++k) {
                    if (weight[k] <= j) {
                        const double cand = redCostX[k] + prev[j - weight[k]];
                        if (cand < best) { best = cand; arg = k; } /* first minimal column */
                    }
                }
                cur[j] = best;
                choice[(size_t)g * W + (size_t)j] = arg;
            }
            double *t = prev; prev = cur; cur = t;
        }
        const double zb = prev[cap];
        if (G == 0 || !(zb < INFINITY)) { /* no group, or no feasible choice */
            colNnz[b] = 0;
            colRedCost[b] = G == 0 ? 0.0 - alpha[b] : INFINITY;
            colOrigCost[b] = 0.0;
            if (z) z[b] = G == 0 ? 0.0 : INFINITY;
        } else {
            int j = cap;
            for (int g = G - 1; g >= 0; --g) {
                const int k = choice[(size_t)g * W + (size_t)j];
                if (ind) ind[g] = k; /* groups ascend, so do the columns */
                j -= weight[k];
            }
            double rc = 0.0, oc = 0.0;
            for (int g = 0; g < G; ++g) {
                const int k = ind ? ind[g] : 0;
                rc += redCostX[k];
                oc += origCost[k];
            }
            colNnz[b] = G;
            colRedCost[b] = rc - alpha[b]; /* DecompAlgo.cpp:6350-6356 */
            colOrigCost[b] = oc;
            if (z) z[b] = zb;
        }
        free(prev);
        free(cur);
        free(choice);
    }
    int kept = 0;
    double sum = 0.0;
    for (int b = 0; b < m; ++b) {
        double mn = 0.0;
        if (colRedCost[b] < mn) mn = colRedCost[b];
        mostNegRC[b] = mn;
        colKeep[b] = colRedCost[b] < -redCostEps ? 1 : 0;
        kept += colKeep[b];
    }
    for (int b = 0; b < m; ++b) sum += mostNegRC[b];
    if (mostNegReducedCost) *mostNegReducedCost = sum;
    if (cellsOut) *cellsOut = cellsTotal;
    return kept;
}

/* ------------------------------------------- integer (unbounded) knapsack blocks -- */

/*
 * kp(obj, weights, capacity): the pricing knapsack of DIP's cutting-stock example,
 * Dip/src/dippy/tests/test_cutting_stock.py:154-179 (the copy with `if zyes > zbest`; the copy in
 * Dip/src/dippy/examples/csp/cutting_stock_func.py:92-118 compares the other way round and returns 0).
 * The reference is a plain recursion on the capacity:
 *     kp(0) = (0, zeros);   kp(c): best = kp(c-1);  for i = 0..n-1: if w_i <= c:
 *         zyes = kp(c - w_i).z + obj_i;  if zyes > zbest: best = (zyes, kp(c - w_i).sol + e_i)
 * i.e. F[c] = max(F[c-1], max_i F[c-w_i] + obj_i) with the FIRST candidate in the order
 * (skip, item 0, item 1, ...) that reaches the maximum, and the value of a cell is the chain sum
 * "value of the source cell + obj_i".  Memoised over the capacity that is this table; the counts
 * follow the recorded choices back from `capacity`.  Weights must be >= 1 (the recursion does not
 * terminate on a zero weight).  n == 0 returns 0 and no counts (:158-159).
 */
double dip_oracle_intknap_kp(int n, const double *obj, const int32_t *weights, int capacity, int32_t *counts)
{
    for (int i = 0; i < n; ++i) counts[i] = 0;
    if (n == 0 || capacity <= 0) return 0.0;
    double *F = (double *)malloc(sizeof(double) * ((size_t)capacity + 1));
    int32_t *choice = (int32_t *)malloc(sizeof(int32_t) * ((size_t)capacity + 1));
    F[0] = 0.0;
    choice[0] = -1;
    for (int c = 1; c <= capacity; ++c) {
        double zbest = F[c - 1]; /* :167 "don't include item" */
        int arg = -1;
        for (int i = 0; i < n; ++i) {
            if (weights[i] <= c) {                       /* :170 */
                const double zyes = F[c - weights[i]] + obj[i]; /* :171-173 */
                if (zyes > zbest) { zbest = zyes; arg = i; }    /* :175-177, strict */
            }
        }
        F[c] = zbest;
        choice[c] = arg;
    }
    int c = capacity;
    while (c > 0) {
        const int i = choice[c];
        if (i < 0) { --c; continue; }
        counts[i] += 1;
        c -= weights[i];
    }
    const double z = F[capacity];
    free(F);
    free(choice);
    return z;
}

/*
 * Blocks priced by kp: relaxed_solver of Dip/src/dippy/tests/test_cutting_stock.py:98-151 in the
 * convention of the C++ plugin surface (the variable's reduced cost comes back WITHOUT the convexity
 * dual, DecompAlgo::solveRelaxed subtracts it, Dip/src/DecompAlgo.cpp:6350-6356).  Block b owns the
 * item columns [blockColStart[b], blockColStart[b+1]) and optionally a "use" column useCol[b] (-1:
 * none): items = columns with redCost < 0 in ascending order (:103-104), obj = -redCost (:106),
 * (z, counts) = kp; when something is cut the column is {item: count} + {use: 1} with reduced cost
 * -z + redCost[use] (:129); otherwise an empty column with reduced cost 0 (:131).  origCost = sum of
 * c_j * x_j in ascending column order.  The filter is the one of generateVars.
 * Output per block: colNnz (entries), colIndOut / colCntOut (ascending x-space index, multiplicity),
 * rows of maxEntries.
 */
int dip_oracle_solve_intknap_blocks(int m, const int32_t *blockColStart, const int32_t *weight,
                                    const int32_t *capacity, const int32_t *useCol, const double *redCostX,
                                    const double *origCost, const double *alpha, double redCostEps,
                                    int maxEntries, int32_t *colNnz, double *colRedCost, double *colOrigCost,
                                    int32_t *colKeep, double *mostNegRC, double *z, int32_t *colIndOut,
                                    int32_t *colCntOut, double *mostNegReducedCost, int64_t *cellsOut)
{
    int64_t cellsTotal = 0;
    for (int b = 0; b < m; ++b) {
        const int cs = blockColStart[b], ce = blockColStart[b + 1];
        const int nAll = ce - cs;
        double *obj = (double *)malloc(sizeof(double) * (size_t)(nAll > 0 ? nAll : 1));
        int32_t *w = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nAll > 0 ? nAll : 1));
        int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nAll > 0 ? nAll : 1));
        int32_t *cnt = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nAll > 0 ? nAll : 1));
        int n = 0;
        for (int j = cs; j < ce; ++j) {
            if (redCostX[j] < 0.0) { obj[n] = -redCostX[j]; w[n] = weight[j]; idx[n] = j; ++n; }
        }
        const double zb = dip_oracle_intknap_kp(n, obj, w, capacity[b], cnt);
        cellsTotal += (int64_t)n * ((int64_t)capacity[b] + 1);
        int total = 0;
        for (int i = 0; i < n; ++i) total += cnt[i];
        int32_t *ind = colIndOut + (size_t)b * (size_t)maxEntries;
        int32_t *mul = colCntOut + (size_t)b * (size_t)maxEntries;
        int e = 0;
        double rc = 0.0, oc = 0.0;
        if (total > 0) {
            const int use = useCol ? useCol[b] : -1;
            int usePlaced = use < 0;
            for (int i = 0; i < n; ++i) {
                if (cnt[i] == 0) continue;
                if (!usePlaced && use < idx[i]) { ind[e] = use; mul[e] = 1; ++e; oc += origCost[use] * 1.0; usePlaced = 1; }
                ind[e] = idx[i]; mul[e] = cnt[i]; ++e;
                oc += origCost[idx[i]] * (double)cnt[i];
            }
            if (!usePlaced) { ind[e] = use; mul[e] = 1; ++e; oc += origCost[use] * 1.0; }
            rc = -zb;
            if (use >= 0) rc = rc + redCostX[use]; /* :129 */
        }
        colNnz[b] = e;
        colRedCost[b] = rc - alpha[b];
        colOrigCost[b] = oc;
        if (z) z[b] = zb;
        free(obj); free(w); free(idx); free(cnt);
    }
    int kept = 0;
    double sum = 0.0;
    for (int b = 0; b < m; ++b) {
        double mn = 0.0;
        if (colRedCost[b] < mn) mn = colRedCost[b];
        mostNegRC[b] = mn;
        colKeep[b] = colRedCost[b] < -redCostEps ? 1 : 0;
        kept += colKeep[b];
    }
    for (int b = 0; b < m; ++b) sum += mostNegRC[b];
    if (mostNegReducedCost) *mostNegReducedCost = sum;
    if (cellsOut) *cellsOut = cellsTotal;
    return kept;
}

/* ---- CPU-baseline variant (bench.py only): the reduced costs on several threads.  DIP's own sweep is the
 * single-threaded row-major scatter above; this one walks a column-ordered copy (entries of a column in DESCENDING
 * row order, so every column adds the same products in the same order and the result is bit-equal), the columns
 * dealt to the threads in contiguous ranges. */
void dip_oracle_calc_redcost_csc(int N, const int64_t *colStart, const int32_t *rowInd, const double *val,
                                 const double *uAdj, const double *c, int phase1, int nThreads, double *redCost)
{
    (void)nThreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nThreads > 0 ? nThreads : 1)
#endif
    for (int j = 0; j < N; ++j) {
        double y = 0.0;
        for (int64_t e = colStart[j]; e < colStart[j + 1]; ++e) {
            const double ui = uAdj[rowInd[e]];
            if (ui != 0.0) y += ui * val[e];
        }
        redCost[j] = phase1 ? -y : c[j] - y;
    }
}

/*
 * dip_oracle.h -- CPU restatement of DIP's column-generation pricing round.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (dip_b200/, include/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Parity pinning status (see DESIGN.md "Oracle"):
 *   - knapsack01 DP: PINNED against the reference's own Python function
 *     (Dip/src/dippy/examples/gap/gap_func.py:132-183), executed unchanged via
 *     AST extraction by oracle/gen_golden.py -> tests/golden/knapsack01_*.json,
 *     and cross-checked on integer-profit cases against the reference's
 *     Horowitz-Sahni B&B compiled from Dip/src/UtilKnapsack.cpp (oracle/_ref).
 *   - reduced costs: restates DecompAlgo.cpp:4506-4619; the inner product is
 *     CoinUtils' CoinPackedMatrix::transposeTimes (CoinUtils 2.11, NOT vendored
 *     in /root/reference) -> summation ORDER is "parity unpinned"; values are
 *     checked to 1e-12 relative, which any order satisfies.
 *   - Pisinger-mode argmax: combo.c is not vendored -> "parity unpinned" for
 *     the argmax on ties; optimal VALUES are unique (integer profits).
 *
 * All paths below are relative to /root/reference.
 */
#ifndef DIP_ORACLE_H
#define DIP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DecompAlgo::generateVarsAdjustDuals, Dip/src/DecompAlgo.cpp:4550-4619.
 * uOld has nMasterRows entries laid out [core rows 0..nBase) | numConvex
 * convexity rows | cut rows]; uNew gets nMasterRows - numConvex entries. */
void dip_oracle_adjust_duals(const double *uOld, int nMasterRows, int nBaseCoreRows,
                             int numConvex, double *uNew);

/* DecompAlgo::generateVarsCalcRedCost, Dip/src/DecompAlgo.cpp:4506-4547, with
 * CoinPackedMatrix::transposeTimes restated as the row-major scatter
 * (rows descending, zero duals skipped).  phase1 != 0 -> c is taken as 0. */
void dip_oracle_calc_redcost(int R, int N, const int64_t *rowStart, const int32_t *colInd,
                             const double *val, const double *uAdj, const double *c,
                             int phase1, double *redCostX);

/* knapsack01, Dip/src/dippy/examples/gap/gap_func.py:132-183.
 * max sum obj*x, sum w*x <= capacity; strict '>' take rule, backtrack from
 * (n-1, capacity).  solution[] receives the chosen item indices in the
 * reference's (descending) order, *nSol their count.  added (optional,
 * n*(capacity+1) bytes) receives the decision table.  Returns z = c[n-1][cap].
 * Documented divergence: n==1 is evaluated as a standard 0-1 DP (the Python
 * aliases row -1 onto row 0 for n==1). */
double dip_oracle_knapsack01(int n, const double *obj, const int32_t *weights, int capacity,
                             int32_t *solution, int *nSol, uint8_t *added);

/* GAP_KnapPisinger::calcScaleFactor, Dip/examples/GAP/GAP_KnapPisinger.h:89-143
 * (UtilFracPart/UtilIsZero from Dip/src/UtilMacros.h:521-552). */
int dip_oracle_calc_scale_factor(int n, const double *redCost, double *work);

/* One block of pricing = GAP_DecompApp::solveRelaxed
 * (Dip/examples/GAP/GAP_DecompApp.cpp:235-285) around a knapsack solve.
 *   profitMode 0: items = {j : redCost[j] < 0} ascending, obj = -redCost (fp64),
 *                 solved by knapsack01 (gap_func.py:96-130 solve_subproblem).
 *   profitMode 1: GAP_KnapPisinger::solve (GAP_KnapPisinger.h:146-273): scaled
 *                 integer profits, n==1 / all-fit / n==2 shortcuts, else exact
 *                 solve (combo.c is not vendored: the DP stands in; same z*).
 * redCost/origCost/weight point at the block's slice (length n).  colOffset is
 * added to the local index to form the x-space index.  ind[] (capacity n) gets
 * the chosen indices ASCENDING; varRedCost / varOrigCost are summed in that
 * order and exclude alpha.  *zOut = knapsack optimum (in DP profit units).
 * Returns the column's nnz. */
int dip_oracle_solve_block(int n, const double *redCost, const double *origCost,
                           const int32_t *weight, int capacity, int colOffset, int profitMode,
                           int32_t *ind, double *varRedCost, double *varOrigCost, double *zOut,
                           int64_t *cellsOut);

/* A full pricing round = DecompAlgo::generateVars, Dip/src/DecompAlgo.cpp:4622-5260
 * for an application whose blocks are knapsacks over contiguous column ranges.
 *   u: master duals, [0,nBase) core rows | m convexity | cuts  (length R + m)
 *   outputs (caller-allocated):
 *     redCostX[N]; per block: colNnz[m], colRedCost[m] (= sum - alpha,
 *     DecompAlgo.cpp:6350-6356), colOrigCost[m], colKeep[m] (rc < -eps, :5216),
 *     mostNegRC[m] (= min(0, rc), :4761 + :5212-5214), z[m];
 *     colInd: m * maxBlockCols entries, block b's indices start at
 *     b*maxBlockCols.  *mostNegReducedCost = sum_b mostNegRC[b] in block order
 *     (:5229-5232).  nThreads > 1 runs blocks under an OpenMP dynamic,1
 *     schedule like DecompAlgo.cpp:4815.  Returns number of kept columns. */
int dip_oracle_price_round(int R, int N, const int64_t *rowStart, const int32_t *colInd,
                           const double *val, const double *c, const double *u, int nBase,
                           int phase1, int m, const int32_t *blockColStart,
                           const int32_t *weight, const int32_t *capacity, int profitMode,
                           double redCostEps, int nThreads, int maxBlockCols,
                           double *redCostX, int32_t *colNnz, double *colRedCost,
                           double *colOrigCost, int32_t *colKeep, double *mostNegRC,
                           double *z, int32_t *colIndOut, double *mostNegReducedCost,
                           int64_t *cellsOut);

/* Stage (b)+(c) only, given redCostX and alpha (the loop of
 * DecompAlgo::solveRelaxed calls + filter).  Same outputs as above. */
int dip_oracle_solve_blocks(int m, const int32_t *blockColStart, const int32_t *weight,
                            const int32_t *capacity, int profitMode, const double *redCostX,
                            const double *origCost, const double *alpha, double redCostEps,
                            int nThreads, int maxBlockCols, int32_t *colNnz,
                            double *colRedCost, double *colOrigCost, int32_t *colKeep,
                            double *mostNegRC, double *z, int32_t *colIndOut,
                            double *mostNegReducedCost, int64_t *cellsOut);

/* Master column of a variable = DecompAlgo::addVarsToPool, Dip/src/DecompAlgo.cpp:5452-5558:
 *   fillDenseArr (Dip/src/DecompVar.cpp:58-70) -> modelCore->M->times(s, denseCol) on the row-ordered
 *   A'' (y[i] = sum over row i's stored entries, in storage order, of val * s[col]) -> the cut rows
 *   (core rows >= nBase) move behind the numConvex convexity rows (:5532-5534), 1.0 in the block's
 *   convexity row (:5540) -> UtilPackedVectorFromDense keeps |v| > tolZero, ascending row
 *   (Dip/src/UtilMacrosDecomp.cpp:46-60).  els == NULL means all elements are 1.0.
 * outRow/outVal need R + numConvex entries.  Returns the number of nonzeros. */
int dip_oracle_expand_column(int R, int N, const int64_t *rowStart, const int32_t *colInd,
                             const double *val, int nBase, int numConvex, int nInd,
                             const int32_t *ind, const double *els, int blockId, double tolZero,
                             int32_t *outRow, double *outVal);

/* UtilCreateStringHash, Dip/src/UtilHash.cpp:52-68 (ind_el_ at 6 significant
 * digits) for a 0/1 column: writes a NUL-terminated string, returns its length
 * (excluding NUL), or -1 if bufLen is too small. */
int dip_oracle_string_hash(int len, const int32_t *ind, const double *els, char *buf,
                           int bufLen);

/* Stage (b)+(c) under node bounds enforced in the subproblem (BranchEnforceInSubProb,
 * Dip/src/DecompParam.h:656): DecompAlgo::setSubProbBounds stores the node's column bounds
 * (Dip/src/DecompAlgo.cpp:2352-2371) and DecompAlgo::solveRelaxed imposes them on the block's
 * subproblem (setActiveColBounds, :6403-6405), which then is
 *     min sum redCost_j x_j  s.t.  sum w_j x_j <= cap,  colLB_j <= x_j <= colUB_j,  x binary.
 * Restated exactly: columns with colLB > 0.5 are in (whatever their reduced cost), their weight comes
 * off the capacity; columns with colUB < 0.5 are out; the rest is knapsack01 over the columns with
 * redCost < 0 (fp64 profit mode).  Column = fixed-to-1 + chosen, ascending; z excludes the fixed
 * columns.  A block whose fixed columns do not fit (or with colLB > colUB) is infeasible: colNnz 0,
 * colRedCost +inf, mostNegRC 0, not kept, z -inf (the reference's MILP pushes no variable).  Which
 * optimal solution a MILP solver returns on ties is "parity unpinned"; ours is the DP's.
 * colLB / colUB are indexed by x-space column.  Other arguments as dip_oracle_solve_blocks. */
int dip_oracle_solve_blocks_bounded(int m, const int32_t *blockColStart, const int32_t *weight,
                                    const int32_t *capacity, const double *redCostX,
                                    const double *origCost, const double *alpha, const double *colLB,
                                    const double *colUB, double redCostEps, int maxBlockCols,
                                    int32_t *colNnz, double *colRedCost, double *colOrigCost,
                                    int32_t *colKeep, double *mostNegRC, double *z, int32_t *colIndOut,
                                    double *mostNegReducedCost, int64_t *cellsOut);

/* Multiple-choice knapsack blocks = MMKP_DecompApp::solveRelaxed with the MCKP0 relaxation
 * (Dip/examples/MMKP/MMKP_DecompApp.cpp:493-557 -> MMKP_MCKnap::solveMCKnap, MMKP_MCKnap.cpp:172-383):
 * block b = groups [blockGroupStart[b], blockGroupStart[b+1]); group g = the contiguous x-space columns
 * [groupColStart[g], groupColStart[g+1]); EXACTLY one column per group, sum of weights <= capacity,
 * minimise sum redCost (:186-196 negates and offsets to feed a max solver).  The reference's solver is
 * Pisinger's mcknap (fetched by get_pisinger.sh, NOT vendored) behind a lossy double -> int scaling of
 * the costs (UtilScaleDblToIntArr with MCKP_EPSILON 1e-3, :213) -- "parity unpinned".  Restated as the
 * exact fp64 DP over the groups in order:
 *     c[g][j] = min over the group's columns k with w_k <= j of redCost_k + c[g-1][j - w_k],  c[-1][.] = 0
 * first minimal column wins ('<'); backtrack from (last group, capacity).  A block with no feasible
 * choice returns no column (colNnz 0, colRedCost +inf, mostNegRC 0, z +inf; the reference asserts).
 * z = c[G-1][cap] = the column's sum of reduced costs (alpha excluded).  cells = columns * (cap + 1).
 * Outputs as dip_oracle_solve_blocks; colIndOut holds the chosen columns ascending. */
int dip_oracle_solve_mckp_blocks(int m, const int32_t *blockGroupStart, const int32_t *groupColStart,
                                 const int32_t *weight, const int32_t *capacity, const double *redCostX,
                                 const double *origCost, const double *alpha, double redCostEps,
                                 int maxBlockCols, int32_t *colNnz, double *colRedCost,
                                 double *colOrigCost, int32_t *colKeep, double *mostNegRC, double *z,
                                 int32_t *colIndOut, double *mostNegReducedCost, int64_t *cellsOut);

/* DecompVarPool::setReducedCosts -> DecompWaitingCol::setReducedCost,
 * Dip/src/DecompVarPool.cpp:185-203, :22-40 (called once per price call, DecompAlgo.cpp:1890-1908):
 *   feasible:  redCost_s = origCost_s - m_col.dotProduct(u)        (:29)
 *   dual ray:  redCost_s =            - m_col.dotProduct(u)        (:36)
 * m_col = the waiting column in MASTER row space (incl. the convexity 1.0), u = the whole master
 * dual vector.  CoinPackedVectorBase::dotProduct is CoinUtils 2.11 (not vendored): restated as
 * dp = 0; for i = n-1..0: dp += els[i]*u[ind[i]]  (order "parity unpinned"; any order meets 1e-12).
 * redCost[nCols] receives the values; returns 1 iff some redCost <= -1e-10 (:30, :37, :199). */
int dip_oracle_pool_reduced_costs(int nCols, const int64_t *colStart, const int32_t *rowInd,
                                  const double *val, const double *origCost, const double *u,
                                  int feasible, double *redCost);

/* DecompAlgoPC::adjustMasterDualSolution, Dip/src/DecompAlgoPC.cpp:126-212 (Wentges smoothing):
 *   dualST[r] = (alpha * dual[r]) + ((1.0 - alpha) * dualRM[r])    (:153-154, :179-181) */
void dip_oracle_smooth_duals(int nRows, double alpha, const double *dual, const double *dualRM,
                             double *dualST);

#ifdef __cplusplus
}
#endif
#endif

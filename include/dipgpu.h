/*
 * dipgpu.h -- C ABI of libdipgpu: the B200 (sm_100a) pricing round behind DIP's
 * DecompApp::solveRelaxed / DecompAlgo::generateVars plugin surface.
 *
 * Plain C: opaque context, caller-owned host buffers, plain pointers and sizes,
 * integer status codes (never throws).  One context drives one GPU (dipgpu_create)
 * or several GPUs of one box (dipgpu_create_multi) from one host thread.  There is
 * NO CPU fallback: every entry point that computes
 * returns DIPGPU_ERR_CUDA when no sm_100 device / driver is usable.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * the coin-or/Dip tree).  INTEGRATION.md shows the reference-side binding.
 */
#ifndef DIPGPU_H
#define DIPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the library is built with -fvisibility=hidden; only this header is exported */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define DIPGPU_ABI_VERSION 4

typedef struct dipgpu_ctx dipgpu_ctx;

enum {
    DIPGPU_OK = 0,
    DIPGPU_ERR_INVALID = 1,     /* bad argument (NULL, negative size, unsorted starts, w < 0 ...) */
    DIPGPU_ERR_CUDA = 2,        /* CUDA runtime / driver error, or no usable sm_100 device */
    DIPGPU_ERR_NOMEM = 3,       /* device or pinned-host allocation failed */
    DIPGPU_ERR_STATE = 4,       /* call order: model / blocks / objective not set yet */
    DIPGPU_ERR_UNSUPPORTED = 5, /* e.g. node bounds on multiple-choice blocks, more than 2^31 stored entries */
    DIPGPU_ERR_OVERFLOW = 6     /* caller's output buffers too small; required sizes are reported */
};

/* Profit arithmetic of the block knapsacks. */
enum {
    /* fp64 profits = -redCost, the reference's own DP: knapsack01(),
     * Dip/src/dippy/examples/gap/gap_func.py:132-183 (+ :96-130). */
    DIPGPU_PROFIT_FP64 = 0,
    /* scaled-integer profits + n==1 / all-fit / n==2 shortcuts of
     * GAP_KnapPisinger::solve, Dip/examples/GAP/GAP_KnapPisinger.h:89-273. */
    DIPGPU_PROFIT_PISINGER_INT = 1
};

/* DecompPhase as far as pricing cares (Dip/src/Decomp.h:170-176):
 * phase 1 prices with c = 0 (Dip/src/DecompAlgo.cpp:4538-4541). */
enum { DIPGPU_PHASE_PRICE1 = 1, DIPGPU_PHASE_PRICE2 = 2 };

/*
 * Result of one pricing round: what DecompAlgo::generateVars hands back
 * (Dip/src/DecompAlgo.cpp:4622-5260) in flat arrays.  All pointers are
 * caller-owned host memory; capacities are inputs, counts are outputs.
 *
 * Per block b (arrays of length >= m, may be NULL to skip):
 *   blockRedCost[b]  = sum_j redCost[j] x_j - alpha_b   (DecompAlgo.cpp:6350-6356)
 *   blockMostNegRC[b]= min(0, blockRedCost[b])          (:4761, :5212-5214)
 *   blockOrigCost[b] = sum_j c[j] x_j                   (GAP_KnapPisinger.h:268)
 *   blockObj[b]      = knapsack optimum in DP profit units
 *   blockNnz[b]      = |column b|
 * Kept columns (rc < -redCostEps, :5216), in block order:
 *   colBlock[k] (DecompVar::setBlockId), colRedCost[k], colOrigCost[k],
 *   colStart[k]..colStart[k+1] into colInd: x-space indices ascending; colEl (may
 *   be NULL) receives the elements: 1.0 for the 0-1 and multiple-choice blocks
 *   (GAP_KnapPisinger.h:265-270), the multiplicities for integer knapsack blocks.
 * mostNegReducedCost = sum_b blockMostNegRC[b], added in block order (:5229-5232).
 */
typedef struct dipgpu_cols {
    /* in: capacities of the arrays below */
    int32_t capBlocks; /* length of the block* arrays (>= m) */
    int32_t capCols;   /* length of colBlock/colRedCost/colOrigCost; colStart has capCols+1 */
    int64_t capInd;    /* length of colInd */
    /* out: per block */
    double *blockRedCost;
    double *blockMostNegRC;
    double *blockOrigCost;
    double *blockObj;
    int32_t *blockNnz;
    /* out: kept columns */
    int32_t nCols;
    int32_t *colBlock;
    double *colRedCost;
    double *colOrigCost;
    int64_t *colStart;
    int32_t *colInd;
    int64_t nInd; /* total indices of the kept columns (required capInd on overflow) */
    double mostNegReducedCost;
    int64_t cells; /* DP cell updates of the reference's DP: sum over active items of (capacity+1) */
    /* cell updates actually evaluated: small shards first drop the items that no optimal solution
     * can contain (a Lagrangian bound against a greedy lower bound, the reduction Pisinger's combo
     * performs before its DP); optimum, column and ties are those of the full DP. */
    int64_t cellsEvaluated;
    /* out, optional (NULL to skip): element of every kept index, parallel to colInd (capInd entries) */
    double *colEl;
} dipgpu_cols;

/* ---------------------------------------------------------------- lifetime */

int dipgpu_abi_version(void);

/* Opens CUDA device `device`, requires compute capability 10.x. */
int dipgpu_create(dipgpu_ctx **ctx, int device);
int dipgpu_destroy(dipgpu_ctx *ctx);

/*
 * One context over nDevices GPUs of one box ("int dipgpu_create(dipgpu_ctx**, const int* devices,
 * int ndev)"), driven by the SAME caller thread through the SAME entry points -- DIP is one process whose
 * generateVars loops over the independent blocks (under OpenMP with SubProbParallel,
 * Dip/src/DecompAlgo.cpp:4815-4848) and whose DecompAlgoPC prices whole ladders of smoothed dual vectors
 * (Dip/src/DecompAlgo.cpp:5642-5655).  devices[0] is the primary device (an ordinal may be listed more than once:
 * its shards then share that GPU).  The model calls (set_core_csr,
 * set_objective, set_knapsack_blocks, append_core_rows, set_col_bounds, set_dual_support, set_option) reach
 * every device; the pricing calls are partitioned:
 *   dipgpu_price_round, dipgpu_solve_blocks, dipgpu_calc_redcost, dipgpu_price_round_expand
 *       the BLOCKS (contiguous ranges balanced by n_b*(cap_b+1), dipgpu_shard_ranges): device d computes the
 *       reduced costs of its blocks' columns only, solves its blocks, filters; the result is merged in block
 *       order on the host (mostNegReducedCost added in block order, Dip/src/DecompAlgo.cpp:5229-5232).
 *       By default (option "shard_rounds" = -1) the blocks are partitioned only when that can pay: a shape whose
 *       blocks all run side by side on ONE GPU (one CTA per block, all co-resident: GAP-E's 80 blocks on 148 SMs)
 *       is a set of equally long latency chains -- no partition shortens it, each one adds the NVLink exchange --
 *       so such rounds stay on the primary device; "shard_rounds" = k >= 1 forces min(k, nDevices) devices;
 *   dipgpu_price_rounds_batch, dipgpu_price_round_stab, dipgpu_price_round_stab_ladder
 *       the DUAL VECTORS: device d prices all blocks for its share of the vectors;
 *   everything else (waiting-column pool, dipgpu_price_round_which, u == NULL rounds) runs on the primary.
 * Exchange per call: every device reads the duals it needs out of the caller's buffer (u[support] when a dual
 * support is declared: 34 KB on GAP-E instead of 2 MB) and writes its part of the packed result into mapped host
 * memory itself; no collective.  Results are bit-identical to the one-GPU context.  Internally one worker
 * thread per additional device issues that device's launches.
 */
int dipgpu_create_multi(dipgpu_ctx **ctx, const int32_t *devices, int32_t nDevices);
/* Devices behind a context (1 for dipgpu_create); peerAccess = 1 when the primary and every other device can
 * address each other's memory (NVLink / PCIe peer-to-peer). */
int dipgpu_device_count(const dipgpu_ctx *ctx, int32_t *nDevices, int32_t *peerAccess);
/* Block range [firstBlock[d], firstBlock[d] + blockCount[d]) of device d (arrays of length cap). */
int dipgpu_shard_ranges(const dipgpu_ctx *ctx, int32_t cap, int32_t *firstBlock, int32_t *blockCount);

/* Copies the last error text of this context (NUL-terminated) into buf. */
int dipgpu_last_error(const dipgpu_ctx *ctx, char *buf, size_t bufLen);

/* Pinned host memory for the dual vector / outputs (optional: any host pointer
 * is accepted, pinned ones are copied without a staging pass). */
int dipgpu_host_alloc(void **ptr, size_t bytes);
int dipgpu_host_free(void *ptr);

/* ------------------------------------------------------------------- model */

/*
 * Core constraint matrix A'' in the layout DIP holds it: row-ordered
 * CoinPackedMatrix (DecompConstraintSet::M, Dip/src/DecompConstraintSet.h:32,
 * forced row-major at DecompConstraintSet.cpp:41-43).  R rows incl. the
 * 2*nInt branch rows appended by coreMatrixAppendColBounds
 * (Dip/src/DecompAlgo.cpp:1399-1502); nBaseRows = modelCore->nBaseRows
 * (rows >= nBaseRows are cuts).  numConvex = m_numConvexCon: the master dual
 * vector is laid out [0,nBaseRows) | numConvex convexity rows | cut rows
 * (Dip/src/DecompAlgo.cpp:4550-4619), and that is how `u` is read below.
 */
int dipgpu_set_core_csr(dipgpu_ctx *ctx, int32_t R, int32_t N, const int64_t *rowStart,
                        const int32_t *colInd, const double *val, int32_t nBaseRows,
                        int32_t numConvex);

/* Cut rows appended to A'' during the run (DecompConstraintSet::appendRow,
 * Dip/src/DecompConstraintSet.h:143-163; DecompAlgo.cpp:6074).  rowStart has
 * nNew+1 entries starting at 0. */
int dipgpu_append_core_rows(dipgpu_ctx *ctx, int32_t nNew, const int64_t *rowStart,
                            const int32_t *colInd, const double *val);

/* Original objective c (DecompApp::m_objective, Dip/src/DecompApp.h:85). */
int dipgpu_set_objective(dipgpu_ctx *ctx, const double *c, int32_t N);

/*
 * The relaxation: m independent 0-1 knapsacks, block b owning the contiguous
 * x-space columns [blockColStart[b], blockColStart[b+1]) with integer weights
 * and capacity -- what GAP_DecompApp builds one GAP_KnapPisinger per block from
 * (Dip/examples/GAP/GAP_DecompApp.cpp:52-58, GAP_KnapPisinger.h:276-294).
 * weight is indexed by x-space column (length blockColStart[m]).
 */
int dipgpu_set_knapsack_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockColStart,
                               const int32_t *weight, const int32_t *capacity,
                               int32_t profitMode);

/*
 * The relaxation as m independent MULTIPLE-CHOICE knapsacks -- the MCKP0 relaxation of DIP's MMKP example
 * (Dip/examples/MMKP/MMKP_DecompApp.cpp:493-557 -> MMKP_MCKnap::solveMCKnap,
 * Dip/examples/MMKP/MMKP_MCKnap.cpp:172-383): block b = the groups [blockGroupStart[b], blockGroupStart[b+1]),
 * group g = the contiguous x-space columns [groupColStart[g], groupColStart[g+1]) (1..65534 of them);
 * exactly one column per group, sum of weights <= capacity[b], minimise the sum of reduced costs.  One
 * column (one index per group) comes back per block; a block without a feasible choice returns none
 * (blockRedCost = +inf, blockMostNegRC = 0, blockObj = +inf, blockNnz = 0).  blockObj = the minimum (alpha
 * excluded).  Replaces dipgpu_set_knapsack_blocks (one kind of block per context); every pricing entry point
 * then solves these blocks.  The reference's solver (Pisinger's mcknap behind an integer scaling of the
 * costs) is not vendored: the exact fp64 DP over the groups takes its place, first minimal column on ties.
 */
int dipgpu_set_mckp_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockGroupStart,
                           const int32_t *groupColStart, const int32_t *weight, const int32_t *capacity);

/*
 * The relaxation as m independent INTEGER (unbounded) knapsacks -- the pricing problem of DIP's cutting-stock
 * example: relaxed_solver and kp of Dip/src/dippy/tests/test_cutting_stock.py:98-151, :154-179 (the copy of kp in
 * Dip/src/dippy/examples/csp/cutting_stock_func.py:92-118 compares the wrong way round and returns 0).  Block b owns
 * the item columns [blockColStart[b], blockColStart[b+1]) with weights >= 1 (the piece lengths) and, optionally, a
 * "use" column useCol[b] anywhere else in x-space (-1 or useCol == NULL: none): the roll of
 * sum_i w_i x_i <= capacity_b * use_b.  Items are the columns with redCost < 0, profit -redCost;
 *     F[c] = max(F[c-1], max_i F[c - w_i] + profit_i),  first maximum in the order (skip, item 0, item 1, ...),
 * evaluated cell by cell in that order so that value AND multiplicities are the recursion's, bit for bit.  When
 * something is cut the column is {item: multiplicity} plus {use: 1} and varRedCost = -z + redCost[use]; otherwise the
 * column is empty with varRedCost 0 (then blockRedCost = -alpha_b).  blockObj = z; blockOrigCost = sum c_j x_j in
 * ascending column order; colEl carries the multiplicities.  The reduced cost excludes the convexity dual as on the
 * C++ plugin surface (the Python relaxed_solver subtracts convexDual itself).  Replaces the other block kinds of the
 * context; dipgpu_expand_columns and dipgpu_set_col_bounds do not apply (DIPGPU_ERR_UNSUPPORTED).
 */
int dipgpu_set_intknap_blocks(dipgpu_ctx *ctx, int32_t m, const int32_t *blockColStart, const int32_t *weight,
                              const int32_t *capacity, const int32_t *useCol);

/*
 * Node bounds of the x-space columns, enforced inside the block subproblems: what
 * DecompAlgo::setSubProbBounds stores in m_colLBNode / m_colUBNode (Dip/src/DecompAlgo.cpp:2352-2371)
 * and DecompAlgo::solveRelaxed hands to the subproblem with BranchEnforceInSubProb
 * (setActiveColBounds, Dip/src/DecompAlgo.cpp:6403-6405; Dip/src/DecompModel.h:149-187) -- DIP's
 * default branching mode (Dip/src/DecompParam.h:656).  For the 0-1 knapsack blocks a column with
 * colUB < 0.5 never enters a column, a column with colLB > 0.5 is in every column of its block
 * (whatever its reduced cost) and its weight comes off the capacity; the block's optimum
 * min sum redCost*x over the remaining columns is the same DP.  A block whose fixed columns do not
 * fit returns no column: blockRedCost = +inf, blockMostNegRC = 0, blockObj = -inf, blockNnz = 0
 * (the reference's subproblem MILP is infeasible and pushes no variable).  blockObj excludes the
 * fixed columns' profit.  Length N >= the block columns; both NULL removes the bounds.  Call after
 * dipgpu_set_knapsack_blocks (which clears them).  In Pisinger mode the scale factor, the n <= 2 / all-fit cases
 * and the integer profits of GAP_KnapPisinger::solve are those of the FREE columns against the capacity left.
 * 0-1 knapsack blocks only (DIPGPU_ERR_UNSUPPORTED for multiple-choice blocks).
 */
int dipgpu_set_col_bounds(dipgpu_ctx *ctx, const double *colLB, const double *colUB, int32_t N);

/* ----------------------------------------------------------------- pricing */

/*
 * Stage (a): generateVarsAdjustDuals + generateVarsCalcRedCost
 * (Dip/src/DecompAlgo.cpp:4550-4619, 4506-4547).  u = master duals of length
 * R + numConvex; redCostX (length N) receives c - A''^T u (phase 2) or
 * -A''^T u (phase 1).
 */
int dipgpu_calc_redcost(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase,
                        double *redCostX);

/*
 * Stages (b)+(c) for a caller-supplied reduced-cost vector: the loop of
 * DecompApp::solveRelaxed(whichBlock, redCostX, alpha_b, varList) calls
 * (Dip/src/DecompApp.h:344-349; GAP_DecompApp.cpp:235-285) over all blocks,
 * followed by the filter of DecompAlgo.cpp:5151-5232.  alpha has m entries.
 * redCostEps = DecompParam::RedCostEpsilon (1e-4); pass -INFINITY to keep every
 * candidate (per-block plugin surface).
 */
int dipgpu_solve_blocks(dipgpu_ctx *ctx, const double *redCostX, const double *alpha,
                        double redCostEps, dipgpu_cols *out);

/*
 * The whole pricing round = DecompAlgo::generateVars
 * (Dip/src/DecompAlgo.h:430, DecompAlgo.cpp:4622-5260): (a) then (b) then (c)
 * in one submission.  alpha_b = u[nBaseRows + b] (:4820).  redCostXOpt (length
 * N) may be NULL.
 */
int dipgpu_price_round(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase,
                       double redCostEps, dipgpu_cols *out, double *redCostXOpt);
/* u == NULL prices with the dual vector already resident on the device: the one uploaded by the
 * last host-buffer call (dipgpu_pool_set_reduced_costs, dipgpu_calc_redcost, dipgpu_price_round) --
 * DIP re-prices the waiting columns and generates new ones with the same duals
 * (Dip/src/DecompAlgo.cpp:1901 then :1914), so the vector crosses PCIe once. */

/*
 * Rows of the master dual vector that can be nonzero (ascending, unique; nIdx = 0 removes the support).
 * DIP keeps one pair of branch rows per integer column in the master and relaxes all of them to free rows at
 * the root, at the other nodes all but the rows of the bounds branched on (Dip/src/AlpsDecompTreeNode.cpp:215-238,
 * DecompAlgo::setMasterBounds Dip/src/DecompAlgo.cpp:2373-2520); a free row's dual is 0.  Between two
 * setMasterBounds calls the duals therefore live on the original rows, the few bounded branch rows, the
 * convexity rows and the cuts -- on GAP-E about 4 000 of 257 680 entries.  With a support set, every entry
 * point that takes host duals gathers u[idx] into pinned staging, uploads only that and scatters it on the
 * device (the device vector stays zero elsewhere): 34 KB instead of 2 MB per round.  The caller guarantees
 * u == 0 outside the support.  Cleared by dipgpu_set_core_csr / dipgpu_append_core_rows (the vector changes).
 */
int dipgpu_set_dual_support(dipgpu_ctx *ctx, int32_t nIdx, const int32_t *idx);

/* ------------------------------------------- a subset of the blocks, per-block user duals
 *
 * DecompAlgo::generateVars' round-robin branch (Dip/src/DecompAlgo.cpp:4929-5149; taken in phase 2 while
 * m_rrIterSinceAll < RoundRobinInterval, :4755): DecompApp::solveRelaxedWhich (Dip/src/DecompApp.h:339-342)
 * names the blocks to solve this round and may hand a different master dual vector to some of them
 * (userDualsByBlock; the user-replaced vector of getDualForGenerateVars, DecompApp.h:330-332, is simply `u`).
 *   whichBlocks[i]  global block ids (this context's block range; no duplicates), any order
 *   userDuals       NULL, or nWhich pointers: userDuals[i] == NULL -> block i is priced with u, else with that
 *                   vector (length uLen, DIP's layout [core rows | convexity rows]): reduced costs c - A''^T uhat
 *                   for the block's columns and alpha = uhat[nBaseRows + block]
 * The result has the layout of a normal round: only the listed blocks carry values (the others: no column,
 * zeros), mostNegReducedCost sums min(0, rc) over the listed blocks (mostNegRCvec starts at 0, :4761).
 * If no listed block prices out (rc < -redCostEps), DIP solves ALL blocks with u (:5076-5148): the call then
 * returns that full round, *solvedAll = 1, and a block priced with user duals contributes the smaller of its
 * two reduced costs to mostNegReducedCost, as the reference's filter loop does (:5212-5214).
 * Differences, stated: kept columns come in ascending block order (the reference: list order); the reference
 * reads alpha from the ADJUSTED user vector at [nBaseRows + block], an index past that array's convexity-free
 * length (:5032) -- here alpha is the user vector's convexity dual, which is what the block's subproblem needs.
 * Cost: one batch submission with one vector per distinct dual pointer; the blocks of a shard run side by
 * side, so pricing a subset takes as long as pricing them all.  The composed round becomes the context's
 * "last round" (dipgpu_expand_columns).
 */
int dipgpu_price_round_which(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double redCostEps,
                             int32_t nWhich, const int32_t *whichBlocks, const double *const *userDuals,
                             dipgpu_cols *out, int32_t *solvedAll);

/* ------------------------------------------- several dual vectors in one submission
 *
 * nVec master dual vectors (row v at U + v*uLen), each priced exactly as dipgpu_price_round would:
 * outs[v] receives the round of vector v.  One reduced-cost launch, one DP launch (the fused kernel
 * with gridDim.y = nVec) and one strided D2H copy for the whole batch, so the latency-bound GAP
 * rounds fill the GPU.  Sources of such batches in DIP: the smoothing ladder below, and the dual
 * vectors of several open nodes.  dipgpu_select_batch_result(which) makes result `which` the
 * context's "last round" (for dipgpu_expand_columns).
 */
int dipgpu_price_rounds_batch(dipgpu_ctx *ctx, int32_t nVec, const double *U, int32_t uLen,
                              int32_t phase, double redCostEps, dipgpu_cols *outs);
int dipgpu_select_batch_result(dipgpu_ctx *ctx, int32_t which);

/* ------------------------------------------------------- dual stabilisation (Wentges)
 *
 * DecompAlgoPC::adjustMasterDualSolution (Dip/src/DecompAlgoPC.cpp:126-212) followed by
 * generateVars with the smoothed duals (DecompAlgoPC.h:99-108):
 *     dualST = alpha * dual + (1 - alpha) * dualRM                       (:179-181)
 * `dual` (m_dual, the stability centre) lives on the device: dipgpu_stab_set_center uploads it
 * (first price call / first phase-2 call, :163-177, where the reference copies dualRM or calls
 * DecompApp::initDualVector); dipgpu_stab_accept(step) copies smoothed vector `step` of the last
 * stabilised call into it (DecompAlgoPC::setObjBound, DecompAlgoPC.h:124-133, when the bound
 * improved).  Without a centre the first stabilised call uses dualRM itself (:172).
 *
 * dipgpu_price_round_stab: one smoothing parameter; dualSTOpt (length uLen, may be NULL) receives
 * the smoothed duals the round was priced with (the host needs them for updateObjBound,
 * Dip/src/DecompAlgo.cpp:3728).
 *
 * dipgpu_price_round_stab_ladder: when every new column is a duplicate the reference multiplies
 * DualStabAlpha by 0.90 and prices again (Dip/src/DecompAlgo.cpp:5642-5655).  The ladder
 * alpha_k = alpha * shrink^k, k < nSteps (repeated multiplication, as the reference) is priced
 * speculatively in ONE submission from ONE upload of dualRM; outs[k] is the round of step k, and
 * the host takes the first step with a column that is not a duplicate.  alphasOut (nSteps, may be
 * NULL) receives the alpha_k.  nSteps <= 64.
 */
int dipgpu_stab_set_center(dipgpu_ctx *ctx, const double *dual, int32_t uLen);
int dipgpu_stab_accept(dipgpu_ctx *ctx, int32_t step);
int dipgpu_price_round_stab(dipgpu_ctx *ctx, const double *uRM, int32_t uLen, double alpha,
                            int32_t phase, double redCostEps, dipgpu_cols *out, double *dualSTOpt);
int dipgpu_price_round_stab_ladder(dipgpu_ctx *ctx, const double *uRM, int32_t uLen, double alpha,
                                   double shrink, int32_t nSteps, int32_t phase, double redCostEps,
                                   dipgpu_cols *outs, double *alphasOut);
int dipgpu_get_smoothed_duals(dipgpu_ctx *ctx, int32_t step, double *dualST, int32_t uLen);

/* ------------------------------------------------------- waiting-column pool
 *
 * DecompVarPool (Dip/src/DecompVarPool.h) as far as the dual vector touches it: the master
 * columns of the waiting variables live in device memory, and once per price call
 * DecompVarPool::setReducedCosts (Dip/src/DecompVarPool.cpp:185-203 -> :22-40; call site
 * Dip/src/DecompAlgo.cpp:1890-1908) re-prices all of them:
 *     feasible != 0:  redCost_s = origCost_s - col_s . u        (:29)
 *     feasible == 0:  redCost_s =            - col_s . u        (:36, u = a dual ray)
 * *foundNegRC = 1 iff some redCost <= -1e-10 (:30, :37, :199).  Columns are in MASTER row space
 * (row indices into u, convexity entry included), each added last entry first (the order of
 * CoinPackedVectorBase::dotProduct).  u == NULL uses the resident dual vector (see above).
 *
 * Columns enter from host arrays (dipgpu_pool_append: colStart has nCols+1 entries starting at 0)
 * or device-to-device from the last dipgpu_expand_columns (dipgpu_pool_append_expanded: `which`
 * indexes that expansion's columns, e.g. those that survived the duplicate test), and leave when
 * DecompAlgo::addVarsFromPool moves them into the master (dipgpu_pool_keep: ascending indices of
 * the columns that stay).
 */
int dipgpu_pool_clear(dipgpu_ctx *ctx);
int dipgpu_pool_size(const dipgpu_ctx *ctx, int64_t *nCols, int64_t *nNnz);
int dipgpu_pool_append(dipgpu_ctx *ctx, int32_t nCols, const int64_t *colStart, const int32_t *rowInd,
                       const double *val, const double *origCost);
int dipgpu_pool_append_expanded(dipgpu_ctx *ctx, int32_t nWhich, const int32_t *which);
int dipgpu_pool_keep(dipgpu_ctx *ctx, int32_t nKeep, const int32_t *keepIdx);
int dipgpu_pool_set_reduced_costs(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t feasible,
                                  double *redCost, int32_t *foundNegRC);

/* ------------------------------------------------- after the round: master columns
 *
 * Master-LP columns of the columns kept by the LAST host-buffer round of this context
 * (dipgpu_price_round / dipgpu_solve_blocks): what DecompAlgo::addVarsToPool computes per new
 * variable (Dip/src/DecompAlgo.cpp:5452-5558) -- A''s through the row-ordered matrix, the cut rows
 * moved behind the convexity rows, 1.0 in the block's convexity row, entries with |v| > tolZero
 * (DecompParam::TolZero, 1e-6) in ascending master-row order -- plus a 64-bit fingerprint of
 * (block, x-space indices) per column, to be compared where DecompVarPool::isDuplicate compares
 * the string hash of Dip/src/UtilHash.cpp:52-68 (Dip/src/DecompVarPool.cpp:150-181).  Column k of
 * the output is kept column k of the round.  Needs dipgpu_set_core_csr.
 */
typedef struct dipgpu_mcols {
    /* in: capacities */
    int32_t capCols; /* colStart has capCols+1, hash capCols entries */
    int64_t capNnz;  /* rowInd / val */
    /* out */
    int32_t nCols;
    int64_t nNnz; /* required capNnz on DIPGPU_ERR_OVERFLOW */
    int64_t *colStart;
    int32_t *rowInd; /* master row indices, ascending per column */
    double *val;
    uint64_t *hash;
} dipgpu_mcols;

int dipgpu_expand_columns(dipgpu_ctx *ctx, double tolZero, dipgpu_mcols *out);

/*
 * dipgpu_price_round followed by dipgpu_expand_columns in ONE submission: what DecompAlgo::processNode does
 * back to back in a price call -- generateVars (Dip/src/DecompAlgo.cpp:1914 -> :4622-5260), then
 * addVarsToPool on its result (:1936 -> :5408-5675).  The expansion kernel is enqueued right behind the
 * round, the packed master columns behind it, and the host waits once; `out` and `mout` are filled exactly
 * as the two separate calls fill them.  u == NULL: the resident dual vector (see dipgpu_price_round).
 */
int dipgpu_price_round_expand(dipgpu_ctx *ctx, const double *u, int32_t uLen, int32_t phase, double redCostEps,
                              double tolZero, dipgpu_cols *out, dipgpu_mcols *mout);

/* ---------------------------------------- device-resident variants (multi-GPU)
 *
 * For the block-sharded multi-GPU driver: the dual vector is already in device
 * memory (NCCL broadcast) and the packed result stays in device memory for the
 * gather.  `stream` is a cudaStream_t (NULL = the context's own stream); the
 * call only enqueues work.
 */

/* Restrict pricing to blocks [first, first+count) (a rank's shard). */
int dipgpu_set_block_range(dipgpu_ctx *ctx, int32_t first, int32_t count);

int dipgpu_price_round_dev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase,
                           double redCostEps, void *stream);
int dipgpu_solve_blocks_dev(dipgpu_ctx *ctx, const double *redCostXDev, const double *alphaDev,
                            double redCostEps, void *stream);

/* Stage (a) alone on a device-resident dual vector; the result stays in the
 * library's reduced-cost buffer (dipgpu_redcost_dev). */
int dipgpu_calc_redcost_dev(dipgpu_ctx *ctx, const double *uDev, int32_t uLen, int32_t phase,
                            void *stream);

/* Device pointer + byte size of the packed result of the last *_dev call. */
int dipgpu_result_dev(dipgpu_ctx *ctx, void **devPtr, size_t *bytes);

/*
 * Multi-device contexts, device-resident: dipgpu_price_round_dev(ctx, uDev, ..., stream = NULL) takes the
 * master dual vector from the PRIMARY device's memory; every device pulls what it needs out of it over NVLink
 * (peer loads of u[support]; a peer copy of the whole vector without a dual support) and its fused DP kernel
 * repeats every result store into its slice of a gather buffer in the primary's memory (peer stores) -- the
 * dual broadcast and the column gather of SURVEY 8e without a collective.  The call only enqueues, on every
 * device's own stream; dipgpu_sync waits for all of them.  dipgpu_result_dev then returns the gather buffer
 * (primary device), dipgpu_result_shards where each shard's packed result sits in it, and
 * dipgpu_unpack_results decodes a host copy of it into ONE round (block order, as dipgpu_price_round).
 * With option "stage_timing" every device brackets its part with CUDA events on its own stream;
 * dipgpu_last_device_ms reports them (and their maximum) after dipgpu_sync / after a host-buffer call.
 */
int dipgpu_sync(dipgpu_ctx *ctx);
int dipgpu_result_shards(const dipgpu_ctx *ctx, int32_t cap, int32_t *nShards, size_t *offsets, size_t *sizes);
int dipgpu_unpack_results(const void *packedHost, int32_t nShards, const size_t *offsets, const size_t *sizes,
                          dipgpu_cols *out);
int dipgpu_last_device_ms(const dipgpu_ctx *ctx, int32_t cap, float *ms, float *maxMs);
/* Fixed upper bound of the packed result size for this context's shard. */
int dipgpu_result_max_bytes(const dipgpu_ctx *ctx, size_t *bytes);
/* Decode a packed result (host copy) into `out`; block ids are global. */
int dipgpu_unpack_result(const void *packedHost, size_t bytes, dipgpu_cols *out);

/* Device pointer of the last reduced-cost vector (length N). */
int dipgpu_redcost_dev(dipgpu_ctx *ctx, double **devPtr);

/* ------------------------------------------------------------ introspection */

/* Counters since creation: kernel launches issued by this library, H2D / D2H
 * bytes moved by it. */
int dipgpu_get_counters(const dipgpu_ctx *ctx, int64_t *kernelLaunches, int64_t *h2dBytes,
                        int64_t *d2hBytes);

/* Device time (ms, CUDA events on the context's stream) of the stages of the
 * last host-buffer call: [0] H2D, [1] reduced costs, [2] knapsack DP,
 * [3] backtrack+emit, [4] filter, [5] D2H.  Valid only when timing was switched
 * on with dipgpu_set_option("stage_timing", 1). */
int dipgpu_last_stage_ms(const dipgpu_ctx *ctx, float *ms6);

/* Roofline denominator for the shared-memory-bound knapsack DP: runs a
 * shared-memory copy microbenchmark (every SM, LDS.64 + STS.64 streams) on the
 * context's device and reports bytes moved / CUDA-event time in GB/s. */
int dipgpu_measure_smem_gbs(dipgpu_ctx *ctx, double *gbs);

/* Host-only planning (no device needed): the DP kernel shape the library would
 * pick for a shard of nBlocks knapsacks on a GPU with smCount SMs.
 * out = {threads per knapsack, cells per thread, items per barrier, guard cells,
 * TMA chunk columns, shared-memory bytes, fused tail (1 = backtrack + filter run
 * inside the DP kernel)}.  dipgpu_get_dp_geometry reports the
 * shape in use by a context. */
int dipgpu_plan_dp_geometry(int32_t smCount, int32_t nBlocks, int32_t maxCapacity, int32_t maxBlockCols,
                            int32_t maxUsableWeight, int32_t out[7]);
int dipgpu_get_dp_geometry(const dipgpu_ctx *ctx, int32_t out[7]);

/* Tuning / diagnostics knobs: "stage_timing", "dp_threads" (multiple of 32),
 * "dp_cells_per_thread" (odd, 1..25), "dp_items_per_step" (1, 2, or 3 with guard cells and <= 3 cells per thread), "dp_no_guard", "dp_no_prune" (1 = the fused kernel runs the DP over every active item), "fuse_redcost" (default 1: small shards form the reduced costs of their blocks' columns inside the fused DP kernel -- the whole round is ONE kernel; 0 = the stream SpMV kernel in front of it), "dp_pdl" (1 = programmatic dependent launch of the fused kernel behind the SpMV; measured slower, off by default), "shard_rounds" (multi-device contexts: see dipgpu_create_multi), "batch_pipeline" (1 = batches of >= 8 page-locked dual vectors are priced in chunks, the gathers of u[support] on a second stream beside the previous chunk's DP launch; measured: +5 % at best, off by default),
 * "batch_pull" / "batch_pull_min" (default 1 / 24: batched launches of at least that many dual vectors fetch u[support] from host
 * memory themselves, CTA (vector, block) its slice, instead of a gather kernel in front), "flag_wait" (default 1: a mirrored single
 * round announces its result through a word in mapped host memory, the host entry point spins on it instead of synchronising the
 * stream), "append_delta_min_nnz" (default 2^20: dipgpu_append_core_rows on models with at least that many stored entries keeps the
 * new rows as a delta -- O(new entries) -- instead of rebuilding; "merge_appended_rows" folds the delta in), "pool_compact" (default
 * 1: with a dual support the pool re-pricing pass reads a compacted copy of the pool without the rows that cannot carry a dual),
 * "pool_kernel" (1 persistent, default; 2 persistent with u[support] and a rank table in shared memory; 0 one CTA per 16 tiles),
 * "dp_no_cluster" (1 = DP rows beyond one CTA in global memory even when a thread-block cluster would hold them),
 * "dp_cluster_threads" (1024 or 512), "dp_cluster_max_cells", diagnostics "host_timing", "flag_stats", "dp_cluster_debug",
 * "spmv_lanes" (0 = TMA stream kernel), "spmv_chunk" / "spmv_max_cols" (entry / column budget of
 * a warp tile; re-tiles the matrix), "spmv_stages", "spmv_warps", "spmv_u_smem", "fuse_filter" (-1 auto, 0 = three kernels,
 * 1 = filter in the backtrack tail, 2 = backtrack + filter in the DP tail); 0 restores the automatic
 * choice of the dp_* knobs.  Transfers of the synchronous host entry points, both on by default:
 * "result_zero_copy" (the fused kernel stores the packed result into mapped pinned host memory itself instead
 * of a device-to-host copy after it; 2 = also from the *_dev entry points, diagnostics), "gather_pinned_duals"
 * (with a dual support and `u` in page-locked or managed memory -- e.g. from dipgpu_host_alloc -- the device
 * picks u[support] out of the caller's buffer; a pageable `u` is gathered on the host).  Unknown names return
 * DIPGPU_ERR_INVALID. */
int dipgpu_set_option(dipgpu_ctx *ctx, const char *name, int64_t value);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* DIPGPU_H */

// dipgpu_decomp.hpp -- header-only C++17 adapter between DIP's pricing plugin surface and the
// C ABI of libdipgpu (include/dipgpu.h).
//
// It mirrors, with the reference's names and semantics, the two surfaces a DIP application or
// algorithm subclass can put the GPU pricing round behind:
//
//   * Pricer::generateVars(newVars, mostNegReducedCost)
//       <-> virtual int DecompAlgo::generateVars(DecompVarList&, double&)
//           (Dip/src/DecompAlgo.h:430, Dip/src/DecompAlgo.cpp:4622-5260)
//   * Pricer::solveRelaxed(whichBlock, redCostX, target, varList)
//       <-> virtual DecompSolverStatus DecompApp::solveRelaxed(const int, const double*,
//           const double, DecompVarList&)  (Dip/src/DecompApp.h:344-349;
//           GAP: Dip/examples/GAP/GAP_DecompApp.cpp:235-285)
//
// The variable type is a template parameter: inside DIP pass ::DecompVar / ::DecompVarList
// (anything constructible from (std::vector<int>, std::vector<double>, redCost, origCost) with
// setBlockId(int), Dip/src/DecompVar.h:217-239); without COIN-OR the stand-in dipgpu::DecompVar
// below is used (tests/cpp).  Columns are allocated with `new` in the caller's C++ runtime, because
// DIP takes ownership and deletes the ones it rejects (DecompAlgo.cpp:5225).
//
// Errors: like the reference (UtilException / CoinError) failures throw -- dipgpu::Error carries
// the library's status code and message.  There is no CPU fallback.
#ifndef DIPGPU_DECOMP_HPP
#define DIPGPU_DECOMP_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "dipgpu.h"

namespace dipgpu {

// Dip/src/Decomp.h:213-219
enum DecompSolverStatus {
    DecompSolStatError,
    DecompSolStatOptimal,
    DecompSolStatFeasible,
    DecompSolStatInfeasible,
    DecompSolStatNoSolution
};

// Dip/src/Decomp.h:170-176 (the two pricing phases)
enum DecompPhase { PHASE_PRICE1 = DIPGPU_PHASE_PRICE1, PHASE_PRICE2 = DIPGPU_PHASE_PRICE2 };

// Stand-in for ::DecompVar when COIN-OR is not available: same constructor shape and accessors
// (Dip/src/DecompVar.h:217-239, getOriginalCost/getReducedCost/getBlockId).
struct DecompVar {
    std::vector<int> ind;      // m_s indices, ascending (DecompVar.h:237 sortVar)
    std::vector<double> els;   // m_s elements (1.0 for knapsack blocks)
    double m_origCost;
    double m_redCost;
    int m_blockId;
    DecompVar(const std::vector<int> &i, const std::vector<double> &e, double redCost, double origCost)
        : ind(i), els(e), m_origCost(origCost), m_redCost(redCost), m_blockId(0) {}
    void setBlockId(int b) { m_blockId = b; }
    int getBlockId() const { return m_blockId; }
    double getOriginalCost() const { return m_origCost; }
    double getReducedCost() const { return m_redCost; }
};
typedef std::list<DecompVar *> DecompVarList;  // Dip/src/Decomp.h (std::list<DecompVar*>)

class Error : public std::runtime_error {
public:
    Error(int code, const std::string &what) : std::runtime_error(what), m_code(code) {}
    int code() const { return m_code; }

private:
    int m_code;
};

class Pricer {
public:
    explicit Pricer(int device = 0) : m_ctx(nullptr), m_m(0), m_N(0), m_nBlockCols(0), m_roundValid(false)
    {
        const int rc = dipgpu_create(&m_ctx, device);
        if (rc != DIPGPU_OK)
            throw Error(rc, "dipgpu_create failed: no usable sm_100 GPU (libdipgpu has no CPU fallback)");
    }
    // One pricer over several GPUs of the box (dipgpu_create_multi): the same calls, generateVars partitioned by
    // blocks (DecompAlgo.cpp:4815-4848: the blocks are independent given the duals), the ladder by dual vectors.
    explicit Pricer(const std::vector<int> &devices) : m_ctx(nullptr), m_m(0), m_N(0), m_nBlockCols(0), m_roundValid(false)
    {
        std::vector<int32_t> d(devices.begin(), devices.end());
        const int rc = dipgpu_create_multi(&m_ctx, d.data(), (int32_t)d.size());
        if (rc != DIPGPU_OK)
            throw Error(rc, "dipgpu_create_multi failed: no usable sm_100 GPUs (libdipgpu has no CPU fallback)");
    }
    ~Pricer() { dipgpu_destroy(m_ctx); }
    Pricer(const Pricer &) = delete;
    Pricer &operator=(const Pricer &) = delete;

    dipgpu_ctx *ctx() { return m_ctx; }

    // ---- model (what DecompAlgo holds when generateVars runs) -----------------------------
    // modelCore->M: row-ordered A'' incl. branch rows (DecompConstraintSet.h:32, .cpp:41-43;
    // DecompAlgo.cpp:1399-1502).  Start/index types are templates so CoinBigIndex/int arrays of a
    // CoinPackedMatrix (getVectorStarts/getIndices/getElements) can be passed directly.
    template <class StartT, class IndexT>
    void setModelCore(int R, int N, const StartT *rowStart, const IndexT *colInd, const double *val,
                      int nBaseRows, int numConvex)
    {
        std::vector<int64_t> rs(rowStart, rowStart + R + 1);
        std::vector<int32_t> ci(colInd, colInd + rs[R]);
        check(dipgpu_set_core_csr(m_ctx, R, N, rs.data(), ci.data(), val, nBaseRows, numConvex));
        m_N = N;
        m_roundValid = false;
    }
    template <class StartT, class IndexT>
    void appendCoreRows(int nNew, const StartT *rowStart, const IndexT *colInd, const double *val)
    {
        std::vector<int64_t> rs(rowStart, rowStart + nNew + 1);
        std::vector<int32_t> ci(colInd, colInd + rs[nNew]);
        check(dipgpu_append_core_rows(m_ctx, nNew, rs.data(), ci.data(), val));
        m_roundValid = false;
    }
    void setObjective(const double *c, int N)
    {
        check(dipgpu_set_objective(m_ctx, c, N));
        if (N > m_N) m_N = N;
    }
    // One knapsack per block (GAP_DecompApp.cpp:52-58).
    void setKnapsackBlocks(int m, const int *blockColStart, const int *weight, const int *capacity,
                           int profitMode = DIPGPU_PROFIT_FP64)
    {
        check(dipgpu_set_knapsack_blocks(m_ctx, m, blockColStart, weight, capacity, profitMode));
        m_m = m;
        m_nBlockCols = blockColStart[m];
        if (m_nBlockCols > m_N) m_N = m_nBlockCols;
        allocResult();
        setBlockSlices(m, [&](int b) { return blockColStart[b]; }, nullptr);
    }

    // One integer knapsack per block with an optional use column: the cutting-stock pricing problem (relaxed_solver /
    // kp of Dip/src/dippy/tests/test_cutting_stock.py:98-179); the columns come back with multiplicities as elements.
    void setIntKnapBlocks(int m, const int *blockColStart, const int *weight, const int *capacity, const int *useCol)
    {
        check(dipgpu_set_intknap_blocks(m_ctx, m, blockColStart, weight, capacity, useCol));
        m_m = m;
        m_nBlockCols = blockColStart[m];
        if (useCol)
            for (int b = 0; b < m; ++b) m_nBlockCols = std::max(m_nBlockCols, useCol[b] + 1);
        if (m_nBlockCols > m_N) m_N = m_nBlockCols;
        allocResult();
        setBlockSlices(m, [&](int b) { return blockColStart[b]; }, useCol);
    }

    // One multiple-choice knapsack per block (MMKP_DecompApp.cpp:52-90: one MMKP_MCKnap per block; group g owns
    // the contiguous columns [groupColStart[g], groupColStart[g+1]), block b the groups
    // [blockGroupStart[b], blockGroupStart[b+1])).
    void setMckpBlocks(int m, const int *blockGroupStart, const int *groupColStart, const int *weight, const int *capacity)
    {
        check(dipgpu_set_mckp_blocks(m_ctx, m, blockGroupStart, groupColStart, weight, capacity));
        m_m = m;
        m_nBlockCols = groupColStart[blockGroupStart[m]];
        if (m_nBlockCols > m_N) m_N = m_nBlockCols;
        allocResult();
        setBlockSlices(m, [&](int b) { return groupColStart[blockGroupStart[b]]; }, nullptr);
    }

    // Rows of the master dual vector that can carry a dual at this node (the free branch rows of
    // AlpsDecompTreeNode.cpp:215-238 cannot): only those cross PCIe per round.  Empty removes the support.
    void setDualSupport(const std::vector<int32_t> &rows) { check(dipgpu_set_dual_support(m_ctx, (int32_t)rows.size(), rows.data())); }

    // ---- surface 2: the whole round ------------------------------------------------------------
    // DecompAlgo::generateVars: u = master duals (length R + numConvex; NULL = the vector the last
    // setReducedCosts / calcRedCost call left on the device), returns newVars.size().
    // mostNegReducedCost = sum_b min(0, rc_b) in block order (DecompAlgo.cpp:5229-5232).
    template <class VarT = DecompVar, class ListT = std::list<VarT *>>
    int generateVars(const double *u, int uLen, int phase, double redCostEps, ListT &newVars,
                     double &mostNegReducedCost)
    {
        check(dipgpu_price_round(m_ctx, u, uLen, phase, redCostEps, &m_cols, nullptr));
        mostNegReducedCost = m_cols.mostNegReducedCost;
        for (int k = 0; k < m_cols.nCols; ++k) newVars.push_back(makeVar<VarT>(k, m_colRedCost[k]));
        return m_cols.nCols;
    }

    // generateVars' round-robin branch (DecompAlgo.cpp:4929-5149): blocksToSolve / userDualsByBlock are what
    // DecompApp::solveRelaxedWhich (DecompApp.h:339-342) filled in.  solvedAll: no listed block priced out and
    // all blocks were solved with u instead (:5076-5148) -- the caller resets m_rrIterSinceAll then.
    template <class VarT = DecompVar, class ListT = std::list<VarT *>>
    int generateVarsWhich(const double *u, int uLen, int phase, double redCostEps, const std::vector<int> &blocksToSolve,
                          const std::map<int, std::vector<double>> &userDualsByBlock, ListT &newVars,
                          double &mostNegReducedCost, bool *solvedAll = nullptr)
    {
        std::vector<int32_t> which(blocksToSolve.begin(), blocksToSolve.end());
        std::vector<const double *> ud(which.size(), nullptr);
        for (size_t i = 0; i < which.size(); ++i) {
            auto it = userDualsByBlock.find(which[i]);
            if (it == userDualsByBlock.end()) continue;
            if ((int)it->second.size() != uLen)
                throw Error(DIPGPU_ERR_INVALID, "The size of the user dual vector is not the same as the number of master rows");
            ud[i] = it->second.data();
        }
        int32_t all = 0;
        check(dipgpu_price_round_which(m_ctx, u, uLen, phase, redCostEps, (int32_t)which.size(), which.data(),
                                       userDualsByBlock.empty() ? nullptr : ud.data(), &m_cols, &all));
        if (solvedAll) *solvedAll = all != 0;
        mostNegReducedCost = m_cols.mostNegReducedCost;
        for (int k = 0; k < m_cols.nCols; ++k) newVars.push_back(makeVar<VarT>(k, m_colRedCost[k]));
        m_roundValid = false;
        return m_cols.nCols;
    }

    // ---- surface 1: one block at a time -------------------------------------------------------
    // DecompApp::solveRelaxed.  Exactly one (possibly empty) column is pushed per call, its reduced
    // cost EXCLUDING the convexity dual `target` -- DecompAlgo::solveRelaxed subtracts it
    // (DecompAlgo.cpp:6350-6356) -- and DecompSolStatOptimal is returned
    // (GAP_DecompApp.cpp:278-284).  All blocks of a round are solved in ONE GPU submission the
    // first time a block's reduced costs are seen; the other calls of the round are served from
    // that result.  Block b's column depends on redCostX only through b's own entries, so a call is
    // served from the stored round exactly when those entries are bit-for-bit the ones the round was
    // solved with (one memcmp of the block's slice against the round's copy of redCostX) and b was not
    // served yet -- nothing is sampled or guessed: DIP allocates redCostX afresh every round
    // (DecompAlgo.cpp:4700) and allocators reuse addresses, so the pointer says nothing.  Any other
    // call starts a new round with the vector it brings.  beginRound() forces one.  Thread-safe
    // across different blocks only after the round's first call has returned (serialise that one, or
    // use surface 2).
    template <class VarT = DecompVar, class ListT = std::list<VarT *>>
    DecompSolverStatus solveRelaxed(int whichBlock, const double *redCostX, double /*target*/,
                                    ListT &varList)
    {
        if (whichBlock < 0 || whichBlock >= m_m) throw Error(DIPGPU_ERR_INVALID, "solveRelaxed: bad block");
        if (!m_roundValid || m_served[(size_t)whichBlock] || !sliceUnchanged(whichBlock, redCostX)) {
            std::vector<double> zero((size_t)m_m, 0.0);
            check(dipgpu_solve_blocks(m_ctx, redCostX, zero.data(),
                                      -std::numeric_limits<double>::infinity(), &m_cols));
            m_roundValid = true;
            m_roundX.assign(redCostX, redCostX + m_nBlockCols);
            m_served.assign((size_t)m_m, 0);
            ++m_roundsSolved;
        }
        m_served[(size_t)whichBlock] = 1;
        // eps = -inf keeps every block, in block order: kept column k == block k
        varList.push_back(makeVar<VarT>(whichBlock, m_blockRedCost[(size_t)whichBlock]));
        return DecompSolStatOptimal;
    }
    void beginRound() { m_roundValid = false; }
    // GPU submissions surface 1 has made so far (one per recognised round)
    long roundsSolved() const { return m_roundsSolved; }

    // ---- after the round: DecompAlgo::addVarsToPool's per-variable work (DecompAlgo.cpp:5452-5558)
    // Master columns [A''s ; e_block] of the columns the last generateVars / solveRelaxed round
    // kept, in that order: column k spans [colStart[k], colStart[k+1]) of (rowInd ascending, val);
    // hash[k] fingerprints (block, x-space indices) -- compare it where DecompVarPool::isDuplicate
    // compares DecompVar::getStrHash() (DecompVarPool.cpp:150-181).
    void expandColumns(double tolZero, std::vector<int64_t> &colStart, std::vector<int32_t> &rowInd,
                       std::vector<double> &val, std::vector<uint64_t> &hash)
    {
        dipgpu_mcols mc;
        std::memset(&mc, 0, sizeof(mc));
        colStart.assign((size_t)m_m + 1, 0);
        hash.assign((size_t)m_m, 0);
        if (rowInd.size() < 1024) rowInd.resize(1024);
        if (val.size() < rowInd.size()) val.resize(rowInd.size());
        for (int attempt = 0; attempt < 2; ++attempt) {
            mc.capCols = m_m;
            mc.capNnz = (int64_t)rowInd.size();
            mc.colStart = colStart.data();
            mc.rowInd = rowInd.data();
            mc.val = val.data();
            mc.hash = hash.data();
            const int rc = dipgpu_expand_columns(m_ctx, tolZero, &mc);
            if (rc == DIPGPU_ERR_OVERFLOW && attempt == 0 && mc.nNnz > (int64_t)rowInd.size()) {
                rowInd.resize((size_t)mc.nNnz);
                val.resize((size_t)mc.nNnz);
                continue;
            }
            check(rc);
            break;
        }
        colStart.resize((size_t)mc.nCols + 1);
        hash.resize((size_t)mc.nCols);
        rowInd.resize((size_t)mc.nNnz);
        val.resize((size_t)mc.nNnz);
    }

    // generateVars and the expansion of its columns in one GPU submission (dipgpu_price_round_expand): what a
    // price call of DecompAlgo::processNode does back to back (DecompAlgo.cpp:1914, :1936).
    template <class VarT = DecompVar, class ListT = std::list<VarT *>>
    int generateVarsExpanded(const double *u, int uLen, int phase, double redCostEps, double tolZero, ListT &newVars,
                             double &mostNegReducedCost, std::vector<int64_t> &colStart, std::vector<int32_t> &rowInd,
                             std::vector<double> &val, std::vector<uint64_t> &hash)
    {
        dipgpu_mcols mc;
        std::memset(&mc, 0, sizeof(mc));
        colStart.assign((size_t)m_m + 1, 0);
        hash.assign((size_t)m_m, 0);
        if (rowInd.size() < 1024) rowInd.resize(1024);
        if (val.size() < rowInd.size()) val.resize(rowInd.size());
        mc.capCols = m_m;
        mc.capNnz = (int64_t)rowInd.size();
        mc.colStart = colStart.data();
        mc.rowInd = rowInd.data();
        mc.val = val.data();
        mc.hash = hash.data();
        const int rc = dipgpu_price_round_expand(m_ctx, u, uLen, phase, redCostEps, tolZero, &m_cols, &mc);
        if (rc == DIPGPU_ERR_OVERFLOW && mc.nNnz > (int64_t)rowInd.size()) {
            // the round is done, only the caller's arrays were short: fetch the columns again
            rowInd.resize((size_t)mc.nNnz);
            val.resize((size_t)mc.nNnz);
            expandColumns(tolZero, colStart, rowInd, val, hash);
        } else {
            check(rc);
            colStart.resize((size_t)mc.nCols + 1);
            hash.resize((size_t)mc.nCols);
            rowInd.resize((size_t)mc.nNnz);
            val.resize((size_t)mc.nNnz);
        }
        mostNegReducedCost = m_cols.mostNegReducedCost;
        for (int k = 0; k < m_cols.nCols; ++k) newVars.push_back(makeVar<VarT>(k, m_colRedCost[k]));
        m_roundValid = false;
        return m_cols.nCols;
    }

    // ---- dual stabilisation: DecompAlgoPC::adjustMasterDualSolution + generateVars -------------
    // (Dip/src/DecompAlgoPC.cpp:126-212; DecompAlgoPC.h:99-133).  The stability centre m_dual lives on
    // the device: setDualCentre <-> "Init dual to dualRM" / DecompApp::initDualVector (:163-177),
    // acceptSmoothedDuals <-> the copy of m_dualST into m_dual in DecompAlgoPC::setObjBound.
    void setDualCentre(const double *dual, int uLen) { check(dipgpu_stab_set_center(m_ctx, dual, uLen)); }
    void acceptSmoothedDuals(int step = 0) { check(dipgpu_stab_accept(m_ctx, step)); }
    // generateVars with dualST = alpha*dual + (1-alpha)*dualRM; dualST (length uLen, may be NULL)
    // receives the vector the round was priced with (what getMasterDualSolution() returns under
    // DualStab, DecompAlgoPC.h:99-108).
    template <class VarT = DecompVar, class ListT = std::list<VarT *>>
    int generateVarsStabilised(const double *dualRM, int uLen, double alpha, int phase, double redCostEps,
                               ListT &newVars, double &mostNegReducedCost, double *dualST)
    {
        check(dipgpu_price_round_stab(m_ctx, dualRM, uLen, alpha, phase, redCostEps, &m_cols, dualST));
        mostNegReducedCost = m_cols.mostNegReducedCost;
        for (int k = 0; k < m_cols.nCols; ++k) newVars.push_back(makeVar<VarT>(k, m_colRedCost[k]));
        return m_cols.nCols;
    }
    // One round per smoothing parameter of the ladder alpha, shrink*alpha, ... that addVarsToPool
    // walks when every new column is a duplicate (DecompAlgo.cpp:5642-5655), priced speculatively in
    // one GPU submission.  Round k's columns are materialised on demand with takeLadderVars(k, ...).
    struct Round {
        std::vector<double> blockRedCost, blockMostNeg, blockOrigCost, blockObj, colRedCost, colOrigCost;
        std::vector<int32_t> blockNnz, colBlock, colInd;
        std::vector<int64_t> colStart;
        dipgpu_cols cols;
    };
    void generateVarsLadder(const double *dualRM, int uLen, double alpha, double shrink, int nSteps, int phase,
                            double redCostEps, std::vector<double> &alphas)
    {
        m_ladder.resize((size_t)nSteps);
        std::vector<dipgpu_cols> arr((size_t)nSteps);
        for (int k = 0; k < nSteps; ++k) {
            bindRound(m_ladder[(size_t)k]);
            arr[(size_t)k] = m_ladder[(size_t)k].cols;
        }
        alphas.assign((size_t)nSteps, 0.0);
        check(dipgpu_price_round_stab_ladder(m_ctx, dualRM, uLen, alpha, shrink, nSteps, phase, redCostEps, arr.data(),
                                             alphas.data()));
        for (int k = 0; k < nSteps; ++k) m_ladder[(size_t)k].cols = arr[(size_t)k];
    }
    const Round &ladderRound(int step) const { return m_ladder.at((size_t)step); }
    template <class VarT = DecompVar, class ListT = std::list<VarT *>>
    int takeLadderVars(int step, ListT &newVars, double &mostNegReducedCost)
    {
        const Round &r = m_ladder.at((size_t)step);
        mostNegReducedCost = r.cols.mostNegReducedCost;
        for (int k = 0; k < r.cols.nCols; ++k) {
            const int64_t b = r.colStart[(size_t)k], e = r.colStart[(size_t)k + 1];
            std::vector<int> ind(r.colInd.begin() + b, r.colInd.begin() + e);
            std::vector<double> els(ind.size(), 1.0);
            VarT *v = new VarT(ind, els, r.colRedCost[(size_t)k], r.colOrigCost[(size_t)k]);
            v->setBlockId(r.colBlock[(size_t)k]);
            newVars.push_back(v);
        }
        check(dipgpu_select_batch_result(m_ctx, step));  // expandColumns() now refers to this step
        return r.cols.nCols;
    }

    // ---- the waiting-column pool: DecompVarPool::setReducedCosts (DecompVarPool.cpp:185-203) -----
    // Call site DecompAlgo.cpp:1890-1908, right before generateVars with the same duals: pass the
    // duals here and u == NULL to generateVars (they stay on the device).
    void poolClear() { check(dipgpu_pool_clear(m_ctx)); }
    void poolAppend(int nCols, const int64_t *colStart, const int32_t *rowInd, const double *val, const double *origCost)
    {
        check(dipgpu_pool_append(m_ctx, nCols, colStart, rowInd, val, origCost));
    }
    // columns `which` of the last expandColumns() (e.g. those that passed isDuplicate / isParallel)
    void poolAppendExpanded(const std::vector<int32_t> &which)
    {
        check(dipgpu_pool_append_expanded(m_ctx, (int32_t)which.size(), which.data()));
    }
    // addVarsFromPool moved the others into the master (DecompAlgo.cpp:5678-5900)
    void poolKeep(const std::vector<int32_t> &keepIdx) { check(dipgpu_pool_keep(m_ctx, (int32_t)keepIdx.size(), keepIdx.data())); }
    // returns found_negrc_var; redCost[k] = reduced cost of waiting column k
    bool setReducedCosts(const double *u, int uLen, bool feasible, std::vector<double> &redCost)
    {
        int64_t n = 0;
        check(dipgpu_pool_size(m_ctx, &n, nullptr));
        redCost.assign((size_t)n, 0.0);
        int32_t found = 0;
        check(dipgpu_pool_set_reduced_costs(m_ctx, u, uLen, feasible ? 1 : 0, redCost.data(), &found));
        return found != 0;
    }

    // ---- stage (a) alone: generateVarsCalcRedCost (DecompAlgo.cpp:4506-4547) -------------------
    void calcRedCost(const double *u, int uLen, int phase, double *redCostX)
    {
        check(dipgpu_calc_redcost(m_ctx, u, uLen, phase, redCostX));
    }

    // per-block results of the last round (length m): rc_b - alpha_b, min(0, .), c.s, knapsack optimum
    const std::vector<double> &blockRedCost() const { return m_blockRedCost; }
    const std::vector<double> &blockMostNegRC() const { return m_blockMostNeg; }
    const std::vector<double> &blockOrigCost() const { return m_blockOrigCost; }
    const std::vector<double> &blockObj() const { return m_blockObj; }
    int64_t cells() const { return m_cols.cells; }

private:
    void check(int rc)
    {
        if (rc == DIPGPU_OK) return;
        char buf[512];
        dipgpu_last_error(m_ctx, buf, sizeof(buf));
        throw Error(rc, std::string("libdipgpu: ") + buf);
    }
    void allocResult()
    {
        const size_t m = (size_t)m_m, n = (size_t)(m_nBlockCols > 0 ? m_nBlockCols : 1);
        m_blockRedCost.assign(m, 0.0);
        m_blockMostNeg.assign(m, 0.0);
        m_blockOrigCost.assign(m, 0.0);
        m_blockObj.assign(m, 0.0);
        m_blockNnz.assign(m, 0);
        m_colBlock.assign(m + 1, 0);
        m_colRedCost.assign(m + 1, 0.0);
        m_colOrigCost.assign(m + 1, 0.0);
        m_colStart.assign(m + 2, 0);
        m_colInd.assign(n + m, 0);
        m_colEl.assign(n + m, 1.0);
        m_served.assign(m, 0);
        std::memset(&m_cols, 0, sizeof(m_cols));
        m_cols.capBlocks = m_m;
        m_cols.capCols = m_m;
        m_cols.capInd = (int64_t)(n + m);
        m_cols.blockRedCost = m_blockRedCost.data();
        m_cols.blockMostNegRC = m_blockMostNeg.data();
        m_cols.blockOrigCost = m_blockOrigCost.data();
        m_cols.blockObj = m_blockObj.data();
        m_cols.blockNnz = m_blockNnz.data();
        m_cols.colBlock = m_colBlock.data();
        m_cols.colRedCost = m_colRedCost.data();
        m_cols.colOrigCost = m_colOrigCost.data();
        m_cols.colStart = m_colStart.data();
        m_cols.colInd = m_colInd.data();
        m_cols.colEl = m_colEl.data();
    }
    void bindRound(Round &r) const
    {
        const size_t m = (size_t)m_m, n = (size_t)(m_nBlockCols > 0 ? m_nBlockCols : 1);
        r.blockRedCost.assign(m, 0.0); r.blockMostNeg.assign(m, 0.0); r.blockOrigCost.assign(m, 0.0);
        r.blockObj.assign(m, 0.0); r.blockNnz.assign(m, 0); r.colBlock.assign(m + 1, 0);
        r.colRedCost.assign(m + 1, 0.0); r.colOrigCost.assign(m + 1, 0.0); r.colStart.assign(m + 2, 0);
        r.colInd.assign(n, 0);
        std::memset(&r.cols, 0, sizeof(r.cols));
        r.cols.capBlocks = m_m; r.cols.capCols = m_m; r.cols.capInd = (int64_t)n;
        r.cols.blockRedCost = r.blockRedCost.data(); r.cols.blockMostNegRC = r.blockMostNeg.data();
        r.cols.blockOrigCost = r.blockOrigCost.data(); r.cols.blockObj = r.blockObj.data();
        r.cols.blockNnz = r.blockNnz.data(); r.cols.colBlock = r.colBlock.data();
        r.cols.colRedCost = r.colRedCost.data(); r.cols.colOrigCost = r.colOrigCost.data();
        r.cols.colStart = r.colStart.data(); r.cols.colInd = r.colInd.data();
    }
    template <class VarT>
    VarT *makeVar(int k, double redCost) const
    {
        const int64_t b = m_colStart[(size_t)k], e = m_colStart[(size_t)k + 1];
        std::vector<int> ind(m_colInd.begin() + b, m_colInd.begin() + e);
        // 1.0 for 0-1 blocks (GAP_KnapPisinger.h:265-270), the multiplicities for integer knapsack blocks
        std::vector<double> els(m_colEl.begin() + b, m_colEl.begin() + e);
        VarT *v = new VarT(ind, els, redCost, m_colOrigCost[(size_t)k]);
        v->setBlockId(m_colBlock[(size_t)k]);
        return v;
    }
    // the entries of redCostX block b reads: its columns [lo, hi) and, for an integer knapsack block, its use column
    template <class F>
    void setBlockSlices(int m, F firstCol, const int *useCol)
    {
        m_sliceLo.resize((size_t)m + 1);
        for (int b = 0; b <= m; ++b) m_sliceLo[(size_t)b] = firstCol(b);
        m_sliceExtra.assign((size_t)m, -1);
        if (useCol)
            for (int b = 0; b < m; ++b) m_sliceExtra[(size_t)b] = useCol[b];
        m_roundValid = false;
    }
    bool sliceUnchanged(int b, const double *x) const
    {
        const int lo = m_sliceLo[(size_t)b], hi = m_sliceLo[(size_t)b + 1];
        if (hi > lo && std::memcmp(x + lo, m_roundX.data() + lo, sizeof(double) * (size_t)(hi - lo)) != 0) return false;
        const int e = m_sliceExtra[(size_t)b];
        return e < 0 || std::memcmp(x + e, m_roundX.data() + e, sizeof(double)) == 0;
    }

    dipgpu_ctx *m_ctx;
    int m_m, m_N, m_nBlockCols;
    dipgpu_cols m_cols;
    std::vector<double> m_blockRedCost, m_blockMostNeg, m_blockOrigCost, m_blockObj, m_colRedCost, m_colOrigCost;
    std::vector<int32_t> m_blockNnz, m_colBlock, m_colInd;
    std::vector<double> m_colEl;
    std::vector<int64_t> m_colStart;
    bool m_roundValid;
    std::vector<double> m_roundX;  // the reduced costs the stored round was solved with (block columns)
    std::vector<int> m_sliceLo, m_sliceExtra;
    long m_roundsSolved = 0;
    std::vector<char> m_served;
    std::vector<Round> m_ladder;
};

}  // namespace dipgpu

#endif  // DIPGPU_DECOMP_HPP

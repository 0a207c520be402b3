// test_adapter.cpp -- the C++ host mirror (include/dipgpu_decomp.hpp) exercised the way a DIP
// application would: build a GAP model (Dip/examples/GAP/GAP_DecompApp.cpp:138-232 layout), then
// call generateVars / solveRelaxed and compare with the CPU oracle (TEST INFRASTRUCTURE:
// oracle/dip_oracle.h).  Without a GPU the Pricer constructor must throw (no CPU fallback):
// the program then prints NO_GPU and exits 0.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "../../include/dipgpu_decomp.hpp"
#include "../../oracle/dip_oracle.h"

static uint64_t lcg(uint64_t &s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return s >> 33; }
static double unif(uint64_t &s) { return (double)(lcg(s) % 1000000) / 1000000.0; }

#define CHECK(c)                                                                     \
    do {                                                                             \
        if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

int main()
{
    if (dipgpu_abi_version() != DIPGPU_ABI_VERSION) { std::printf("ABI mismatch\n"); return 1; }
    const int m = 12, n = 90, N = m * n;
    uint64_t seed = 20200;
    std::vector<int> weight(N), capacity(m), blockColStart(m + 1);
    std::vector<double> cost(N);
    for (int i = 0; i < m; ++i) {
        long sum = 0;
        for (int j = 0; j < n; ++j) {
            weight[i * n + j] = 1 + (int)(lcg(seed) % 100);
            cost[i * n + j] = 111 - weight[i * n + j] + (int)(lcg(seed) % 21) - 10;
            sum += weight[i * n + j];
        }
        capacity[i] = (int)(0.8 * sum / m);
        blockColStart[i] = i * n;
    }
    blockColStart[m] = N;
    // A'': n assignment rows, then 2N identity branch rows (DecompAlgo.cpp:1431-1455)
    const int R = n + 2 * N;
    std::vector<int64_t> rowStart(R + 1, 0);
    std::vector<int32_t> colInd;
    std::vector<double> val;
    for (int j = 0; j < n; ++j) {
        for (int i = 0; i < m; ++i) { colInd.push_back(i * n + j); val.push_back(1.0); }
        rowStart[j + 1] = (int64_t)colInd.size();
    }
    for (int r = 0; r < 2 * N; ++r) {
        colInd.push_back(r % N);
        val.push_back(1.0);
        rowStart[n + r + 1] = (int64_t)colInd.size();
    }
    // master duals [core rows | m convexity]
    std::vector<double> u(R + m, 0.0);
    for (int j = 0; j < n; ++j) u[j] = unif(seed) * 120.0;
    for (int r = 0; r < 2 * N; ++r)
        if (lcg(seed) % 50 == 0) u[n + r] = (r < N ? -1.0 : 1.0) * unif(seed) * 3.0;
    for (int b = 0; b < m; ++b) u[R + b] = -10.0 * unif(seed);

    dipgpu::Pricer *pricer = nullptr;
    try {
        pricer = new dipgpu::Pricer(0);
    } catch (const dipgpu::Error &e) {
        if (e.code() == DIPGPU_ERR_CUDA) { std::printf("NO_GPU\n"); return 0; }
        std::printf("unexpected error: %s\n", e.what());
        return 1;
    }
    pricer->setModelCore(R, N, rowStart.data(), colInd.data(), val.data(), R, m);
    pricer->setObjective(cost.data(), N);
    pricer->setKnapsackBlocks(m, blockColStart.data(), weight.data(), capacity.data());

    // ---- oracle
    std::vector<double> redCostX(N), colRedCost(m), colOrigCost(m), mostNeg(m), z(m);
    std::vector<int32_t> colNnz(m), colKeep(m), ind((size_t)m * n);
    double mostNegRef = 0;
    int64_t cells = 0;
    const int kept = dip_oracle_price_round(R, N, rowStart.data(), colInd.data(), val.data(), cost.data(), u.data(),
                                            R, 0, m, blockColStart.data(), weight.data(), capacity.data(), 0, 1e-4,
                                            1, n, redCostX.data(), colNnz.data(), colRedCost.data(),
                                            colOrigCost.data(), colKeep.data(), mostNeg.data(), z.data(), ind.data(),
                                            &mostNegRef, &cells);

    // ---- surface 2: DecompAlgo::generateVars
    dipgpu::DecompVarList newVars;
    double mostNegRC = 1.0;
    const int nNew = pricer->generateVars(u.data(), (int)u.size(), dipgpu::PHASE_PRICE2, 1e-4, newVars, mostNegRC);
    CHECK(nNew == kept && (int)newVars.size() == kept);
    CHECK(mostNegRC == mostNegRef);
    CHECK(pricer->cells() == cells);
    for (int b = 0; b < m; ++b) CHECK(pricer->blockObj()[b] == z[b]);  // bit-exact knapsack optima
    // ---- the step after the round: master columns of the kept variables (addVarsToPool)
    {
        std::vector<int64_t> cs;
        std::vector<int32_t> ri;
        std::vector<double> va;
        std::vector<uint64_t> hs;
        pricer->expandColumns(1e-6, cs, ri, va, hs);
        CHECK((int)hs.size() == kept && (int)cs.size() == kept + 1);
        std::vector<int32_t> oRow(R + m);
        std::vector<double> oVal(R + m);
        int k = 0;
        for (dipgpu::DecompVar *v : newVars) {
            const int nn = dip_oracle_expand_column(R, N, rowStart.data(), colInd.data(), val.data(), R, m,
                                                    (int)v->ind.size(), v->ind.data(), nullptr, v->getBlockId(), 1e-6,
                                                    oRow.data(), oVal.data());
            CHECK(cs[k + 1] - cs[k] == nn);
            for (int q = 0; q < nn; ++q) CHECK(ri[cs[k] + q] == oRow[q] && va[cs[k] + q] == oVal[q]);
            ++k;
        }
        // ... and both steps in one submission: the same variables, the same master columns
        std::vector<int64_t> cs1;
        std::vector<int32_t> ri1(8);     // too short on purpose: the adapter fetches the columns again
        std::vector<double> va1;
        std::vector<uint64_t> hs1;
        dipgpu::DecompVarList vars1;
        double mn1 = 1.0;
        const int n1 = pricer->generateVarsExpanded(u.data(), (int)u.size(), dipgpu::PHASE_PRICE2, 1e-4, 1e-6, vars1, mn1,
                                                    cs1, ri1, va1, hs1);
        CHECK(n1 == kept && mn1 == mostNegRef);
        CHECK(cs1 == cs && ri1 == ri && va1 == va && hs1 == hs);
        auto it = newVars.begin();
        for (dipgpu::DecompVar *v : vars1) {
            CHECK(v->getBlockId() == (*it)->getBlockId() && v->ind == (*it)->ind && v->getReducedCost() == (*it)->getReducedCost());
            ++it;
            delete v;
        }
    }
    for (dipgpu::DecompVar *v : newVars) {
        const int b = v->getBlockId();
        CHECK(colKeep[b] == 1 && v->getReducedCost() < -1e-4);
        CHECK((int)v->ind.size() == colNnz[b]);
        for (int k = 0; k < colNnz[b]; ++k) CHECK(v->ind[k] == ind[(size_t)b * n + k] && v->els[k] == 1.0);
        CHECK(v->getReducedCost() == colRedCost[b] && v->getOriginalCost() == colOrigCost[b]);
        delete v;  // the caller owns the columns, as DIP does (DecompAlgo.cpp:5225)
    }

    // ---- stage (a) alone, then surface 1: DecompApp::solveRelaxed per block on that vector
    std::vector<double> rcx(N);
    pricer->calcRedCost(u.data(), (int)u.size(), dipgpu::PHASE_PRICE2, rcx.data());
    for (int j = 0; j < N; ++j) CHECK(rcx[j] == redCostX[j]);
    for (int b = 0; b < m; ++b) {
        dipgpu::DecompVarList vars;
        const dipgpu::DecompSolverStatus st = pricer->solveRelaxed(b, rcx.data(), u[R + b], vars);
        CHECK(st == dipgpu::DecompSolStatOptimal && vars.size() == 1);
        dipgpu::DecompVar *v = vars.front();
        CHECK(v->getBlockId() == b && (int)v->ind.size() == colNnz[b]);
        // reduced cost WITHOUT alpha; the driver subtracts it (DecompAlgo.cpp:6350-6356)
        CHECK(v->getReducedCost() - u[R + b] == colRedCost[b]);
        delete v;
    }
    int64_t launches = 0;
    dipgpu_get_counters(pricer->ctx(), &launches, nullptr, nullptr);
    // 2 (round) + 1 (expand) + 2 + 1 (+1: the columns fetched again) (round + expansion in one submission) + 1 (SpMV)
    // + <=3 (ONE submission served all 12 solveRelaxed calls)
    CHECK(launches <= 11);
    // ---- surface 1 recognises rounds by content, not by address: DIP allocates redCostX afresh every round
    // (DecompAlgo.cpp:4700) and the allocator may hand back the same address.  ONE entry changes in place -- one the
    // former 64-sample fingerprint did not look at -- and the first call of the new round names a block that the
    // stored round has not served yet: the column must come from the new vector.
    {
        std::vector<double> x(redCostX);
        const long r0 = pricer->roundsSolved();
        dipgpu::DecompVarList v0;
        pricer->solveRelaxed(0, x.data(), u[R + 0], v0);          // every block was served above: a new round
        CHECK(pricer->roundsSolved() == r0 + 1);
        const int j = 3 * n + 1;                                  // block 3, not a multiple of N/64, not the last entry
        CHECK(weight[j] <= capacity[3]);
        x[j] = -1.0e6;                                            // the optimal column of block 3 now contains j
        dipgpu::DecompVarList v3, v4, v3f;
        pricer->solveRelaxed(3, x.data(), u[R + 3], v3);
        CHECK(pricer->roundsSolved() == r0 + 2);                  // block 3's entries changed: not served from the stored round
        CHECK(std::find(v3.front()->ind.begin(), v3.front()->ind.end(), j) != v3.front()->ind.end());
        CHECK(v3.front()->getReducedCost() <= -1.0e6);
        pricer->solveRelaxed(4, x.data(), u[R + 4], v4);          // block 4's entries are the stored round's: served from it
        CHECK(pricer->roundsSolved() == r0 + 2);
        CHECK((int)v4.front()->ind.size() == colNnz[4] && v4.front()->getReducedCost() - u[R + 4] == colRedCost[4]);
        pricer->beginRound();                                     // forced: the same answer
        pricer->solveRelaxed(3, x.data(), u[R + 3], v3f);
        CHECK(pricer->roundsSolved() == r0 + 3);
        CHECK(v3f.front()->ind == v3.front()->ind && v3f.front()->getReducedCost() == v3.front()->getReducedCost());
        for (auto *l : {&v0, &v3, &v4, &v3f})
            for (dipgpu::DecompVar *v : *l) delete v;
    }
    // phase 1 prices with c = 0
    pricer->calcRedCost(u.data(), (int)u.size(), dipgpu::PHASE_PRICE1, rcx.data());
    std::vector<double> uAdj(R);
    dip_oracle_adjust_duals(u.data(), R + m, R, m, uAdj.data());
    dip_oracle_calc_redcost(R, N, rowStart.data(), colInd.data(), val.data(), uAdj.data(), cost.data(), 1,
                            redCostX.data());
    for (int j = 0; j < N; ++j) CHECK(rcx[j] == redCostX[j]);
    // ---- the price call as DecompAlgo::processNode runs it (DecompAlgo.cpp:1890-1914): pool
    // re-pricing, then generateVars with the SAME duals (left on the device), under Wentges smoothing
    {
        dipgpu::DecompVarList vars0;
        double mn0 = 0;
        pricer->generateVars(u.data(), (int)u.size(), dipgpu::PHASE_PRICE2, 1e-4, vars0, mn0);
        std::vector<int64_t> cs;
        std::vector<int32_t> ri;
        std::vector<double> va;
        std::vector<uint64_t> hs;
        pricer->expandColumns(1e-6, cs, ri, va, hs);
        std::vector<int32_t> which;
        std::vector<double> oc;
        for (int k = 0; k < (int)hs.size(); k += 2) which.push_back(k);
        {
            int k = 0;
            for (dipgpu::DecompVar *v : vars0) {
                if (k % 2 == 0) oc.push_back(v->getOriginalCost());
                ++k;
                delete v;
            }
        }
        pricer->poolClear();
        pricer->poolAppendExpanded(which);
        // host copy of the same waiting columns for the oracle
        std::vector<int64_t> pcs(1, 0);
        std::vector<int32_t> pri;
        std::vector<double> pva;
        for (int k : which) {
            pri.insert(pri.end(), ri.begin() + cs[k], ri.begin() + cs[k + 1]);
            pva.insert(pva.end(), va.begin() + cs[k], va.begin() + cs[k + 1]);
            pcs.push_back((int64_t)pri.size());
        }
        std::vector<double> u2(u);
        for (int j = 0; j < n; ++j) u2[j] = unif(seed) * 120.0;
        // smoothed duals: centre = u, dualRM = u2
        std::vector<double> st(u.size()), stRef(u.size());
        pricer->setDualCentre(u.data(), (int)u.size());
        dipgpu::DecompVarList varsS;
        double mnS = 0;
        const int nS = pricer->generateVarsStabilised(u2.data(), (int)u2.size(), 0.1, dipgpu::PHASE_PRICE2, 1e-4, varsS, mnS,
                                                      st.data());
        dip_oracle_smooth_duals((int)u.size(), 0.1, u.data(), u2.data(), stRef.data());
        for (size_t r = 0; r < st.size(); ++r) CHECK(st[r] == stRef[r]);
        const int keptS = dip_oracle_price_round(R, N, rowStart.data(), colInd.data(), val.data(), cost.data(), stRef.data(),
                                                 R, 0, m, blockColStart.data(), weight.data(), capacity.data(), 0, 1e-4, 1, n,
                                                 redCostX.data(), colNnz.data(), colRedCost.data(), colOrigCost.data(),
                                                 colKeep.data(), mostNeg.data(), z.data(), ind.data(), &mostNegRef, &cells);
        CHECK(nS == keptS && mnS == mostNegRef);
        for (dipgpu::DecompVar *v : varsS) {
            CHECK(v->getReducedCost() == colRedCost[v->getBlockId()]);
            delete v;
        }
        // pool re-pricing with the smoothed duals, then generateVars(u == NULL) on the resident vector
        std::vector<double> prc, prcRef(which.size());
        const bool found = pricer->setReducedCosts(stRef.data(), (int)stRef.size(), true, prc);
        const int foundRef = dip_oracle_pool_reduced_costs((int)which.size(), pcs.data(), pri.data(), pva.data(), oc.data(),
                                                           stRef.data(), 1, prcRef.data());
        CHECK(prc.size() == which.size() && found == (foundRef != 0));
        for (size_t k = 0; k < prc.size(); ++k) CHECK(prc[k] == prcRef[k]);
        dipgpu::DecompVarList varsR;
        double mnR = 0;
        const int nR = pricer->generateVars(nullptr, (int)stRef.size(), dipgpu::PHASE_PRICE2, 1e-4, varsR, mnR);
        CHECK(nR == keptS && mnR == mostNegRef);
        for (dipgpu::DecompVar *v : varsR) delete v;
        // the alpha ladder in one submission: step k == the stabilised round with alpha * 0.9^k
        std::vector<double> alphas;
        pricer->generateVarsLadder(u2.data(), (int)u2.size(), 0.1, 0.90, 4, dipgpu::PHASE_PRICE2, 1e-4, alphas);
        double a = 0.1;
        for (int k = 0; k < 4; ++k) {
            CHECK(alphas[(size_t)k] == a);
            dip_oracle_smooth_duals((int)u.size(), a, u.data(), u2.data(), stRef.data());
            const int kk = dip_oracle_price_round(R, N, rowStart.data(), colInd.data(), val.data(), cost.data(), stRef.data(),
                                                  R, 0, m, blockColStart.data(), weight.data(), capacity.data(), 0, 1e-4, 1, n,
                                                  redCostX.data(), colNnz.data(), colRedCost.data(), colOrigCost.data(),
                                                  colKeep.data(), mostNeg.data(), z.data(), ind.data(), &mostNegRef, &cells);
            dipgpu::DecompVarList varsL;
            double mnL = 0;
            CHECK(pricer->takeLadderVars(k, varsL, mnL) == kk && mnL == mostNegRef);
            for (int b = 0; b < m; ++b) CHECK(pricer->ladderRound(k).blockObj[(size_t)b] == z[b]);
            for (dipgpu::DecompVar *v : varsL) delete v;
            a = a * 0.90;
        }
        // round-robin branch: two listed blocks, one with its own dual vector; against two oracle rounds
        {
            std::vector<double> uh(u2);
            for (int r = 0; r < R; ++r) uh[(size_t)r] *= 0.9;
            std::vector<int> colNnzU(m), colKeepU(m);
            std::vector<double> rcU(m), ocU(m), mnU(m), zU(m), rcxU(N);
            std::vector<int> indU((size_t)m * n);
            double tot = 0;
            int64_t cl = 0;
            dip_oracle_price_round(R, N, rowStart.data(), colInd.data(), val.data(), cost.data(), u2.data(), R, 0, m,
                                   blockColStart.data(), weight.data(), capacity.data(), 0, 1e-4, 1, n, redCostX.data(),
                                   colNnz.data(), colRedCost.data(), colOrigCost.data(), colKeep.data(), mostNeg.data(),
                                   z.data(), ind.data(), &tot, &cl);
            dip_oracle_price_round(R, N, rowStart.data(), colInd.data(), val.data(), cost.data(), uh.data(), R, 0, m,
                                   blockColStart.data(), weight.data(), capacity.data(), 0, 1e-4, 1, n, rcxU.data(),
                                   colNnzU.data(), rcU.data(), ocU.data(), colKeepU.data(), mnU.data(), zU.data(),
                                   indU.data(), &tot, &cl);
            const int bStd = 7, bUsr = 2;
            std::map<int, std::vector<double>> userDuals;
            userDuals[bUsr] = uh;
            dipgpu::DecompVarList varsW;
            double mnW = 0;
            bool all = false;
            const int nW = pricer->generateVarsWhich(u2.data(), (int)u2.size(), dipgpu::PHASE_PRICE2, 1e-4, {bStd, bUsr},
                                                     userDuals, varsW, mnW, &all);
            const bool anyNeg = colRedCost[bStd] < -1e-4 || rcU[bUsr] < -1e-4;
            CHECK(all == !anyNeg);
            if (anyNeg) {
                CHECK(nW == (colRedCost[bStd] < -1e-4 ? 1 : 0) + (rcU[bUsr] < -1e-4 ? 1 : 0));
                CHECK(mnW == std::min(0.0, rcU[bUsr]) + std::min(0.0, colRedCost[bStd]));   // block order: 2 then 7
                for (dipgpu::DecompVar *v : varsW) {
                    const int b = v->getBlockId();
                    CHECK(b == bStd || b == bUsr);
                    CHECK(v->getReducedCost() == (b == bStd ? colRedCost[bStd] : rcU[bUsr]));
                }
            }
            for (dipgpu::DecompVar *v : varsW) delete v;
        }
    }
    // ---- one pricer over two GPUs (the same GPU twice on a one-GPU box): generateVars partitioned by blocks,
    // the ladder by dual vectors, the expansion by shards -- against the one-GPU pricer above
    {
        int nd = 0;
        {
            // (the C ABI has no device query of its own: a second context on ordinal 1 tells)
            dipgpu_ctx *probe = nullptr;
            nd = dipgpu_create(&probe, 1) == DIPGPU_OK ? 2 : 1;
            dipgpu_destroy(probe);
        }
        dipgpu::Pricer multi(std::vector<int>{0, nd == 2 ? 1 : 0});
        int32_t nDev = 0, peer = 0;
        dipgpu_device_count(multi.ctx(), &nDev, &peer);
        CHECK(dipgpu_set_option(multi.ctx(), "shard_rounds", 2) == DIPGPU_OK);  // partition the blocks (the default keeps co-resident rounds on one GPU)
        CHECK(nDev == 2);
        multi.setModelCore(R, N, rowStart.data(), colInd.data(), val.data(), R, m);
        multi.setObjective(cost.data(), N);
        multi.setKnapsackBlocks(m, blockColStart.data(), weight.data(), capacity.data());
        std::vector<int32_t> support;
        for (int r = 0; r < R + m; ++r)
            if (r < n || r >= R || u[(size_t)r] != 0.0) support.push_back(r);
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                pricer->setDualSupport(support);
                multi.setDualSupport(support);
            }
            dipgpu::DecompVarList v1, v2;
            double mn1 = 0, mn2 = 0;
            const int n1 = pricer->generateVars(u.data(), (int)u.size(), dipgpu::PHASE_PRICE2, 1e-4, v1, mn1);
            const int n2 = multi.generateVars(u.data(), (int)u.size(), dipgpu::PHASE_PRICE2, 1e-4, v2, mn2);
            CHECK(n1 == n2 && mn1 == mn2 && multi.cells() == pricer->cells());
            for (int b = 0; b < m; ++b) CHECK(multi.blockObj()[b] == pricer->blockObj()[b] && multi.blockRedCost()[b] == pricer->blockRedCost()[b]);
            auto it = v1.begin();
            for (dipgpu::DecompVar *v : v2) {
                CHECK(v->getBlockId() == (*it)->getBlockId() && v->ind == (*it)->ind && v->getReducedCost() == (*it)->getReducedCost() &&
                      v->getOriginalCost() == (*it)->getOriginalCost());
                ++it;
            }
            for (dipgpu::DecompVar *v : v1) delete v;
            for (dipgpu::DecompVar *v : v2) delete v;
            std::vector<int64_t> csA, csB;
            std::vector<int32_t> riA, riB;
            std::vector<double> vaA, vaB;
            std::vector<uint64_t> hsA, hsB;
            pricer->expandColumns(1e-6, csA, riA, vaA, hsA);
            multi.expandColumns(1e-6, csB, riB, vaB, hsB);
            CHECK(csA == csB && riA == riB && vaA == vaB && hsA == hsB);
        }
        std::vector<double> u2(u), alA, alB;
        for (int j = 0; j < n; ++j) u2[(size_t)j] = unif(seed) * 120.0;
        pricer->setDualCentre(u.data(), (int)u.size());
        multi.setDualCentre(u.data(), (int)u.size());
        pricer->generateVarsLadder(u2.data(), (int)u2.size(), 0.2, 0.90, 5, dipgpu::PHASE_PRICE2, 1e-4, alA);
        multi.generateVarsLadder(u2.data(), (int)u2.size(), 0.2, 0.90, 5, dipgpu::PHASE_PRICE2, 1e-4, alB);
        CHECK(alA == alB);
        for (int k = 0; k < 5; ++k) {
            const dipgpu::Pricer::Round &a = pricer->ladderRound(k), &b = multi.ladderRound(k);
            CHECK(a.cols.nCols == b.cols.nCols && a.cols.mostNegReducedCost == b.cols.mostNegReducedCost);
            CHECK(a.blockObj == b.blockObj && a.blockRedCost == b.blockRedCost && a.colBlock == b.colBlock);
            CHECK(std::vector<int32_t>(a.colInd.begin(), a.colInd.begin() + a.cols.nInd) ==
                  std::vector<int32_t>(b.colInd.begin(), b.colInd.begin() + b.cols.nInd));
        }
        pricer->setDualSupport({});
    }
    delete pricer;
    std::printf("PARITY_OK kept=%d mostNegRC=%.9f cells=%lld launches=%lld\n", kept, mostNegRC, (long long)cells,
                (long long)launches);
    return 0;
}

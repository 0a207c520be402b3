"""Parity of the rows SURVEY.md section 8(f) ranks after the round itself, through the C ABI:
  f-2  waiting-column re-pricing   DecompVarPool::setReducedCosts   (Dip/src/DecompVarPool.cpp:185-203, :22-40)
  f-3  Wentges smoothing + ladder  DecompAlgoPC::adjustMasterDualSolution (Dip/src/DecompAlgoPC.cpp:126-212),
       the alpha *= 0.9 retry of DecompAlgo::addVarsToPool (Dip/src/DecompAlgo.cpp:5642-5655),
       and batches of dual vectors priced in one submission.
Every batched / stabilised round must equal, field by field, the single-vector round of the same
dual vector (which the other GPU tests pin against the oracle), and the oracle itself."""
import numpy as np
import pytest

from dip_b200 import instances as I
from dip_b200.pricer import PROFIT_FP64, PROFIT_PISINGER_INT, RoundResult

pytestmark = pytest.mark.gpu


def _same_round(a, b):
    np.testing.assert_array_equal(a.blockObj, b.blockObj)
    np.testing.assert_array_equal(a.blockNnz, b.blockNnz)
    np.testing.assert_array_equal(a.blockRedCost, b.blockRedCost)
    np.testing.assert_array_equal(a.blockOrigCost, b.blockOrigCost)
    np.testing.assert_array_equal(a.blockMostNegRC, b.blockMostNegRC)
    assert a.nCols == b.nCols and a.cells == b.cells
    assert a.mostNegReducedCost == b.mostNegReducedCost
    np.testing.assert_array_equal(a.colBlock[:a.nCols], b.colBlock[:b.nCols])
    np.testing.assert_array_equal(a.colStart[:a.nCols + 1], b.colStart[:b.nCols + 1])
    n = int(a.colStart[a.nCols])
    np.testing.assert_array_equal(a.colInd[:n], b.colInd[:n])


def _oracle_round(oracle, model, u, profit_mode=0, eps=1e-4):
    return oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                              model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                              profit_mode=profit_mode, red_cost_eps=eps)


def _check_vs_oracle(res, ref):
    np.testing.assert_array_equal(res.blockObj, ref.z)
    np.testing.assert_array_equal(res.blockNnz, ref.col_nnz)
    kept = [b for b in range(len(ref.z)) if ref.col_keep[b]]
    assert list(res.colBlock[:res.nCols]) == kept
    for k, b in enumerate(kept):
        assert list(res.column(k)) == list(ref.col_ind[b])
    assert res.cells == ref.cells


@pytest.mark.parametrize("profit_mode", [PROFIT_FP64, PROFIT_PISINGER_INT])
@pytest.mark.parametrize("which", ["gap_d", "gap_c"])
def test_batch_equals_single_rounds(oracle, pricer, which, profit_mode):
    model = I.gap_pricing_model(I.gap_class_d() if which == "gap_d" else I.gap_class_c())
    pricer.setModel(model, profit_mode)
    rng = np.random.default_rng(31)
    U = np.stack([I.gap_duals(model, rng) for _ in range(7)])
    if profit_mode == PROFIT_PISINGER_INT:
        U = np.round(U, 3)   # a few decimals, as calcScaleFactor expects
    batch = pricer.priceRoundsBatch(U)
    assert len(batch) == 7
    for v in range(7):
        single = pricer.priceRound(U[v])
        _same_round(batch[v], single)
        _check_vs_oracle(batch[v], _oracle_round(oracle, model, U[v], profit_mode))
    # a second batch on the same context (publication epochs must carry over), other phase
    batch1 = pricer.priceRoundsBatch(U[:3], phase=1)
    for v in range(3):
        _same_round(batch1[v], pricer.priceRound(U[v], phase=1))


def test_batch_gap_e_and_select_for_expansion(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_e())
    pricer.setModel(model)
    rng = np.random.default_rng(8)
    U = np.stack([I.gap_duals(model, rng) for _ in range(5)])
    batch = pricer.priceRoundsBatch(U)
    _same_round(batch[0], pricer.priceRound(U[0]))
    _same_round(batch[4], pricer.priceRound(U[4]))
    _check_vs_oracle(batch[2], _oracle_round(oracle, model, U[2]))
    # master columns of batch result 3 == those of the single round of vector 3
    batch = pricer.priceRoundsBatch(U)
    pricer.selectBatchResult(3)
    cs, ri, va, hs = pricer.expandColumns()
    pricer.priceRound(U[3])
    cs2, ri2, va2, hs2 = pricer.expandColumns()
    np.testing.assert_array_equal(cs, cs2)
    np.testing.assert_array_equal(ri, ri2)
    np.testing.assert_array_equal(va, va2)
    np.testing.assert_array_equal(hs, hs2)


def test_batch_on_a_large_shard_uses_the_unfused_path(oracle, pricer):
    """More than 256 blocks: no fused kernel, the batch is priced vector after vector."""
    b = I.random_batch(300, 0, 40, 0, 120, seed=3)
    N = int(b.block_col_start[-1])
    m = 300
    # identity-like core: one assignment-style row per column group so that redCost = c - u[row]
    R = 40
    rows = np.arange(N) % R
    order = np.argsort(rows, kind="stable")
    rs = np.zeros(R + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=R), out=rs[1:])
    ci = order.astype(np.int32)
    v = np.ones(N)
    pricer.setModelCore(R, N, rs, ci, v, R, m)
    pricer.setObjective(b.orig_cost)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    rng = np.random.default_rng(5)
    U = rng.uniform(0, 2 * b.orig_cost.mean(), size=(3, R + m))
    U[:, R:] = rng.uniform(-10, 0, size=(3, m))
    batch = pricer.priceRoundsBatch(U)
    for k in range(3):
        _same_round(batch[k], pricer.priceRound(U[k]))


def test_smoothing_ladder_matches_oracle_and_single_rounds(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model)
    rng = np.random.default_rng(12)
    center = I.gap_duals(model, rng)
    uRM = I.gap_duals(model, rng)
    pricer.stabSetCenter(center)
    results, alphas = pricer.priceRoundStabLadder(uRM, 0.1, 6)
    a = 0.1
    for k in range(6):
        assert alphas[k] == a                         # DualStabAlpha *= 0.90, repeated (DecompAlgo.cpp:5646)
        st = oracle.smooth_duals(a, center, uRM)      # DecompAlgoPC.cpp:179-181
        np.testing.assert_array_equal(pricer.smoothedDuals(k), st)
        _same_round(results[k], pricer.priceRound(st))
        _check_vs_oracle(results[k], _oracle_round(oracle, model, st))
        a = a * 0.90
    # accepting step 2 moves the centre there (DecompAlgoPC.h:124-133); the next smoothing starts from it
    results, alphas = pricer.priceRoundStabLadder(uRM, 0.1, 3)
    st2 = oracle.smooth_duals(alphas[2], center, uRM)
    pricer.stabAccept(2)
    uRM2 = I.gap_duals(model, rng)
    res, st = pricer.priceRoundStab(uRM2, 0.25)
    np.testing.assert_array_equal(st, oracle.smooth_duals(0.25, st2, uRM2))
    _same_round(res, pricer.priceRound(st))
    # the single stabilised round is also the context's "last round" for the expansion
    cs, ri, va, hs = pricer.expandColumns()
    assert len(cs) == res.nCols + 1


def test_first_stabilised_call_uses_dual_rm_as_centre(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_c())
    pricer.setModel(model)
    uRM = I.gap_duals(model, np.random.default_rng(2))
    res, st = pricer.priceRoundStab(uRM, 0.1)
    np.testing.assert_array_equal(st, oracle.smooth_duals(0.1, uRM, uRM))   # DecompAlgoPC.cpp:163-173
    _same_round(res, pricer.priceRound(st))


def _expanded_pool(oracle, pricer, model, n_rounds, rng):
    """Prices n_rounds dual vectors, expands the kept columns and appends them to the pool; returns
    the host copy (colStart, rowInd, val, origCost) assembled from the oracle's own expansion."""
    cs_all, ri_all, va_all, oc_all = [0], [], [], []
    for _ in range(n_rounds):
        res = pricer.priceRound(I.gap_duals(model, rng))
        cs, ri, va, hs = pricer.expandColumns()
        which = np.arange(res.nCols, dtype=np.int32)[::2]   # every other column "survives the duplicate test"
        pricer.poolAppendExpanded(which)
        for k in which:
            rr, vv = oracle.expand_column(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows,
                                          model.m, res.column(k), int(res.colBlock[k]))
            ri_all.extend(rr.tolist())
            va_all.extend(vv.tolist())
            cs_all.append(len(ri_all))
            oc_all.append(float(res.colOrigCost[k]))
    return np.array(cs_all, np.int64), np.array(ri_all, np.int32), np.array(va_all), np.array(oc_all)


def test_pool_reduced_costs_from_expanded_columns(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model)
    rng = np.random.default_rng(21)
    cs, ri, va, oc = _expanded_pool(oracle, pricer, model, 4, rng)
    assert pricer.poolSize() == (len(oc), len(ri))
    u = I.gap_duals(model, rng)
    for feasible in (True, False):
        got, found = pricer.poolSetReducedCosts(u, feasible)
        ref, ref_found = oracle.pool_reduced_costs(cs, ri, va, oc, u, feasible)
        np.testing.assert_array_equal(got, ref)          # same products, same order: bit-equal
        assert found == ref_found
    # the duals are resident now: generateVars with u == NULL equals the round of u
    res_resident = pricer.priceRoundResident()
    res = pricer.priceRound(u)
    _same_round(res_resident, res)
    # addVarsFromPool took columns 1, 4, 5 out: the rest keep their values
    keep = np.array([k for k in range(len(oc)) if k not in (1, 4, 5)], np.int32)
    pricer.poolKeep(keep)
    got, _ = pricer.poolSetReducedCosts(u)
    np.testing.assert_array_equal(got, ref_after := oracle.pool_reduced_costs(cs, ri, va, oc, u)[0][keep])
    assert pricer.poolSize()[0] == len(keep) and len(ref_after) == len(keep)
    pricer.poolClear()
    got, found = pricer.poolSetReducedCosts(u)
    assert len(got) == 0 and not found


@pytest.mark.parametrize("pool_kernel", [1, 0])
def test_pool_from_host_columns_ragged_and_large(oracle, pricer, pool_kernel):
    """Host-built waiting columns: empty columns, a column longer than one staging chunk (512), many
    columns; threshold semantics of found_negrc_var (<= -1e-10).  pool_kernel: the persistent kernel (tiles
    streamed across tile borders) and the one-CTA-per-16-tiles kernel."""
    pricer.set_option("pool_kernel", pool_kernel)
    rng = np.random.default_rng(6)
    R, N, m = 3000, 64, 4
    rs = np.zeros(R + 1, np.int64)
    pricer.setModelCore(R, N, rs, np.zeros(0, np.int32), np.zeros(0), R, m)
    lens = np.concatenate([[0, 1, 2, 700, 0, 513, 512, 31, 33], rng.integers(0, 90, size=5000)])
    cs = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=cs[1:])
    ri = np.concatenate([np.sort(rng.choice(R + m, size=l, replace=False)) for l in lens]).astype(np.int32)
    va = rng.uniform(-3, 3, size=len(ri))
    oc = rng.uniform(-5, 50, size=len(lens))
    pricer.poolAppend(cs[:101], ri[:cs[100]], va[:cs[100]], oc[:100])                    # two appends: growth path
    pricer.poolAppend(cs[100:] - cs[100], ri[cs[100]:], va[cs[100]:], oc[100:])
    u = rng.standard_normal(R + m)
    u[rng.random(R + m) < 0.3] = 0.0
    got, found = pricer.poolSetReducedCosts(u)
    ref, ref_found = oracle.pool_reduced_costs(cs, ri, va, oc, u)
    np.testing.assert_array_equal(got, ref)
    assert found == ref_found is True
    # no column prices out
    pricer.poolClear()
    pricer.poolAppend(np.array([0, 1, 1], np.int64), np.array([5], np.int32), np.array([1.0]), np.array([2.0, 0.0]))
    uu = np.zeros(R + m)
    uu[5] = 2.0
    got, found = pricer.poolSetReducedCosts(uu)
    np.testing.assert_array_equal(got, [0.0, 0.0])
    assert not found
    uu[5] = 2.0 + 1e-9
    got, found = pricer.poolSetReducedCosts(uu)
    assert found and got[0] < -1e-10


def test_spmv_tile_shapes_are_bit_equal(oracle, pricer):
    """Every tiling of the stream kernel (entry budget, column budget, ring depth, warps, duals in
    shared / global memory) gives the same bits as the oracle's scatter."""
    rng = np.random.default_rng(14)
    R, N, m, nBase = 900, 20000, 8, 800
    nnz = 150_000
    r = np.sort(rng.integers(0, R, size=nnz))
    rs = np.zeros(R + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=R), out=rs[1:])
    ci = rng.integers(0, N, size=nnz).astype(np.int32)
    v = rng.uniform(-10, 10, size=nnz)
    c = rng.uniform(0, 100, size=N)
    pricer.setModelCore(R, N, rs, ci, v, nBase, m)
    pricer.setObjective(c)
    u = rng.standard_normal(R + m)
    u[rng.random(R + m) < 0.3] = 0.0
    ref = oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(u, nBase, m), c)
    ref1 = oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(u, nBase, m), c, True)
    for chunk, max_cols, stages, warps, usmem in [(0, 0, 0, 0, -1), (64, 32, 2, 1, 1), (96, 64, 3, 5, 0), (512, 128, 4, 32, 1),
                                                   (2048, 128, 2, 7, 1), (200, 96, 8, 3, 0)]:
        pricer.set_option("spmv_chunk", chunk)
        pricer.set_option("spmv_max_cols", max_cols)
        pricer.set_option("spmv_stages", stages)
        pricer.set_option("spmv_warps", warps)
        pricer.set_option("spmv_u_smem", usmem)
        np.testing.assert_array_equal(pricer.calcRedCost(u), ref)
        np.testing.assert_array_equal(pricer.calcRedCost(u, 1), ref1)


# ------------------------------------------------------------------ f-4: node bounds inside the blocks


def _check_bounded(oracle, pricer, b, lb, ub, eps=1e-4):
    res = pricer.solveBlocks(b.red_cost, b.alpha, eps)
    ref = oracle.solve_blocks_bounded(b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha, lb, ub,
                                      red_cost_eps=eps)
    np.testing.assert_array_equal(res.blockObj, ref.z)                  # bit-exact optimum of the free part
    np.testing.assert_array_equal(res.blockNnz, ref.col_nnz)
    np.testing.assert_array_equal(res.blockRedCost, ref.col_red_cost)   # same additions, ascending index order
    np.testing.assert_array_equal(res.blockOrigCost, ref.col_orig_cost)
    np.testing.assert_array_equal(res.blockMostNegRC, ref.most_neg_rc)
    kept = [k for k in range(len(ref.z)) if ref.col_keep[k]]
    assert list(res.colBlock[:res.nCols]) == kept
    for k, blk in enumerate(kept):
        assert list(res.column(k)) == list(ref.col_ind[blk])
    assert res.cells == ref.cells
    assert res.mostNegReducedCost == ref.most_neg_reduced_cost
    return res, ref


@pytest.mark.parametrize("m,nhi,caphi,seed", [(40, 60, 150, 1), (200, 120, 400, 2), (400, 50, 120, 3), (6, 2300, 6000, 4)])
def test_node_bounds_in_the_blocks(oracle, pricer, m, nhi, caphi, seed):
    """Small shards run the fused kernel (parallel merge of the fixed columns), > 256 blocks and
    multi-chunk blocks the three-kernel path (backward merge in the backtrack kernel)."""
    b = I.random_batch(m, 1, nhi, 10, caphi, seed)
    N = int(b.block_col_start[-1])
    rng = np.random.default_rng(100 + seed)
    pricer.setObjective(b.orig_cost)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    lb = (rng.random(N) < 0.03).astype(np.float64)
    ub = 1.0 - (rng.random(N) < 0.10).astype(np.float64)
    ub = np.maximum(ub, lb)
    lb[b.block_col_start[0]:b.block_col_start[1]] = 1.0        # block 0: everything fixed in -> cannot fit
    ub[b.block_col_start[0]:b.block_col_start[1]] = 1.0
    if m > 2:
        j = int(b.block_col_start[2])
        lb[j], ub[j] = 1.0, 0.0                                   # block 2: lb > ub
    pricer.setColBounds(lb, ub)
    res, ref = _check_bounded(oracle, pricer, b, lb, ub)
    assert np.isinf(res.blockRedCost[0]) and res.blockNnz[0] == 0
    # fixed-to-1 columns with non-negative reduced cost are in the returned columns
    forced_pos = [j for j in np.nonzero(lb > 0.5)[0] if b.red_cost[j] >= 0]
    in_cols = set(int(x) for k in range(res.nCols) for x in res.column(k))
    kept_blocks = set(int(x) for x in res.colBlock[:res.nCols])
    for j in forced_pos:
        blk = int(np.searchsorted(b.block_col_start, j, side="right") - 1)
        if blk in kept_blocks:
            assert int(j) in in_cols
    # bounds off again: the plain round
    pricer.setColBounds(None, None)
    ref0 = oracle.solve_blocks(b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha)
    res0 = pricer.solveBlocks(b.red_cost, b.alpha)
    np.testing.assert_array_equal(res0.blockObj, ref0.z)
    np.testing.assert_array_equal(res0.blockNnz, ref0.col_nnz)


def test_node_bounds_through_the_round_and_batches(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model)
    rng = np.random.default_rng(9)
    lb = (rng.random(model.N) < 0.01).astype(np.float64)
    ub = np.maximum(1.0 - (rng.random(model.N) < 0.05).astype(np.float64), lb)
    pricer.setColBounds(lb, ub)
    U = np.stack([I.gap_duals(model, rng) for _ in range(4)])
    batch = pricer.priceRoundsBatch(U)
    for v in range(4):
        res, rcx = pricer.priceRound(U[v], want_redcost=True)
        _same_round(batch[v], res)
        ref = oracle.solve_blocks_bounded(model.block_col_start, model.weight, model.capacity, rcx, model.objective,
                                          U[v][model.n_base_rows:model.n_base_rows + model.m], lb, ub)
        np.testing.assert_array_equal(res.blockObj, ref.z)
        np.testing.assert_array_equal(res.blockRedCost, ref.col_red_cost)
        assert [list(res.column(k)) for k in range(res.nCols)] == [list(ref.col_ind[b]) for b in range(model.m) if ref.col_keep[b]]


def _bounded_pisinger_reference(oracle, b, lb, ub, eps=1e-4):
    """Node bounds in Pisinger mode, composed from pinned pieces: the fixed columns leave the problem (fixed to 1: in the
    column, their weight off the capacity), GAP_KnapPisinger::solve's restatement prices the free columns, the sums run
    over the merged column in ascending index order."""
    m = len(b.capacity)
    rc_free = b.red_cost.copy()
    cap_left = b.capacity.astype(np.int64).copy()
    forced = []
    for k in range(m):
        s, e = b.block_col_start[k], b.block_col_start[k + 1]
        one = lb[s:e] > 0.5
        zero = ub[s:e] < 0.5
        rc_free[s:e][one | zero] = 1.0          # no item
        cap_left[k] -= int(b.weight[s:e][one].sum())
        forced.append((s + np.nonzero(one)[0]).tolist())
    feas = cap_left >= 0
    ref = oracle.solve_blocks(b.block_col_start, b.weight, np.maximum(cap_left, 0).astype(np.int32), rc_free, b.orig_cost,
                              np.zeros(m), profit_mode=1, red_cost_eps=eps)
    cols, rcs, ocs, zs = [], [], [], []
    for k in range(m):
        if not feas[k]:
            cols.append([]); rcs.append(np.inf); ocs.append(0.0); zs.append(-np.inf)
            continue
        col = sorted(list(ref.col_ind[k]) + forced[k])
        r = 0.0
        o = 0.0
        for j in col:
            r += b.red_cost[j]
            o += b.orig_cost[j]
        cols.append(col); rcs.append(r - b.alpha[k]); ocs.append(o); zs.append(ref.z[k])
    return cols, np.array(rcs), np.array(ocs), np.array(zs)


@pytest.mark.parametrize("m,nhi,caphi,seed", [(40, 60, 200, 1), (700, 30, 90, 2), (6, 1400, 600, 3)])
def test_node_bounds_in_pisinger_mode(oracle, pricer, m, nhi, caphi, seed):
    """DecompAlgo::setSubProbBounds (DecompAlgo.cpp:2352-2371) with GAP_KnapPisinger pricing: fused and large shards,
    trivial cases (n <= 2 free items, all fit) and DP blocks, blocks whose fixed columns do not fit."""
    b = I.random_batch(m, 0, nhi, 0, caphi, seed)
    # reduced costs with a few decimals, so that the scale factor varies
    b.red_cost = np.round(b.red_cost, 2)
    rng = np.random.default_rng(seed)
    n = len(b.weight)
    lb = (rng.random(n) < 0.03).astype(np.float64)
    ub = 1.0 - (rng.random(n) < 0.06) * (1 - lb)
    pricer.setObjective(b.orig_cost)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity, PROFIT_PISINGER_INT)
    pricer.setColBounds(lb, ub)
    res = pricer.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
    cols, rcs, ocs, zs = _bounded_pisinger_reference(oracle, b, lb, ub)
    np.testing.assert_array_equal(res.blockObj, zs)
    np.testing.assert_array_equal(res.blockRedCost, rcs)
    np.testing.assert_array_equal(res.blockOrigCost, ocs)
    kept = [k for k in range(m) if rcs[k] < np.inf]
    assert list(res.colBlock[:res.nCols]) == kept
    for q, k in enumerate(kept):
        assert list(res.column(q)) == cols[k], k
    # bounds removed: the plain Pisinger-mode round again
    pricer.setColBounds(None, None)
    res2 = pricer.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
    ref2 = oracle.solve_blocks(b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha, profit_mode=1,
                               red_cost_eps=-np.inf)
    np.testing.assert_array_equal(res2.blockObj, ref2.z)


# ------------------------------------------------------------------ round + expansion in one submission


def test_price_round_expand_equals_the_two_calls(oracle, pricer):
    """dipgpu_price_round_expand == dipgpu_price_round followed by dipgpu_expand_columns, field by field --
    rounds that keep every block, some blocks, and none (the expansion kernel is launched before the host
    knows how many columns the round kept)."""
    model = I.gap_pricing_model(I.gap_class_d(6, 40, seed=5), branch_rows=True)
    pricer.setModel(model)
    u = I.gap_duals(model, np.random.default_rng(21))
    for eps, tiny_cap in ((1e-4, False), (-np.inf, False), (np.inf, False), (1e-4, True)):
        two = pricer.priceRound(u, redCostEps=eps, result=RoundResult(model.m, model.m, model.N))
        cs2, ri2, va2, hs2 = pricer.expandColumns()
        one, (cs1, ri1, va1, hs1) = pricer.priceRoundExpand(u, redCostEps=eps, cap_nnz=4 if tiny_cap else 1 << 16)
        _same_round(one, two)
        np.testing.assert_array_equal(cs1, cs2)
        np.testing.assert_array_equal(ri1, ri2)
        np.testing.assert_array_equal(va1, va2)
        np.testing.assert_array_equal(hs1, hs2)
        if eps == np.inf:
            assert one.nCols == 0 and len(ri1) == 0 and list(cs1) == [0]
    # against the oracle's restatement of addVarsToPool, column by column
    one, (cs1, ri1, va1, hs1) = pricer.priceRoundExpand(u)
    for k in range(one.nCols):
        rows, vals = oracle.expand_column(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows,
                                          model.m, one.column(k), int(one.colBlock[k]))
        np.testing.assert_array_equal(ri1[cs1[k]:cs1[k + 1]], rows)
        np.testing.assert_array_equal(va1[cs1[k]:cs1[k + 1]], vals)


def test_price_round_expand_in_pisinger_mode_and_on_a_shard(oracle, pricer):
    """The one-submission call where the DP is not the fused fp64 kernel's default path: scaled-integer profits,
    and a context restricted to a block range (a multi-GPU shard)."""
    model = I.gap_pricing_model(I.gap_class_d(6, 40, seed=9), branch_rows=True)
    pricer.setModel(model, profit_mode=PROFIT_PISINGER_INT)
    u = I.gap_duals(model, np.random.default_rng(33))
    two = pricer.priceRound(u, result=RoundResult(model.m, model.m, model.N))
    cs2, ri2, va2, hs2 = pricer.expandColumns()
    one, (cs1, ri1, va1, hs1) = pricer.priceRoundExpand(u)
    _same_round(one, two)
    for x, y in ((cs1, cs2), (ri1, ri2), (va1, va2), (hs1, hs2)):
        np.testing.assert_array_equal(x, y)
    pricer.setModel(model)
    pricer.setBlockRange(2, 3)
    two = pricer.priceRound(u, result=RoundResult(model.m, model.m, model.N))
    cs2, ri2, va2, hs2 = pricer.expandColumns()
    one, (cs1, ri1, va1, hs1) = pricer.priceRoundExpand(u)
    _same_round(one, two)
    assert all(2 <= int(b) < 5 for b in one.colBlock[:one.nCols])
    for x, y in ((cs1, cs2), (ri1, ri2), (va1, va2), (hs1, hs2)):
        np.testing.assert_array_equal(x, y)
    # a block list on the shard: only blocks of the range are accepted
    from dip_b200.capi import DipGpuError, ERR_INVALID
    res, _ = pricer.priceRoundWhich(u, [4, 2])
    assert all(int(b) in (2, 4) for b in res.colBlock[:res.nCols])
    with pytest.raises(DipGpuError) as e:
        pricer.priceRoundWhich(u, [1])
    assert e.value.code == ERR_INVALID


# ------------------------------------------------------------------ a10: block subsets, per-block user duals


def _check_which(oracle, pricer, model, u, which, user=None, eps=1e-4):
    res, solved_all = pricer.priceRoundWhich(u, which, user, redCostEps=eps)
    ref = oracle.price_round_which(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                                   model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                                   which, user, red_cost_eps=eps)
    assert solved_all == ref["solved_all"]
    want = sorted(ref["new_vars"], key=lambda v: v[0])       # the reference pushes in list order, we in block order
    assert res.nCols == len(want)
    for k, (b, ind, rc, oc) in enumerate(want):
        assert int(res.colBlock[k]) == b and list(res.column(k)) == list(ind)
        assert res.colRedCost[k] == rc and res.colOrigCost[k] == oc
    np.testing.assert_array_equal(res.blockMostNegRC, ref["most_neg_vec"])
    assert res.mostNegReducedCost == ref["most_neg_reduced_cost"]
    return res, solved_all


def test_block_subset_and_user_duals(oracle, pricer):
    """generateVars' round-robin branch (DecompAlgo.cpp:4929-5149): the listed blocks only, some with their own
    master dual vector; the all-blocks fallback when none of them prices out."""
    from dip_b200.capi import DipGpuError, ERR_INVALID
    model = I.gap_pricing_model(I.gap_class_d(6, 40, seed=3), branch_rows=False)
    pricer.setModel(model)
    rng = np.random.default_rng(70)
    u = I.gap_duals(model, np.random.default_rng(12))
    full = pricer.priceRound(u, result=RoundResult(model.m, model.m, model.N))
    assert full.nCols >= 3
    # one block, standard duals
    res, all_ = _check_which(oracle, pricer, model, u, [3])
    assert not all_ and res.nCols == 1 and res.colRedCost[0] == full.blockRedCost[3]
    # several blocks in any order, two of them with the same user vector, one with another
    ua = u * rng.uniform(0.8, 1.2, size=len(u))
    ub = u * rng.uniform(0.5, 1.5, size=len(u))
    res, all_ = _check_which(oracle, pricer, model, u, [5, 1, 2, 0], {1: ua, 2: ua, 0: ub})
    assert not all_
    cs, ri, va, hs = pricer.expandColumns()                   # the composed round is the "last round"
    assert len(cs) == res.nCols + 1
    for k in range(res.nCols):
        rows, vals = oracle.expand_column(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows,
                                          model.m, res.column(k), int(res.colBlock[k]))
        np.testing.assert_array_equal(ri[cs[k]:cs[k + 1]], rows)
        np.testing.assert_array_equal(va[cs[k]:cs[k + 1]], vals)
    # no listed block prices out -> all blocks at the standard duals; block 4's user vector has a hopeless alpha,
    # block 2 does not price out under u either
    u2 = u.copy()
    u2[model.n_base_rows + 2] = -1e6
    uc = u2.copy()
    uc[model.n_base_rows + 4] = -1e6
    res, all_ = _check_which(oracle, pricer, model, u2, [4, 2], {4: uc})
    assert all_ and 4 in list(res.colBlock[:res.nCols]) and 2 not in list(res.colBlock[:res.nCols])
    # ... and a user vector that prices worse than u does: mostNegRC of that block is the smaller of the two
    ud = u2.copy()
    ud[model.n_base_rows + 4] = u2[model.n_base_rows + 4] - 1e-3    # alpha lower -> rc higher, may still be < 0
    _check_which(oracle, pricer, model, u2, [4], {4: ud})
    # empty list: nothing prices out -> the full round
    res, all_ = _check_which(oracle, pricer, model, u, [])
    assert all_
    _same_round(res, full)
    for bad in ([1, 1], [model.m], [-1]):
        with pytest.raises(DipGpuError) as e:
            pricer.priceRoundWhich(u, bad)
        assert e.value.code == ERR_INVALID


def test_block_subset_in_pisinger_mode_and_large_shards(oracle, pricer):
    """The same branch where the batch runs vector after vector (non-fused shapes) and in the scaled-integer mode."""
    model = I.gap_pricing_model(I.gap_class_d(4, 30, seed=8), branch_rows=False)
    pricer.setModel(model, profit_mode=PROFIT_PISINGER_INT)
    u = I.gap_duals(model, np.random.default_rng(5))
    ua = u * np.random.default_rng(3).uniform(0.9, 1.1, size=len(u))
    res, solved_all = pricer.priceRoundWhich(u, [2, 0], {0: ua})
    ref = oracle.price_round_which(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                                   model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                                   [2, 0], {0: ua}, profit_mode=PROFIT_PISINGER_INT)
    assert solved_all == ref["solved_all"]
    want = sorted(ref["new_vars"], key=lambda v: v[0])
    assert [int(b) for b in res.colBlock[:res.nCols]] == [v[0] for v in want]
    for k, (b, ind, rc, oc) in enumerate(want):
        assert list(res.column(k)) == list(ind) and res.colRedCost[k] == rc


# ------------------------------------------------------------------ sparse upload of the duals


def test_dual_support_upload_equals_dense(oracle, pricer):
    """With the rows that can carry a dual declared (free branch rows have none), only those cross PCIe:
    rounds, batches, the smoothing ladder and the pool pass must equal their dense-upload results."""
    from dip_b200.capi import DipGpuError, ERR_INVALID
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model)
    n = model.meta["n"]
    rng = np.random.default_rng(40)
    nb = model.n_base_rows - n
    branch = np.sort(rng.choice(nb, size=nb // 50, replace=False)) + n
    support = np.concatenate([np.arange(n), branch, np.arange(model.n_base_rows, model.u_len)]).astype(np.int32)

    def duals():
        u = np.zeros(model.u_len)
        u[:n] = rng.uniform(0, 2 * model.objective.mean(), size=n)
        u[branch] = rng.uniform(-3, 3, size=len(branch))
        u[model.n_base_rows:] = rng.uniform(-10, 0, size=model.m)
        return u
    U = np.stack([duals() for _ in range(5)])
    dense = []
    for v in range(5):
        r, rcx = pricer.priceRound(U[v], want_redcost=True, result=RoundResult(model.m, model.m, model.N))
        dense.append((r, rcx))
    h0 = pricer.counters()["h2d_bytes"]
    pricer.setDualSupport(support)
    for v in (3, 0, 4):     # any order: entries of the previous vector must not linger
        r, rcx = pricer.priceRound(U[v], want_redcost=True)
        np.testing.assert_array_equal(rcx, dense[v][1])
        _same_round(r, dense[v][0])
    assert pricer.counters()["h2d_bytes"] - h0 == 3 * 8 * len(support)
    batch = pricer.priceRoundsBatch(U)
    for v in range(5):
        _same_round(batch[v], dense[v][0])
    pricer.stabSetCenter(U[0])
    results, alphas = pricer.priceRoundStabLadder(U[1], 0.2, 3)
    for k in range(3):
        st = oracle.smooth_duals(alphas[k], U[0], U[1])
        np.testing.assert_array_equal(pricer.smoothedDuals(k), st)
    batch = pricer.priceRoundsBatch(U[:2])      # after the ladder wrote dense vectors into the batch buffer
    _same_round(batch[1], dense[1][0])
    # dense again
    pricer.setDualSupport(None)
    u_dense = duals()
    u_dense[n + 7] = 1.25                        # outside the old support
    r, rcx = pricer.priceRound(u_dense, want_redcost=True)
    ref = oracle.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val,
                              oracle.adjust_duals(u_dense, model.n_base_rows, model.m), model.objective)
    np.testing.assert_array_equal(rcx, ref)
    with pytest.raises(DipGpuError) as e:
        pricer.setDualSupport(np.array([3, 2], np.int32))
    assert e.value.code == ERR_INVALID


def test_pinned_duals_and_host_mirror_paths(pricer):
    """Two transfer shortcuts of the host entry points must not change a bit: with a dual support and u in
    page-locked memory the kernel picks u[support] out of the caller's buffer (no host gather), and the fused
    kernel stores the packed result into mapped host memory itself (no D2H copy).  All four combinations of
    the options against the plain path (pageable u, both off)."""
    from dip_b200 import capi
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model)
    n = model.meta["n"]
    rng = np.random.default_rng(41)
    nb = model.n_base_rows - n
    branch = np.sort(rng.choice(nb, size=nb // 40, replace=False)) + n
    support = np.concatenate([np.arange(n), branch, np.arange(model.n_base_rows, model.u_len)]).astype(np.int32)
    U = np.zeros((4, model.u_len))
    for v in range(4):
        U[v, :n] = rng.uniform(0, 2 * model.objective.mean(), size=n)
        U[v, branch] = rng.uniform(-3, 3, size=len(branch))
        U[v, model.n_base_rows:] = rng.uniform(-10, 0, size=model.m)
    pricer.set_option("gather_pinned_duals", 0)
    pricer.set_option("result_zero_copy", 0)
    plain = [pricer.priceRound(U[v], result=RoundResult(model.m, model.m, model.N)) for v in range(4)]
    pin = capi.PinnedArray(U.shape, np.float64)
    pin.array[:] = U
    pricer.setDualSupport(support)
    for gather, mirror in ((1, 1), (1, 0), (0, 1), (0, 0)):
        pricer.set_option("gather_pinned_duals", gather)
        pricer.set_option("result_zero_copy", mirror)
        for v in (2, 0, 3, 1):
            _same_round(pricer.priceRound(pin.array[v]), plain[v])
            _same_round(pricer.priceRound(U[v]), plain[v])            # pageable u: always the host gather
        r = pricer.priceRound(pin.array[1])
        cs, ri, va, hs = pricer.expandColumns()                        # reads the device copy of the result
        assert len(cs) == r.nCols + 1
        batch = pricer.priceRoundsBatch(pin.array)
        for v in range(4):
            _same_round(batch[v], plain[v])
    pin.free()


@pytest.mark.parametrize("pool_kernel,compact", [(1, 1), (1, 0), (2, 0), (0, 0)])
def test_pool_with_dual_support_bitmap(oracle, pricer, pool_kernel, compact):
    """A large pool (>= 32 Ki columns) with a declared dual support runs the kernel variant that skips the
    gathers of rows without a dual (shared-memory bitmap); bit-equal to the oracle and to the plain variant.
    pool_kernel 1: persistent kernel, gathers from L2; 2: u[support] compacted in shared memory and found by rank
    in the bitmap; 0: the one-CTA-per-16-tiles kernel."""
    pricer.set_option("pool_kernel", pool_kernel)
    pricer.set_option("pool_compact", compact)   # 1: the pass reads a copy of the pool without the rows that cannot carry a dual
    rng = np.random.default_rng(61)
    R, m = 40_000, 16
    pricer.setModelCore(R, 8, np.zeros(R + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), R, m)
    n_cols = 40_000
    lens = rng.integers(0, 60, size=n_cols)
    cs = np.zeros(n_cols + 1, np.int64)
    np.cumsum(lens, out=cs[1:])
    ri = rng.integers(0, R + m, size=cs[-1]).astype(np.int32)
    va = rng.uniform(-2, 2, size=cs[-1])
    oc = rng.uniform(-1, 30, size=n_cols)
    pricer.poolAppend(cs, ri, va, oc)
    support = np.sort(rng.choice(R + m, size=(R + m) // 10, replace=False)).astype(np.int32)
    u = np.zeros(R + m)
    u[support] = rng.standard_normal(len(support))
    plain, found0 = pricer.poolSetReducedCosts(u)
    ref, ref_found = oracle.pool_reduced_costs(cs, ri, va, oc, u)
    np.testing.assert_array_equal(plain, ref)
    pricer.setDualSupport(support)
    got, found = pricer.poolSetReducedCosts(u)
    np.testing.assert_array_equal(got, ref)
    assert found == ref_found == found0
    got_ray, _ = pricer.poolSetReducedCosts(None, feasible=False)     # resident duals, dual-ray form
    np.testing.assert_array_equal(got_ray, oracle.pool_reduced_costs(cs, ri, va, oc, u, False)[0])
    # the pool and the support change: the compacted copy follows (columns removed, columns appended, another support)
    keep = np.arange(0, n_cols, 3).astype(np.int32)
    pricer.poolKeep(keep)
    extra = 500
    lens2 = rng.integers(0, 40, size=extra)
    cs2 = np.zeros(extra + 1, np.int64)
    np.cumsum(lens2, out=cs2[1:])
    ri2 = rng.integers(0, R + m, size=cs2[-1]).astype(np.int32)
    va2 = np.round(rng.uniform(-2, 2, size=cs2[-1]), 1)
    oc2 = rng.uniform(-1, 30, size=extra)
    pricer.poolAppend(cs2, ri2, va2, oc2)
    lens_k = (cs[1:] - cs[:-1])[keep]
    csk = np.zeros(len(keep) + 1, np.int64)
    np.cumsum(lens_k, out=csk[1:])
    rik = np.concatenate([ri[cs[j]:cs[j + 1]] for j in keep])
    vak = np.concatenate([va[cs[j]:cs[j + 1]] for j in keep])
    csA = np.concatenate([csk, csk[-1] + cs2[1:]])
    riA, vaA, ocA = np.concatenate([rik, ri2]), np.concatenate([vak, va2]), np.concatenate([oc[keep], oc2])
    got, _ = pricer.poolSetReducedCosts(u)
    np.testing.assert_array_equal(got, oracle.pool_reduced_costs(csA, riA, vaA, ocA, u)[0])
    support2 = np.sort(rng.choice(R + m, size=(R + m) // 7, replace=False)).astype(np.int32)
    u2 = np.zeros(R + m)
    u2[support2] = rng.standard_normal(len(support2))
    pricer.setDualSupport(support2)
    got, _ = pricer.poolSetReducedCosts(u2)
    np.testing.assert_array_equal(got, oracle.pool_reduced_costs(csA, riA, vaA, ocA, u2)[0])


# ------------------------------------------------------------------ f-4: multiple-choice knapsack blocks (MMKP)


def _mckp_instance(seed, m, g_lo, g_hi, k_lo, k_hi, cap_lo, cap_hi, w_hi):
    rng = np.random.default_rng(seed)
    groups_per_block = rng.integers(g_lo, g_hi + 1, size=m)
    sizes = rng.integers(k_lo, k_hi + 1, size=int(groups_per_block.sum()))
    gcs = np.zeros(len(sizes) + 1, np.int32)
    np.cumsum(sizes, out=gcs[1:])
    bgs = np.zeros(m + 1, np.int32)
    np.cumsum(groups_per_block, out=bgs[1:])
    N = int(gcs[-1])
    w = rng.integers(0, w_hi + 1, size=N).astype(np.int32)
    cap = rng.integers(cap_lo, cap_hi + 1, size=m).astype(np.int32)
    rc = rng.uniform(-20, 20, size=N)
    tie = rng.random(N) < 0.1
    rc[tie] = np.round(rc[tie])                                      # some ties
    oc = rng.integers(1, 100, size=N).astype(np.float64)
    alpha = rng.uniform(-10, 10, size=m)
    return bgs, gcs, w, cap, rc, oc, alpha


def _check_mckp(oracle, pricer, bgs, gcs, w, cap, rc, oc, alpha, eps=1e-4):
    pricer.setObjective(oc)
    pricer.setMckpBlocks(bgs, gcs, w, cap)
    res = pricer.solveBlocks(rc, alpha, eps)
    ref = oracle.solve_mckp_blocks(bgs, gcs, w, cap, rc, oc, alpha, red_cost_eps=eps)
    np.testing.assert_array_equal(res.blockObj, ref.z)                  # the same additions in the same order
    np.testing.assert_array_equal(res.blockNnz, ref.col_nnz)
    np.testing.assert_array_equal(res.blockRedCost, ref.col_red_cost)
    np.testing.assert_array_equal(res.blockOrigCost, ref.col_orig_cost)
    np.testing.assert_array_equal(res.blockMostNegRC, ref.most_neg_rc)
    kept = [b for b in range(len(cap)) if ref.col_keep[b]]
    assert list(res.colBlock[:res.nCols]) == kept
    for k, b in enumerate(kept):
        assert list(res.column(k)) == list(ref.col_ind[b])
    assert res.cells == ref.cells and res.mostNegReducedCost == ref.most_neg_reduced_cost
    return res, ref


@pytest.mark.parametrize("seed,m,g_hi,k_hi,cap_hi,w_hi", [(1, 50, 12, 10, 120, 30), (2, 400, 6, 5, 60, 20),
                                                         (3, 8, 60, 30, 3000, 150), (4, 30, 10, 8, 25, 15),
                                                         (5, 6, 5, 700, 900, 60)])   # groups beyond the staging buffer
def test_mckp_blocks(oracle, pricer, seed, m, g_hi, k_hi, cap_hi, w_hi):
    """Exactly one column per group (MMKP_MCKnap::solveMCKnap): optimum, column, sums, infeasible blocks."""
    inst = _mckp_instance(seed, m, 0 if seed == 4 else 1, g_hi, 1, k_hi, cap_hi // 3, cap_hi, w_hi)
    res, ref = _check_mckp(oracle, pricer, *inst)
    if seed == 4:
        assert np.isinf(res.blockRedCost).any()       # tight capacities: some block has no feasible choice


def test_mckp_through_the_round_and_in_batches(oracle, pricer):
    """The MMKP structure end to end: reduced costs from the SpMV, then the multiple-choice blocks."""
    bgs, gcs, w, cap, _, oc, _ = _mckp_instance(9, 12, 4, 10, 3, 9, 40, 120, 25)
    m, N = 12, int(gcs[-1])
    rng = np.random.default_rng(10)
    R = 30                                              # dense-ish side constraints over all columns
    rows = np.repeat(np.arange(R), 40)
    ci = np.concatenate([np.sort(rng.choice(N, size=40, replace=False)) for _ in range(R)]).astype(np.int32)
    rs = np.arange(R + 1, dtype=np.int64) * 40
    v = rng.uniform(1, 9, size=len(ci))
    pricer.setModelCore(R, N, rs, ci, v, R, m)
    pricer.setObjective(oc)
    pricer.setMckpBlocks(bgs, gcs, w, cap)
    U = rng.uniform(-2, 2, size=(3, R + m))
    batch = pricer.priceRoundsBatch(U)
    for k in range(3):
        res, rcx = pricer.priceRound(U[k], want_redcost=True)
        ref_rc = oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(U[k], R, m), oc)
        np.testing.assert_array_equal(rcx, ref_rc)
        ref = oracle.solve_mckp_blocks(bgs, gcs, w, cap, rcx, oc, U[k][R:R + m])
        np.testing.assert_array_equal(res.blockObj, ref.z)
        assert [list(res.column(j)) for j in range(res.nCols)] == [list(ref.col_ind[b]) for b in range(m) if ref.col_keep[b]]
        _same_round(batch[k], res)
        cs, ri, va, hs = pricer.expandColumns()          # master columns of the multiple-choice columns
        assert len(cs) == res.nCols + 1

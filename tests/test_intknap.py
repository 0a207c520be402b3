"""Integer knapsack blocks = the pricing problem of DIP's cutting-stock example (third variant):
kp / relaxed_solver of Dip/src/dippy/tests/test_cutting_stock.py:98-179.  The oracle against the reference's own kp
(tests/golden/intknap_kp_cases.json, made by oracle/gen_golden.py from the function executed as is), the host-side
driver on the reference's instance (objective 2.0, :34), and -- on the GPU, through the C ABI -- dipgpu_set_intknap_blocks
against the oracle: values, multiplicities, sums, filter."""
import numpy as np
import pytest

from examples import csp_colgen as CSP


# ------------------------------------------------------------------ oracle (CPU)
def test_oracle_kp_matches_reference_function(oracle, golden_intknap):
    assert len(golden_intknap) >= 150 and sum(c["plain"] for c in golden_intknap) >= 50
    for c in golden_intknap:
        z, cnt = oracle.intknap_kp(c["obj"], c["w"], c["cap"])
        assert z == c["z"], c["tag"]                       # bit-exact value (the chain sums of the recursion)
        assert list(cnt) == c["counts"], c["tag"]          # the multiplicities of the recursion's solution


def _relaxed_solver_restated(red_cost, orig_cost, items, lengths, use, cap, kp):
    """relaxed_solver of test_cutting_stock.py:98-151 on index lists, reduced cost without the convexity dual."""
    idx = [j for j in items if red_cost[j] < 0]
    z, sol = kp([-red_cost[j] for j in idx], [lengths[j] for j in idx], cap)
    if sum(sol) > 0:
        col = {j: s for j, s in zip(idx, sol) if s > 0}
        col[use] = 1
        return -z + red_cost[use], col
    return 0.0, {}


def test_oracle_blocks_match_restated_relaxed_solver(oracle):
    rng = np.random.default_rng(5)
    for trial in range(60):
        m = int(rng.integers(1, 6))
        n_items = rng.integers(0, 7, size=m)
        bcs = np.concatenate([[0], np.cumsum(n_items)]).astype(np.int32)
        n_it = int(bcs[-1])
        use = (n_it + rng.permutation(m)).astype(np.int32)
        N = n_it + m
        w = rng.integers(1, 9, size=n_it).astype(np.int32)
        cap = rng.integers(0, 30, size=m).astype(np.int32)
        rc = np.round(rng.uniform(-3, 1.5, size=N), 3)
        oc = rng.integers(0, 3, size=N).astype(np.float64)
        alpha = rng.uniform(-1, 1, size=m)
        ref = oracle.solve_intknap_blocks(bcs, w, cap, use, rc, oc, alpha, red_cost_eps=1e-4)
        for b in range(m):
            items = list(range(bcs[b], bcs[b + 1]))
            r, col = _relaxed_solver_restated(rc, oc, items, {j: int(w[j]) for j in items}, int(use[b]), int(cap[b]),
                                              lambda o, ww, c: oracle.intknap_kp(o, ww, c))
            assert ref.col_red_cost[b] == r - alpha[b]
            assert dict(zip(ref.col_ind[b].tolist(), ref.col_cnt[b].tolist())) == col
            assert list(ref.col_ind[b]) == sorted(col)
            assert ref.col_orig_cost[b] == sum(oc[j] * col[j] for j in sorted(col))
            assert bool(ref.col_keep[b]) == (ref.col_red_cost[b] < -1e-4)


def test_cutting_stock_instance_reaches_2_with_the_oracle_backend(oracle):
    """The reference pins this instance at objective 2.0 (test_cutting_stock.py:20, :34)."""
    model = CSP.csp_model()
    lb, best, tr, cols = CSP.solve(model, CSP.oracle_pricer(model))
    assert lb == 2 and best == 2.0
    assert abs(tr.lower_bound[3] - 1.875) < 1e-9           # the root's Dantzig-Wolfe bound


# ------------------------------------------------------------------ GPU
def _blocks_from_cases(cases, with_use, rng):
    """One block per golden case: item columns with redCost = -obj (plus a few non-negative ones in between)."""
    bcs, w, rc, cap, keep_idx = [0], [], [], [], []
    for c in cases:
        idx = []
        for p, ww in zip(c["obj"], c["w"]):
            if rng.random() < 0.2:                          # a column that is no item (redCost >= 0)
                rc.append(float(rng.uniform(0, 2)))
                w.append(int(rng.integers(1, 9)))
            idx.append(len(rc))
            rc.append(-float(p))
            w.append(int(ww))
        keep_idx.append(idx)
        bcs.append(len(rc))
        cap.append(c["cap"])
    m = len(cases)
    n_it = len(rc)
    use = None
    if with_use:
        use = (n_it + rng.permutation(m)).astype(np.int32)
        rc += list(np.round(rng.uniform(-1, 2, size=m), 3))
    return (np.array(bcs, np.int32), np.array(w, np.int32), np.array(cap, np.int32), use, np.array(rc, np.float64),
            keep_idx)


def _check_int_round(res, ref, m):
    np.testing.assert_array_equal(res.blockObj, ref.z)
    np.testing.assert_array_equal(res.blockRedCost, ref.col_red_cost)
    np.testing.assert_array_equal(res.blockOrigCost, ref.col_orig_cost)
    np.testing.assert_array_equal(res.blockNnz, ref.col_nnz)
    np.testing.assert_array_equal(res.blockMostNegRC, ref.most_neg_rc)
    assert res.mostNegReducedCost == ref.most_neg_reduced_cost
    kept = [b for b in range(m) if ref.col_keep[b]]
    assert list(res.colBlock[:res.nCols]) == kept
    for k, b in enumerate(kept):
        assert list(res.column(k)) == list(ref.col_ind[b]), b
        assert list(res.elements(k)) == [float(v) for v in ref.col_cnt[b]], b
    assert res.cells == ref.cells


@pytest.mark.gpu
@pytest.mark.parametrize("with_use", [False, True])
def test_gpu_kp_golden_cases(oracle, pricer, golden_intknap, with_use):
    """Every golden case of the reference's kp as one block of one round: value and multiplicities bit-exact."""
    rng = np.random.default_rng(3)
    bcs, w, cap, use, rc, idx = _blocks_from_cases(golden_intknap, with_use, rng)
    m = len(golden_intknap)
    oc = rng.integers(0, 4, size=len(rc)).astype(np.float64)
    alpha = np.round(rng.uniform(-0.5, 0.5, size=m), 3)
    pricer.setObjective(oc)
    pricer.setIntKnapBlocks(bcs, w, cap, use)
    res = pricer.solveBlocks(rc, alpha, redCostEps=-np.inf)
    for b, c in enumerate(golden_intknap):
        assert res.blockObj[b] == c["z"], c["tag"]
    kept = list(res.colBlock[:res.nCols])
    assert kept == list(range(m))
    for b, c in enumerate(golden_intknap):
        got = dict(zip(res.column(b).tolist(), res.elements(b).tolist()))
        want = {idx[b][i]: float(n) for i, n in enumerate(c["counts"]) if n > 0}
        if with_use and want:
            want[int(use[b])] = 1.0
        assert got == want, c["tag"]
    ref = oracle.solve_intknap_blocks(bcs, w, cap, use, rc, oc, alpha, red_cost_eps=-np.inf)
    _check_int_round(res, ref, m)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,m,n_hi,cap_hi,w_lo,w_hi", [(1, 300, 12, 60, 1, 9), (2, 40, 200, 2000, 1, 300),
                                                          (3, 8, 700, 12000, 40, 900), (4, 500, 3, 25, 5, 9),
                                                          (5, 3, 50, 20000, 1, 50)])
def test_gpu_intknap_blocks_random(oracle, pricer, seed, m, n_hi, cap_hi, w_lo, w_hi):
    """Ragged blocks incl. empty ones and zero capacities, tables in shared memory and (seed 5) in global scratch,
    the filter at DIP's RedCostEpsilon."""
    rng = np.random.default_rng(seed)
    n_items = rng.integers(0, n_hi + 1, size=m)
    n_items[rng.integers(0, m)] = 0
    bcs = np.concatenate([[0], np.cumsum(n_items)]).astype(np.int32)
    n_it = int(bcs[-1])
    use = (n_it + rng.permutation(m)).astype(np.int32)
    use[rng.integers(0, m)] = -1
    N = n_it + m
    w = rng.integers(w_lo, w_hi + 1, size=n_it).astype(np.int32)
    cap = rng.integers(cap_hi // 2, cap_hi + 1, size=m).astype(np.int32)
    cap[rng.integers(0, m)] = 0
    rc = rng.uniform(-1.0, 0.3, size=N) * np.concatenate([w.astype(np.float64) ** 0.8, np.ones(m)])
    if seed == 4:
        rc = np.round(rc, 1)                                # many ties
    oc = rng.integers(0, 3, size=N).astype(np.float64)
    alpha = rng.uniform(-2, 2, size=m)
    pricer.setObjective(oc)
    pricer.setIntKnapBlocks(bcs, w, cap, use)
    for eps in (1e-4, -np.inf):
        res = pricer.solveBlocks(rc, alpha, redCostEps=eps)
        ref = oracle.solve_intknap_blocks(bcs, w, cap, use, rc, oc, alpha, red_cost_eps=eps)
        _check_int_round(res, ref, m)


@pytest.mark.gpu
def test_gpu_cutting_stock_search_equals_oracle_search(oracle):
    """The reference's instance end to end: reduced costs from the SpMV over the core (demand, ordering and branch
    rows), integer knapsack blocks, filter -- the GPU-driven branch-and-price and the oracle-driven one agree on every
    round's duals, columns and bounds, and reach the objective the reference pins (2.0)."""
    model = CSP.csp_model()
    price, p = CSP.gpu_pricer(model)
    try:
        lb, best, tr, cols = CSP.solve(model, price)
    finally:
        p.close()
    lb2, best2, tr2, cols2 = CSP.solve(model, CSP.oracle_pricer(model))
    assert (lb, best) == (2, 2.0) and (lb2, best2) == (2, 2.0)
    assert tr.nodes == tr2.nodes and len(tr.z_rmp) == len(tr2.z_rmp)
    for a, b in zip(tr.duals, tr2.duals):
        np.testing.assert_array_equal(a, b)
    assert tr.columns == tr2.columns
    assert tr.lower_bound == tr2.lower_bound


@pytest.mark.gpu
def test_gpu_intknap_errors_and_batches(oracle, pricer):
    from dip_b200 import capi
    bcs = np.array([0, 2, 4], np.int32)
    with pytest.raises(capi.DipGpuError) as e:
        pricer.setIntKnapBlocks(bcs, np.array([3, 0, 2, 2], np.int32), np.array([9, 9], np.int32))   # weight 0
    assert e.value.code == capi.ERR_INVALID
    with pytest.raises(capi.DipGpuError) as e:
        pricer.setIntKnapBlocks(bcs, np.array([3, 1, 2, 2], np.int32), np.array([9, 9], np.int32), np.array([1, 5], np.int32))
    assert e.value.code == capi.ERR_INVALID                 # the use column of block 0 is one of its items
    # a small model through the round and in a batch of dual vectors
    model = CSP.csp_model(lengths=(11, 8, 6, 5), demand=(2, 3, 2, 4), roll_len=31)
    price, p = CSP.gpu_pricer(model)
    try:
        rng = np.random.default_rng(8)
        U = np.zeros((3, model.R + model.m))
        U[:, :model.meta["n_own"]] = rng.uniform(0, 0.6, size=(3, model.meta["n_own"]))
        U[:, model.R:] = rng.uniform(-0.2, 0.2, size=(3, model.m))
        batch = p.priceRoundsBatch(U)
        for k in range(3):
            res, rcx = p.priceRound(U[k], want_redcost=True)
            ref_rc = oracle.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val,
                                         oracle.adjust_duals(U[k], model.R, model.m), model.objective)
            np.testing.assert_array_equal(rcx, ref_rc)
            ref = oracle.solve_intknap_blocks(model.block_col_start, model.weight, model.capacity, model.use_col, rcx,
                                              model.objective, U[k][model.R:])
            _check_int_round(res, ref, model.m)
            _check_int_round(batch[k], ref, model.m)
        with pytest.raises(capi.DipGpuError) as e:
            p.expandColumns()
        assert e.value.code == capi.ERR_UNSUPPORTED
    finally:
        p.close()

"""The CPU oracle against the reference's own outputs (tests/golden, made by oracle/gen_golden.py
from gap_func.knapsack01 executed unchanged, KnapsackOptimizeHS compiled from the reference, and
the in-tree instance gap0515-2.dat)."""
import numpy as np

from dip_b200.instances import GapInstance, gap_pricing_model


def test_knapsack01_matches_reference_function(oracle, golden_knapsack):
    assert len(golden_knapsack) >= 150
    for c in golden_knapsack:
        z, sol = oracle.knapsack01(c["obj"], c["w"], c["cap"])
        assert z == c["z"], c["tag"]                      # bit-exact optimum
        assert list(sol) == c["solution"], c["tag"]       # same argmax, same (descending) order


def test_knapsack01_decision_table_backtrack(oracle, golden_knapsack):
    for c in golden_knapsack[:40]:
        z, sol, added = oracle.knapsack01(c["obj"], c["w"], c["cap"], want_table=True)
        j, i, walk = c["cap"], len(c["w"]) - 1, []
        while i >= 0 and j >= 0:
            if added[i, j]:
                walk.append(i)
                j -= int(c["w"][i])
            i -= 1
        assert walk == list(sol)


def test_integer_cases_match_reference_hs(oracle, golden_knapsack):
    """Integer profits: z* is unique, so the reference's B&B must equal the DP value."""
    n = 0
    for c in golden_knapsack:
        if c["hs_z"] is None or c["hs_status"] != 0:
            continue
        z, _ = oracle.knapsack01(c["obj"], c["w"], c["cap"])
        assert z == c["hs_z"], c["tag"]
        n += 1
    assert n >= 20


def test_live_reference_hs_if_built(oracle):
    if not oracle.have_ref():
        import pytest
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(7)
    for _ in range(50):
        n = int(rng.integers(3, 30))
        w = rng.integers(1, 30, size=n)
        p = rng.integers(1, 60, size=n).astype(np.float64)
        cap = int(rng.integers(10, 120))
        st, zhs, x = oracle.ref_knapsack_hs(p, w, cap)
        z, sol = oracle.knapsack01(p, w.astype(np.int32), cap)
        if st == 0:
            assert z == zhs
            assert (w * x).sum() <= cap


def _gap0515(g):
    inst = GapInstance(g["m"], g["n"], np.array(g["cost"]).reshape(g["m"], g["n"]),
                       np.array(g["weight"], np.int32).reshape(g["m"], g["n"]),
                       np.array(g["capacity"], np.int32), "gap0515-2")
    return gap_pricing_model(inst)


def test_gap0515_rounds(oracle, golden_gap0515):
    g = golden_gap0515
    model = _gap0515(g)
    assert model.R == g["n"] + 2 * g["m"] * g["n"]
    for r in g["rounds"]:
        res = oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective,
                                 r["u"], model.n_base_rows, model.m, model.block_col_start, model.weight,
                                 model.capacity)
        np.testing.assert_array_equal(res.red_cost_x, r["red_cost"])   # same order -> bit-exact
        alpha = r["u"][model.n_base_rows:model.n_base_rows + model.m]
        for b, gb in enumerate(r["blocks"]):
            assert res.z[b] == gb["z"]
            assert list(res.col_ind[b]) == gb["ind"]
            rc = sum(r["red_cost"][j] for j in gb["ind"]) - alpha[b]
            assert abs(res.col_red_cost[b] - rc) <= 1e-12 * max(1.0, abs(rc))
            assert res.most_neg_rc[b] == min(0.0, res.col_red_cost[b])
            assert bool(res.col_keep[b]) == (res.col_red_cost[b] < -1e-4)
        assert res.most_neg_reduced_cost == float(np.add.reduce(res.most_neg_rc)) or \
            abs(res.most_neg_reduced_cost - res.most_neg_rc.sum()) < 1e-12


def test_adjust_duals_and_phase1(oracle):
    u = np.arange(10, dtype=np.float64)
    out = oracle.adjust_duals(u, n_base=4, m=3)
    np.testing.assert_array_equal(out, [0, 1, 2, 3, 7, 8, 9])
    # 2 rows, 3 cols
    rs = np.array([0, 2, 3], np.int64)
    ci = np.array([0, 2, 1], np.int32)
    v = np.array([1.0, 2.0, 3.0])
    c = np.array([10.0, 20.0, 30.0])
    ua = np.array([0.5, -1.0])
    np.testing.assert_array_equal(oracle.calc_redcost(2, 3, rs, ci, v, ua, c), [9.5, 23.0, 29.0])
    np.testing.assert_array_equal(oracle.calc_redcost(2, 3, rs, ci, v, ua, c, phase1=True), [-0.5, 3.0, -1.0])


def test_string_hash(oracle):
    assert oracle.string_hash([3, 17, 42], [1.0, 1.0, 1.0]) == "3_1_17_1_42_1_"
    assert oracle.string_hash([5], [0.123456789]) == "5_0.123457_"


# ---------------------------------------------------------------- rows 8(f): pool, smoothing, node bounds


def test_pool_reduced_costs_and_smoothing_restatements(oracle):
    """DecompVarPool::setReducedCosts / DecompAlgoPC::adjustMasterDualSolution restated in C against
    their formulas in numpy (Dip/src/DecompVarPool.cpp:22-40, Dip/src/DecompAlgoPC.cpp:179-181)."""
    rng = np.random.default_rng(3)
    nrows = 500
    lens = rng.integers(0, 40, size=200)
    cs = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=cs[1:])
    ri = rng.integers(0, nrows, size=cs[-1]).astype(np.int32)
    va = rng.uniform(-2, 2, size=cs[-1])
    oc = rng.uniform(-1, 5, size=len(lens))
    u = rng.standard_normal(nrows)
    for feasible in (True, False):
        got, found = oracle.pool_reduced_costs(cs, ri, va, oc, u, feasible)
        for s in range(len(lens)):
            dp = 0.0
            for k in range(cs[s + 1] - 1, cs[s] - 1, -1):   # CoinPackedVectorBase::dotProduct, last entry first
                dp += va[k] * u[ri[k]]
            assert got[s] == (oc[s] - dp if feasible else -dp)
        assert found == bool(np.any(got <= -1e-10))
    a = 0.1 * 0.9 * 0.9
    d, rm = rng.standard_normal(1000), rng.standard_normal(1000)
    np.testing.assert_array_equal(oracle.smooth_duals(a, d, rm), (a * d) + ((1.0 - a) * rm))


def test_bounded_blocks_against_enumeration(oracle):
    """Node bounds enforced in the block (DecompAlgo.cpp:6403-6405): the restatement (fixed columns +
    knapsack01 on the rest) reaches the optimum of min sum rc*x s.t. knapsack, lb <= x <= ub by
    enumeration, returns fixed-to-1 columns whatever their reduced cost, and flags infeasible blocks."""
    rng = np.random.default_rng(11)
    for trial in range(60):
        m = 4
        n = int(rng.integers(1, 11))
        bcs = (np.arange(m + 1) * n).astype(np.int32)
        w = rng.integers(1, 12, size=m * n).astype(np.int32)
        cap = rng.integers(0, 30, size=m).astype(np.int32)
        rc = np.round(rng.uniform(-10, 6, size=m * n), 2)
        oc = rng.uniform(0, 10, size=m * n)
        alpha = rng.uniform(-3, 0, size=m)
        lb = (rng.random(m * n) < 0.15).astype(np.float64)
        ub = 1.0 - (rng.random(m * n) < 0.2).astype(np.float64)
        if trial % 3:
            ub = np.maximum(ub, lb)          # no clashes in most trials
        res = oracle.solve_blocks_bounded(bcs, w, cap, rc, oc, alpha, lb, ub)
        for b in range(m):
            s = slice(b * n, (b + 1) * n)
            best = None
            for mask in range(1 << n):
                x = np.array([(mask >> j) & 1 for j in range(n)], np.float64)
                if np.any(x < lb[s]) or np.any(x > ub[s]) or np.dot(x, w[s]) > cap[b]:
                    continue
                val = float(np.dot(x, rc[s]))
                if best is None or val < best - 1e-9:
                    best = val
            if best is None:
                assert res.col_nnz[b] == 0 and np.isinf(res.col_red_cost[b]) and res.most_neg_rc[b] == 0.0
                assert not res.col_keep[b] and res.z[b] == -np.inf
                continue
            ind = res.col_ind[b] - b * n
            x = np.zeros(n)
            x[ind] = 1
            assert np.all(x >= lb[s]) and np.all(x <= ub[s]) and np.dot(x, w[s]) <= cap[b]
            assert abs((res.col_red_cost[b] + alpha[b]) - best) <= 1e-9
            assert list(ind) == sorted(ind)
        assert res.most_neg_reduced_cost == float(np.sum(res.most_neg_rc))


def test_mckp_blocks_against_enumeration(oracle):
    """Multiple-choice knapsack blocks (MMKP_MCKnap::solveMCKnap, MMKP_MCKnap.cpp:172-383): the exact DP over
    the groups reaches the optimum of min sum rc s.t. exactly one column per group, weights <= capacity, found
    by enumeration; infeasible blocks return no column."""
    import itertools
    rng = np.random.default_rng(23)
    for trial in range(40):
        m = 3
        groups_per_block = rng.integers(1, 5, size=m)
        sizes = rng.integers(1, 5, size=int(groups_per_block.sum()))
        gcs = np.zeros(len(sizes) + 1, np.int32)
        np.cumsum(sizes, out=gcs[1:])
        bgs = np.zeros(m + 1, np.int32)
        np.cumsum(groups_per_block, out=bgs[1:])
        N = int(gcs[-1])
        w = rng.integers(0, 9, size=N).astype(np.int32)
        cap = rng.integers(0, 18, size=m).astype(np.int32)
        rc = np.round(rng.uniform(-8, 8, size=N), 3)
        oc = rng.uniform(0, 10, size=N)
        alpha = rng.uniform(-3, 0, size=m)
        res = oracle.solve_mckp_blocks(bgs, gcs, w, cap, rc, oc, alpha)
        for b in range(m):
            gl = [range(gcs[g], gcs[g + 1]) for g in range(bgs[b], bgs[b + 1])]
            best = None
            for pick in itertools.product(*gl):
                if sum(w[k] for k in pick) <= cap[b]:
                    v = sum(rc[k] for k in pick)
                    if best is None or v < best - 1e-12:
                        best = v
            if best is None:
                assert res.col_nnz[b] == 0 and np.isinf(res.col_red_cost[b]) and not res.col_keep[b]
                continue
            assert abs(res.z[b] - best) <= 1e-9
            assert len(res.col_ind[b]) == len(gl) and all(k in gl[i] for i, k in enumerate(res.col_ind[b]))
            assert sum(w[k] for k in res.col_ind[b]) <= cap[b]
            assert abs(res.col_red_cost[b] + alpha[b] - best) <= 1e-9


def test_round_robin_branch_restatement(oracle):
    """price_round_which (DecompAlgo.cpp:4929-5149 + the filter loop): listing every block with the standard
    duals is the plain round; an empty list falls back to it; a user vector only changes its own block."""
    from dip_b200 import instances as I
    model = I.gap_pricing_model(I.gap_class_d(5, 30, seed=2), branch_rows=False)
    u = I.gap_duals(model, np.random.default_rng(4))
    args = (model.R, model.N, model.row_start, model.col_ind, model.val, model.objective)
    blk = (model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity)
    plain = oracle.price_round(*args, u, *blk)
    kept = [b for b in range(model.m) if plain.col_keep[b]]
    assert kept
    everything = oracle.price_round_which(*args, u, *blk, list(range(model.m)))
    assert not everything["solved_all"] and [v[0] for v in everything["new_vars"]] == kept
    assert everything["most_neg_reduced_cost"] == plain.most_neg_reduced_cost
    nothing = oracle.price_round_which(*args, u, *blk, [])
    assert nothing["solved_all"] and [v[0] for v in nothing["new_vars"]] == kept
    assert nothing["most_neg_reduced_cost"] == plain.most_neg_reduced_cost
    b = kept[0]
    uh = u.copy()
    uh[model.n_base_rows + b] -= 0.5                       # alpha lower by 0.5 -> that block's rc higher by 0.5
    one = oracle.price_round_which(*args, u, *blk, [b], {b: uh})
    if not one["solved_all"]:
        assert len(one["new_vars"]) == 1 and one["new_vars"][0][0] == b
        assert abs(one["new_vars"][0][2] - (plain.col_red_cost[b] + 0.5)) <= 1e-9
        assert one["most_neg_reduced_cost"] == one["new_vars"][0][2]
    reorder = oracle.price_round_which(*args, u, *blk, kept[::-1])
    assert [v[0] for v in reorder["new_vars"]] == kept[::-1]   # the reference's push order = list order

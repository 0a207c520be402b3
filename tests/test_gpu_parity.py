"""Parity of the CUDA pricing path (through the C ABI) against the CPU oracle and the committed
golden vectors.  Bar (BASELINE.json north_star): knapsack optima BIT-EXACT, column sets identical,
reduced costs within 1e-12 relative."""
import numpy as np
import pytest

from dip_b200 import instances as I

pytestmark = pytest.mark.gpu
RC_RTOL = 1e-12  # north_star: "Reduced costs must agree within 1e-12 relative"


def _batch_from_cases(cases):
    n = np.array([len(c["w"]) for c in cases])
    bcs = np.zeros(len(cases) + 1, np.int32)
    np.cumsum(n, out=bcs[1:])
    w = np.concatenate([c["w"] for c in cases]).astype(np.int32)
    rc = -np.concatenate([c["obj"] for c in cases])
    cap = np.array([c["cap"] for c in cases], np.int32)
    return bcs, w, cap, rc


def _check_against_oracle(oracle, pricer, bcs, w, cap, rc, oc, alpha, eps=1e-4, profit_mode=0):
    m = len(cap)
    pricer.setObjective(oc)
    pricer.setKnapsackBlocks(bcs, w, cap, profit_mode)
    res = pricer.solveBlocks(rc, alpha, eps)
    ref = oracle.solve_blocks(bcs, w, cap, rc, oc, alpha, red_cost_eps=eps, profit_mode=profit_mode)
    np.testing.assert_array_equal(res.blockObj, ref.z)                 # bit-exact optimum
    np.testing.assert_array_equal(res.blockNnz, ref.col_nnz)
    np.testing.assert_allclose(res.blockRedCost, ref.col_red_cost, rtol=RC_RTOL, atol=1e-12)
    np.testing.assert_allclose(res.blockOrigCost, ref.col_orig_cost, rtol=RC_RTOL, atol=1e-12)
    np.testing.assert_array_equal(res.blockMostNegRC, np.minimum(0.0, res.blockRedCost))
    kept = [b for b in range(m) if ref.col_keep[b]]
    # a block whose rc sits within rounding of -eps may legitimately flip; none in these inputs
    assert res.nCols == len(kept)
    assert list(res.colBlock[:res.nCols]) == kept
    for k, b in enumerate(kept):
        assert list(res.column(k)) == list(ref.col_ind[b]), f"block {b}"
        assert res.colRedCost[k] == res.blockRedCost[b]
        assert res.colOrigCost[k] == res.blockOrigCost[b]
    assert res.cells == ref.cells
    assert abs(res.mostNegReducedCost - ref.most_neg_reduced_cost) <= RC_RTOL * max(1.0, abs(ref.most_neg_reduced_cost))
    return res, ref


def test_golden_knapsack01_cases(pricer, golden_knapsack):
    """Every golden case (outputs of the reference's own knapsack01) as one block of one batch."""
    cases = [c for c in golden_knapsack if len(c["w"]) >= 2]  # n==1: documented divergence of the Python
    bcs, w, cap, rc = _batch_from_cases(cases)
    pricer.setObjective(np.zeros(len(w)))
    pricer.setKnapsackBlocks(bcs, w, cap)
    res = pricer.solveBlocks(rc, np.zeros(len(cases)), redCostEps=-np.inf)
    assert res.nCols == len(cases)
    for k, c in enumerate(cases):
        assert res.blockObj[k] == c["z"], c["tag"]
        got = [int(j - bcs[k]) for j in res.column(k)]
        assert got == sorted(c["solution"]), c["tag"]


def test_golden_gap0515_rounds(pricer, golden_gap0515):
    g = golden_gap0515
    inst = I.GapInstance(g["m"], g["n"], np.array(g["cost"]).reshape(g["m"], g["n"]),
                         np.array(g["weight"], np.int32).reshape(g["m"], g["n"]),
                         np.array(g["capacity"], np.int32))
    model = I.gap_pricing_model(inst)
    pricer.setModel(model)
    for r in g["rounds"]:
        res, rcx = pricer.priceRound(r["u"], redCostEps=-np.inf, want_redcost=True)
        # the stream SpMV adds each column in the reference's scatter order, unfused: bit-equal
        np.testing.assert_array_equal(rcx, r["red_cost"])
        for b, gb in enumerate(r["blocks"]):
            assert res.blockObj[b] == gb["z"]
            assert list(res.column(b)) == gb["ind"]
        # stage (b) alone on the golden reduced costs is always bit-exact
        alpha = r["u"][model.n_base_rows:model.n_base_rows + model.m]
        res2 = pricer.solveBlocks(r["red_cost"], alpha, redCostEps=-np.inf)
        for b, gb in enumerate(r["blocks"]):
            assert res2.blockObj[b] == gb["z"]
            assert list(res2.column(b)) == gb["ind"]


@pytest.mark.parametrize("seed,m,nhi,caphi", [(1, 64, 40, 90), (2, 300, 130, 700), (3, 33, 700, 2500),
                                              (4, 7, 2300, 9000), (5, 1, 50, 12000)])
def test_random_ragged_batches(oracle, pricer, seed, m, nhi, caphi):
    b = I.random_batch(m, 0, nhi, 0, caphi, seed)
    _check_against_oracle(oracle, pricer, b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha)


@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_filter_fused_and_standalone(oracle, pricer, fuse):
    """Stage (c) as its own kernels (0), as the tail of the backtrack kernel (1), and backtrack +
    filter as the tail of the DP kernel (2, the automatic choice for small shards)."""
    pricer.set_option("fuse_filter", fuse)
    b = I.random_batch(700, 0, 60, 0, 200, seed=77)
    _check_against_oracle(oracle, pricer, b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha)
    # twice: the ticket counter must have been reset
    _check_against_oracle(oracle, pricer, b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha)


def test_tie_heavy_integer_profits(oracle, pricer):
    b = I.random_batch(200, 1, 60, 1, 60, seed=11, int_profits=True)
    _check_against_oracle(oracle, pricer, b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha)


def test_edge_blocks(oracle, pricer):
    """Empty block, all-inactive block, capacity 0, zero-weight items, item heavier than capacity."""
    bcs = np.array([0, 0, 3, 6, 9, 12], np.int32)
    w = np.array([1, 2, 3, 1, 1, 1, 0, 0, 5, 9, 9, 2], np.int32)
    cap = np.array([5, 4, 0, 3, 4], np.int32)
    rc = np.array([1.0, 0.0, 2.0, -1.0, -2.0, -3.0, -1.5, -2.5, -9.0, -7.0, -8.0, -1.0])
    oc = np.arange(12, dtype=np.float64)
    alpha = np.array([0.0, -1.0, 0.0, 0.5, -0.25])
    res, ref = _check_against_oracle(oracle, pricer, bcs, w, cap, rc, oc, alpha)
    assert res.blockNnz[0] == 0 and res.blockNnz[1] == 0
    assert res.blockRedCost[1] == 1.0  # empty column: 0 - alpha


@pytest.mark.parametrize("T,S,items,noguard", [(32, 7, 1, 0), (64, 9, 2, 0), (256, 5, 2, 1), (1024, 3, 1, 1),
                                               (512, 13, 1, 0), (96, 1, 2, 0), (160, 25, 2, 0),
                                               (192, 1, 3, 0), (64, 3, 3, 0), (1024, 1, 3, 0)])
def test_forced_dp_geometries(oracle, pricer, T, S, items, noguard):
    """Same results whatever the shape of the DP row: threads x cells/thread, one or two items per
    barrier, guard cells or predicated shifted reads."""
    pricer.set_option("dp_threads", T)
    pricer.set_option("dp_cells_per_thread", S)
    pricer.set_option("dp_items_per_step", items)
    pricer.set_option("dp_no_guard", noguard)
    b = I.random_batch(40, 0, 200, 0, T * S - 1, seed=100 + T + S)
    _check_against_oracle(oracle, pricer, b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, b.alpha)


def _check_round(oracle, pricer, model, u, phase1=False):
    from dip_b200.pricer import PHASE_PRICE1, PHASE_PRICE2
    res, rcx = pricer.priceRound(u, PHASE_PRICE1 if phase1 else PHASE_PRICE2, want_redcost=True)
    ref = oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                             model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                             phase1=phase1)
    np.testing.assert_array_equal(rcx, ref.red_cost_x)   # same summation order as the oracle's scatter
    assert res.nCols == ref.n_kept and res.mostNegReducedCost == ref.most_neg_reduced_cost
    # stages (b)+(c) on identical inputs (the GPU's reduced costs) must be bit-exact
    alpha = u[model.n_base_rows:model.n_base_rows + model.m]
    ref2 = oracle.solve_blocks(model.block_col_start, model.weight, model.capacity, rcx, model.objective, alpha)
    np.testing.assert_array_equal(res.blockObj, ref2.z)
    np.testing.assert_array_equal(res.blockNnz, ref2.col_nnz)
    kept = [b for b in range(model.m) if ref2.col_keep[b]]
    assert list(res.colBlock[:res.nCols]) == kept
    for k, b in enumerate(kept):
        assert list(res.column(k)) == list(ref2.col_ind[b])
    np.testing.assert_allclose(res.blockRedCost, ref2.col_red_cost, rtol=RC_RTOL, atol=1e-12)
    return res, ref


def test_gap_d_round(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model)
    rng = np.random.default_rng(5)
    for r in range(3):
        _check_round(oracle, pricer, model, I.gap_duals(model, rng))
    _check_round(oracle, pricer, model, I.gap_duals(model, rng), phase1=True)


def test_gap_e_round(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_e())
    pricer.setModel(model)
    rng = np.random.default_rng(6)
    for r in range(2):
        res, _ = _check_round(oracle, pricer, model, I.gap_duals(model, rng))
        assert res.nCols > 0


def test_gap_e_round_with_programmatic_dependent_launch(oracle, pricer):
    """Option dp_pdl: the fused kernel is launched behind the SpMV with programmatic stream serialization
    and waits (griddepcontrol.wait) before it reads the reduced costs; same round."""
    model = I.gap_pricing_model(I.gap_class_e())
    pricer.setModel(model)
    pricer.set_option("dp_pdl", 1)
    rng = np.random.default_rng(6)
    for _ in range(3):
        _check_round(oracle, pricer, model, I.gap_duals(model, rng))


@pytest.mark.parametrize("flag_wait", [1, 0])
def test_round_completion_word_and_stream_synchronise_agree(oracle, pricer, flag_wait):
    """The host entry point learns that a mirrored round is complete from a word the kernel's last CTA stores into
    mapped host memory (option flag_wait, default on) or from cudaStreamSynchronize: the same rounds, many in a row
    (the word carries a serial: a stale word must never end a wait), with and without the reduced costs handed back
    (that call copies device memory and always synchronises)."""
    pricer.set_option("flag_wait", flag_wait)
    model = I.gap_pricing_model(I.gap_class_d())
    rng = np.random.default_rng(77)
    branch_rows = I.gap_branch_rows(model, rng)
    pricer.setModel(model)
    pricer.setDualSupport(I.gap_dual_support(model, branch_rows))
    duals = [I.gap_duals(model, rng, branch_rows=branch_rows) for _ in range(6)]
    refs = [oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                               model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity) for u in duals]
    for it in range(120):
        k = it % 6
        if it % 17 == 5:
            res, rcx = pricer.priceRound(duals[k], want_redcost=True)
            np.testing.assert_array_equal(rcx, refs[k].red_cost_x)
        else:
            res = pricer.priceRound(duals[k])
        np.testing.assert_array_equal(res.blockObj, refs[k].z)
        np.testing.assert_array_equal(res.blockRedCost, refs[k].col_red_cost)
        assert res.nCols == int(np.sum(refs[k].col_keep)) and res.mostNegReducedCost == refs[k].most_neg_reduced_cost


def test_gap12_class_c_round(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_c())
    pricer.setModel(model)
    _check_round(oracle, pricer, model, I.gap_duals(model, np.random.default_rng(12)))


def test_spmv_random_csr_with_cuts(oracle, pricer):
    """General sparse A'' (not GAP-shaped), zero duals skipped, then cut rows appended."""
    rng = np.random.default_rng(9)
    R, N, m, nBase = 300, 5000, 10, 260
    nnz = 40_000
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
    for phase1 in (False, True):
        got = pricer.calcRedCost(u, 1 if phase1 else 2)
        ref = oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(u, nBase, m), c, phase1)
        np.testing.assert_array_equal(got, ref)
        # the lane-cooperative fallback kernel (used when a column exceeds one tile) reorders the
        # sum: north_star tolerance 1e-12 relative to the magnitude of the terms
        pricer.set_option("spmv_lanes", 8)
        got2 = pricer.calcRedCost(u, 1 if phase1 else 2)
        pricer.set_option("spmv_lanes", 0)
        assert np.max(np.abs(got2 - ref)) <= 1e-12 * np.max(np.abs(ref)) * 50
    # append 25 cut rows
    k = 25
    nn = 900
    r2 = np.sort(rng.integers(0, k, size=nn))
    rs2 = np.zeros(k + 1, np.int64)
    np.cumsum(np.bincount(r2, minlength=k), out=rs2[1:])
    ci2 = rng.integers(0, N, size=nn).astype(np.int32)
    v2 = rng.uniform(-5, 5, size=nn)
    pricer.appendCoreRows(rs2, ci2, v2)
    u2 = rng.standard_normal(R + k + m)
    got = pricer.calcRedCost(u2)
    rsA = np.concatenate([rs, rs[-1] + rs2[1:]])
    ref = oracle.calc_redcost(R + k, N, rsA, np.concatenate([ci, ci2]), np.concatenate([v, v2]),
                              oracle.adjust_duals(u2, nBase, m), c)
    np.testing.assert_array_equal(got, ref)


def test_spmv_long_column_fallback(oracle, pricer):
    """A column with more stored entries than one warp tile (384) switches the whole sweep to the
    lane-cooperative kernel; parity is then the 1e-12 tolerance."""
    rng = np.random.default_rng(10)
    R, N, m = 5000, 64, 2
    dense_col = 7
    rows, cols = [], []
    for i in range(R):
        cs = set(rng.integers(0, N, size=3).tolist())
        if i < 4000:
            cs.add(dense_col)
        for cc in sorted(cs):
            rows.append(i)
            cols.append(cc)
    rs = np.zeros(R + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=R), out=rs[1:])
    ci = np.array(cols, np.int32)
    v = rng.uniform(-3, 3, size=len(ci))
    c = rng.uniform(0, 10, size=N)
    pricer.setModelCore(R, N, rs, ci, v, R, m)
    pricer.setObjective(c)
    u = rng.standard_normal(R + m)
    got = pricer.calcRedCost(u)
    ref = oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(u, R, m), c)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.sum(np.abs(v[ci == dense_col])) * np.max(np.abs(u))


def test_milpblock_spmv_full_size(oracle, pricer):
    """BASELINE config 5 at full size (256 blocks, 10^7 nnz): bit-equal to the oracle's scatter,
    plus linearity redCost(u1+u2) - c == (redCost(u1) - c) + (redCost(u2) - c) on exact data."""
    model, u = I.milpblock_model()
    pricer.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
    pricer.setObjective(model.objective)
    got = pricer.calcRedCost(u)
    ref = oracle.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val,
                              oracle.adjust_duals(u, model.n_base_rows, model.m), model.objective)
    np.testing.assert_array_equal(got, ref)
    got1 = pricer.calcRedCost(u, 1)   # phase 1: -A''^T u
    ref1 = oracle.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val,
                               oracle.adjust_duals(u, model.n_base_rows, model.m), model.objective, True)
    np.testing.assert_array_equal(got1, ref1)
    assert np.max(np.abs((model.objective - got) + got1)) <= 1e-9


def test_solve_relaxed_plugin_surface(oracle, pricer):
    """DecompApp::solveRelaxed semantics: one (possibly empty) column per block, redCost without alpha."""
    from dip_b200.pricer import DecompSolStatOptimal
    b = I.random_batch(12, 0, 50, 5, 90, seed=21)
    pricer.setObjective(b.orig_cost)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    ref = oracle.solve_blocks(b.block_col_start, b.weight, b.capacity, b.red_cost, b.orig_cost, np.zeros(12))
    for blk in range(12):
        st, vs = pricer.solveRelaxed(blk, b.red_cost, target=b.alpha[blk])
        assert st == DecompSolStatOptimal and len(vs) == 1
        v = vs[0]
        assert v.blockId == blk and list(v.ind) == list(ref.col_ind[blk])
        assert abs(v.redCost - ref.col_red_cost[blk]) <= 1e-12 * max(1, abs(ref.col_red_cost[blk]))
        assert abs(v.origCost - ref.col_orig_cost[blk]) <= 1e-12 * max(1, abs(ref.col_orig_cost[blk]))
    assert pricer.counters()["kernel_launches"] <= 4  # one submission served all 12 calls


def test_generate_vars_surface(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_d(6, 40, seed=3))
    pricer.setModel(model)
    u = I.gap_duals(model, np.random.default_rng(1))
    new_vars, most_neg = pricer.generateVars(u)
    ref = oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                             model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity)
    assert len(new_vars) == ref.n_kept
    assert abs(most_neg - ref.most_neg_reduced_cost) <= 1e-12 * max(1, abs(ref.most_neg_reduced_cost))
    for v in new_vars:
        assert v.redCost < -1e-4
        assert list(v.ind) == list(ref.col_ind[v.blockId])


def test_error_codes(pricer):
    from dip_b200.capi import DipGpuError, ERR_STATE, ERR_INVALID
    with pytest.raises(DipGpuError) as e:
        pricer.priceRound(np.zeros(4))
    assert e.value.code == ERR_STATE
    with pytest.raises(DipGpuError) as e:
        pricer.setKnapsackBlocks([0, 2], [1, -1], [3])
    assert e.value.code == ERR_INVALID
    # any capacity is priced (rows beyond one CTA live in global memory): no UNSUPPORTED below 10^6 and beyond
    pricer.setObjective(np.array([3.0, 4.0]))
    pricer.setKnapsackBlocks([0, 2], [700_000, 400_000], [1_000_000])
    r = pricer.solveBlocks(np.array([-2.5, -1.25]), np.zeros(1), redCostEps=-np.inf)
    assert r.blockObj[0] == 2.5 and list(r.column(0)) == [0]


@pytest.mark.parametrize("no_cluster", [0, 1])
@pytest.mark.parametrize("profit_mode", [0, 1])
@pytest.mark.parametrize("cap,nhi,m,seed", [(13_000, 300, 5, 1), (100_000, 500, 3, 2), (40_000, 90, 12, 3), (29_000, 120, 40, 4),
                                            (200_000, 60, 2, 5), (230_000, 40, 2, 6)])
def test_capacity_beyond_one_cta(oracle, pricer, profit_mode, cap, nhi, m, seed, no_cluster):
    """Rows of more than 12 800 cells (GAP_KnapPisinger.h:243-249 prices any capacity): the row distributed over the
    shared memories of a thread-block cluster (2, 4, 8 or 16 CTAs; up to 212 992 cells) or, beyond that and with
    option dp_no_cluster, in global memory; bit-exact optimum, identical columns; with node bounds as well."""
    pricer.set_option("dp_no_cluster", no_cluster)
    rng = np.random.default_rng(seed)
    n = rng.integers(nhi // 2, nhi + 1, size=m)
    bcs = np.zeros(m + 1, np.int32)
    np.cumsum(n, out=bcs[1:])
    N = int(bcs[-1])
    w = rng.integers(1, cap // 8, size=N).astype(np.int32)
    rc = np.round(rng.uniform(-60, 20, size=N), 2 if profit_mode else 9)
    oc = rng.uniform(1, 100, size=N)
    capv = rng.integers(cap // 2, cap + 1, size=m).astype(np.int32)
    capv[0] = cap
    alpha = rng.uniform(-5, 0, size=m)
    res, ref = _check_against_oracle(oracle, pricer, bcs, w, capv, rc, oc, alpha, profit_mode=profit_mode)
    geo = pricer.dpGeometry()
    assert (geo["cells_per_thread"] == 0) == (no_cluster == 1 or cap + 1 > 16 * 13 * 1024)
    if profit_mode == 0:
        lb = (rng.random(N) < 0.01).astype(np.float64)
        ub = 1.0 - (rng.random(N) < 0.05) * (1 - lb)
        pricer.setColBounds(lb, ub)
        r2 = pricer.solveBlocks(rc, alpha)
        ref2 = oracle.solve_blocks_bounded(bcs, w, capv, rc, oc, alpha, lb, ub)
        np.testing.assert_array_equal(r2.blockObj, ref2.z)
        np.testing.assert_array_equal(r2.blockRedCost, ref2.col_red_cost)
        assert [list(r2.column(k)) for k in range(r2.nCols)] == [list(ref2.col_ind[b]) for b in range(m) if ref2.col_keep[b]]


@pytest.mark.parametrize("cap,m,n,seed", [(30_000, 6, 40, 11), (100_000, 3, 36, 12), (150_000, 2, 30, 13)])
def test_cluster_rows_with_weights_spanning_several_ctas(oracle, pricer, cap, m, n, seed):
    """The cluster kernel reads c[i-1][j-w] from the CTA that owns cell j-w: with weights up to the capacity that is
    any lower-ranked CTA of the cluster (not just the neighbour), cells below w have no candidate, and weights that
    are exact multiples of a CTA's slice land on slice borders.  Bit-exact optimum and identical columns."""
    rng = np.random.default_rng(seed)
    bcs = (np.arange(m + 1) * n).astype(np.int32)
    N = m * n
    w = rng.integers(1, cap + 1, size=N).astype(np.int32)          # any weight up to the capacity
    small = rng.random(N) < 0.5
    w[small] = rng.integers(1, cap // 20, size=int(small.sum()))    # ... mixed with small ones, so that columns have several items
    for slice_cells in (13 * 1024, 7 * 1024, 2 * 13 * 1024):        # multiples of possible slice lengths
        w[rng.integers(0, N, size=3)] = min(slice_cells, cap)
    rc = -rng.uniform(0.5, 50.0, size=N)
    oc = rng.uniform(1, 100, size=N)
    capv = np.full(m, cap, np.int32)
    capv[1:] = rng.integers(cap // 2, cap + 1, size=m - 1)
    res, ref = _check_against_oracle(oracle, pricer, bcs, w, capv, rc, oc, np.zeros(m))
    assert pricer.dpGeometry()["cells_per_thread"] > 0     # the cluster kernel ran
    pricer.set_option("dp_no_cluster", 1)
    _check_against_oracle(oracle, pricer, bcs, w, capv, rc, oc, np.zeros(m))


def test_csp_full_size_properties(oracle, pricer):
    """BASELINE config 4 at full size (10^4 x 500 x 10^4): size-independent properties on every
    block + bit-exact oracle parity on a sample of blocks."""
    b = I.csp_batch()
    pricer.setObjective(b.orig_cost)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    res = pricer.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
    assert res.nCols == b.m and res.cells == b.cells
    w64 = b.weight.astype(np.int64)
    for k in range(b.m):
        ind = res.column(k)
        assert ind.size == 0 or (np.all(np.diff(ind) > 0) and ind[0] >= b.block_col_start[k]
                                 and ind[-1] < b.block_col_start[k + 1])
        assert w64[ind].sum() <= b.capacity[k]                      # feasible
        z = -b.red_cost[ind].sum()
        assert abs(z - res.blockObj[k]) <= 1e-9 * max(1.0, z)        # column value == DP optimum
    rng = np.random.default_rng(0)
    for k in rng.choice(b.m, size=24, replace=False):
        s, e = b.block_col_start[k], b.block_col_start[k + 1]
        z, sol = oracle.knapsack01(-b.red_cost[s:e], b.weight[s:e], int(b.capacity[k]))
        assert z == res.blockObj[k]
        assert sorted(int(s + i) for i in sol) == list(res.column(k))


# ------------------------------------------------------------------ Pisinger (scaled-integer) mode


def _decimal_batch(m, nhi, caphi, seed, digits):
    """Reduced costs with a bounded number of decimals, so calcScaleFactor lands on 1..10^4."""
    rng = np.random.default_rng(seed)
    b = I.random_batch(m, 0, nhi, 1, caphi, seed)
    n = np.diff(b.block_col_start)
    dig = np.repeat(rng.integers(0, digits + 1, size=m), n)   # per block: 0..digits decimals
    q = 10.0 ** dig
    b.red_cost = np.where(b.red_cost < 0, -np.maximum(np.round(-b.red_cost * q), 0) / q, b.red_cost)
    return b


@pytest.mark.parametrize("seed,m,nhi,caphi,digits", [(31, 200, 40, 120, 5), (32, 64, 300, 900, 3),
                                                     (33, 500, 6, 40, 4), (34, 20, 1500, 3000, 2)])
def test_pisinger_mode_batches(oracle, pricer, seed, m, nhi, caphi, digits):
    """GAP_KnapPisinger::solve semantics: scale factor, integer profits, trivial cases; optimum and
    column against the oracle's restatement (the DP stands in for combo.c on both sides)."""
    from dip_b200.pricer import PROFIT_PISINGER_INT
    b = _decimal_batch(m, nhi, caphi, seed, digits)
    res, ref = _check_against_oracle(oracle, pricer, b.block_col_start, b.weight, b.capacity, b.red_cost,
                                     b.orig_cost, b.alpha, profit_mode=PROFIT_PISINGER_INT)
    # integer profits: the optimum is an integer
    assert np.all(res.blockObj == np.round(res.blockObj))


def test_pisinger_scale_factor_and_shortcuts(oracle, pricer):
    from dip_b200.pricer import PROFIT_PISINGER_INT
    # block 0: one item that fits; 1: one item that does not fit; 2: all fit (incl. a profit that
    # rounds to 0); 3: two items, tie on profit (reference prefers item 1); 4: two items, both fit
    # is impossible, unequal profits; 5: scale 10 (0.5 steps); 6: scale 10^4 (too many decimals);
    # 7: no active item; 8: three items -> DP
    bcs = np.array([0, 2, 4, 8, 10, 12, 15, 18, 20, 23], np.int32)
    w = np.array([3, 1, 9, 1, 1, 2, 1, 1, 4, 4, 4, 5, 2, 2, 2, 3, 3, 3, 1, 1, 3, 3, 3], np.int32)
    cap = np.array([5, 5, 10, 5, 6, 3, 4, 5, 5], np.int32)
    rc = np.array([-2.5, 1.0, -7.0, 0.0, -1.0, -2.0, -0.00001, 3.0, -3.0, -3.0, -2.0, -4.5,
                   -1.5, -2.5, -3.5, -0.12345, -0.5, -0.25, 1.0, 2.0, -1.0, -2.0, -3.0])
    oc = np.arange(len(w), dtype=np.float64) + 1
    alpha = np.zeros(9)
    res, ref = _check_against_oracle(oracle, pricer, bcs, w, cap, rc, oc, alpha, eps=-np.inf,
                                     profit_mode=PROFIT_PISINGER_INT)
    assert list(res.column(0)) == [0] and list(res.column(1)) == []
    assert list(res.column(2)) == [4, 5, 6]
    assert list(res.column(3)) == [9]                # tie: {1} wins over {0}
    for blk in range(9):
        s, e = bcs[blk], bcs[blk + 1]
        assert oracle.calc_scale_factor(rc[s:e]) in (1, 10, 100, 1000, 10000)


def test_pisinger_matches_reference_hs_values(oracle, pricer, golden_knapsack):
    """Integer profits make the optimum unique: the GPU's Pisinger-mode value must equal what the
    reference's own Horowitz-Sahni B&B (compiled from Dip/src/UtilKnapsack.cpp) returned."""
    from dip_b200.pricer import PROFIT_PISINGER_INT
    cases = [c for c in golden_knapsack if c["hs_z"] is not None and c["hs_status"] == 0]
    assert len(cases) >= 20
    bcs, w, cap, rc = _batch_from_cases(cases)
    pricer.setObjective(np.zeros(len(w)))
    pricer.setKnapsackBlocks(bcs, w, cap, PROFIT_PISINGER_INT)
    res = pricer.solveBlocks(rc, np.zeros(len(cases)), redCostEps=-np.inf)
    for k, c in enumerate(cases):
        assert res.blockObj[k] == c["hs_z"], c["tag"]


def test_gap_d_round_pisinger(oracle, pricer):
    from dip_b200.pricer import PROFIT_PISINGER_INT
    model = I.gap_pricing_model(I.gap_class_d())
    pricer.setModel(model, PROFIT_PISINGER_INT)
    u = I.gap_duals(model, np.random.default_rng(8))
    res, rcx = pricer.priceRound(u, want_redcost=True)
    ref = oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                             model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                             profit_mode=1)
    np.testing.assert_array_equal(rcx, ref.red_cost_x)
    np.testing.assert_array_equal(res.blockObj, ref.z)
    np.testing.assert_array_equal(res.blockRedCost, ref.col_red_cost)
    for k in range(res.nCols):
        assert list(res.column(k)) == list(ref.col_ind[int(res.colBlock[k])])


# ------------------------------------------------- after the round: master columns


def _check_expand(oracle, pricer, model, res, tol=1e-6):
    col_start, row_ind, val, hsh = pricer.expandColumns(tol)
    assert len(col_start) == res.nCols + 1 and len(hsh) == res.nCols
    for k in range(res.nCols):
        b = int(res.colBlock[k])
        rr, vv = oracle.expand_column(pricer.R, model.N, model._rs, model._ci, model._v, model.n_base_rows, model.m,
                                      res.column(k), b, tol)
        got_r = row_ind[col_start[k]:col_start[k + 1]]
        got_v = val[col_start[k]:col_start[k + 1]]
        np.testing.assert_array_equal(got_r, rr)
        np.testing.assert_array_equal(got_v, vv)   # same additions in the same order: bit-equal
    return col_start, row_ind, val, hsh


def test_expand_columns_gap_d(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_d())
    model._rs, model._ci, model._v = model.row_start, model.col_ind, model.val
    pricer.setModel(model)
    u = I.gap_duals(model, np.random.default_rng(2))
    res = pricer.priceRound(u)
    assert res.nCols > 0
    cs, ri, va, hs = _check_expand(oracle, pricer, model, res)
    # every master column carries its block's convexity row with 1.0
    for k in range(res.nCols):
        rows = ri[cs[k]:cs[k + 1]]
        assert model.n_base_rows + int(res.colBlock[k]) in rows
    # fingerprints: same partition as the reference's string hash (UtilHash.cpp:52-68)
    strs = [oracle.string_hash(res.column(k), np.ones(len(res.column(k)))) + f"|{int(res.colBlock[k])}"
            for k in range(res.nCols)]
    assert len(set(strs)) == len(set(int(h) for h in hs))
    # the same round again: identical fingerprints (dedup would reject every column)
    res2 = pricer.priceRound(u)
    _, _, _, hs2 = pricer.expandColumns()
    np.testing.assert_array_equal(hs, hs2)


def test_expand_columns_general_matrix_with_cuts_and_cancellation(oracle, pricer):
    """Non-GAP A'' with several entries per (row, block), negative coefficients that cancel
    (dropped by TolZero) and cut rows appended behind the convexity rows."""
    rng = np.random.default_rng(17)
    m, n = 6, 40
    N = m * n
    R0, ncut = 50, 12
    rows, cols, vals = [], [], []
    for i in range(R0 + ncut):
        k = int(rng.integers(3, 30))
        cc = rng.choice(N, size=k, replace=False)
        cc.sort()
        for c in cc:
            rows.append(i)
            cols.append(int(c))
            vals.append(float(rng.integers(-2, 3)))   # small integers: plenty of exact cancellation
    rows = np.array(rows)
    rs = np.zeros(R0 + ncut + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=R0 + ncut), out=rs[1:])
    ci = np.array(cols, np.int32)
    v = np.array(vals)
    b = I.random_batch(m, n, n, 30, 200, seed=5)
    pricer.setModelCore(R0, N, rs[:R0 + 1], ci[:rs[R0]], v[:rs[R0]], R0, m)
    pricer.appendCoreRows(rs[R0:] - rs[R0], ci[rs[R0]:], v[rs[R0]:])   # the cut rows
    pricer.setObjective(b.orig_cost)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    res = pricer.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)

    class M:  # what _check_expand needs
        pass
    model = M()
    model.N, model.n_base_rows, model.m = N, R0, m
    model._rs, model._ci, model._v = rs, ci, v
    cs, ri, va, hs = _check_expand(oracle, pricer, model, res)
    assert (ri >= R0 + m).any()          # cut rows sit behind the convexity rows
    assert np.all(np.abs(va) > 1e-6)


def test_expand_columns_gap_e(oracle, pricer):
    model = I.gap_pricing_model(I.gap_class_e())
    model._rs, model._ci, model._v = model.row_start, model.col_ind, model.val
    pricer.setModel(model)
    res = pricer.priceRound(I.gap_duals(model, np.random.default_rng(4)))
    cs, ri, va, hs = pricer.expandColumns()
    assert len(cs) == res.nCols + 1 and cs[-1] == len(ri)
    for k in list(range(0, res.nCols, 7)):
        rr, vv = oracle.expand_column(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows,
                                      model.m, res.column(k), int(res.colBlock[k]))
        np.testing.assert_array_equal(ri[cs[k]:cs[k + 1]], rr)
        np.testing.assert_array_equal(va[cs[k]:cs[k + 1]], vv)


# ------------------------------------------------------------------ the fused kernel's reduction pre-pass


def _reduction_batch(seed, m, n, cap_lo, cap_hi, kind):
    """Blocks big enough (>= 1024 columns) for the fused kernel's reduction to engage."""
    rng = np.random.default_rng(seed)
    bcs = (np.arange(m + 1) * n).astype(np.int32)
    N = m * n
    cap = rng.integers(cap_lo, cap_hi + 1, size=m).astype(np.int32)
    if kind == "gap":          # small weights, spread-out profits: a handful of items matter
        w = np.floor(1 - 10 * np.log(rng.uniform(1e-9, 1, size=N))).astype(np.int32)
        rc = -(rng.uniform(0, 1000, size=N) / w)
    elif kind == "ties":       # small integer profits and weights: many optimal solutions
        w = rng.integers(1, 12, size=N).astype(np.int32)
        rc = -rng.integers(1, 7, size=N).astype(np.float64)
    elif kind == "equal_eff":  # every item has the same efficiency: nothing can be dropped, one fat bucket
        w = rng.integers(1, 40, size=N).astype(np.int32)
        rc = -3.0 * w
    elif kind == "subset_sum":
        w = rng.integers(1, 200, size=N).astype(np.int32)
        rc = -w.astype(np.float64) - rng.uniform(0, 1e-9, size=N)
    else:                      # "wide": efficiencies over 30 octaves, zero weights, items that do not fit
        w = rng.integers(0, 60, size=N).astype(np.int32)
        rc = -np.exp(rng.uniform(-12, 12, size=N))
        w[rng.random(N) < 0.02] = cap.max() + 7
    rc = np.where(rng.random(N) < 0.25, rng.uniform(0, 5, size=N), rc)   # inactive columns
    oc = rng.integers(1, 100, size=N).astype(np.float64)
    alpha = rng.uniform(-10, 10, size=m)
    return bcs, w, cap, rc, oc, alpha


@pytest.mark.parametrize("kind,n,cap_lo,cap_hi", [("gap", 1600, 100, 250), ("gap", 2048, 20, 60), ("ties", 1200, 30, 90),
                                                  ("equal_eff", 1100, 50, 400), ("subset_sum", 1024, 200, 600),
                                                  ("wide", 1500, 10, 120), ("gap", 1030, 1, 5)])
def test_reduction_keeps_optimum_column_and_ties(oracle, pricer, kind, n, cap_lo, cap_hi):
    """Dropping the items no optimal solution can contain must not change anything the reference's DP
    returns -- optimum bit for bit, the column (also among tied optima), sums, cells -- and switching the
    pre-pass off gives the same round."""
    seed = {"gap": 1, "ties": 2, "equal_eff": 3, "subset_sum": 4, "wide": 5}[kind] * 10000 + n
    bcs, w, cap, rc, oc, alpha = _reduction_batch(seed, 40, n, cap_lo, cap_hi, kind)
    res, ref = _check_against_oracle(oracle, pricer, bcs, w, cap, rc, oc, alpha)
    assert pricer.dpGeometry()["fused_tail"] == 1
    ev_on = res.cellsEvaluated
    assert ev_on <= res.cells
    if kind == "gap":
        assert ev_on < res.cells               # the pre-pass engaged
    cols_on = [list(res.column(k)) for k in range(res.nCols)]
    obj_on = res.blockObj.copy()
    pricer.set_option("dp_no_prune", 1)
    res2, _ = _check_against_oracle(oracle, pricer, bcs, w, cap, rc, oc, alpha)
    assert res2.cellsEvaluated == res2.cells
    assert [list(res2.column(k)) for k in range(res2.nCols)] == cols_on
    np.testing.assert_array_equal(res2.blockObj, obj_on)


def test_reduction_in_pisinger_mode(oracle, pricer):
    from dip_b200.pricer import PROFIT_PISINGER_INT
    bcs, w, cap, rc, oc, alpha = _reduction_batch(77, 24, 1400, 60, 200, "gap")
    rc = np.round(rc, 2)
    _check_against_oracle(oracle, pricer, bcs, w, cap, rc, oc, alpha, profit_mode=PROFIT_PISINGER_INT)
    res = pricer.solveBlocks(rc, alpha)
    assert res.cellsEvaluated < res.cells


def test_gap_e_many_rounds_match_oracle(oracle, pricer):
    """48 GAP-E 80x1600 rounds (3 840 blocks through the reduction pre-pass), singly and in batches of 16:
    optimum bit for bit, columns, per-block reduced costs and the reference's cell count, every round."""
    model = I.gap_pricing_model(I.gap_class_e())
    pricer.setModel(model)
    rng = np.random.default_rng(2024)
    U = np.stack([I.gap_duals(model, rng, branch_density=0.002 * (1 + v % 7)) for v in range(48)])
    U[5, :model.meta["n"]] *= 0.05          # almost no item prices out
    U[6, :model.meta["n"]] *= 3.0           # nearly every item does
    U[7, :model.meta["n"]] = np.round(U[7, :model.meta["n"]])   # integer duals: ties
    refs = [oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, U[v],
                               model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                               n_threads=8) for v in range(48)]
    evaluated = 0
    for lo in range(0, 48, 16):
        batch = pricer.priceRoundsBatch(U[lo:lo + 16])
        for k, res in enumerate(batch):
            ref = refs[lo + k]
            np.testing.assert_array_equal(res.blockObj, ref.z)
            np.testing.assert_array_equal(res.blockNnz, ref.col_nnz)
            np.testing.assert_array_equal(res.blockRedCost, ref.col_red_cost)
            kept = [b for b in range(model.m) if ref.col_keep[b]]
            assert list(res.colBlock[:res.nCols]) == kept
            for j, b in enumerate(kept):
                assert list(res.column(j)) == list(ref.col_ind[b])
            assert res.cells == ref.cells
            evaluated += res.cellsEvaluated
    for v in (0, 5, 6, 7, 47):
        res = pricer.priceRound(U[v])
        np.testing.assert_array_equal(res.blockObj, refs[v].z)
        assert res.mostNegReducedCost == refs[v].most_neg_reduced_cost
    assert evaluated < 0.5 * sum(r.cells for r in refs)


# ------------------------------------------------- cut rows appended in place (DecompConstraintSet::appendRow)


def _random_rows(rng, k, N, nn, lo=-5, hi=5):
    r = np.sort(rng.integers(0, k, size=nn))
    rs = np.zeros(k + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=k), out=rs[1:])
    return rs, rng.integers(0, N, size=nn).astype(np.int32), rng.uniform(lo, hi, size=nn)


@pytest.mark.parametrize("spmv_u_smem", [0, 1])
def test_appended_rows_stay_a_delta_and_match_a_full_rebuild(oracle, pricer, spmv_u_smem):
    """dipgpu_append_core_rows on a model above "append_delta_min_nnz": nothing is re-transposed or re-tiled, the new
    rows are a small CSC whose sums START every column's addition chain (they are the first terms of the reference's
    scatter order).  Reduced costs stay bit-equal to the oracle on the whole matrix through three appends (single
    vector, both phases, a batch), the expansion walks the new rows, and the explicit merge changes nothing."""
    rng = np.random.default_rng(90 + spmv_u_smem)
    m, n = 8, 300
    N = m * n
    R, nBase = 400, 330
    rs, ci, v = _random_rows(rng, R, N, 30_000, -10, 10)
    c = rng.uniform(0, 100, size=N)
    pricer.set_option("append_delta_min_nnz", 0)
    pricer.set_option("spmv_u_smem", spmv_u_smem)
    pricer.setModelCore(R, N, rs, ci, v, nBase, m)
    pricer.setObjective(c)
    b = I.random_batch(m, n, n, 30, 400, seed=6)
    pricer.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    u = rng.standard_normal(R + m)
    np.testing.assert_array_equal(pricer.calcRedCost(u), oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(u, nBase, m), c))
    l0 = pricer.counters()["kernel_launches"]
    for step, (k, nn) in enumerate(((25, 700), (1, 40), (10, 0))):   # (an append of empty rows too)
        rs2, ci2, v2 = _random_rows(rng, k, N, nn)
        if step == 1:
            v2[:] = np.round(v2)      # integers: exact cancellations inside a column's leading sum
        pricer.appendCoreRows(rs2, ci2, v2)
        rs = np.concatenate([rs, rs[-1] + rs2[1:]])
        ci = np.concatenate([ci, ci2])
        v = np.concatenate([v, v2])
        R += k
        assert pricer.R == R
        u = rng.standard_normal(R + m)
        u[rng.random(R + m) < 0.2] = 0.0
        for phase in (2, 1):
            ref = oracle.calc_redcost(R, N, rs, ci, v, oracle.adjust_duals(u, nBase, m), c, phase == 1)
            l1 = pricer.counters()["kernel_launches"]
            np.testing.assert_array_equal(pricer.calcRedCost(u, phase), ref)
            # the delta kernel in front of the stream kernel: the appended rows were NOT folded into the tiles
            assert pricer.counters()["kernel_launches"] - l1 == (2 if len(ci) > 30_000 else 1)
        # whole rounds and a batch of them on the grown matrix
        U = np.stack([rng.standard_normal(R + m) for _ in range(3)])
        batch = pricer.priceRoundsBatch(U)
        for q in range(3):
            refq = oracle.price_round(R, N, rs, ci, v, c, U[q], nBase, m, b.block_col_start, b.weight, b.capacity)
            np.testing.assert_array_equal(batch[q].blockObj, refq.z)
            np.testing.assert_array_equal(batch[q].blockRedCost, refq.col_red_cost)
            single = pricer.priceRound(U[q])
            np.testing.assert_array_equal(single.blockObj, refq.z)
            assert single.nCols == batch[q].nCols == int(np.sum(refq.col_keep))
    # the expansion of the last round's columns touches the appended rows
    res = pricer.priceRound(U[2], redCostEps=-np.inf)   # keeps every block's column

    class M:
        pass
    model = M()
    model.N, model.n_base_rows, model.m = N, nBase, m
    model._rs, model._ci, model._v = rs, ci, v
    cs, ri, va, hs = _check_expand(oracle, pricer, model, res)
    assert (ri >= R - 36 + m).any()
    # folding the rows in gives the same bits with one launch
    before = pricer.calcRedCost(u)
    pricer.set_option("merge_appended_rows", 1)
    l1 = pricer.counters()["kernel_launches"]
    np.testing.assert_array_equal(pricer.calcRedCost(u), before)
    assert pricer.counters()["kernel_launches"] - l1 == 1
    _check_expand(oracle, pricer, model, pricer.priceRound(U[2], redCostEps=-np.inf))
    assert l0 > 0

"""Host logic: synthetic instance generators and the GAP model layout (no GPU)."""
import numpy as np

from dip_b200 import instances as I


def test_gap_model_layout():
    g = I.gap_class_d(4, 7, seed=1)
    mdl = I.gap_pricing_model(g)
    n, N = 7, 28
    assert mdl.R == n + 2 * N and mdl.n_base_rows == mdl.R and mdl.nnz == N + 2 * N
    # assignment row j holds x[i, j] for every machine i (GAP_DecompApp.cpp:69-135)
    for j in range(n):
        cols = mdl.col_ind[mdl.row_start[j]:mdl.row_start[j + 1]]
        assert list(cols) == [i * n + j for i in range(4)]
    # identity branch rows (DecompAlgo.cpp:1431-1455)
    assert list(mdl.col_ind[mdl.row_start[n]:mdl.row_start[n + N]]) == list(range(N))
    assert list(mdl.col_ind[mdl.row_start[n + N]:]) == list(range(N))
    assert mdl.u_len == mdl.R + 4
    assert mdl.spmv_bytes() == 12 * mdl.nnz + 20 * N + 8 * mdl.R


def test_gap_e_shape():
    g = I.gap_class_e()
    assert g.cost.shape == (80, 1600) and g.weight.min() >= 1
    assert (g.capacity >= g.weight.max(axis=1)).all()
    assert 100 < g.capacity.mean() < 260
    mdl = I.gap_pricing_model(g)
    assert mdl.N == 128_000 and mdl.R == 257_600 and mdl.nnz == 384_000
    u = I.gap_duals(mdl, np.random.default_rng(0))
    assert len(u) == 257_680
    rc = mdl.objective - np.tile(u[:1600], 80)
    assert 0.2 < (rc < 0).mean() < 0.8


def test_parse_gap_roundtrip():
    g = I.gap_class_c(3, 5, seed=2)
    txt = "3 5\n" + " ".join(map(str, g.cost.reshape(-1))) + "\n" + " ".join(map(str, g.weight.reshape(-1))) + \
        "\n" + " ".join(map(str, g.capacity)) + "\n"
    h = I.parse_gap(txt)
    np.testing.assert_array_equal(h.cost, g.cost)
    np.testing.assert_array_equal(h.weight, g.weight)
    np.testing.assert_array_equal(h.capacity, g.capacity)


def test_csp_batch_cells():
    b = I.csp_batch(m=10, n=50, cap=100, seed=1)
    assert b.cells == 10 * 50 * 101
    assert (b.red_cost < 0).all()


def test_random_batch_is_ragged():
    b = I.random_batch(50, 0, 40, 0, 90, seed=3)
    n = np.diff(b.block_col_start)
    assert n.min() == 0 or n.min() < 5
    assert (b.red_cost == 0).any() and (b.red_cost > 0).any()


def test_gap_duals_of_one_node_share_their_support():
    """bench.py prices 16 dual vectors of ONE node: the same branch rows carry a bound in all of them, and every
    vector is zero outside the declared support (what dipgpu_set_dual_support relies on)."""
    import numpy as np
    from dip_b200 import instances as I
    model = I.gap_pricing_model(I.gap_class_d())
    rng = np.random.default_rng(3)
    rows = I.gap_branch_rows(model, rng, branch_density=0.02)
    assert len(rows) == int(0.02 * (model.n_base_rows - model.meta["n"])) and np.all(np.diff(rows) > 0)
    support = I.gap_dual_support(model, rows)
    assert np.all(np.diff(support) > 0) and support[-1] == model.u_len - 1
    outside = np.ones(model.u_len, bool)
    outside[support] = False
    for _ in range(4):
        u = I.gap_duals(model, rng, branch_rows=rows)
        assert not np.any(u[outside]) and np.count_nonzero(u[support]) > 0.9 * len(support)

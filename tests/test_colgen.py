"""End-to-end: Dantzig-Wolfe column generation on the reference's in-tree instance gap0515-2
(optimum 269, Dip/data/GAP/gap.opt:9) with the pricing round on the GPU, against the same loop
driven by the CPU oracle.  Same duals in -> same columns out, round after round, and the final DW
bound sits between the LP relaxation and the known integer optimum."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dip_b200 import instances as I  # noqa: E402
from examples import gap_colgen as cg  # noqa: E402


def _oracle_pricer(model, oracle):
    def price(u):
        r = oracle.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                               model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity)
        out = []
        for b in range(model.m):
            if r.col_keep[b]:
                rows, vals = oracle.expand_column(model.R, model.N, model.row_start, model.col_ind, model.val,
                                                  model.n_base_rows, model.m, r.col_ind[b], b)
                out.append((b, r.col_ind[b], float(r.col_red_cost[b]), float(r.col_orig_cost[b]), rows, vals))
        return out, r.most_neg_reduced_cost
    return price


def _gap0515(golden_gap0515):
    g = golden_gap0515
    inst = I.GapInstance(g["m"], g["n"], np.array(g["cost"]).reshape(g["m"], g["n"]),
                         np.array(g["weight"], np.int32).reshape(g["m"], g["n"]),
                         np.array(g["capacity"], np.int32), "gap0515-2")
    return I.gap_pricing_model(inst, branch_rows=False)


def test_colgen_cpu_driver_reaches_a_valid_bound(oracle, golden_gap0515):
    """The driver itself (host logic), with the oracle as pricing backend."""
    model = _gap0515(golden_gap0515)
    tr, cols = cg.run(model, _oracle_pricer(model, oracle))
    assert tr.n_new[-1] == 0 or tr.lower_bound[-1] >= tr.z_rmp[-1] - 1e-4
    z = tr.z_rmp[-1]
    # the DW bound of this instance is tight: it equals the reference's known optimum 269
    # (Dip/data/GAP/gap.opt:9), and the last round proves it (LB == z_RMP)
    assert abs(z - golden_gap0515["known_optimum_min"]) <= 1e-6
    assert abs(tr.lower_bound[-1] - z) <= 1e-6
    assert all(tr.lower_bound[k] <= z + 1e-6 for k in range(len(tr.z_rmp)))   # valid lower bounds
    assert np.all(np.diff(tr.z_rmp) <= 1e-7)                          # the RMP value never increases


@pytest.mark.gpu
def test_colgen_gpu_matches_oracle_round_by_round(oracle, golden_gap0515):
    model = _gap0515(golden_gap0515)
    tr_cpu, cols_cpu = cg.run(model, _oracle_pricer(model, oracle))
    price, p = cg.gpu_pricer(model)
    tr_gpu, cols_gpu = cg.run(model, price)
    p.close()
    assert len(tr_gpu.z_rmp) == len(tr_cpu.z_rmp)
    for k in range(len(tr_cpu.z_rmp)):
        np.testing.assert_array_equal(tr_gpu.duals[k], tr_cpu.duals[k])
        assert tr_gpu.columns[k] == tr_cpu.columns[k], f"round {k}"
        assert tr_gpu.z_rmp[k] == tr_cpu.z_rmp[k] and tr_gpu.lower_bound[k] == tr_cpu.lower_bound[k]
    assert [(c.block, c.ind) for c in cols_gpu] == [(c.block, c.ind) for c in cols_cpu]


@pytest.mark.gpu
def test_colgen_gap_d_gpu(oracle):
    model = I.gap_pricing_model(I.gap_class_d(8, 60, seed=7), branch_rows=False)
    price, p = cg.gpu_pricer(model)
    tr, cols = cg.run(model, price)
    p.close()
    tr_cpu, _ = cg.run(model, _oracle_pricer(model, oracle))
    assert tr.z_rmp == tr_cpu.z_rmp and tr.columns == tr_cpu.columns


def _oracle_stab_pricer(model, oracle):
    """The stabilised backend restated with the oracle: the centre is a host vector."""
    plain = _oracle_pricer(model, oracle)
    state = {"centre": None, "st": []}

    def price_ladder(u_rm, alpha0, n_steps):
        if state["centre"] is None:
            state["centre"] = u_rm.copy()          # first call: dual := dualRM (DecompAlgoPC.cpp:163-173)
        out, a, state["st"] = [], alpha0, []
        for _ in range(n_steps):
            u_st = oracle.smooth_duals(a, state["centre"], u_rm)
            cols, mn = plain(u_st)
            out.append((cols, mn, u_st))
            state["st"].append(u_st)
            a = a * 0.90
        return out

    def accept(step):
        state["centre"] = state["st"][step].copy()
    return price_ladder, plain, accept


def test_stabilised_colgen_cpu_driver_reaches_the_known_optimum(oracle, golden_gap0515):
    model = _gap0515(golden_gap0515)
    ladder, plain, accept = _oracle_stab_pricer(model, oracle)
    tr, cols = cg.run_stabilised(model, ladder, plain, accept)
    assert abs(tr.z_rmp[-1] - golden_gap0515["known_optimum_min"]) <= 1e-6
    assert tr.lower_bound[-1] <= tr.z_rmp[-1] + 1e-6 and tr.z_rmp[-1] - tr.lower_bound[-1] <= 1e-4
    assert np.all(np.diff(tr.lower_bound) >= -1e-12)      # the best bound only moves up


@pytest.mark.gpu
def test_stabilised_colgen_gpu_matches_oracle_round_by_round(oracle, golden_gap0515):
    """Wentges smoothing with the centre on the device, the alpha ladder priced in one submission, the chosen
    step expanded on the GPU: same smoothed duals, same columns, same bounds as the oracle-driven loop."""
    model = _gap0515(golden_gap0515)
    ladder, plain, accept = _oracle_stab_pricer(model, oracle)
    tr_cpu, cols_cpu = cg.run_stabilised(model, ladder, plain, accept)
    gl, gp, ga, p = cg.gpu_stab_pricer(model)
    tr_gpu, cols_gpu = cg.run_stabilised(model, gl, gp, ga)
    p.close()
    assert len(tr_gpu.z_rmp) == len(tr_cpu.z_rmp)
    for k in range(len(tr_cpu.z_rmp)):
        np.testing.assert_array_equal(tr_gpu.duals[k], tr_cpu.duals[k])
        assert tr_gpu.columns[k] == tr_cpu.columns[k], f"round {k}"
        assert tr_gpu.z_rmp[k] == tr_cpu.z_rmp[k] and tr_gpu.lower_bound[k] == tr_cpu.lower_bound[k]
    assert abs(tr_gpu.z_rmp[-1] - golden_gap0515["known_optimum_min"]) <= 1e-6

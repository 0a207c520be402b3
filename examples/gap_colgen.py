#!/usr/bin/env python3
"""Mini price-and-cut driver: Dantzig-Wolfe column generation for a GAP instance with the pricing
round on the GPU (dip_b200) and a stand-in restricted master LP (HiGHS via scipy -- COIN-OR's Clp is
not installable here).  It mimics the CONTRACT of DIP's loop
(Dip/src/DecompAlgo.cpp:1782-2293): solve the restricted master, hand the duals in DIP's row layout
[core rows | convexity rows] to generateVars, turn the returned variables into master columns
(addVarsToPool), repeat until no column prices out; the lower bound after each round is
z_RMP + mostNegReducedCost (updateObjBound, :3757).

    python examples/gap_colgen.py [path/to/gapXXXX.dat]        (default: synthetic class D 8x60)

The pricing backend is injected (`price`), so the tests can drive the same loop with the CPU oracle
and compare iterates.  Big-M artificials replace DIP's two-phase start.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize import linprog
from scipy.sparse import csc_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dip_b200 import instances as I  # noqa: E402

BIG_M = 1.0e4


@dataclass
class Column:
    block: int
    ind: tuple          # x-space indices
    cost: float         # c . s
    rows: np.ndarray    # master rows of [A''s ; e_b]
    vals: np.ndarray


@dataclass
class Trace:
    z_rmp: list = field(default_factory=list)
    lower_bound: list = field(default_factory=list)
    n_new: list = field(default_factory=list)
    duals: list = field(default_factory=list)
    columns: list = field(default_factory=list)   # per round: [(block, ind), ...] added


def solve_master(model, cols):
    """Restricted master: min sum cost*lambda + M*sum(art) s.t. assignment rows == 1 (with one
    artificial each), convexity rows == 1.  Returns (z, u in DIP layout [core | convex], lambda)."""
    n_core, m = model.n_base_rows, model.m
    nrows = n_core + m
    data, ri, ci = [], [], []
    for k, c in enumerate(cols):
        data.extend(c.vals)
        ri.extend(c.rows)
        ci.extend([k] * len(c.rows))
    ncol = len(cols)
    for j in range(n_core):  # artificials on the core rows
        data.append(1.0)
        ri.append(j)
        ci.append(ncol + j)
    A = csc_matrix((data, (ri, ci)), shape=(nrows, ncol + n_core))
    cost = np.concatenate([[c.cost for c in cols], np.full(n_core, BIG_M)])
    res = linprog(cost, A_eq=A, b_eq=np.ones(nrows), bounds=(0, None), method="highs-ds")
    assert res.status == 0, res.message
    u = np.asarray(res.eqlin.marginals, np.float64)   # d z / d rhs: rc = c - A^T u
    return float(res.fun), u, res.x[:ncol]


def run(model, price, max_rounds=200, eps=1e-4, log=None, init_cols=()):
    """price(u) -> (list of (block, ind tuple, redCost, origCost, rows, vals), mostNegReducedCost).
    init_cols: columns (block, ind, origCost, rows, vals) the master starts with beside the empty ones -- what
    DecompAlgo::generateInitVars provides (Dip/src/DecompAlgo.cpp:3400-3560)."""
    m = model.m
    # the empty column of every block (cost 0, only its convexity row), as GAP_DecompApp returns
    # when nothing prices out (GAP_DecompApp.cpp:278-284)
    cols = [Column(b, (), 0.0, np.array([model.n_base_rows + b], np.int32), np.array([1.0])) for b in range(m)]
    seen = {(b, ()) for b in range(m)}
    for (b, ind, oc, rows, vals) in init_cols:
        key = (int(b), tuple(int(i) for i in ind))
        if key not in seen:
            seen.add(key)
            cols.append(Column(key[0], key[1], float(oc), np.asarray(rows, np.int32), np.asarray(vals, np.float64)))
    tr = Trace()
    for it in range(max_rounds):
        z, u, lam = solve_master(model, cols)
        new, most_neg = price(u)
        added = []
        for (b, ind, rc, oc, rows, vals) in new:
            key = (b, tuple(int(i) for i in ind))
            if key in seen:          # DecompVarPool::isDuplicate
                continue
            seen.add(key)
            cols.append(Column(b, key[1], oc, np.asarray(rows, np.int32), np.asarray(vals, np.float64)))
            added.append(key)
        tr.z_rmp.append(z)
        tr.lower_bound.append(z + most_neg)
        tr.n_new.append(len(added))
        tr.duals.append(u)
        tr.columns.append(added)
        if log:
            log(f"round {it:3d}  z_RMP {z:14.6f}  LB {z + most_neg:14.6f}  new cols {len(added):3d}")
        if not added or most_neg > -eps:
            break
    return tr, cols


def run_stabilised(model, price_ladder, price_plain, accept, alpha0=0.1, n_steps=6, max_rounds=300, eps=1e-4, log=None):
    """The same loop under Wentges smoothing, following DecompAlgoPC's contract (Dip/src/DecompAlgoPC.cpp:126-212,
    DecompAlgoPC.h:99-133, Dip/src/DecompAlgo.cpp:5642-5655):
      * the round is priced with dualST = alpha*dual + (1-alpha)*dualRM, dual = the stability centre (first
        call: dualRM itself);
      * the bound of the round is the Lagrangian bound AT dualST: b.dualST + mostNegReducedCost
        (updateObjBound reads getMasterDualSolution(), DecompAlgo.cpp:3728-3758); when it improves the best bound the
        centre moves to dualST (setObjBound);
      * if every new column is a duplicate, alpha *= 0.9 and the round is priced again -- here the whole ladder
        alpha0 * 0.9^k, k < n_steps, is priced in ONE submission and the first step that yields a new column is
        taken; if none does, the round is priced at dualRM itself (alpha = 0), and no new column there ends it.
    price_ladder(uRM, alpha0, n_steps) -> [(cols, mostNeg, dualST)] per step; price_plain(uRM) -> (cols, mostNeg);
    accept(step) moves the centre to that step's dualST (step None: to dualRM)."""
    m = model.m
    cols = [Column(b, (), 0.0, np.array([model.n_base_rows + b], np.int32), np.array([1.0])) for b in range(m)]
    seen = {(b, ()) for b in range(m)}
    rhs = np.ones(model.n_base_rows + m)
    tr = Trace()
    best_lb = -np.inf
    for it in range(max_rounds):
        z, u_rm, lam = solve_master(model, cols)
        steps = price_ladder(u_rm, alpha0, n_steps)
        added, used, most_neg, u_used = [], None, 0.0, u_rm
        for k, (new, mn, u_st) in enumerate(steps):
            fresh = [c for c in new if (c[0], tuple(int(i) for i in c[1])) not in seen]
            lb = float(np.dot(rhs, u_st)) + mn
            if lb > best_lb + 1e-9:          # the bound improved: this smoothed vector becomes the centre
                best_lb = lb
                accept(k)
            if fresh:
                added, used, most_neg, u_used = fresh, k, mn, u_st
                break
        if not added:                        # the ladder produced only duplicates: price at the master duals
            new, mn = price_plain(u_rm)
            best_lb = max(best_lb, z + mn)
            added = [c for c in new if (c[0], tuple(int(i) for i in c[1])) not in seen]
            most_neg, u_used = mn, u_rm
        keys = []
        for (b, ind, rc, oc, rows, vals) in added:
            key = (b, tuple(int(i) for i in ind))
            seen.add(key)
            cols.append(Column(b, key[1], oc, np.asarray(rows, np.int32), np.asarray(vals, np.float64)))
            keys.append(key)
        tr.z_rmp.append(z)
        tr.lower_bound.append(best_lb)
        tr.n_new.append(len(keys))
        tr.duals.append(u_used)
        tr.columns.append(keys)
        if log:
            log(f"round {it:3d}  z_RMP {z:14.6f}  best LB {best_lb:14.6f}  step {used}  new cols {len(keys):3d}")
        if not keys or z - best_lb <= eps:
            break
    return tr, cols


def gpu_stab_pricer(model):
    """The stabilised backend on the GPU: dipgpu_price_round_stab_ladder + dipgpu_select_batch_result +
    dipgpu_expand_columns per step, dipgpu_stab_accept / dipgpu_stab_set_center for the centre."""
    from dip_b200.pricer import GpuPricer
    p = GpuPricer(0)
    p.setModel(model)
    state = {"u_rm": None}

    def cols_of(res):
        cs, ri, va, hs = p.expandColumns()
        return [(int(res.colBlock[k]), res.column(k).copy(), float(res.colRedCost[k]), float(res.colOrigCost[k]),
                 ri[cs[k]:cs[k + 1]].copy(), va[cs[k]:cs[k + 1]].copy()) for k in range(res.nCols)]

    def price_ladder(u_rm, alpha0, n_steps):
        state["u_rm"] = u_rm
        results, alphas = p.priceRoundStabLadder(u_rm, alpha0, n_steps)
        out = []
        for k, res in enumerate(results):
            p.selectBatchResult(k)
            out.append((cols_of(res), res.mostNegReducedCost, p.smoothedDuals(k)))
        return out

    def price_plain(u_rm):
        res = p.priceRound(u_rm)
        return cols_of(res), res.mostNegReducedCost

    def accept(step):
        p.stabAccept(step)
    return price_ladder, price_plain, accept, p


def gpu_pricer(model):
    from dip_b200.pricer import GpuPricer
    p = GpuPricer(0)
    p.setModel(model)

    def price(u):
        res = p.priceRound(u)
        cs, ri, va, hs = p.expandColumns()
        out = []
        for k in range(res.nCols):
            out.append((int(res.colBlock[k]), res.column(k).copy(), float(res.colRedCost[k]),
                        float(res.colOrigCost[k]), ri[cs[k]:cs[k + 1]].copy(), va[cs[k]:cs[k + 1]].copy()))
        return out, res.mostNegReducedCost
    return price, p


def main():
    if len(sys.argv) > 1:
        g = I.parse_gap(open(sys.argv[1]).read(), os.path.basename(sys.argv[1]))
    else:
        g = I.gap_class_d(8, 60, seed=7)
    model = I.gap_pricing_model(g, branch_rows=False)
    price, p = gpu_pricer(model)
    tr, cols = run(model, price, log=print)
    print(f"{g.name}: DW bound {tr.z_rmp[-1]:.6f} after {len(tr.z_rmp)} rounds, {len(cols)} columns")
    p.close()


if __name__ == "__main__":
    main()

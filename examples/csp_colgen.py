#!/usr/bin/env python3
"""Mini branch-and-price driver for DIP's cutting-stock example with the pricing round on the GPU.

The instance is the one of Dip/src/dippy/tests/test_cutting_stock.py:44-62 (lengths 9/7/5, demands 1/1/4, rolls of
length 20, one block per candidate roll), whose optimum the reference pins at 2.0 (:20, :34).  Model as the test
builds it (:64-96): x = cut(p, i) for every roll p and item i, then use(p); core rows = demand rows
sum_p cut(p, i) >= d_i and ordering rows use(p) - use(p+1) >= 0; block p = {length . cut(p, .) <= L * use(p)}, priced
by the integer knapsack kp (:154-179) -- dipgpu_set_intknap_blocks.

Every node follows the CONTRACT of DIP's loop (Dip/src/DecompAlgo.cpp:1782-2293) with a stand-in restricted master
(HiGHS via scipy; Clp is not installable here): master duals in DIP's layout [core rows | convexity rows] ->
generateVars -> master columns [A''s ; e_b] -> until nothing prices out (lower bound z_RMP + mostNegReducedCost,
:3757; the objective counts rolls, so ceil(bound) is valid).  Branching is DIP's BranchEnforceInMaster: the core
carries one x_j <= ub_j and one x_j >= lb_j identity row per column (:1431-1455), free until a node bounds them, so
a node changes right-hand sides only and the duals of its branch rows reach the pricing problem through the
reduced costs -- the GPU model never changes.  The integer restricted master over the columns generated so far is
the primal heuristic.  The pricing backend is injected so that the tests drive the same search with the CPU oracle
and compare every round.

    python examples/csp_colgen.py
"""
from __future__ import annotations

import math
import os
import sys
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize import Bounds, LinearConstraint, linprog, milp
from scipy.sparse import csr_matrix, hstack, vstack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BIG_M = 1.0e3


@dataclass
class CspModel:
    """x-space: cut(p, i) at p * nItems + i, use(p) at P * nItems + p."""
    lengths: np.ndarray
    demand: np.ndarray
    roll_len: int
    P: int
    R: int = 0
    N: int = 0
    m: int = 0
    n_base_rows: int = 0
    row_start: np.ndarray = None
    col_ind: np.ndarray = None
    val: np.ndarray = None
    objective: np.ndarray = None
    block_col_start: np.ndarray = None
    weight: np.ndarray = None
    capacity: np.ndarray = None
    use_col: np.ndarray = None
    rhs: np.ndarray = None    # of the core rows (all >=)
    meta: dict = field(default_factory=dict)


def csp_model(lengths=(9, 7, 5), demand=(1, 1, 4), roll_len=20, P=None, cut_ub=10) -> CspModel:
    """Core rows in DIP's order (Dip/src/DecompAlgo.cpp:1431-1455): the model's own rows, then one x_j <= ub_j row and
    one x_j >= lb_j row per (integer) column -- the identity rows through which BranchEnforceInMaster branches; they are
    relaxed to free rows until a node bounds them (Dip/src/AlpsDecompTreeNode.cpp:215-238)."""
    lengths = np.asarray(lengths, np.int32)
    demand = np.asarray(demand, np.int32)
    nI = len(lengths)
    P = int(demand.sum()) if P is None else P          # total_patterns, test_cutting_stock.py:57
    N = P * nI + P
    rows = []
    for i in range(nI):                                # demand rows, :78-80
        rows.append(([p * nI + i for p in range(P)], [1.0] * P))
    for p in range(P - 1):                             # ordering rows, :83-85
        rows.append(([P * nI + p, P * nI + p + 1], [1.0, -1.0]))
    n_own = len(rows)
    for _ in range(2):                                 # x <= ub rows, then x >= lb rows
        for j in range(N):
            rows.append(([j], [1.0]))
    R = len(rows)
    rs = np.zeros(R + 1, np.int64)
    ci, va = [], []
    for r, (c, v) in enumerate(rows):
        ci.extend(c)
        va.extend(v)
        rs[r + 1] = len(ci)
    obj = np.zeros(N)
    obj[P * nI:] = 1.0                                 # :75
    mdl = CspModel(lengths, demand, roll_len, P, R=R, N=N, m=P, n_base_rows=R, row_start=rs,
                   col_ind=np.asarray(ci, np.int32), val=np.asarray(va, np.float64), objective=obj,
                   block_col_start=np.arange(P + 1, dtype=np.int32) * nI, weight=np.tile(lengths, P).astype(np.int32),
                   capacity=np.full(P, roll_len, np.int32), use_col=(P * nI + np.arange(P)).astype(np.int32),
                   rhs=np.concatenate([demand.astype(np.float64), np.zeros(P - 1)]))
    mdl.meta = dict(n_own=n_own, col_ub=np.concatenate([np.full(P * nI, float(cut_ub)), np.ones(P)]))  # :69-71
    return mdl


_CSR = {}


def master_column(model, block, ind, els):
    """[A''s ; e_b]: what DecompAlgo::addVarsToPool forms per variable (Dip/src/DecompAlgo.cpp:5452-5558)."""
    A = _CSR.get(id(model))
    if A is None:
        A = _CSR[id(model)] = csr_matrix((model.val, model.col_ind, model.row_start), shape=(model.R, model.N))
    s = np.zeros(model.N)
    s[np.asarray(ind, np.int64)] = els
    col = A @ s
    rows = np.nonzero(np.abs(col) > 1e-6)[0]
    return (np.concatenate([rows, [model.n_base_rows + block]]).astype(np.int32),
            np.concatenate([col[rows], [1.0]]))


@dataclass
class Column:
    block: int
    ind: tuple
    els: tuple
    cost: float
    rows: np.ndarray
    vals: np.ndarray


def _node_rows(model, node_lb, node_ub):
    """Master rows of a node: (row ids of the >= rows, their rhs, row ids of the <= rows, their rhs).  A branch row is
    in the LP only while the node bounds it; a free row's dual is 0."""
    n_own, N = model.meta["n_own"], model.N
    ge = list(range(n_own))
    ge_rhs = list(model.rhs)
    le, le_rhs = [], []
    for j in range(N):
        if node_ub[j] < model.meta["col_ub"][j]:
            le.append(n_own + j)
            le_rhs.append(node_ub[j])
        if node_lb[j] > 0:
            ge.append(n_own + N + j)
            ge_rhs.append(node_lb[j])
    return np.asarray(ge), np.asarray(ge_rhs, np.float64), np.asarray(le, np.int64), np.asarray(le_rhs, np.float64)


def _master_matrix(model, cols):
    nrows = model.R + model.m
    data, ri, ci = [], [], []
    for k, c in enumerate(cols):
        data.extend(c.vals)
        ri.extend(c.rows)
        ci.extend([k] * len(c.rows))
    return csr_matrix((data, (ri, ci)), shape=(nrows, len(cols)))


def solve_master(model, cols, node_lb, node_ub):
    """min cost.lambda + M.art  s.t.  >= rows (an artificial on each with a positive rhs), <= rows, convexity == 1.
    Returns (z, u in DIP's layout [core rows | convexity rows], lambda, sum of the artificials)."""
    A = _master_matrix(model, cols)
    R, nc = model.R, len(cols)
    ge, ge_rhs, le, le_rhs = _node_rows(model, node_lb, node_ub)
    art = [k for k in range(len(ge)) if ge_rhs[k] > 0]
    Aart = csr_matrix((np.ones(len(art)), (art, np.arange(len(art)))), shape=(len(ge), len(art)))
    Age = hstack([A[ge], Aart]).tocsr()
    Aub = -Age
    bub = -ge_rhs
    if len(le):
        Ale = hstack([A[le], csr_matrix((len(le), len(art)))]).tocsr()
        Aub = vstack([Aub, Ale]).tocsr()
        bub = np.concatenate([bub, le_rhs])
    Aeq = hstack([A[R:], csr_matrix((model.m, len(art)))]).tocsr()
    cost = np.concatenate([[c.cost for c in cols], np.full(len(art), BIG_M)])
    res = linprog(cost, A_ub=Aub, b_ub=bub, A_eq=Aeq, b_eq=np.ones(model.m), bounds=(0, None), method="highs-ds")
    assert res.status == 0, res.message
    mg = np.asarray(res.ineqlin.marginals)
    u = np.zeros(R + model.m)                 # rc = c - A^T u
    u[ge] = -mg[:len(ge)]
    if len(le):
        u[le] = mg[len(ge):]
    u[R:] = np.asarray(res.eqlin.marginals)
    return float(res.fun), u, res.x[:nc], float(res.x[nc:].sum())


def solve_master_integer(model, cols):
    """The restricted master over the generated columns with integral lambdas: a primal heuristic (any integral
    combination of generated columns that meets the model's own rows is a solution of the original problem)."""
    A = _master_matrix(model, cols)
    R, n_own = model.R, model.meta["n_own"]
    cons = [LinearConstraint(A[:n_own], lb=model.rhs, ub=np.inf), LinearConstraint(A[R:], lb=1.0, ub=1.0)]
    res = milp(np.array([c.cost for c in cols]), constraints=cons, integrality=np.ones(len(cols)), bounds=Bounds(0, 1))
    if res.status != 0:
        return np.inf, None
    return float(res.fun), res.x


@dataclass
class Trace:
    z_rmp: list = field(default_factory=list)
    lower_bound: list = field(default_factory=list)
    duals: list = field(default_factory=list)
    columns: list = field(default_factory=list)
    nodes: int = 0


class Pool:
    def __init__(self, model):
        self.cols = []
        self.seen = set()
        for b in range(model.m):   # the empty pattern of every roll (use = 0): relaxed_solver's else branch, :131-147
            self.cols.append(Column(b, (), (), 0.0, np.array([model.n_base_rows + b], np.int32), np.array([1.0])))
            self.seen.add((b, (), ()))


def price_node(model, price, pool, node_lb, node_ub, tr, max_rounds=200, eps=1e-4, log=None):
    """Column generation at one node.  Returns (z_RMP, lower bound, lambda, infeasible)."""
    for it in range(max_rounds):
        z, u, lam, art = solve_master(model, pool.cols, node_lb, node_ub)
        new, most_neg = price(u)
        added = []
        for (b, ind, els, rc, oc) in new:
            key = (int(b), tuple(int(i) for i in ind), tuple(float(e) for e in els))
            if key in pool.seen:                             # DecompVarPool::isDuplicate
                continue
            pool.seen.add(key)
            rows, vals = master_column(model, key[0], key[1], key[2])
            pool.cols.append(Column(key[0], key[1], key[2], float(oc), rows, vals))
            added.append(key)
        tr.z_rmp.append(z)
        tr.lower_bound.append(z + most_neg)
        tr.duals.append(u)
        tr.columns.append(added)
        if log:
            log(f"  round {it:3d}  z_RMP {z:12.6f}  LB {z + most_neg:12.6f}  new cols {len(added):3d}")
        if not added or most_neg > -eps:
            return z, z + most_neg, lam, art > 1e-7
    raise RuntimeError("column generation did not converge")


def solve(model, price, log=None, max_nodes=200):
    """Depth-first branch-and-price on the original columns through the master's branch rows, the integer restricted
    master as primal heuristic.  Returns (global lower bound, best integer objective, trace, columns)."""
    pool = Pool(model)
    tr = Trace()
    best = np.inf
    root_lb = None
    stack = [(np.zeros(model.N), model.meta["col_ub"].copy())]
    while stack and tr.nodes < max_nodes:
        node_lb, node_ub = stack.pop()
        tr.nodes += 1
        if log:
            log(f"node {tr.nodes}: {int((node_lb > 0).sum())} lower / {int((node_ub < model.meta['col_ub']).sum())} upper bounds")
        z, lb, lam, infeasible = price_node(model, price, pool, node_lb, node_ub, tr, log=log)
        bound = math.ceil(lb - 1e-6)                         # the objective counts rolls: integral
        if root_lb is None:
            root_lb = bound
        if infeasible or bound >= best:
            continue
        z_int, _ = solve_master_integer(model, pool.cols)
        best = min(best, z_int)
        if best <= root_lb:
            break
        x = np.zeros(model.N)
        for c, l in zip(pool.cols, lam):
            if l > 1e-9 and c.ind:
                x[list(c.ind)] += l * np.asarray(c.els)
        frac = np.abs(x - np.round(x))
        if frac.max() < 1e-6:
            best = min(best, z)
            continue
        j = int(np.argmax(frac > frac.max() - 1e-9))         # most fractional, lowest index
        up_lb, dn_ub = node_lb.copy(), node_ub.copy()
        up_lb[j] = math.ceil(x[j])
        dn_ub[j] = math.floor(x[j])
        stack.append((up_lb, node_ub))
        stack.append((node_lb, dn_ub))                       # the down branch first
    return root_lb, best, tr, pool.cols


def gpu_pricer(model, device=0):
    from dip_b200.pricer import GpuPricer
    p = GpuPricer(device)
    p.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
    p.setObjective(model.objective)
    p.setIntKnapBlocks(model.block_col_start, model.weight, model.capacity, model.use_col)

    def price(u):
        res = p.priceRound(u)
        out = [(int(res.colBlock[k]), res.column(k).copy(), res.elements(k).copy(), float(res.colRedCost[k]),
                float(res.colOrigCost[k])) for k in range(res.nCols)]
        return out, res.mostNegReducedCost
    return price, p


def oracle_pricer(model):
    """The same backend on the CPU oracle (TEST INFRASTRUCTURE: used by tests/ only)."""
    from oracle import dip_oracle as orc

    def price(u):
        rc = orc.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val,
                              orc.adjust_duals(u, model.n_base_rows, model.m), model.objective)
        ref = orc.solve_intknap_blocks(model.block_col_start, model.weight, model.capacity, model.use_col, rc,
                                       model.objective, u[model.n_base_rows:model.n_base_rows + model.m])
        out = [(b, ref.col_ind[b], ref.col_cnt[b].astype(np.float64), float(ref.col_red_cost[b]),
                float(ref.col_orig_cost[b])) for b in range(model.m) if ref.col_keep[b]]
        return out, ref.most_neg_reduced_cost
    return price


def main():
    model = csp_model()
    price, p = gpu_pricer(model)
    lb, z_int, tr, cols = solve(model, price, log=print)
    print(f"cutting stock {model.lengths.tolist()} x {model.demand.tolist()} / {model.roll_len}: lower bound {lb}, "
          f"best integer solution {z_int:.1f} ({tr.nodes} nodes, {len(tr.z_rmp)} pricing rounds, {len(cols)} columns)")
    p.close()


if __name__ == "__main__":
    main()

"""ctypes/numpy front end of the CPU oracle (oracle/dip_oracle.c).

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under dip_b200/ does.

Also exposes the reference's own Horowitz-Sahni B&B (``ref_knapsack_hs``) when
oracle/_ref/libdip_ref_knap.so has been built from
/root/reference/Dip/src/UtilKnapsack.cpp (see oracle/Makefile).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile libdip_oracle.so (and oracle/_ref when the reference is present)."""
    so = os.path.join(_HERE, "libdip_oracle.so")
    src = os.path.join(_HERE, "dip_oracle.c")
    stale = (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale:
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["make", "-C", _HERE, "libdip_oracle.so"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    ref = os.path.join(_HERE, "_ref", "libdip_ref_knap.so")
    if (force or not os.path.exists(ref)) and os.path.exists("/root/reference/Dip/src/UtilKnapsack.cpp"):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, env=env,
                       stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    build()
    lib = C.CDLL(os.path.join(_HERE, "libdip_oracle.so"))
    lib.dip_oracle_adjust_duals.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, _f64p]
    lib.dip_oracle_adjust_duals.restype = None
    lib.dip_oracle_calc_redcost.argtypes = [C.c_int, C.c_int, _i64p, _i32p, _f64p, _f64p, _f64p,
                                            C.c_int, _f64p]
    lib.dip_oracle_calc_redcost.restype = None
    lib.dip_oracle_knapsack01.argtypes = [C.c_int, _f64p, _i32p, C.c_int, _i32p,
                                          C.POINTER(C.c_int), C.c_void_p]
    lib.dip_oracle_knapsack01.restype = C.c_double
    lib.dip_oracle_calc_scale_factor.argtypes = [C.c_int, _f64p, _f64p]
    lib.dip_oracle_calc_scale_factor.restype = C.c_int
    lib.dip_oracle_solve_blocks.argtypes = [
        C.c_int, _i32p, _i32p, _i32p, C.c_int, _f64p, _f64p, _f64p, C.c_double, C.c_int, C.c_int,
        _i32p, _f64p, _f64p, _i32p, _f64p, _f64p, C.c_void_p, C.POINTER(C.c_double),
        C.POINTER(C.c_int64)]
    lib.dip_oracle_solve_blocks.restype = C.c_int
    lib.dip_oracle_price_round.argtypes = [
        C.c_int, C.c_int, _i64p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_int, _i32p,
        _i32p, _i32p, C.c_int, C.c_double, C.c_int, C.c_int, _f64p, _i32p, _f64p, _f64p, _i32p,
        _f64p, _f64p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.dip_oracle_price_round.restype = C.c_int
    lib.dip_oracle_expand_column.argtypes = [C.c_int, C.c_int, _i64p, _i32p, _f64p, C.c_int, C.c_int, C.c_int,
                                             _i32p, C.c_void_p, C.c_int, C.c_double, _i32p, _f64p]
    lib.dip_oracle_expand_column.restype = C.c_int
    lib.dip_oracle_string_hash.argtypes = [C.c_int, _i32p, _f64p, C.c_char_p, C.c_int]
    lib.dip_oracle_string_hash.restype = C.c_int
    lib.dip_oracle_solve_blocks_bounded.argtypes = [
        C.c_int, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p, C.c_double, C.c_int,
        _i32p, _f64p, _f64p, _i32p, _f64p, _f64p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.dip_oracle_solve_blocks_bounded.restype = C.c_int
    lib.dip_oracle_solve_mckp_blocks.argtypes = [
        C.c_int, _i32p, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_double, C.c_int,
        _i32p, _f64p, _f64p, _i32p, _f64p, _f64p, _i32p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.dip_oracle_solve_mckp_blocks.restype = C.c_int
    lib.dip_oracle_set_hs.argtypes = [C.c_void_p]
    lib.dip_oracle_set_hs.restype = None
    lib.dip_oracle_calc_redcost_csc.argtypes = [C.c_int, _i64p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_int, _f64p]
    lib.dip_oracle_calc_redcost_csc.restype = None
    lib.dip_oracle_intknap_kp.argtypes = [C.c_int, _f64p, _i32p, C.c_int, _i32p]
    lib.dip_oracle_intknap_kp.restype = C.c_double
    lib.dip_oracle_solve_intknap_blocks.argtypes = [
        C.c_int, _i32p, _i32p, _i32p, C.c_void_p, _f64p, _f64p, _f64p, C.c_double, C.c_int,
        _i32p, _f64p, _f64p, _i32p, _f64p, _f64p, _i32p, _i32p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.dip_oracle_solve_intknap_blocks.restype = C.c_int
    lib.dip_oracle_pool_reduced_costs.argtypes = [C.c_int, _i64p, _i32p, _f64p, _f64p, _f64p, C.c_int, _f64p]
    lib.dip_oracle_pool_reduced_costs.restype = C.c_int
    lib.dip_oracle_smooth_duals.argtypes = [C.c_int, C.c_double, _f64p, _f64p, _f64p]
    lib.dip_oracle_smooth_duals.restype = None
    _LIB = lib
    return lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libdip_ref_knap.so"))


def _ref():
    global _REF
    if _REF is None:
        lib = C.CDLL(os.path.join(_HERE, "_ref", "libdip_ref_knap.so"))
        # int KnapsackOptimizeHS(int n, double c, double* p, double* w, int* x, double* z, int* st)
        fn = lib._Z18KnapsackOptimizeHSidPdS_PiS_S0_
        fn.argtypes = [C.c_int, C.c_double, _f64p, _f64p, _i32p, C.POINTER(C.c_double),
                       C.POINTER(C.c_int)]
        fn.restype = C.c_int
        _REF = fn
    return _REF


# --------------------------------------------------------------------------- API


def adjust_duals(u: np.ndarray, n_base: int, m: int) -> np.ndarray:
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.empty(len(u) - m, dtype=np.float64)
    _lib().dip_oracle_adjust_duals(u, len(u), n_base, m, out)
    return out


def calc_redcost(R, N, row_start, col_ind, val, u_adj, c, phase1=False) -> np.ndarray:
    out = np.empty(N, dtype=np.float64)
    _lib().dip_oracle_calc_redcost(
        R, N, np.ascontiguousarray(row_start, np.int64), np.ascontiguousarray(col_ind, np.int32),
        np.ascontiguousarray(val, np.float64), np.ascontiguousarray(u_adj, np.float64),
        np.ascontiguousarray(c, np.float64), int(bool(phase1)), out)
    return out


def knapsack01(obj, weights, capacity, want_table=False):
    """Returns (z, solution_descending[, added])."""
    obj = np.ascontiguousarray(obj, np.float64)
    weights = np.ascontiguousarray(weights, np.int32)
    n = len(obj)
    sol = np.zeros(max(n, 1), np.int32)
    ns = C.c_int(0)
    tab = np.zeros((max(n, 1), capacity + 1), np.uint8) if want_table else None
    z = _lib().dip_oracle_knapsack01(n, obj, weights, int(capacity), sol, C.byref(ns),
                                     tab.ctypes.data if want_table else None)
    if want_table:
        return z, sol[: ns.value].copy(), tab[:n]
    return z, sol[: ns.value].copy()


def calc_scale_factor(red_cost) -> int:
    red_cost = np.ascontiguousarray(red_cost, np.float64)
    work = np.zeros(max(len(red_cost), 1), np.float64)
    return _lib().dip_oracle_calc_scale_factor(len(red_cost), red_cost, work)


@dataclass
class RoundResult:
    red_cost_x: np.ndarray | None
    col_nnz: np.ndarray
    col_red_cost: np.ndarray      # includes -alpha
    col_orig_cost: np.ndarray
    col_keep: np.ndarray
    most_neg_rc: np.ndarray
    z: np.ndarray
    col_ind: list                 # per block: ascending x-space indices
    most_neg_reduced_cost: float
    n_kept: int
    cells: int


def _alloc(m, max_cols, want_ind):
    return (np.zeros(m, np.int32), np.zeros(m, np.float64), np.zeros(m, np.float64),
            np.zeros(m, np.int32), np.zeros(m, np.float64), np.zeros(m, np.float64),
            np.zeros((m, max_cols), np.int32) if want_ind else None)


def solve_blocks(block_col_start, weight, capacity, red_cost_x, orig_cost, alpha,
                 profit_mode=0, red_cost_eps=1e-4, n_threads=1, want_ind=True) -> RoundResult:
    bcs = np.ascontiguousarray(block_col_start, np.int32)
    m = len(bcs) - 1
    max_cols = int(np.max(np.diff(bcs))) if m > 0 else 1
    max_cols = max(max_cols, 1)
    nnz, rc, oc, keep, mn, z, ind = _alloc(m, max_cols, want_ind)
    tot = C.c_double(0.0)
    cells = C.c_int64(0)
    kept = _lib().dip_oracle_solve_blocks(
        m, bcs, np.ascontiguousarray(weight, np.int32), np.ascontiguousarray(capacity, np.int32),
        int(profit_mode), np.ascontiguousarray(red_cost_x, np.float64),
        np.ascontiguousarray(orig_cost, np.float64), np.ascontiguousarray(alpha, np.float64),
        float(red_cost_eps), int(n_threads), max_cols, nnz, rc, oc, keep, mn, z,
        ind.ctypes.data if want_ind else None, C.byref(tot), C.byref(cells))
    cols = [ind[b, : nnz[b]].copy() for b in range(m)] if want_ind else []
    return RoundResult(None, nnz, rc, oc, keep, mn, z, cols, tot.value, kept, cells.value)


def solve_blocks_bounded(block_col_start, weight, capacity, red_cost_x, orig_cost, alpha, col_lb, col_ub,
                         red_cost_eps=1e-4, want_ind=True) -> RoundResult:
    """Stage (b)+(c) with the node's column bounds enforced in the block subproblems."""
    bcs = np.ascontiguousarray(block_col_start, np.int32)
    m = len(bcs) - 1
    max_cols = max(int(np.max(np.diff(bcs))) if m > 0 else 1, 1)
    nnz, rc, oc, keep, mn, z, ind = _alloc(m, max_cols, want_ind)
    tot = C.c_double(0.0)
    cells = C.c_int64(0)
    kept = _lib().dip_oracle_solve_blocks_bounded(
        m, bcs, np.ascontiguousarray(weight, np.int32), np.ascontiguousarray(capacity, np.int32),
        np.ascontiguousarray(red_cost_x, np.float64), np.ascontiguousarray(orig_cost, np.float64),
        np.ascontiguousarray(alpha, np.float64), np.ascontiguousarray(col_lb, np.float64),
        np.ascontiguousarray(col_ub, np.float64), float(red_cost_eps), max_cols, nnz, rc, oc, keep, mn, z,
        ind.ctypes.data if want_ind else None, C.byref(tot), C.byref(cells))
    cols = [ind[b, : nnz[b]].copy() for b in range(m)] if want_ind else []
    return RoundResult(None, nnz, rc, oc, keep, mn, z, cols, tot.value, kept, cells.value)


def solve_mckp_blocks(block_group_start, group_col_start, weight, capacity, red_cost_x, orig_cost, alpha,
                      red_cost_eps=1e-4) -> RoundResult:
    """Multiple-choice knapsack blocks (MMKP MCKP0 relaxation): exactly one column per group."""
    bgs = np.ascontiguousarray(block_group_start, np.int32)
    gcs = np.ascontiguousarray(group_col_start, np.int32)
    m = len(bgs) - 1
    max_cols = max(int(np.max(np.diff(bgs))) if m > 0 else 1, 1)   # one column per group
    nnz, rc, oc, keep, mn, z, ind = _alloc(m, max_cols, True)
    tot = C.c_double(0.0)
    cells = C.c_int64(0)
    kept = _lib().dip_oracle_solve_mckp_blocks(
        m, bgs, gcs, np.ascontiguousarray(weight, np.int32), np.ascontiguousarray(capacity, np.int32),
        np.ascontiguousarray(red_cost_x, np.float64), np.ascontiguousarray(orig_cost, np.float64),
        np.ascontiguousarray(alpha, np.float64), float(red_cost_eps), max_cols, nnz, rc, oc, keep, mn, z, ind,
        C.byref(tot), C.byref(cells))
    cols = [ind[b, : nnz[b]].copy() for b in range(m)]
    return RoundResult(None, nnz, rc, oc, keep, mn, z, cols, tot.value, kept, cells.value)


def intknap_kp(obj, weights, capacity):
    """kp of Dip/src/dippy/tests/test_cutting_stock.py:154-179: (z, counts per item)."""
    obj = np.ascontiguousarray(obj, np.float64)
    w = np.ascontiguousarray(weights, np.int32)
    cnt = np.zeros(max(len(obj), 1), np.int32)
    z = _lib().dip_oracle_intknap_kp(len(obj), obj if len(obj) else np.zeros(1), w if len(w) else np.zeros(1, np.int32),
                                     int(capacity), cnt)
    return z, cnt[: len(obj)].copy()


def solve_intknap_blocks(block_col_start, weight, capacity, use_col, red_cost_x, orig_cost, alpha,
                         red_cost_eps=1e-4) -> RoundResult:
    """Integer-knapsack blocks (cutting-stock pricing): col_ind[b] ascending x-space indices, col_cnt[b] their
    multiplicities (attribute added to the RoundResult)."""
    bcs = np.ascontiguousarray(block_col_start, np.int32)
    m = len(bcs) - 1
    max_e = max(int(np.max(np.diff(bcs))) if m > 0 else 1, 1) + 1
    nnz, rc, oc, keep, mn, z, ind = _alloc(m, max_e, True)
    cnt = np.zeros((m, max_e), np.int32)
    tot = C.c_double(0.0)
    cells = C.c_int64(0)
    use = None if use_col is None else np.ascontiguousarray(use_col, np.int32)
    kept = _lib().dip_oracle_solve_intknap_blocks(
        m, bcs, np.ascontiguousarray(weight, np.int32), np.ascontiguousarray(capacity, np.int32),
        None if use is None else use.ctypes.data, np.ascontiguousarray(red_cost_x, np.float64),
        np.ascontiguousarray(orig_cost, np.float64), np.ascontiguousarray(alpha, np.float64), float(red_cost_eps),
        max_e, nnz, rc, oc, keep, mn, z, ind, cnt, C.byref(tot), C.byref(cells))
    res = RoundResult(None, nnz, rc, oc, keep, mn, z, [ind[b, : nnz[b]].copy() for b in range(m)], tot.value, kept,
                      cells.value)
    res.col_cnt = [cnt[b, : nnz[b]].copy() for b in range(m)]
    return res


def price_round(R, N, row_start, col_ind, val, c, u, n_base, m, block_col_start, weight,
                capacity, phase1=False, profit_mode=0, red_cost_eps=1e-4, n_threads=1,
                want_ind=True) -> RoundResult:
    bcs = np.ascontiguousarray(block_col_start, np.int32)
    max_cols = max(int(np.max(np.diff(bcs))) if m > 0 else 1, 1)
    nnz, rc, oc, keep, mn, z, ind = _alloc(m, max_cols, want_ind)
    rcx = np.zeros(N, np.float64)
    tot = C.c_double(0.0)
    cells = C.c_int64(0)
    kept = _lib().dip_oracle_price_round(
        R, N, np.ascontiguousarray(row_start, np.int64), np.ascontiguousarray(col_ind, np.int32),
        np.ascontiguousarray(val, np.float64), np.ascontiguousarray(c, np.float64),
        np.ascontiguousarray(u, np.float64), int(n_base), int(bool(phase1)), m, bcs,
        np.ascontiguousarray(weight, np.int32), np.ascontiguousarray(capacity, np.int32),
        int(profit_mode), float(red_cost_eps), int(n_threads), max_cols, rcx, nnz, rc, oc, keep,
        mn, z, ind.ctypes.data if want_ind else None, C.byref(tot), C.byref(cells))
    cols = [ind[b, : nnz[b]].copy() for b in range(m)] if want_ind else []
    return RoundResult(rcx, nnz, rc, oc, keep, mn, z, cols, tot.value, kept, cells.value)


def price_round_which(R, N, row_start, col_ind, val, c, u, n_base, m, block_col_start, weight, capacity,
                      which, user_duals=None, profit_mode=0, red_cost_eps=1e-4):
    """DecompAlgo::generateVars' round-robin branch (Dip/src/DecompAlgo.cpp:4929-5149) followed by its
    filter loop (:5151-5232), restated on top of price_round (blocks are independent given a dual vector,
    so "solve block b with vector uhat" = block b of the round priced with uhat).
      which       blocksToSolve, in the order solveRelaxedWhich returned them
      user_duals  {block: master dual vector} = userDualsByBlock
    Returns dict(new_vars=[(block, ind, redCost, origCost)] in the reference's push order, most_neg_vec,
    most_neg_reduced_cost, solved_all).  alpha of a user-dual block is that vector's convexity dual (the
    reference indexes the convexity-free adjusted copy at nBaseCoreRows + b, :5032, past its end)."""
    user_duals = user_duals or {}
    rounds = {}

    def round_of(vec_id, vec):
        if vec_id not in rounds:
            rounds[vec_id] = price_round(R, N, row_start, col_ind, val, c, vec, n_base, m, block_col_start, weight,
                                         capacity, profit_mode=profit_mode, red_cost_eps=red_cost_eps)
        return rounds[vec_id]
    potential = []                       # potentialVars: (block, ind, redCost, origCost)
    for b in which:                      # :4952-5070
        b = int(b)
        r = round_of(id(user_duals[b]), user_duals[b]) if b in user_duals else round_of("u", u)
        potential.append((b, r.col_ind[b], float(r.col_red_cost[b]), float(r.col_orig_cost[b])))
    found = any(v[2] < -red_cost_eps for v in potential)     # :5063-5069
    if not found:                        # :5076-5148: all blocks, standard duals
        r = round_of("u", u)
        for b in range(m):
            potential.append((b, r.col_ind[b], float(r.col_red_cost[b]), float(r.col_orig_cost[b])))
    most_neg_vec = np.zeros(m)           # :4761
    new_vars = []
    for (b, ind, rc, oc) in potential:   # :5151-5232
        if rc < most_neg_vec[b]:
            most_neg_vec[b] = rc
        if rc < -red_cost_eps:
            new_vars.append((b, ind, rc, oc))
    tot = 0.0
    for b in range(m):
        tot += most_neg_vec[b]
    return dict(new_vars=new_vars, most_neg_vec=most_neg_vec, most_neg_reduced_cost=tot, solved_all=not found)


def expand_column(R, N, row_start, col_ind, val, n_base, num_convex, ind, block_id, tol_zero=1e-6):
    """Master column (rows ascending, values) of the 0/1 variable with x-space indices `ind`."""
    ind = np.ascontiguousarray(ind, np.int32)
    out_row = np.zeros(R + num_convex, np.int32)
    out_val = np.zeros(R + num_convex, np.float64)
    n = _lib().dip_oracle_expand_column(
        R, N, np.ascontiguousarray(row_start, np.int64), np.ascontiguousarray(col_ind, np.int32),
        np.ascontiguousarray(val, np.float64), int(n_base), int(num_convex), len(ind),
        ind if len(ind) else np.zeros(1, np.int32), None, int(block_id), float(tol_zero), out_row, out_val)
    return out_row[:n].copy(), out_val[:n].copy()


def pool_reduced_costs(col_start, row_ind, val, orig_cost, u, feasible=True):
    """DecompVarPool::setReducedCosts: (redCost per waiting column, found_negrc_var)."""
    cs = np.ascontiguousarray(col_start, np.int64)
    n = len(cs) - 1
    out = np.zeros(max(n, 1), np.float64)
    ri = np.ascontiguousarray(row_ind, np.int32)
    vv = np.ascontiguousarray(val, np.float64)
    if len(ri) == 0:
        ri, vv = np.zeros(1, np.int32), np.zeros(1, np.float64)
    oc = np.ascontiguousarray(orig_cost, np.float64)
    if len(oc) == 0:
        oc = np.zeros(1, np.float64)
    found = _lib().dip_oracle_pool_reduced_costs(n, cs, ri, vv, oc, np.ascontiguousarray(u, np.float64),
                                                 int(bool(feasible)), out)
    return out[:n].copy(), bool(found)


def smooth_duals(alpha, dual, dual_rm) -> np.ndarray:
    """DecompAlgoPC::adjustMasterDualSolution: alpha*dual + (1-alpha)*dualRM."""
    dual = np.ascontiguousarray(dual, np.float64)
    dual_rm = np.ascontiguousarray(dual_rm, np.float64)
    out = np.empty_like(dual)
    _lib().dip_oracle_smooth_duals(len(dual), float(alpha), dual, dual_rm, out)
    return out


def string_hash(ind, els) -> str:
    ind = np.ascontiguousarray(ind, np.int32)
    els = np.ascontiguousarray(els, np.float64)
    buf = C.create_string_buffer(32 * max(len(ind), 1) + 16)
    k = _lib().dip_oracle_string_hash(len(ind), ind, els, buf, len(buf))
    assert k >= 0
    return buf.value.decode()


PROFIT_REF_HS = 2   # bench.py CPU baseline: blocks priced by the reference's compiled KnapsackOptimizeHS


def enable_ref_hs() -> bool:
    """Hands the reference's KnapsackOptimizeHS (oracle/_ref) to the C oracle for profit_mode=PROFIT_REF_HS."""
    if not have_ref():
        return False
    fn = _ref()
    _lib().dip_oracle_set_hs(C.cast(fn, C.c_void_p))
    return True


def csc_descending(R, N, row_start, col_ind, val):
    """Column-ordered copy of a row-ordered matrix, the entries of a column in DESCENDING row order."""
    rs = np.asarray(row_start, np.int64)
    rows = np.repeat(np.arange(R, dtype=np.int32), np.diff(rs))
    ci = np.asarray(col_ind, np.int64)
    order = np.lexsort((-rows.astype(np.int64), ci))
    cs = np.zeros(N + 1, np.int64)
    np.cumsum(np.bincount(ci, minlength=N), out=cs[1:])
    return cs, np.ascontiguousarray(rows[order]), np.ascontiguousarray(np.asarray(val, np.float64)[order])


def calc_redcost_csc(N, col_start, row_ind, val, u_adj, c, phase1=False, n_threads=1) -> np.ndarray:
    out = np.zeros(N, np.float64)
    _lib().dip_oracle_calc_redcost_csc(N, np.ascontiguousarray(col_start, np.int64), np.ascontiguousarray(row_ind, np.int32),
                                       np.ascontiguousarray(val, np.float64), np.ascontiguousarray(u_adj, np.float64),
                                       np.ascontiguousarray(c, np.float64), int(bool(phase1)), int(n_threads), out)
    return out


def ref_knapsack_hs(p, w, capacity):
    """The reference's KnapsackOptimizeHS (Dip/src/UtilKnapsack.cpp:246-502).

    Items are sorted by non-increasing p/w first (its documented precondition).
    Returns (status, z, x_in_original_order).  n <= 2 is handled by the caller:
    the reference's n==1 branch writes x[1] (UtilKnapsack.cpp:314)."""
    p = np.asarray(p, np.float64)
    w = np.asarray(w, np.float64)
    n = len(p)
    order = np.argsort(-(p / w), kind="stable")
    ps = np.zeros(n + 1, np.float64)
    ws = np.zeros(n + 1, np.float64)
    ps[:n] = p[order]
    ws[:n] = w[order]
    x = np.zeros(n + 2, np.int32)
    z = C.c_double(0.0)
    st = C.c_int(0)
    rc = _ref()(n, float(capacity), ps, ws, x, C.byref(z), C.byref(st))
    xo = np.zeros(n, np.int32)
    xo[order] = x[:n]
    return rc, z.value, xo

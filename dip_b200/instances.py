"""Synthetic instances of the shapes BASELINE.json names.

The reference ships no generators; class definitions follow the GAP literature cited
in Dip/data/GAP/ReadMe.txt:10-13 (Chu-Beasley / Yagiura classes C, D, E).  Everything
here is host-side numpy: model data that DIP itself would hold (A'' in row-ordered CSR,
the objective, block structure) and master dual vectors in DIP's row layout.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class GapInstance:
    """A GAP as Dip/examples/GAP reads it (GAP_Instance.cpp:20-73): m machines, n tasks."""
    m: int
    n: int
    cost: np.ndarray      # (m, n) int64, minimised (GAP_DecompApp.cpp:187-189)
    weight: np.ndarray    # (m, n) int32
    capacity: np.ndarray  # (m,)  int32
    name: str = "gap"


@dataclass
class PricingModel:
    """What DecompAlgo holds when generateVars runs (Dip/src/DecompAlgo.cpp:4622-4726).

    A'' rows: [orig core rows | nInt x<=ub rows | nInt x>=lb rows] (DecompAlgo.cpp:1431-1455),
    nBaseRows = all of those; the master dual vector is [0,nBaseRows) | m convexity | cuts.
    """
    R: int
    N: int
    row_start: np.ndarray  # int64 (R+1)
    col_ind: np.ndarray    # int32
    val: np.ndarray        # float64
    n_base_rows: int
    objective: np.ndarray  # float64 (N)
    m: int
    block_col_start: np.ndarray  # int32 (m+1)
    weight: np.ndarray     # int32 (N), indexed by x-space column
    capacity: np.ndarray   # int32 (m)
    name: str = "model"
    meta: dict = field(default_factory=dict)

    @property
    def nnz(self) -> int:
        return int(self.row_start[-1])

    @property
    def u_len(self) -> int:
        return self.R + self.m

    def spmv_bytes(self) -> int:
        """Algorithmic HBM bytes of one reduced-cost sweep: 12 nnz + 20 N + 8 R."""
        return 12 * self.nnz + 20 * self.N + 8 * self.R


# ----------------------------------------------------------------------------- GAP


def gap_class_c(m=10, n=60, seed=1060) -> GapInstance:
    rng = np.random.default_rng(seed)
    w = rng.integers(5, 26, size=(m, n))
    c = rng.integers(10, 51, size=(m, n))
    cap = np.floor(0.8 * w.sum(axis=1) / m).astype(np.int64)
    return GapInstance(m, n, c.astype(np.int64), w.astype(np.int32), cap.astype(np.int32), f"gapC_{m}x{n}")


def gap_class_d(m=20, n=200, seed=20200) -> GapInstance:
    rng = np.random.default_rng(seed)
    w = rng.integers(1, 101, size=(m, n))
    c = 111 - w + rng.integers(-10, 11, size=(m, n))
    cap = np.floor(0.8 * w.sum(axis=1) / m).astype(np.int64)
    return GapInstance(m, n, c.astype(np.int64), w.astype(np.int32), cap.astype(np.int32), f"gapD_{m}x{n}")


def gap_class_e(m=80, n=1600, seed=801600) -> GapInstance:
    rng = np.random.default_rng(seed)
    u = 1.0 - rng.random(size=(m, n))  # (0, 1]
    w = np.floor(1.0 - 10.0 * np.log(u)).astype(np.int64)
    w = np.maximum(w, 1)
    c = np.floor(1000.0 / w - 10.0 * rng.random(size=(m, n))).astype(np.int64)
    cap = np.maximum(np.floor(0.8 * w.sum(axis=1) / m).astype(np.int64), w.max(axis=1))
    return GapInstance(m, n, c, w.astype(np.int32), cap.astype(np.int32), f"gapE_{m}x{n}")


def parse_gap(text: str, name="gap") -> GapInstance:
    """OR-Library single-instance format as Dip/examples/GAP/GAP_Instance.cpp:22-74 reads it:
    m n, then m*n profits (machine-major), m*n weights, m capacities."""
    tok = text.split()
    m, n = int(tok[0]), int(tok[1])
    v = np.array([int(t) for t in tok[2:2 + 2 * m * n + m]], np.int64)
    return GapInstance(m, n, v[:m * n].reshape(m, n), v[m * n:2 * m * n].reshape(m, n).astype(np.int32),
                       v[2 * m * n:].astype(np.int32), name)


def gap_pricing_model(g: GapInstance, branch_rows=True) -> PricingModel:
    """The GAP DW model of GAP_DecompApp::createModels (GAP_DecompApp.cpp:138-232):
    core = n assignment rows sum_i x[i,j] = 1 (:69-135), x-space index i*n + j, block i =
    columns [i*n, (i+1)*n) (:252-253).  With BranchEnforceInMaster (GAP_Main.cpp:40-41) the
    algorithm appends 2N identity rows to A'' (DecompAlgo.cpp:1399-1502)."""
    m, n = g.m, g.n
    N = m * n
    rows_ap = np.arange(n, dtype=np.int64)
    # assignment row j holds columns i*n + j, i ascending
    ap_cols = (np.arange(m, dtype=np.int64)[None, :] * n + rows_ap[:, None]).reshape(-1)
    col_ind = [ap_cols]
    counts = [np.full(n, m, np.int64)]
    if branch_rows:
        ident = np.arange(N, dtype=np.int64)
        col_ind += [ident, ident]
        counts += [np.ones(2 * N, np.int64)]
    col_ind = np.concatenate(col_ind).astype(np.int32)
    cnt = np.concatenate(counts)
    row_start = np.zeros(len(cnt) + 1, np.int64)
    np.cumsum(cnt, out=row_start[1:])
    R = len(cnt)
    return PricingModel(
        R=R, N=N, row_start=row_start, col_ind=col_ind, val=np.ones(len(col_ind), np.float64),
        n_base_rows=R, objective=g.cost.astype(np.float64).reshape(-1), m=m,
        block_col_start=(np.arange(m + 1, dtype=np.int64) * n).astype(np.int32),
        weight=g.weight.astype(np.int32).reshape(-1), capacity=g.capacity.astype(np.int32),
        name=g.name, meta=dict(n=n, branch_rows=branch_rows))


def gap_branch_rows(model: PricingModel, rng: np.random.Generator, branch_density=0.01) -> np.ndarray:
    """The branch rows (positions relative to the first branch row) that carry a bound at one node of the
    tree: `branch_density` of them, ascending.  Every other branch row is a free row (zero dual)."""
    nb = model.n_base_rows - model.meta["n"]
    return np.sort(rng.choice(nb, size=int(nb * branch_density), replace=False))


def gap_dual_support(model: PricingModel, branch_rows: np.ndarray) -> np.ndarray:
    """Master rows that can carry a dual at that node: assignment rows, the bounded branch rows, convexity rows."""
    n = model.meta["n"]
    return np.concatenate([np.arange(n), n + np.asarray(branch_rows), np.arange(model.n_base_rows, model.u_len)]).astype(np.int32)


def gap_duals(model: PricingModel, rng: np.random.Generator, branch_density=0.01, branch_rows=None) -> np.ndarray:
    """A master dual vector in DIP's row layout for a GAP model: assignment duals
    U[0, 2 mean(c)] (about half of the items get a negative reduced cost), branch-row duals 0
    except `branch_density` random entries (or exactly the rows `branch_rows`, see gap_branch_rows),
    alpha_b ~ U[-10, 0]."""
    n = model.meta["n"]
    N = model.N
    u = np.zeros(model.u_len, np.float64)
    u[:n] = rng.uniform(0.0, 2.0 * model.objective.mean(), size=n)
    nb = model.n_base_rows - n
    if nb > 0 and (branch_density > 0 or branch_rows is not None):
        k = int(nb * branch_density) if branch_rows is None else len(branch_rows)
        idx = rng.choice(nb, size=k, replace=False) if branch_rows is None else np.asarray(branch_rows)
        # x <= ub rows carry duals <= 0, x >= lb rows duals >= 0 (min-form)
        vals = rng.uniform(0.0, 3.0, size=k)
        u[n + idx] = np.where(idx < N, -vals, vals)
    u[model.n_base_rows:model.n_base_rows + model.m] = rng.uniform(-10.0, 0.0, size=model.m)
    return u


# ------------------------------------------------------------------- knapsack batches


@dataclass
class KnapsackBatch:
    """m independent 0-1 knapsacks in block layout + one reduced-cost vector (profit = -rc)."""
    m: int
    block_col_start: np.ndarray
    weight: np.ndarray
    capacity: np.ndarray
    red_cost: np.ndarray
    orig_cost: np.ndarray
    alpha: np.ndarray
    name: str = "batch"

    @property
    def cells(self) -> int:
        n = np.diff(self.block_col_start).astype(np.int64)
        return int((n * (self.capacity.astype(np.int64) + 1)).sum())


def csp_batch(m=10_000, n=500, cap=10_000, seed=500, wmax=1000) -> KnapsackBatch:
    """Cutting-stock-style pricing batch (BASELINE.json configs[3]): weakly correlated fp64
    profits p = U(0,1) w (1 +- 0.1), w ~ U{1..wmax}; every item active (rc = -p < 0)."""
    rng = np.random.default_rng(seed)
    w = rng.integers(1, wmax + 1, size=m * n).astype(np.int32)
    p = rng.random(size=m * n) * w * (1.0 + rng.uniform(-0.1, 0.1, size=m * n))
    p = np.maximum(p, 1e-9)
    return KnapsackBatch(m, (np.arange(m + 1, dtype=np.int64) * n).astype(np.int32), w,
                         np.full(m, cap, np.int32), -p, p.copy(), np.zeros(m, np.float64),
                         f"csp_{m}x{n}x{cap}")


def random_batch(m, n_lo, n_hi, cap_lo, cap_hi, seed, frac_active=0.7, int_profits=False) -> KnapsackBatch:
    """Ragged batch for parity tests: block sizes in [n_lo, n_hi] (0 allowed), capacities in
    [cap_lo, cap_hi], some items with rc >= 0 (inactive) and some heavier than the capacity."""
    rng = np.random.default_rng(seed)
    n = rng.integers(n_lo, n_hi + 1, size=m)
    bcs = np.zeros(m + 1, np.int64)
    np.cumsum(n, out=bcs[1:])
    N = int(bcs[-1])
    cap = rng.integers(cap_lo, cap_hi + 1, size=m).astype(np.int32)
    capcol = np.repeat(cap, n)
    w = rng.integers(0, np.maximum(capcol // 2, 2) + 1, size=N)
    heavy = rng.random(N) < 0.03
    w = np.where(heavy, capcol + rng.integers(1, 5, size=N), w).astype(np.int32)
    if int_profits:
        rc = -rng.integers(1, 9, size=N).astype(np.float64)
    else:
        rc = -rng.uniform(0.01, 50.0, size=N)
    inactive = rng.random(N) >= frac_active
    rc = np.where(inactive, rng.uniform(0.0, 5.0, size=N), rc)
    rc[rng.random(N) < 0.01] = 0.0  # exact zeros are NOT active (strict rc < 0)
    oc = rng.integers(1, 100, size=N).astype(np.float64)
    alpha = rng.uniform(-10, 10, size=m)
    return KnapsackBatch(m, bcs.astype(np.int32), w, cap, rc, oc, alpha, f"rand_{m}_{seed}")


# ------------------------------------------------------------------ MILPBlock SpMV


def milpblock_model(blocks=256, cols_per_block=4096, rows=16_384, nnz=10_000_000, seed=256,
                    zero_dual_frac=0.3) -> tuple[PricingModel, np.ndarray]:
    """Block-angular master of BASELINE.json configs[4]: `rows` linking rows over
    N = blocks*cols_per_block columns, nnz placed uniformly, val ~ U[-10,10]; returns the model
    and a dual vector u ~ N(0,1) with 30 % exact zeros.  Blocks carry dummy knapsack data
    (only stage (a) is exercised on this shape)."""
    rng = np.random.default_rng(seed)
    N = blocks * cols_per_block
    r = np.sort(rng.integers(0, rows, size=nnz, dtype=np.int64))
    cnt = np.bincount(r, minlength=rows).astype(np.int64)
    row_start = np.zeros(rows + 1, np.int64)
    np.cumsum(cnt, out=row_start[1:])
    col = rng.integers(0, N, size=nnz, dtype=np.int64).astype(np.int32)
    val = rng.uniform(-10.0, 10.0, size=nnz)
    c = rng.uniform(0.0, 100.0, size=N)
    model = PricingModel(
        R=rows, N=N, row_start=row_start, col_ind=col, val=val, n_base_rows=rows, objective=c,
        m=blocks, block_col_start=(np.arange(blocks + 1, dtype=np.int64) * cols_per_block).astype(np.int32),
        weight=np.ones(N, np.int32), capacity=np.full(blocks, 8, np.int32),
        name=f"milpblock_{blocks}x{cols_per_block}_nnz{nnz}")
    u = rng.standard_normal(model.u_len)
    u[rng.random(model.u_len) < zero_dual_frac] = 0.0
    return model, u

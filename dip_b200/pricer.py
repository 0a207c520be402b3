"""Host-side mirror of DIP's pricing plugin surface over libdipgpu (Python front end used by
the tests and bench.py; the C++ adapter with the same shape is include/dipgpu_decomp.hpp).

Names follow the reference:
  * ``GpuPricer.generateVars(u, phase)``      <-> DecompAlgo::generateVars
    (Dip/src/DecompAlgo.h:430, DecompAlgo.cpp:4622-5260)
  * ``GpuPricer.solveRelaxed(whichBlock, redCostX, target)`` <-> DecompApp::solveRelaxed
    (Dip/src/DecompApp.h:344-349; GAP_DecompApp.cpp:235-285)
  * ``DecompVar`` <-> Dip/src/DecompVar.h:217-239 (sparse 0/1 column, origCost, redCost, blockId)

The compute is entirely in the CUDA library; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import DipGpuError, PHASE_PRICE1, PHASE_PRICE2, PROFIT_FP64, PROFIT_PISINGER_INT  # noqa: F401

DecompSolStatOptimal = 1   # Dip/src/Decomp.h:213-219 (DecompSolverStatus)
DecompSolStatNoSolution = 4


@dataclass
class DecompVar:
    """Dip/src/DecompVar.h:217-239: column in x-space; all elements are 1.0 for knapsack blocks."""
    ind: np.ndarray          # m_s indices (ascending, DecompVar.h:237)
    origCost: float          # m_origCost
    redCost: float           # m_redCost
    blockId: int             # m_blockId
    elements: np.ndarray | None = None  # m_s elements when they are not all 1.0 (integer knapsack blocks)

    @property
    def els(self) -> np.ndarray:
        return np.ones(len(self.ind), np.float64) if self.elements is None else self.elements


class RoundResult:
    """Flat view of one pricing round (struct dipgpu_cols)."""

    def __init__(self, m: int, cap_cols: int, cap_ind: int):
        self.m = m
        self.blockRedCost = np.zeros(m, np.float64)
        self.blockMostNegRC = np.zeros(m, np.float64)
        self.blockOrigCost = np.zeros(m, np.float64)
        self.blockObj = np.zeros(m, np.float64)
        self.blockNnz = np.zeros(m, np.int32)
        self.colBlock = np.zeros(max(cap_cols, 1), np.int32)
        self.colRedCost = np.zeros(max(cap_cols, 1), np.float64)
        self.colOrigCost = np.zeros(max(cap_cols, 1), np.float64)
        self.colStart = np.zeros(max(cap_cols, 1) + 1, np.int64)
        self.colInd = np.zeros(max(cap_ind, 1), np.int32)
        self.colEl = np.zeros(max(cap_ind, 1), np.float64)
        s = capi.Cols()
        s.capBlocks, s.capCols, s.capInd = m, cap_cols, cap_ind
        s.blockRedCost, s.blockMostNegRC = capi.ptr(self.blockRedCost), capi.ptr(self.blockMostNegRC)
        s.blockOrigCost, s.blockObj, s.blockNnz = (capi.ptr(self.blockOrigCost), capi.ptr(self.blockObj),
                                                   capi.ptr(self.blockNnz))
        s.colBlock, s.colRedCost, s.colOrigCost = (capi.ptr(self.colBlock), capi.ptr(self.colRedCost),
                                                   capi.ptr(self.colOrigCost))
        s.colStart, s.colInd = capi.ptr(self.colStart), capi.ptr(self.colInd)
        s.colEl = capi.ptr(self.colEl)
        self.c = s

    @property
    def nCols(self) -> int:
        return self.c.nCols

    @property
    def mostNegReducedCost(self) -> float:
        return self.c.mostNegReducedCost

    @property
    def cells(self) -> int:
        return self.c.cells

    @property
    def cellsEvaluated(self) -> int:
        return self.c.cellsEvaluated

    def column(self, k: int) -> np.ndarray:
        return self.colInd[self.colStart[k]:self.colStart[k + 1]]

    def elements(self, k: int) -> np.ndarray:
        """Elements of kept column k (1.0 except for integer knapsack blocks: the multiplicities)."""
        return self.colEl[self.colStart[k]:self.colStart[k + 1]]

    def vars(self) -> list[DecompVar]:
        return [DecompVar(self.column(k).copy(), float(self.colOrigCost[k]), float(self.colRedCost[k]),
                          int(self.colBlock[k]), self.elements(k).copy()) for k in range(self.nCols)]


class GpuPricer:
    """One libdipgpu context = one GPU driven from one host thread."""

    def __init__(self, device=0):
        """device: one CUDA ordinal (dipgpu_create), or a list of them -- one context over several GPUs of the box
        (dipgpu_create_multi): single rounds are partitioned by blocks, batches / ladders by dual vectors."""
        self._lib = capi.lib()
        self._ctx = C.c_void_p()
        if isinstance(device, (list, tuple, np.ndarray)):
            devs = np.ascontiguousarray(device, np.int32)
            rc = self._lib.dipgpu_create_multi(C.byref(self._ctx), capi.ptr(devs), len(devs))
            self.devices = [int(d) for d in devs]
        else:
            rc = self._lib.dipgpu_create(C.byref(self._ctx), device)
            self.devices = [int(device)]
        if rc != capi.OK:
            raise DipGpuError(rc, f"dipgpu_create(device={device}) failed: no usable sm_100 GPU "
                                  "(libdipgpu has no CPU fallback)")
        self.m = 0
        self.N = 0
        self.R = 0
        self.numConvex = 0
        self.nBaseRows = 0
        self.nBlockCols = 0
        self._keep = []           # arrays that must outlive the C calls
        self._result = None
        self._relax_cache = None  # (fingerprint, RoundResult) for solveRelaxed

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc != capi.OK:
            buf = C.create_string_buffer(512)
            self._lib.dipgpu_last_error(self._ctx, buf, len(buf))
            raise DipGpuError(rc, buf.value.decode(errors="replace"))

    def close(self):
        if self._ctx:
            self._lib.dipgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, name: str, value: int):
        self._check(self._lib.dipgpu_set_option(self._ctx, name.encode(), int(value)))

    def counters(self) -> dict:
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.dipgpu_get_counters(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return dict(kernel_launches=a.value, h2d_bytes=b.value, d2h_bytes=c.value)

    def dpGeometry(self) -> dict:
        out = (C.c_int32 * 7)()
        self._check(self._lib.dipgpu_get_dp_geometry(self._ctx, out))
        return dict(zip(["threads", "cells_per_thread", "items_per_step", "guard", "chunk_cols", "smem_bytes", "fused_tail"],
                        list(out)))

    def last_stage_ms(self) -> dict:
        ms = (C.c_float * 6)()
        self._check(self._lib.dipgpu_last_stage_ms(self._ctx, ms))
        return dict(zip(["h2d", "redcost", "knapsack_dp", "backtrack_emit", "filter", "d2h"], list(ms)))

    # ------------------------------------------------------------ several devices
    def deviceCount(self) -> tuple[int, bool]:
        n, p = C.c_int32(), C.c_int32()
        self._check(self._lib.dipgpu_device_count(self._ctx, C.byref(n), C.byref(p)))
        return n.value, bool(p.value)

    def shardRanges(self) -> list[tuple[int, int]]:
        n = len(self.devices)
        f, c = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self._check(self._lib.dipgpu_shard_ranges(self._ctx, n, capi.ptr(f), capi.ptr(c)))
        return [(int(a), int(b)) for a, b in zip(f, c)]

    def sync(self):
        self._check(self._lib.dipgpu_sync(self._ctx))

    def lastDeviceMs(self) -> tuple[list[float], float]:
        n = len(self.devices)
        ms = np.zeros(n, np.float32)
        mx = C.c_float()
        self._check(self._lib.dipgpu_last_device_ms(self._ctx, n, capi.ptr(ms), C.byref(mx)))
        return [float(x) for x in ms], float(mx.value)

    def resultShards(self) -> tuple[np.ndarray, np.ndarray]:
        n = C.c_int32()
        cap = max(len(self.devices), 1)
        off, siz = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
        self._check(self._lib.dipgpu_result_shards(self._ctx, cap, C.byref(n), capi.ptr(off), capi.ptr(siz)))
        return off[:n.value], siz[:n.value]

    # --------------------------------------------------------------------- model
    def setModelCore(self, R, N, row_start, col_ind, val, n_base_rows, num_convex):
        """DecompConstraintSet::M, row-ordered (Dip/src/DecompConstraintSet.h:32, .cpp:41-43)."""
        rs = np.ascontiguousarray(row_start, np.int64)
        ci = np.ascontiguousarray(col_ind, np.int32)
        v = np.ascontiguousarray(val, np.float64)
        self._check(self._lib.dipgpu_set_core_csr(self._ctx, R, N, capi.ptr(rs), capi.ptr(ci), capi.ptr(v),
                                                  n_base_rows, num_convex))
        self.R, self.N, self.nBaseRows, self.numConvex = R, N, n_base_rows, num_convex

    def appendCoreRows(self, row_start, col_ind, val):
        """DecompConstraintSet::appendRow (Dip/src/DecompConstraintSet.h:143-163) for cut rows."""
        rs = np.ascontiguousarray(row_start, np.int64)
        ci = np.ascontiguousarray(col_ind, np.int32)
        v = np.ascontiguousarray(val, np.float64)
        self._check(self._lib.dipgpu_append_core_rows(self._ctx, len(rs) - 1, capi.ptr(rs), capi.ptr(ci),
                                                      capi.ptr(v)))
        self.R += len(rs) - 1

    def setObjective(self, c):
        c = np.ascontiguousarray(c, np.float64)
        self._check(self._lib.dipgpu_set_objective(self._ctx, capi.ptr(c), len(c)))
        self.N = max(self.N, len(c))

    def setKnapsackBlocks(self, block_col_start, weight, capacity, profit_mode=PROFIT_FP64):
        """One knapsack per block (GAP_DecompApp.cpp:52-58, GAP_KnapPisinger.h:276-294)."""
        bcs = np.ascontiguousarray(block_col_start, np.int32)
        w = np.ascontiguousarray(weight, np.int32)
        cap = np.ascontiguousarray(capacity, np.int32)
        m = len(bcs) - 1
        self._check(self._lib.dipgpu_set_knapsack_blocks(self._ctx, m, capi.ptr(bcs), capi.ptr(w),
                                                         capi.ptr(cap), profit_mode))
        self.m = m
        self.nBlockCols = int(bcs[-1])
        self.N = max(self.N, self.nBlockCols)
        self._result = None
        self._relax_cache = None

    def setMckpBlocks(self, block_group_start, group_col_start, weight, capacity):
        """Multiple-choice knapsack blocks (MMKP MCKP0 relaxation, MMKP_MCKnap.cpp:172-383): exactly one
        column per group, weights <= capacity."""
        bgs = np.ascontiguousarray(block_group_start, np.int32)
        gcs = np.ascontiguousarray(group_col_start, np.int32)
        w = np.ascontiguousarray(weight, np.int32)
        cap = np.ascontiguousarray(capacity, np.int32)
        m = len(bgs) - 1
        self._check(self._lib.dipgpu_set_mckp_blocks(self._ctx, m, capi.ptr(bgs), capi.ptr(gcs), capi.ptr(w), capi.ptr(cap)))
        self.m = m
        self.nBlockCols = int(gcs[-1])
        self.N = max(self.N, self.nBlockCols)
        self._result = None
        self._relax_cache = None

    def setIntKnapBlocks(self, block_col_start, weight, capacity, use_col=None):
        """Integer (unbounded) knapsack blocks with an optional use column per block: the cutting-stock pricing
        problem (kp / relaxed_solver of Dip/src/dippy/tests/test_cutting_stock.py:98-179)."""
        bcs = np.ascontiguousarray(block_col_start, np.int32)
        w = np.ascontiguousarray(weight, np.int32)
        cap = np.ascontiguousarray(capacity, np.int32)
        use = None if use_col is None else np.ascontiguousarray(use_col, np.int32)
        m = len(bcs) - 1
        self._check(self._lib.dipgpu_set_intknap_blocks(self._ctx, m, capi.ptr(bcs), capi.ptr(w), capi.ptr(cap),
                                                        None if use is None else capi.ptr(use)))
        self.m = m
        self.nBlockCols = max(int(bcs[-1]), int(use.max()) + 1 if use is not None and len(use) else 0)
        self.N = max(self.N, self.nBlockCols)
        self._result = None
        self._relax_cache = None

    def setModel(self, model, profit_mode=PROFIT_FP64):
        """Upload a dip_b200.instances.PricingModel."""
        self.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows,
                          model.m)
        self.setObjective(model.objective)
        self.setKnapsackBlocks(model.block_col_start, model.weight, model.capacity, profit_mode)

    def setColBounds(self, col_lb=None, col_ub=None):
        """DecompAlgo::setSubProbBounds (DecompAlgo.cpp:2352-2371) under BranchEnforceInSubProb: node
        bounds of the x-space columns, enforced inside the block knapsacks.  None, None removes them."""
        if col_lb is None and col_ub is None:
            self._check(self._lib.dipgpu_set_col_bounds(self._ctx, None, None, 0))
            return
        lb = np.ascontiguousarray(col_lb, np.float64)
        ub = np.ascontiguousarray(col_ub, np.float64)
        self._check(self._lib.dipgpu_set_col_bounds(self._ctx, capi.ptr(lb), capi.ptr(ub), len(lb)))
        self._relax_cache = None

    def setDualSupport(self, idx=None):
        """Rows of the master dual vector that can be nonzero (free branch rows have zero duals,
        AlpsDecompTreeNode.cpp:215-238): only those are uploaded per round.  None removes the support."""
        if idx is None:
            self._check(self._lib.dipgpu_set_dual_support(self._ctx, 0, None))
            return
        ix = np.ascontiguousarray(idx, np.int32)
        self._check(self._lib.dipgpu_set_dual_support(self._ctx, len(ix), capi.ptr(ix)))

    def setBlockRange(self, first: int, count: int):
        self._check(self._lib.dipgpu_set_block_range(self._ctx, first, count))

    def _res(self) -> RoundResult:
        if self._result is None:
            self._result = RoundResult(self.m, self.m, max(self.nBlockCols + self.m, 1))
        return self._result

    # ------------------------------------------------------------------- pricing
    def calcRedCost(self, u, phase=PHASE_PRICE2) -> np.ndarray:
        """generateVarsAdjustDuals + generateVarsCalcRedCost (DecompAlgo.cpp:4506-4619)."""
        u = np.ascontiguousarray(u, np.float64)
        out = np.empty(self.N, np.float64)
        self._check(self._lib.dipgpu_calc_redcost(self._ctx, capi.ptr(u), len(u), phase, capi.ptr(out)))
        return out

    def solveBlocks(self, redCostX, alpha, redCostEps=1e-4, result: RoundResult | None = None) -> RoundResult:
        """The loop of DecompApp::solveRelaxed over all blocks + the filter of
        DecompAlgo.cpp:5151-5232, for a caller-supplied reduced-cost vector."""
        rc = np.ascontiguousarray(redCostX, np.float64)
        al = np.ascontiguousarray(alpha, np.float64)
        if len(rc) < self.nBlockCols or len(al) < self.m:
            raise ValueError("redCostX / alpha shorter than the block structure")
        res = result or self._res()
        self._check(self._lib.dipgpu_solve_blocks(self._ctx, capi.ptr(rc), capi.ptr(al), float(redCostEps),
                                                  C.byref(res.c)))
        return res

    def priceRound(self, u, phase=PHASE_PRICE2, redCostEps=1e-4, want_redcost=False,
                   result: RoundResult | None = None):
        """One whole pricing round in one submission; returns RoundResult (and redCostX)."""
        u = np.ascontiguousarray(u, np.float64)
        res = result or self._res()
        rcx = np.empty(self.N, np.float64) if want_redcost else None
        self._check(self._lib.dipgpu_price_round(self._ctx, capi.ptr(u), len(u), phase, float(redCostEps),
                                                 C.byref(res.c), capi.ptr(rcx)))
        return (res, rcx) if want_redcost else res

    # ---- several dual vectors in one submission ------------------------------
    def _batch_results(self, n: int):
        res = [RoundResult(self.m, self.m, max(self.nBlockCols + self.m, 1)) for _ in range(n)]
        arr = (capi.Cols * n)()
        for k in range(n):
            arr[k] = res[k].c
        return res, arr

    @staticmethod
    def _batch_finish(res, arr):
        for k, r in enumerate(res):
            r.c = arr[k]      # counts / sums written by the library; the array pointers are unchanged
        return res

    def priceRoundsBatch(self, U, phase=PHASE_PRICE2, redCostEps=1e-4) -> list[RoundResult]:
        """nVec dual vectors (rows of U), each priced as priceRound would (dipgpu_price_rounds_batch)."""
        U = np.ascontiguousarray(U, np.float64)
        n, u_len = U.shape
        res, arr = self._batch_results(n)
        self._check(self._lib.dipgpu_price_rounds_batch(self._ctx, n, capi.ptr(U), u_len, phase, float(redCostEps), arr))
        return self._batch_finish(res, arr)

    def priceRoundsBatchRaw(self, n: int, u_ptr: int, u_len: int, phase: int, eps: float, arr):
        rc = self._lib.dipgpu_price_rounds_batch(self._ctx, n, u_ptr, u_len, phase, eps, arr)
        if rc != capi.OK:
            self._check(rc)

    def priceRoundWhich(self, u, which, userDuals=None, phase=PHASE_PRICE2, redCostEps=1e-4, result=None):
        """generateVars' round-robin branch (Dip/src/DecompAlgo.cpp:4929-5149): only the blocks DecompApp::
        solveRelaxedWhich names are solved, `userDuals` = {block: dual vector} for the blocks priced with their
        own master duals (userDualsByBlock).  Returns (RoundResult, solvedAll): solvedAll is True when no listed
        block priced out and the full round at `u` was returned instead (DecompAlgo.cpp:5076-5148)."""
        u = np.ascontiguousarray(u, np.float64)
        which = np.ascontiguousarray(which, np.int32)
        keep = []                                   # the vectors must outlive the call
        ptrs = None
        if userDuals:
            ptrs = (C.c_void_p * max(len(which), 1))()
            for i, b in enumerate(which):
                ub = userDuals.get(int(b))
                if ub is not None:
                    ub = np.ascontiguousarray(ub, np.float64)
                    if len(ub) != len(u):
                        raise ValueError("The size of the user dual vector is not the same as the number of master rows")
                    keep.append(ub)
                    ptrs[i] = ub.ctypes.data
        res = result if result is not None else RoundResult(self.m, self.m, max(self.nBlockCols + self.m, 1))
        all_ = C.c_int32(0)
        self._check(self._lib.dipgpu_price_round_which(self._ctx, capi.ptr(u), len(u), phase, float(redCostEps), len(which),
                                                       capi.ptr(which), ptrs, C.byref(res.c), C.byref(all_)))
        return res, bool(all_.value)

    def selectBatchResult(self, which: int):
        self._check(self._lib.dipgpu_select_batch_result(self._ctx, int(which)))

    # ---- dual stabilisation (DecompAlgoPC::adjustMasterDualSolution) ------------
    def stabSetCenter(self, dual):
        """m_dual := dual (first price call / first phase-2 call, DecompAlgoPC.cpp:163-177)."""
        d = np.ascontiguousarray(dual, np.float64)
        self._check(self._lib.dipgpu_stab_set_center(self._ctx, capi.ptr(d), len(d)))

    def stabAccept(self, step: int = 0):
        """m_dual := m_dualST (DecompAlgoPC::setObjBound, DecompAlgoPC.h:124-133)."""
        self._check(self._lib.dipgpu_stab_accept(self._ctx, int(step)))

    def priceRoundStab(self, uRM, alpha, phase=PHASE_PRICE2, redCostEps=1e-4, want_duals=True):
        """adjustMasterDualSolution + generateVars with the smoothed duals; returns (RoundResult, dualST)."""
        u = np.ascontiguousarray(uRM, np.float64)
        res = RoundResult(self.m, self.m, max(self.nBlockCols + self.m, 1))
        st = np.empty(len(u), np.float64) if want_duals else None
        self._check(self._lib.dipgpu_price_round_stab(self._ctx, capi.ptr(u), len(u), float(alpha), phase,
                                                      float(redCostEps), C.byref(res.c), capi.ptr(st)))
        return res, st

    def priceRoundStabLadder(self, uRM, alpha, n_steps, shrink=0.90, phase=PHASE_PRICE2, redCostEps=1e-4):
        """The smoothing ladder alpha, shrink*alpha, ... (DecompAlgo.cpp:5642-5655) priced speculatively in
        one submission; returns ([RoundResult per step], alphas)."""
        u = np.ascontiguousarray(uRM, np.float64)
        res, arr = self._batch_results(n_steps)
        alphas = np.zeros(n_steps, np.float64)
        self._check(self._lib.dipgpu_price_round_stab_ladder(self._ctx, capi.ptr(u), len(u), float(alpha), float(shrink),
                                                             n_steps, phase, float(redCostEps), arr, capi.ptr(alphas)))
        return self._batch_finish(res, arr), alphas

    def priceRoundStabLadderRaw(self, u_ptr: int, u_len: int, alpha: float, shrink: float, n_steps: int, phase: int,
                                eps: float, arr):
        rc = self._lib.dipgpu_price_round_stab_ladder(self._ctx, u_ptr, u_len, alpha, shrink, n_steps, phase, eps, arr, None)
        if rc != capi.OK:
            self._check(rc)

    def smoothedDuals(self, step: int = 0) -> np.ndarray:
        out = np.empty(self.R + self.numConvex, np.float64)
        self._check(self._lib.dipgpu_get_smoothed_duals(self._ctx, int(step), capi.ptr(out), len(out)))
        return out

    # ---- waiting-column pool (DecompVarPool) -----------------------------------
    def poolClear(self):
        self._check(self._lib.dipgpu_pool_clear(self._ctx))

    def poolSize(self) -> tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        self._check(self._lib.dipgpu_pool_size(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def poolAppend(self, col_start, row_ind, val, orig_cost):
        cs = np.ascontiguousarray(col_start, np.int64)
        ri = np.ascontiguousarray(row_ind, np.int32)
        v = np.ascontiguousarray(val, np.float64)
        oc = np.ascontiguousarray(orig_cost, np.float64)
        self._check(self._lib.dipgpu_pool_append(self._ctx, len(cs) - 1, capi.ptr(cs), capi.ptr(ri), capi.ptr(v), capi.ptr(oc)))

    def poolAppendExpanded(self, which):
        """Columns `which` of the last expandColumns() enter the pool device-to-device."""
        w = np.ascontiguousarray(which, np.int32)
        self._check(self._lib.dipgpu_pool_append_expanded(self._ctx, len(w), capi.ptr(w)))

    def poolKeep(self, keep_idx):
        k = np.ascontiguousarray(keep_idx, np.int32)
        self._check(self._lib.dipgpu_pool_keep(self._ctx, len(k), capi.ptr(k)))

    def poolSetReducedCosts(self, u=None, feasible=True):
        """DecompVarPool::setReducedCosts(u, stat): returns (redCost per waiting column, found_negrc_var).
        u=None re-uses the dual vector already on the device."""
        n, _ = self.poolSize()
        out = np.zeros(max(n, 1), np.float64)
        found = C.c_int32(0)
        if u is not None:
            u = np.ascontiguousarray(u, np.float64)
        u_len = len(u) if u is not None else self.R + self.numConvex
        self._check(self._lib.dipgpu_pool_set_reduced_costs(self._ctx, capi.ptr(u), u_len, int(bool(feasible)),
                                                            capi.ptr(out), C.byref(found)))
        return out[:n], bool(found.value)

    def priceRoundResident(self, phase=PHASE_PRICE2, redCostEps=1e-4, result: RoundResult | None = None):
        """priceRound with the dual vector already on the device (u == NULL)."""
        res = result or self._res()
        self._check(self._lib.dipgpu_price_round(self._ctx, None, self.R + self.numConvex, phase, float(redCostEps),
                                                 C.byref(res.c), None))
        return res

    def priceRoundRaw(self, u_ptr: int, u_len: int, phase: int, eps: float, res: RoundResult):
        """Same call with a raw host pointer (pinned buffers in bench.py): no numpy marshalling."""
        rc = self._lib.dipgpu_price_round(self._ctx, u_ptr, u_len, phase, eps, C.byref(res.c), None)
        if rc != capi.OK:
            self._check(rc)

    def expandColumns(self, tolZero=1e-6):
        """Master columns of the columns kept by the last round = the per-variable work of
        DecompAlgo::addVarsToPool (DecompAlgo.cpp:5452-5558).  Returns (colStart, rowInd, val, hash):
        column k = kept column k; rowInd ascending master rows; hash = 64-bit fingerprint of
        (block, x-space indices) for duplicate detection."""
        cap_cols = max(self.m, 1)
        cap_nnz = 1 << 16
        while True:
            col_start = np.zeros(cap_cols + 1, np.int64)
            row_ind = np.zeros(cap_nnz, np.int32)
            val = np.zeros(cap_nnz, np.float64)
            hsh = np.zeros(cap_cols, np.uint64)
            mc = capi.MCols()
            mc.capCols, mc.capNnz = cap_cols, cap_nnz
            mc.colStart, mc.rowInd, mc.val, mc.hash = (capi.ptr(col_start), capi.ptr(row_ind), capi.ptr(val),
                                                       capi.ptr(hsh))
            rc = self._lib.dipgpu_expand_columns(self._ctx, float(tolZero), C.byref(mc))
            if rc == capi.ERR_OVERFLOW and mc.nNnz > cap_nnz:
                cap_nnz = int(mc.nNnz)
                continue
            self._check(rc)
            n = mc.nCols
            return col_start[:n + 1], row_ind[:mc.nNnz], val[:mc.nNnz], hsh[:n]

    def priceRoundExpand(self, u, phase=PHASE_PRICE2, redCostEps=1e-4, tolZero=1e-6, result=None, cap_nnz=1 << 16):
        """priceRound + expandColumns in one submission (dipgpu_price_round_expand): returns
        (RoundResult, (colStart, rowInd, val, hash))."""
        u = np.ascontiguousarray(u, np.float64)
        res = result if result is not None else RoundResult(self.m, self.m, max(self.nBlockCols + self.m, 1))
        cap_cols = max(self.m, 1)
        first = True
        while True:
            col_start = np.zeros(cap_cols + 1, np.int64)
            row_ind = np.zeros(cap_nnz, np.int32)
            val = np.zeros(cap_nnz, np.float64)
            hsh = np.zeros(cap_cols, np.uint64)
            mc = capi.MCols()
            mc.capCols, mc.capNnz = cap_cols, cap_nnz
            mc.colStart, mc.rowInd, mc.val, mc.hash = (capi.ptr(col_start), capi.ptr(row_ind), capi.ptr(val),
                                                       capi.ptr(hsh))
            if first:
                rc = self._lib.dipgpu_price_round_expand(self._ctx, capi.ptr(u), len(u), phase, float(redCostEps),
                                                         float(tolZero), C.byref(res.c), C.byref(mc))
            else:        # only the output arrays were too small: the round itself is done
                rc = self._lib.dipgpu_expand_columns(self._ctx, float(tolZero), C.byref(mc))
            if rc == capi.ERR_OVERFLOW and mc.nNnz > cap_nnz:
                cap_nnz, first = int(mc.nNnz), False
                continue
            self._check(rc)
            n = mc.nCols
            return res, (col_start[:n + 1], row_ind[:mc.nNnz], val[:mc.nNnz], hsh[:n])

    # ---- the reference's two plugin surfaces ---------------------------------
    def generateVars(self, u, phase=PHASE_PRICE2, redCostEps=1e-4):
        """DecompAlgo::generateVars(newVars, mostNegReducedCost): returns (newVars, mostNegRC)."""
        res = self.priceRound(u, phase, redCostEps)
        return res.vars(), res.mostNegReducedCost

    def solveRelaxed(self, whichBlock: int, redCostX, target: float):
        """DecompApp::solveRelaxed(whichBlock, redCostX, target, varList): returns
        (DecompSolStatOptimal, [DecompVar]).  As in GAP_DecompApp.cpp:278-284 exactly one
        (possibly empty) column comes back per block, its redCost EXCLUDING the convexity dual
        (the caller subtracts it, DecompAlgo.cpp:6350-6356).  All blocks of a round are solved in
        one GPU submission the first time a new redCostX is seen; later calls are served from it."""
        rc = np.ascontiguousarray(redCostX, np.float64)
        fp = (rc.ctypes.data, hash(rc.tobytes()))
        if self._relax_cache is None or self._relax_cache[0] != fp:
            res = RoundResult(self.m, self.m, max(self.nBlockCols + self.m, 1))
            self.solveBlocks(rc, np.zeros(self.m), redCostEps=-np.inf, result=res)
            self._relax_cache = (fp, res)
        res = self._relax_cache[1]
        # every block is "kept" with eps = -inf, in block order
        k = int(whichBlock)
        assert res.colBlock[k] == k
        var = DecompVar(res.column(k).copy(), float(res.colOrigCost[k]), float(res.colRedCost[k]), k)
        return DecompSolStatOptimal, [var]

    # ---- device-resident variants (multi-GPU driver) ---------------------------
    def priceRoundDev(self, u_dev_ptr: int, u_len: int, phase=PHASE_PRICE2, redCostEps=1e-4, stream: int = 0):
        self._check(self._lib.dipgpu_price_round_dev(self._ctx, u_dev_ptr, u_len, phase, float(redCostEps),
                                                     stream or None))

    def calcRedCostDev(self, u_dev_ptr: int, u_len: int, phase=PHASE_PRICE2, stream: int = 0):
        self._check(self._lib.dipgpu_calc_redcost_dev(self._ctx, u_dev_ptr, u_len, phase, stream or None))

    def measureSmemGBs(self) -> float:
        g = C.c_double()
        self._check(self._lib.dipgpu_measure_smem_gbs(self._ctx, C.byref(g)))
        return g.value

    def solveBlocksDev(self, rc_dev_ptr: int, alpha_dev_ptr: int, redCostEps=1e-4, stream: int = 0):
        self._check(self._lib.dipgpu_solve_blocks_dev(self._ctx, rc_dev_ptr, alpha_dev_ptr, float(redCostEps),
                                                      stream or None))

    def resultDev(self) -> tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._lib.dipgpu_result_dev(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def redCostDev(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.dipgpu_redcost_dev(self._ctx, C.byref(p)))
        return p.value


def unpack_results(packed: np.ndarray, offsets, sizes, m_total: int, cap_ind: int, result: RoundResult | None = None) -> RoundResult:
    """dipgpu_unpack_results on a host copy of a multi-device context's gather buffer: the shards' packed results
    decoded into ONE round, in block order."""
    packed = np.ascontiguousarray(packed, np.uint8)
    off = np.ascontiguousarray(offsets, np.uint64)
    siz = np.ascontiguousarray(sizes, np.uint64)
    res = result or RoundResult(m_total, m_total, cap_ind)
    rc = capi.lib().dipgpu_unpack_results(capi.ptr(packed), len(off), capi.ptr(off), capi.ptr(siz), C.byref(res.c))
    if rc != capi.OK:
        raise DipGpuError(rc, "dipgpu_unpack_results failed")
    return res


def unpack_result(packed: np.ndarray, m_total: int, cap_ind: int, result: RoundResult | None = None) -> RoundResult:
    """dipgpu_unpack_result on a host copy of a shard's packed result (block ids are global).
    Block arrays of `result` are filled at the shard's global positions; kept columns are
    those of this shard only."""
    packed = np.ascontiguousarray(packed, np.uint8)
    res = result or RoundResult(m_total, m_total, cap_ind)
    rc = capi.lib().dipgpu_unpack_result(capi.ptr(packed), packed.nbytes, C.byref(res.c))
    if rc != capi.OK:
        raise DipGpuError(rc, "dipgpu_unpack_result failed")
    return res

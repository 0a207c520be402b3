"""ctypes binding of libdipgpu.so (include/dipgpu.h) -- the only way Python reaches the GPU path.

There is no CPU fallback: importing this module without the built library raises,
and every compute entry point raises ``DipGpuError`` when no sm_100 device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdipgpu.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_UNSUPPORTED, ERR_OVERFLOW = range(7)
PROFIT_FP64, PROFIT_PISINGER_INT = 0, 1
PHASE_PRICE1, PHASE_PRICE2 = 1, 2

_ERR_NAMES = {
    ERR_INVALID: "DIPGPU_ERR_INVALID", ERR_CUDA: "DIPGPU_ERR_CUDA", ERR_NOMEM: "DIPGPU_ERR_NOMEM",
    ERR_STATE: "DIPGPU_ERR_STATE", ERR_UNSUPPORTED: "DIPGPU_ERR_UNSUPPORTED",
    ERR_OVERFLOW: "DIPGPU_ERR_OVERFLOW",
}

# every symbol include/dipgpu.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "dipgpu_abi_version", "dipgpu_create", "dipgpu_destroy", "dipgpu_last_error",
    "dipgpu_host_alloc", "dipgpu_host_free", "dipgpu_set_core_csr", "dipgpu_append_core_rows",
    "dipgpu_set_objective", "dipgpu_set_knapsack_blocks", "dipgpu_calc_redcost",
    "dipgpu_solve_blocks", "dipgpu_price_round", "dipgpu_set_block_range",
    "dipgpu_price_round_dev", "dipgpu_solve_blocks_dev", "dipgpu_result_dev",
    "dipgpu_result_max_bytes", "dipgpu_unpack_result", "dipgpu_redcost_dev",
    "dipgpu_get_counters", "dipgpu_last_stage_ms", "dipgpu_set_option",
    "dipgpu_calc_redcost_dev", "dipgpu_measure_smem_gbs", "dipgpu_plan_dp_geometry",
    "dipgpu_get_dp_geometry", "dipgpu_expand_columns",
    "dipgpu_price_rounds_batch", "dipgpu_select_batch_result", "dipgpu_stab_set_center", "dipgpu_stab_accept",
    "dipgpu_price_round_stab", "dipgpu_price_round_stab_ladder", "dipgpu_get_smoothed_duals",
    "dipgpu_pool_clear", "dipgpu_pool_size", "dipgpu_pool_append", "dipgpu_pool_append_expanded",
    "dipgpu_pool_keep", "dipgpu_pool_set_reduced_costs", "dipgpu_set_col_bounds", "dipgpu_set_dual_support", "dipgpu_set_mckp_blocks",
    "dipgpu_set_intknap_blocks",
    "dipgpu_price_round_which", "dipgpu_price_round_expand",
    "dipgpu_create_multi", "dipgpu_device_count", "dipgpu_shard_ranges", "dipgpu_sync", "dipgpu_result_shards",
    "dipgpu_unpack_results", "dipgpu_last_device_ms",
]


class DipGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Cols(C.Structure):
    """struct dipgpu_cols (include/dipgpu.h)."""
    _fields_ = [
        ("capBlocks", C.c_int32), ("capCols", C.c_int32), ("capInd", C.c_int64),
        ("blockRedCost", C.c_void_p), ("blockMostNegRC", C.c_void_p), ("blockOrigCost", C.c_void_p),
        ("blockObj", C.c_void_p), ("blockNnz", C.c_void_p),
        ("nCols", C.c_int32), ("colBlock", C.c_void_p), ("colRedCost", C.c_void_p),
        ("colOrigCost", C.c_void_p), ("colStart", C.c_void_p), ("colInd", C.c_void_p),
        ("nInd", C.c_int64), ("mostNegReducedCost", C.c_double), ("cells", C.c_int64),
        ("cellsEvaluated", C.c_int64), ("colEl", C.c_void_p),
    ]


class MCols(C.Structure):
    """struct dipgpu_mcols (include/dipgpu.h)."""
    _fields_ = [
        ("capCols", C.c_int32), ("capNnz", C.c_int64), ("nCols", C.c_int32), ("nNnz", C.c_int64),
        ("colStart", C.c_void_p), ("rowInd", C.c_void_p), ("val", C.c_void_p), ("hash", C.c_void_p),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dip_b200._build` "
            "(the pricing path is CUDA-only; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "dipgpu_abi_version": ([], C.c_int),
        "dipgpu_create": ([C.POINTER(vp), C.c_int], C.c_int),
        "dipgpu_destroy": ([vp], C.c_int),
        "dipgpu_last_error": ([vp, C.c_char_p, C.c_size_t], C.c_int),
        "dipgpu_host_alloc": ([C.POINTER(vp), C.c_size_t], C.c_int),
        "dipgpu_host_free": ([vp], C.c_int),
        "dipgpu_set_core_csr": ([vp, i32, i32, vp, vp, vp, i32, i32], C.c_int),
        "dipgpu_append_core_rows": ([vp, i32, vp, vp, vp], C.c_int),
        "dipgpu_set_objective": ([vp, vp, i32], C.c_int),
        "dipgpu_set_knapsack_blocks": ([vp, i32, vp, vp, vp, i32], C.c_int),
        "dipgpu_calc_redcost": ([vp, vp, i32, i32, vp], C.c_int),
        "dipgpu_solve_blocks": ([vp, vp, vp, dbl, C.POINTER(Cols)], C.c_int),
        "dipgpu_price_round": ([vp, vp, i32, i32, dbl, C.POINTER(Cols), vp], C.c_int),
        "dipgpu_set_block_range": ([vp, i32, i32], C.c_int),
        "dipgpu_price_round_dev": ([vp, vp, i32, i32, dbl, vp], C.c_int),
        "dipgpu_calc_redcost_dev": ([vp, vp, i32, i32, vp], C.c_int),
        "dipgpu_measure_smem_gbs": ([vp, C.POINTER(dbl)], C.c_int),
        "dipgpu_plan_dp_geometry": ([i32, i32, i32, i32, i32, C.POINTER(i32)], C.c_int),
        "dipgpu_get_dp_geometry": ([vp, C.POINTER(i32)], C.c_int),
        "dipgpu_expand_columns": ([vp, dbl, C.POINTER(MCols)], C.c_int),
        "dipgpu_solve_blocks_dev": ([vp, vp, vp, dbl, vp], C.c_int),
        "dipgpu_result_dev": ([vp, C.POINTER(vp), C.POINTER(C.c_size_t)], C.c_int),
        "dipgpu_result_max_bytes": ([vp, C.POINTER(C.c_size_t)], C.c_int),
        "dipgpu_unpack_result": ([vp, C.c_size_t, C.POINTER(Cols)], C.c_int),
        "dipgpu_redcost_dev": ([vp, C.POINTER(vp)], C.c_int),
        "dipgpu_get_counters": ([vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "dipgpu_last_stage_ms": ([vp, C.POINTER(C.c_float)], C.c_int),
        "dipgpu_set_option": ([vp, C.c_char_p, i64], C.c_int),
        "dipgpu_price_rounds_batch": ([vp, i32, vp, i32, i32, dbl, C.POINTER(Cols)], C.c_int),
        "dipgpu_select_batch_result": ([vp, i32], C.c_int),
        "dipgpu_price_round_expand": ([vp, vp, i32, i32, dbl, dbl, C.POINTER(Cols), C.POINTER(MCols)], C.c_int),
        "dipgpu_price_round_which": ([vp, vp, i32, i32, dbl, i32, vp, vp, C.POINTER(Cols), C.POINTER(C.c_int32)], C.c_int),
        "dipgpu_stab_set_center": ([vp, vp, i32], C.c_int),
        "dipgpu_stab_accept": ([vp, i32], C.c_int),
        "dipgpu_price_round_stab": ([vp, vp, i32, dbl, i32, dbl, C.POINTER(Cols), vp], C.c_int),
        "dipgpu_price_round_stab_ladder": ([vp, vp, i32, dbl, dbl, i32, i32, dbl, C.POINTER(Cols), vp], C.c_int),
        "dipgpu_get_smoothed_duals": ([vp, i32, vp, i32], C.c_int),
        "dipgpu_pool_clear": ([vp], C.c_int),
        "dipgpu_pool_size": ([vp, C.POINTER(i64), C.POINTER(i64)], C.c_int),
        "dipgpu_pool_append": ([vp, i32, vp, vp, vp, vp], C.c_int),
        "dipgpu_pool_append_expanded": ([vp, i32, vp], C.c_int),
        "dipgpu_pool_keep": ([vp, i32, vp], C.c_int),
        "dipgpu_pool_set_reduced_costs": ([vp, vp, i32, i32, vp, C.POINTER(i32)], C.c_int),
        "dipgpu_set_col_bounds": ([vp, vp, vp, i32], C.c_int),
        "dipgpu_set_dual_support": ([vp, i32, vp], C.c_int),
        "dipgpu_set_mckp_blocks": ([vp, i32, vp, vp, vp, vp], C.c_int),
        "dipgpu_set_intknap_blocks": ([vp, i32, vp, vp, vp, vp], C.c_int),
        "dipgpu_create_multi": ([C.POINTER(vp), vp, i32], C.c_int),
        "dipgpu_device_count": ([vp, C.POINTER(i32), C.POINTER(i32)], C.c_int),
        "dipgpu_shard_ranges": ([vp, i32, vp, vp], C.c_int),
        "dipgpu_sync": ([vp], C.c_int),
        "dipgpu_result_shards": ([vp, i32, C.POINTER(i32), vp, vp], C.c_int),
        "dipgpu_unpack_results": ([vp, i32, vp, vp, C.POINTER(Cols)], C.c_int),
        "dipgpu_last_device_ms": ([vp, i32, vp, C.POINTER(C.c_float)], C.c_int),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data


class PinnedArray:
    """A numpy view over page-locked host memory from dipgpu_host_alloc."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        self._ptr = C.c_void_p()
        rc = lib().dipgpu_host_alloc(C.byref(self._ptr), max(n * self.dtype.itemsize, 1))
        if rc != OK:
            raise DipGpuError(rc, "dipgpu_host_alloc failed")
        buf = (C.c_char * max(n * self.dtype.itemsize, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=n).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            lib().dipgpu_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def plan_dp_geometry(n_blocks, max_capacity, max_block_cols, max_usable_weight, sm_count=148) -> dict:
    """Host-only: the DP kernel shape the library picks (dipgpu_plan_dp_geometry)."""
    out = (C.c_int32 * 7)()
    rc = lib().dipgpu_plan_dp_geometry(sm_count, n_blocks, max_capacity, max_block_cols, max_usable_weight, out)
    if rc != OK:
        raise DipGpuError(rc, "dipgpu_plan_dp_geometry")
    return dict(zip(["threads", "cells_per_thread", "items_per_step", "guard", "chunk_cols", "smem_bytes", "fused_tail"], list(out)))

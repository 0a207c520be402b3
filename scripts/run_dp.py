#!/usr/bin/env python3
"""Run the batched knapsack DP a few times on one workload with forced geometry options
(for ncu):  python scripts/run_dp.py <workload> [S] [items] [noguard] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dip_b200.pricer import GpuPricer  # noqa: E402
from scripts.dp_sweep import batch  # noqa: E402

name = sys.argv[1]
S = int(sys.argv[2]) if len(sys.argv) > 2 else 0
items = int(sys.argv[3]) if len(sys.argv) > 3 else 0
noguard = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
b = batch(name)
with GpuPricer(0) as p:
    p.setObjective(b.orig_cost)
    p.set_option("dp_cells_per_thread", S)
    p.set_option("dp_items_per_step", items)
    p.set_option("dp_no_guard", noguard)
    p.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    p.set_option("stage_timing", 1)
    for _ in range(reps):
        res = p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
        print(p.last_stage_ms(), res.cells)
    if os.environ.get("DP_TIMES"):
        p.set_option("dp_debug_times", 1)
        res = p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
        res = p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
        tt = np.zeros((b.m, 8), np.uint64)
        p.set_option("dp_debug_times", tt.ctypes.data)
        t0 = tt[:, 0].min()
        rel = (tt.astype(np.int64) - np.int64(t0)) / 1e3
        names = ["start", "staged", "compacted", "dp_done", "backtracked", "filtered", "polled", "prefixed"]
        print(p.dpGeometry())
        for k, nm in enumerate(names):
            col = rel[:, k]
            print(f"{nm:12s} min {col.min():8.2f}  mean {col.mean():8.2f}  max {col.max():8.2f} us")
        if os.environ.get("DP_TIMES") == "2":
            for blk in range(b.m):
                print(blk, " ".join(f"{rel[blk, k]:7.2f}" for k in (3, 7, 4, 6, 5)))
        print("last CTA (max filtered) block", int(rel[:, 5].argmax()), "stamps", rel[int(rel[:, 5].argmax()), :6])

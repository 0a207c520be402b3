import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
from dip_b200.pricer import GpuPricer
from scripts.dp_sweep import batch
b = batch("gape")
for T in (0, 256, 384, 512, 768):
    with GpuPricer(0) as p:
        p.setObjective(b.orig_cost)
        if T: p.set_option("dp_threads", T)
        p.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
        p.set_option("stage_timing", 1)
        ts = []
        for _ in range(12):
            res = p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
            ts.append(p.last_stage_ms()["knapsack_dp"])
        p.set_option("dp_debug_times", 1)
        p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf); p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
        tt = np.zeros((b.m, 8), np.uint64)
        p.set_option("dp_debug_times", tt.ctypes.data)
        rel = (tt.astype(np.int64) - np.int64(tt[:, 0].min())) / 1e3
        print(T, p.dpGeometry(), "stage_dp_us %.1f" % (1e3 * float(np.mean(ts[3:]))), "stamps(mean):", " ".join("%.1f" % rel[:, k].mean() for k in range(6)), "max filtered %.1f" % rel[:, 5].max(), flush=True)

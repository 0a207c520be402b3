#!/usr/bin/env python3
"""Capacity-1e5 knapsacks on the cluster DP kernel: 1 024 threads x <= 13 cells against 512 threads x <= 26 cells.
    python scripts/cluster_probe.py [m]"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dip_b200 import instances as I  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 60
b = I.csp_batch(m=m, n=500, cap=100_000, seed=77, wmax=10_000)
ref = None
dbg_modes = [int(a) for a in sys.argv[2:]] or [0]
for threads, maxc, dbg in [(1024, 0, d) for d in dbg_modes]:
    with GpuPricer(0) as p:
        p.set_option("dp_cluster_threads", threads)
        p.set_option("dp_cluster_max_cells", maxc)
        p.set_option("dp_cluster_debug", dbg)
        p.setObjective(b.orig_cost)
        p.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
        p.set_option("stage_timing", 1)
        res = p.solveBlocks(b.red_cost, np.zeros(m), redCostEps=-np.inf)
        ts = []
        for _ in range(3):
            p.solveBlocks(b.red_cost, np.zeros(m), redCostEps=-np.inf, result=res)
            ts.append(p.last_stage_ms()["knapsack_dp"])
        z = res.blockObj.copy()
        if ref is None and dbg == 0:
            ref = z
        if dbg == 0:
            assert np.array_equal(z, ref)
        cells = m * 500 * 100_001
        print(threads, maxc, 'dbg', dbg, p.dpGeometry(), f"{min(ts):.3f} ms  {cells / min(ts) / 1e6:.0f} GCUPS", flush=True)

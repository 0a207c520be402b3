#!/usr/bin/env python3
"""Sweep the knapsack-DP geometry (cells/thread S, items per barrier, guard) on a B200 and print
the DP kernel time per configuration next to the library's automatic choice.
    python scripts/dp_sweep.py [gape|gapd|csp|csp_small] ...
Output is the evidence behind the cost model in dip_b200/csrc/knapsack.cu (geomCost)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dip_b200 import instances as I  # noqa: E402
from dip_b200.capi import DipGpuError  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402


def batch(name):
    if name in ("gape", "gapd"):
        model = I.gap_pricing_model(I.gap_class_e() if name == "gape" else I.gap_class_d())
        u = I.gap_duals(model, np.random.default_rng(801600))
        rc = model.objective - np.tile(u[:model.meta["n"]], model.m)
        return I.KnapsackBatch(model.m, model.block_col_start, model.weight, model.capacity, rc, model.objective,
                               np.zeros(model.m), name)
    if name == "csp":
        return I.csp_batch()
    if name == "csp_small":
        return I.csp_batch(m=148 * 8)
    if name == "csp_few":
        return I.csp_batch(m=64)
    raise SystemExit(f"unknown workload {name}")


def time_dp(p, b, reps=4):
    best = None
    for _ in range(reps):
        res = p.solveBlocks(b.red_cost, b.alpha, redCostEps=-np.inf)
        ms = p.last_stage_ms()
        if best is None or ms["knapsack_dp"] < best["knapsack_dp"]:
            best = ms
    return best, res


SWEEP_S = tuple(int(x) for x in os.environ.get('SWEEP_S', '').split(',') if x)


def main():
    names = sys.argv[1:] or ["gape", "csp_small"]
    for name in names:
        b = batch(name)
        with GpuPricer(0) as p:
            p.setObjective(b.orig_cost)
            p.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
            p.set_option("stage_timing", 1)
            ms, res = time_dp(p, b)
            cells = res.cells
            ref_obj = res.blockObj.copy()
            print(f"## {name}: m={b.m} maxcap={int(b.capacity.max())} cells={cells}")
            print(f"auto: dp={ms['knapsack_dp']*1e3:9.1f} us  backtrack={ms['backtrack_emit']*1e3:8.1f} us  "
                  f"gcups={cells/ms['knapsack_dp']/1e6:8.1f}")
            print("   S     T items guard   dp_us     gcups")
            for S in (SWEEP_S or (1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 25)):
                for items in (1, 2, 3):
                    for noguard in (0, 1):
                        try:
                            p.set_option("dp_cells_per_thread", S)
                            p.set_option("dp_items_per_step", items)
                            p.set_option("dp_no_guard", noguard)
                        except DipGpuError:
                            continue
                        ms, res = time_dp(p, b, reps=3)
                        assert np.array_equal(res.blockObj, ref_obj), (S, items, noguard)
                        T = ((int(b.capacity.max()) + 1 + S - 1) // S + 31) // 32 * 32
                        print(f"{S:4d} {T:5d} {items:5d} {1-noguard:5d} {ms['knapsack_dp']*1e3:9.1f} "
                              f"{cells/ms['knapsack_dp']/1e6:9.1f}", flush=True)
            for k in ("dp_cells_per_thread", "dp_items_per_step", "dp_no_guard"):
                try:
                    p.set_option(k, 0)
                except DipGpuError:
                    pass


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Record the master dual vectors of real column-generation runs (examples/gap_colgen.py: restricted master in
HiGHS, DIP's row layout) for bench.py's replay legs: tests/golden/dual_trace_gap_d.npz (the whole root-node run of
GAP-D 20x200, every 4th round kept) and dual_trace_gap_e.npz (the first 16 rounds of GAP-E 80x1600 from a greedy primal
start -- the role of generateInitVars.  That is where the committed trace ends: from round 17 on the restricted master
takes HiGHS 15, 30, 60, 130 ... seconds per solve (measured with DIP_TRACE_LOG=1; z_RMP has not left the greedy
start's value by then and the bound is still of big-M scale), so the GAP-E replay prices EARLY-round duals;
DIP_TRACE_ROUNDS / DIP_TRACE_KEEP set the cut).  The core is built without branch rows (root node: they
are free rows with zero duals), a vector is stored as [assignment duals | convexity duals]; bench.py scatters it
into the full layout.  Pricing here runs on the CPU oracle (bit-equal to the GPU path, tests/test_colgen.py), so
the script needs no GPU.

    python scripts/record_dual_trace.py [gap_d] [gap_e]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dip_b200 import instances as I  # noqa: E402
from examples import gap_colgen as G  # noqa: E402
from oracle import dip_oracle as orc  # noqa: E402


def oracle_price(model):
    n = model.meta["n"]

    def price(u):
        ref = orc.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                              model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                              n_threads=os.cpu_count() or 1)
        out = []
        for b in range(model.m):
            if ref.col_keep[b]:
                ind = ref.col_ind[b]
                rows = np.concatenate([np.sort(ind % n), [model.n_base_rows + b]]).astype(np.int32)  # [A''s ; e_b] of a GAP column
                out.append((b, ind, ref.col_red_cost[b], ref.col_orig_cost[b], rows, np.ones(len(rows))))
        return out, ref.most_neg_reduced_cost
    return price


def greedy_start(g, model):
    """A primal start for the restricted master, the role of DecompAlgo::generateInitVars: jobs in order of decreasing
    regret go to the machine where they weigh least (ties: cheapest) that still has room; every machine's job set is one column.  With (nearly) all
    jobs placed the artificials leave the basis at once and the duals are on the scale of the costs from round 0 --
    without it the first dozens of rounds are priced under big-M duals."""
    m, n = g.m, g.n
    left = g.capacity.astype(np.int64).copy()
    cost = g.cost.astype(np.float64)
    srt = np.sort(cost, axis=0)
    order = np.argsort(-(srt[1] - srt[0]), kind="stable")
    jobs = [[] for _ in range(m)]
    placed = 0
    for j in order:
        # (class E costs fall with the weight: the cheapest machine is the heaviest -- take the lightest fit, ties by cost)
        for i in np.lexsort((cost[:, j], g.weight[:, j])):
            if g.weight[i, j] <= left[i]:
                left[i] -= g.weight[i, j]
                jobs[i].append(int(j))
                placed += 1
                break
    cols = []
    for i in range(m):
        js = sorted(jobs[i])
        ind = [i * n + j for j in js]
        rows = np.array(js + [model.n_base_rows + i], np.int32)
        cols.append((i, ind, float(sum(cost[i, j] for j in js)), rows, np.ones(len(rows))))
    return cols, placed


def record(name, g, max_rounds, keep_every, start=False):
    model = I.gap_pricing_model(g, branch_rows=False)
    t0 = time.perf_counter()
    init, placed = greedy_start(g, model) if start else ((), 0)
    if start:
        print(f"{name}: greedy start places {placed} of {g.n} jobs")
    log = (lambda msg: print(f"[{time.perf_counter() - t0:7.1f} s] {msg}", flush=True)) if os.environ.get("DIP_TRACE_LOG") else None
    tr, cols = G.run(model, oracle_price(model), max_rounds=max_rounds, init_cols=init, log=log)
    U = np.stack(tr.duals)[::keep_every]
    path = os.path.join(ROOT, "tests", "golden", f"dual_trace_{name}.npz")
    np.savez_compressed(path, duals=U, rounds=np.arange(len(tr.duals))[::keep_every], z_rmp=np.array(tr.z_rmp),
                        lower_bound=np.array(tr.lower_bound), m=model.m, n=model.meta["n"],
                        converged=bool(tr.n_new[-1] == 0 or tr.lower_bound[-1] >= tr.z_rmp[-1] - 1e-6))
    print(f"{name}: {len(tr.duals)} rounds in {time.perf_counter() - t0:.1f} s, {len(U)} vectors kept -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB); z_RMP {tr.z_rmp[-1]:.4f}, bound {tr.lower_bound[-1]:.4f}")


def main():
    which = sys.argv[1:] or ["gap_d", "gap_e"]
    if "gap_d" in which:
        record("gap_d", I.gap_class_d(), 600, 4)
    if "gap_e" in which:
        # (DIP_TRACE_ROUNDS: a master solve of this size takes tens of seconds in HiGHS once a few thousand columns are in)
        record("gap_e", I.gap_class_e(), int(os.environ.get("DIP_TRACE_ROUNDS", "150")), int(os.environ.get("DIP_TRACE_KEEP", "3")), start=True)


if __name__ == "__main__":
    main()

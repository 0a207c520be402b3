#!/usr/bin/env python3
"""Randomised parity of the cut-row path: random models, random sequences of appends (as a delta), merges, objective
changes, dual supports, phases, batches -- every sweep and every round compared with the oracle on the whole matrix.
    python scripts/fuzz_append.py [models] [seed]"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dip_b200 import instances as I  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402
from oracle import dip_oracle as orc  # noqa: E402

models = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
orc.build()
sweeps = rounds = 0
for mi in range(models):
    m = int(rng.integers(2, 12))
    n = int(rng.integers(20, 400))
    N = m * n
    R = int(rng.integers(40, 600))
    nBase = int(rng.integers(R // 2, R + 1))
    nnz = int(rng.integers(R, 40 * R))
    r = np.sort(rng.integers(0, R, size=nnz))
    rs = np.zeros(R + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=R), out=rs[1:])
    ci = rng.integers(0, N, size=nnz).astype(np.int32)
    v = np.round(rng.uniform(-9, 9, size=nnz), int(rng.integers(0, 4)))
    c = rng.uniform(0, 100, size=N)
    b = I.random_batch(m, n, n, 20, 300, seed=int(rng.integers(1 << 30)))
    with GpuPricer(0) as p:
        p.set_option("append_delta_min_nnz", int(rng.choice([0, 0, 1 << 40])))
        p.setModelCore(R, N, rs, ci, v, nBase, m)
        p.setObjective(c)
        p.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
        for step in range(int(rng.integers(3, 9))):
            op = int(rng.integers(0, 6))
            if op <= 2:   # append cut rows
                k = int(rng.integers(1, 30))
                nn = int(rng.integers(0, 25 * k))
                r2 = np.sort(rng.integers(0, k, size=nn))
                rs2 = np.zeros(k + 1, np.int64)
                np.cumsum(np.bincount(r2, minlength=k), out=rs2[1:])
                ci2 = rng.integers(0, N, size=nn).astype(np.int32)
                v2 = np.round(rng.uniform(-5, 5, size=nn), int(rng.integers(0, 3)))
                p.appendCoreRows(rs2, ci2, v2)
                rs = np.concatenate([rs, rs[-1] + rs2[1:]])
                ci = np.concatenate([ci, ci2])
                v = np.concatenate([v, v2])
                R += k
            elif op == 3:
                p.set_option("merge_appended_rows", 1)
            elif op == 4:
                c = rng.uniform(0, 100, size=N)
                p.setObjective(c)
            u = rng.standard_normal(R + m)
            u[rng.random(R + m) < 0.25] = 0.0
            phase = int(rng.choice([2, 2, 1]))
            ref = orc.calc_redcost(R, N, rs, ci, v, orc.adjust_duals(u, nBase, m), c, phase == 1)
            assert np.array_equal(p.calcRedCost(u, phase), ref), (mi, step, "sweep")
            sweeps += 1
            rr = orc.price_round(R, N, rs, ci, v, c, u, nBase, m, b.block_col_start, b.weight, b.capacity)
            res = p.priceRound(u)
            assert np.array_equal(res.blockObj, rr.z) and np.array_equal(res.blockRedCost, rr.col_red_cost), (mi, step, "round")
            rounds += 1
            if rng.random() < 0.3:
                U = np.stack([rng.standard_normal(R + m) for _ in range(3)])
                for q, g in enumerate(p.priceRoundsBatch(U)):
                    rq = orc.price_round(R, N, rs, ci, v, c, U[q], nBase, m, b.block_col_start, b.weight, b.capacity)
                    assert np.array_equal(g.blockObj, rq.z) and np.array_equal(g.blockRedCost, rq.col_red_cost), (mi, step, "batch", q)
                    rounds += 1
print(f"{models} random models: {sweeps} sweeps and {rounds} rounds bit-equal to the oracle through appends, merges and objective changes")

#!/usr/bin/env python3
"""Randomised parity of the rows-beyond-one-CTA kernels against the oracle: capacities 12 801 .. 240 000 (cluster sizes
2..16 and the global-memory kernel), few to many items, weight distributions from tiny to the whole capacity, both
profit modes, node bounds now and then.   python scripts/fuzz_big.py [cases] [seed]"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dip_b200.pricer import GpuPricer  # noqa: E402
from oracle import dip_oracle as orc  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
orc.build()
kinds = {}
for case in range(cases):
    cap = int(rng.choice([rng.integers(12_801, 30_000), rng.integers(30_000, 110_000), rng.integers(110_000, 240_000)]))
    m = int(rng.integers(1, 5))
    n = rng.integers(1, 70, size=m)
    bcs = np.zeros(m + 1, np.int32)
    np.cumsum(n, out=bcs[1:])
    N = int(bcs[-1])
    wkind = int(rng.integers(0, 4))
    hi = [cap // 50, cap // 6, cap, cap + cap // 10][wkind]
    w = rng.integers(0 if wkind == 3 else 1, max(hi, 2) + 1, size=N).astype(np.int32)
    mode = int(rng.integers(0, 2))
    rc = np.round(rng.uniform(-80, 15, size=N), 2 if mode else 9)
    oc = rng.uniform(1, 100, size=N)
    capv = rng.integers(12_801, cap + 1, size=m).astype(np.int32)
    capv[0] = cap
    alpha = rng.uniform(-5, 0, size=m)
    no_cluster = int(rng.random() < 0.25)
    with GpuPricer(0) as p:
        p.set_option("dp_no_cluster", no_cluster)
        p.setObjective(oc)
        p.setKnapsackBlocks(bcs, w, capv, mode)
        res = p.solveBlocks(rc, alpha, redCostEps=-np.inf)
        ref = orc.solve_blocks(bcs, w, capv, rc, oc, alpha, red_cost_eps=-np.inf, profit_mode=mode)
        assert np.array_equal(res.blockObj, ref.z), (case, cap, m, wkind, mode)
        for k in range(res.nCols):
            assert list(res.column(k)) == list(ref.col_ind[int(res.colBlock[k])]), (case, k)
        assert res.cells == ref.cells
        geo = p.dpGeometry()
        key = ("global" if geo["cells_per_thread"] == 0 else f"cluster S={geo['cells_per_thread']}")
        kinds[key] = kinds.get(key, 0) + 1
        if mode == 0 and rng.random() < 0.3:
            lb = (rng.random(N) < 0.03).astype(np.float64)
            ub = 1.0 - (rng.random(N) < 0.08) * (1 - lb)
            p.setColBounds(lb, ub)
            r2 = p.solveBlocks(rc, alpha)
            ref2 = orc.solve_blocks_bounded(bcs, w, capv, rc, oc, alpha, lb, ub)
            assert np.array_equal(r2.blockObj, ref2.z), (case, "bounds")
print(f"{cases} random cases bit-exact against the oracle; kernels used: {kinds}")

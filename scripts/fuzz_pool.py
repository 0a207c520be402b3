#!/usr/bin/env python3
"""Randomised parity of the waiting-column pool: random pools (large enough for the bitmap / compacted-copy variants
and small ones), random sequences of appends, removals, support changes and dual vectors -- every re-pricing pass
bit-equal to the oracle over the whole pool.    python scripts/fuzz_pool.py [cases] [seed]"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dip_b200.pricer import GpuPricer  # noqa: E402
from oracle import dip_oracle as orc  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 11)
orc.build()
passes = 0
for case in range(cases):
    R = int(rng.integers(2_000, 60_000))
    m = int(rng.integers(1, 40))
    uLen = R + m
    big = rng.random() < 0.7
    with GpuPricer(0) as p:
        p.set_option("pool_kernel", int(rng.choice([1, 1, 1, 2, 0])))
        p.set_option("pool_compact", int(rng.choice([1, 1, 0])))
        p.setModelCore(R, 8, np.zeros(R + 1, np.int64), np.zeros(0, np.int32), np.zeros(0), R, m)
        cs = np.zeros(1, np.int64)
        ri = np.zeros(0, np.int32)
        va = np.zeros(0)
        oc = np.zeros(0)
        support = None
        for step in range(int(rng.integers(3, 8))):
            op = int(rng.integers(0, 5))
            if op <= 1 or len(oc) == 0:     # append columns
                n = int(rng.integers(30_000, 60_000)) if big and len(oc) < 40_000 else int(rng.integers(1, 3_000))
                lens = rng.integers(0, int(rng.choice([8, 40, 90])), size=n)
                cs2 = np.zeros(n + 1, np.int64)
                np.cumsum(lens, out=cs2[1:])
                ri2 = rng.integers(0, uLen, size=cs2[-1]).astype(np.int32)
                va2 = np.round(rng.uniform(-3, 3, size=cs2[-1]), int(rng.integers(0, 3)))
                oc2 = rng.uniform(-5, 40, size=n)
                p.poolAppend(cs2, ri2, va2, oc2)
                cs = np.concatenate([cs, cs[-1] + cs2[1:]])
                ri, va, oc = np.concatenate([ri, ri2]), np.concatenate([va, va2]), np.concatenate([oc, oc2])
            elif op == 2:                    # remove columns
                keep = np.flatnonzero(rng.random(len(oc)) < 0.7).astype(np.int32)
                p.poolKeep(keep)
                lens = (cs[1:] - cs[:-1])[keep]
                parts_r = [ri[cs[j]:cs[j + 1]] for j in keep]
                parts_v = [va[cs[j]:cs[j + 1]] for j in keep]
                ri = np.concatenate(parts_r).astype(np.int32) if parts_r else np.zeros(0, np.int32)
                va = np.concatenate(parts_v) if parts_v else np.zeros(0)
                oc = oc[keep]
                cs = np.zeros(len(keep) + 1, np.int64)
                np.cumsum(lens, out=cs[1:])
            elif op == 3:                    # another support (or none)
                if rng.random() < 0.25:
                    support = None
                    p.setDualSupport(None)
                else:
                    support = np.sort(rng.choice(uLen, size=max(1, uLen // int(rng.integers(5, 40))), replace=False)).astype(np.int32)
                    p.setDualSupport(support)
            u = np.zeros(uLen)
            if support is None:
                u = rng.standard_normal(uLen)
                u[rng.random(uLen) < 0.3] = 0.0
            else:
                u[support] = rng.standard_normal(len(support))
            feasible = bool(rng.random() < 0.8)
            got, found = p.poolSetReducedCosts(u, feasible)
            ref, ref_found = orc.pool_reduced_costs(cs, ri, va, oc, u, feasible)
            assert np.array_equal(got, ref), (case, step)
            assert found == ref_found
            passes += 1
print(f"{cases} random pools: {passes} re-pricing passes bit-equal to the oracle through appends, removals and support changes")

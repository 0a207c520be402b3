#!/usr/bin/env python3
"""Soak of the end-to-end GAP-E round (host-buffer call, completion through the done word): N rounds back to back,
every result compared with the first pass over the same 16 dual vectors; reports how many waits ended through the
word and how many fell back to cudaStreamSynchronize.   python scripts/soak_rounds.py [rounds]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
model, duals = bench.workload()
u_pin = capi.PinnedArray((16, model.u_len), np.float64)
for k in range(16):
    u_pin.array[k] = duals[k]
ptrs = [u_pin.array[k].ctypes.data for k in range(16)]
res = RoundResult(model.m, model.m, model.N)
with GpuPricer(0) as p:
    p.setModel(model)
    p.setDualSupport(model.dual_support)
    ref = []
    for k in range(16):
        p.priceRoundRaw(ptrs[k], model.u_len, capi.PHASE_PRICE2, 1e-4, res)
        ref.append((res.nCols, res.blockObj.tobytes(), res.blockRedCost.tobytes(), res.colInd[:res.c.nInd].tobytes(), res.mostNegReducedCost))
    t0 = time.perf_counter()
    bad = 0
    for it in range(n):
        k = it % 16
        p.priceRoundRaw(ptrs[k], model.u_len, capi.PHASE_PRICE2, 1e-4, res)
        if (res.nCols, res.blockObj.tobytes(), res.blockRedCost.tobytes(), res.colInd[:res.c.nInd].tobytes(), res.mostNegReducedCost) != ref[k]:
            bad += 1
    dt = time.perf_counter() - t0
    st = np.zeros(2, np.int64)
    p.set_option("flag_stats", st.ctypes.data)
    print(f"{n} rounds in {dt:.1f} s ({1e6 * dt / n:.1f} us per round incl. the comparison, L2 warm), mismatches {bad}, "
          f"completed through the done word {int(st[0])}, fell back to the synchronise {int(st[1])}")
    assert bad == 0

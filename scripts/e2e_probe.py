#!/usr/bin/env python3
"""End-to-end round (dipgpu_price_round, host buffers) on the bench workload with an option toggled:
    python scripts/e2e_probe.py result_zero_copy            (prints us/round with the option 0 and 1, twice)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

opt = sys.argv[1] if len(sys.argv) > 1 else "result_zero_copy"
model, duals = bench.workload()
dev = torch.device("cuda", 0)
wb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rb = torch.ones(64 << 20, dtype=torch.float32, device=dev)
u_pin = capi.PinnedArray((len(duals), model.u_len), np.float64)
for k, u in enumerate(duals):
    u_pin.array[k] = u
ptrs = [u_pin.array[k].ctypes.data for k in range(len(duals))]
res = RoundResult(model.m, model.m, model.N)
ref = None
vals = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 0, 1]
for val in vals:
    with GpuPricer(0) as pr:
        pr.setModel(model)
        pr.setDualSupport(model.dual_support)
        pr.set_option(opt, val)
        tot, n = 0.0, 300
        for k in range(n + 10):
            wb.fill_(1)
            rb.sum()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pr.priceRoundRaw(ptrs[k % len(duals)], model.u_len, capi.PHASE_PRICE2, bench.RC_EPS, res)
            if k >= 10:
                tot += time.perf_counter() - t0
        sig = (res.nCols, float(res.mostNegReducedCost), res.column(0).tolist(), float(res.colRedCost[:res.nCols].sum()))
        ref = ref or sig
        assert sig == ref, (sig, ref)
        print(f"{opt}={val}: {1e6 * tot / n:.2f} us/round", flush=True)

#!/usr/bin/env python3
"""A few GAP-E rounds of the bench workload (dual support declared: the whole round is ONE fused kernel), for ncu:
    python scripts/run_round.py [dev|host] [reps]
dev = device-resident duals (dipgpu_price_round_dev), host = the host-buffer call (the kernel pulls u[support])."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "dev"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
model, duals = bench.workload()
dev = torch.device("cuda", 0)
u_dev = [torch.from_numpy(u).to(dev) for u in duals[:4]]
u_pin = capi.PinnedArray((4, model.u_len), np.float64)
for k in range(4):
    u_pin.array[k] = duals[k]
res = RoundResult(model.m, model.m, model.N)
with GpuPricer(0) as p:
    p.setModel(model)
    p.setDualSupport(model.dual_support)
    for k in range(reps):
        if mode == "dev":
            p.priceRoundDev(u_dev[k % 4].data_ptr(), model.u_len, redCostEps=bench.RC_EPS)
            torch.cuda.synchronize()
        else:
            p.priceRoundRaw(u_pin.array[k % 4].ctypes.data, model.u_len, capi.PHASE_PRICE2, bench.RC_EPS, res)
    print(mode, "rounds", reps, "kernel launches", p.counters()["kernel_launches"])

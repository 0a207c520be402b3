#!/usr/bin/env python3
"""dipgpu_expand_columns / dipgpu_price_round_expand on the bench workload with the host mirror + done word of the
expansion on (flag_wait = 1) and off (D2H copy + stream synchronise): wall clock per call, L2 flushed."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

model, duals = bench.workload()
dev = torch.device("cuda", 0)
fl = bench.Flusher(torch, [0])
u_pin = capi.PinnedArray((16, model.u_len), np.float64)
for k in range(16):
    u_pin.array[k] = duals[k]
ptrs = [u_pin.array[k].ctypes.data for k in range(16)]
for flag in (1, 0, 1, 0):
    with GpuPricer(0) as pr:
        pr.set_option("flag_wait", flag)
        pr.setModel(model)
        pr.setDualSupport(model.dual_support)
        res = RoundResult(model.m, model.m, model.N)
        out = bench.bench_expand(pr, model, duals, ptrs, res, lambda: fl(sync=False), torch, 1)
        print("flag_wait", flag, {k: round(v, 4) for k, v in out.items() if isinstance(v, float)}, flush=True)

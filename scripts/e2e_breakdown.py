#!/usr/bin/env python3
"""Where the end-to-end round's time goes on the bench workload (wall clock, L2 flushed between rounds):
full call / resident duals (no upload) / device entry point + sync (no upload, no result read) / stage events."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

model, duals = bench.workload()
dev = torch.device("cuda", 0)
wb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rb = torch.ones(64 << 20, dtype=torch.float32, device=dev)
u_pin = capi.PinnedArray((len(duals), model.u_len), np.float64)
for k, u in enumerate(duals):
    u_pin.array[k] = u
ptrs = [u_pin.array[k].ctypes.data for k in range(len(duals))]
u_dev = [torch.from_numpy(u).to(dev) for u in duals]
res = RoundResult(model.m, model.m, model.N)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def timed(fn, n=300):
    tot = 0.0
    for k in range(n + 10):
        wb.fill_(1)
        rb.sum()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn(k)
        if k >= 10:
            tot += time.perf_counter() - t0
    return 1e6 * tot / n


with GpuPricer(0) as pr:
    pr.setModel(model)
    pr.setDualSupport(model.dual_support)
    print("full call            %.2f us" % timed(lambda k: pr.priceRoundRaw(ptrs[k % 16], model.u_len, capi.PHASE_PRICE2, bench.RC_EPS, res)))
    print("resident duals       %.2f us" % timed(lambda k: pr.priceRoundRaw(None, model.u_len, capi.PHASE_PRICE2, bench.RC_EPS, res)))

    def devsync(k):
        pr.priceRoundDev(u_dev[k % 16].data_ptr(), model.u_len, redCostEps=bench.RC_EPS, stream=stream.cuda_stream)
        stream.synchronize()
    print("device entry + sync  %.2f us" % timed(devsync))

    def empty(k):
        stream.synchronize()
    print("sync only            %.2f us" % timed(empty))
    pr.set_option("stage_timing", 1)
    acc = {}
    for k in range(110):
        wb.fill_(1)
        rb.sum()
        torch.cuda.synchronize()
        pr.priceRoundRaw(ptrs[k % 16], model.u_len, capi.PHASE_PRICE2, bench.RC_EPS, res)
        if k >= 10:
            for n, v in pr.last_stage_ms().items():
                acc[n] = acc.get(n, 0.0) + v
    print("stage events (us):", {n: round(10 * v, 2) for n, v in acc.items()})

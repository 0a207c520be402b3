#!/usr/bin/env python3
"""The end-to-end GAP-E round (host-buffer C-ABI call, dual support declared) taken apart: wall clock of the call,
the in-kernel %globaltimer stamps of the fused DP kernel on that path (the kernel pulls u[support] from host memory),
and the same for the device-resident entry point.  Options on the command line: name=value ..."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

opts = dict(a.split("=") for a in sys.argv[1:])
model, duals = bench.workload()
dev = torch.device("cuda", 0)
wb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rb = torch.ones(64 << 20, dtype=torch.float32, device=dev)
u_pin = capi.PinnedArray((16, model.u_len), np.float64)
for k in range(16):
    u_pin.array[k] = duals[k]
ptrs = [u_pin.array[k].ctypes.data for k in range(16)]
u_dev = [torch.from_numpy(u).to(dev) for u in duals[:16]]
res = RoundResult(model.m, model.m, model.N)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
names = ["start", "staged", "compacted", "dp_done", "backtracked", "filtered", "polled", "prefixed"]


def flush():
    wb.fill_(1)
    rb.sum()
    torch.cuda.synchronize()


def stamps(p, fn):
    p.set_option("dp_debug_times", 1)
    acc = []
    for k in range(8):
        flush()
        fn(k)
        torch.cuda.synchronize()
        tt = np.zeros((model.m, 8), np.uint64)
        p.set_option("dp_debug_times", tt.ctypes.data)
        p.set_option("dp_debug_times", 1)
        if k >= 2:
            acc.append((tt.astype(np.int64) - np.int64(tt[:, 0].min())) / 1e3)
    p.set_option("dp_debug_times", 0)
    rel = np.mean(acc, axis=0)
    return " ".join(f"{nm} {rel[:, k].mean():.1f}/{rel[:, k].max():.1f}" for k, nm in enumerate(names))


with GpuPricer(0) as p:
    for k, v in opts.items():
        p.set_option(k, int(v))
    p.setModel(model)
    p.setDualSupport(model.dual_support)

    def full(k):
        p.priceRoundRaw(ptrs[k % 16], model.u_len, capi.PHASE_PRICE2, bench.RC_EPS, res)

    def devr(k):
        p.priceRoundDev(u_dev[k % 16].data_ptr(), model.u_len, redCostEps=bench.RC_EPS, stream=stream.cuda_stream)
        stream.synchronize()

    p.set_option("host_timing", 1)
    for nm, fn in (("host-buffer call", full), ("device entry + sync", devr)):
        ts, hs = [], []
        for k in range(210):
            flush()
            t0 = time.perf_counter()
            fn(k)
            if k >= 10:
                ts.append(time.perf_counter() - t0)
                h = np.zeros(4)
                p.set_option("host_timing", h.ctypes.data)
                hs.append(h)
        if fn is full:
            h = np.mean(hs, axis=0)
            print(f"   host stamps (us after entry): DP launch call begins {h[3]:.2f}  enqueued {h[0]:.2f}  result seen {h[1]:.2f}  unpacked {h[2]:.2f}")
        print(f"{nm}: mean {1e6 * np.mean(ts):.2f} us  median {1e6 * np.median(ts):.2f}  p10 {1e6 * np.percentile(ts, 10):.2f}  p90 {1e6 * np.percentile(ts, 90):.2f}")
        print("   in-kernel (mean/max over blocks, us after the first CTA's start):", stamps(p, fn), flush=True)

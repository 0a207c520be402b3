#!/usr/bin/env python3
"""Device-side cost of mirroring the packed result into mapped host memory: the device-resident round
(CUDA events, L2 flushed) with result_zero_copy = 1 (no mirror from this entry point) and 2 (mirror always)."""
import sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402
model, duals = bench.workload()
dev = torch.device('cuda', 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
wb = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rb = torch.ones(64 << 20, dtype=torch.float32, device=dev)
u_dev = [torch.from_numpy(u).to(dev) for u in duals]
for val in (1, 2, 1, 2):
    with GpuPricer(0) as p:
        p.setModel(model)
        p.set_option("result_zero_copy", val)
        ts = []
        for k in range(210):
            wb.fill_(1)
            rb.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            p.priceRoundDev(u_dev[k % 16].data_ptr(), model.u_len, redCostEps=1e-4, stream=stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            if k >= 10:
                ts.append(a.elapsed_time(b))
        print("result_zero_copy", val, "us/round %.2f" % (1e3 * float(np.mean(ts))), flush=True)

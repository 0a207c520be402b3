#!/usr/bin/env python3
"""Device-resident GAP-E round (CUDA events, L2 flushed) under the round-2 knobs:
fuse_redcost (reduced costs inside the DP kernel) x dp_coop (cooperative launch vs tickets) x dual support,
with the in-kernel %globaltimer stamps of the last round."""
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
u_dev = [torch.from_numpy(u).to(dev) for u in duals[:16]]
names = ["start", "staged", "compacted", "dp_done", "backtracked", "filtered", "polled", "prefixed"]
for fuse, coop, sup in ((0, 1, 0), (1, 1, 0), (1, 0, 0), (1, 1, 1), (1, 0, 1), (0, 1, 1)):
    with GpuPricer(0) as p:
        p.set_option("fuse_redcost", fuse)
        p.set_option("dp_coop", coop)
        p.setModel(model)
        if sup:
            p.setDualSupport(model.dual_support)
        ts = []
        for k in range(160):
            wb.fill_(1)
            rb.sum()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            p.priceRoundDev(u_dev[k % 16].data_ptr(), model.u_len, redCostEps=1e-4, stream=stream.cuda_stream)
            b.record(stream)
            torch.cuda.synchronize()
            if k >= 10:
                ts.append(a.elapsed_time(b))
        p.set_option("dp_debug_times", 1)
        for k in range(3):
            wb.fill_(1)
            rb.sum()
            p.priceRoundDev(u_dev[k].data_ptr(), model.u_len, redCostEps=1e-4, stream=stream.cuda_stream)
            torch.cuda.synchronize()
        tt = np.zeros((model.m, 8), np.uint64)
        p.set_option("dp_debug_times", tt.ctypes.data)
        rel = (tt.astype(np.int64) - np.int64(tt[:, 0].min())) / 1e3
        print(f"fuse_redcost={fuse} dp_coop={coop} support={sup}: {1e3 * float(np.mean(ts)):.2f} us/round (median {1e3 * float(np.median(ts)):.2f}) | " +
              " ".join(f"{nm} {rel[:, k].mean():.1f}" for k, nm in enumerate(names)), flush=True)

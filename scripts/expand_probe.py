import sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from dip_b200.pricer import GpuPricer, RoundResult
model, duals = bench.workload()
wb = torch.empty(256 << 20, dtype=torch.uint8, device='cuda'); rb = torch.ones(64 << 20, dtype=torch.float32, device='cuda')
with GpuPricer(0) as p:
    p.setModel(model)
    res = RoundResult(model.m, model.m, model.N)
    tot = 0
    for k in range(60):
        p.priceRound(duals[k % 16], result=res)
        wb.fill_(1); rb.sum(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        cs, ri, va, hs = p.expandColumns()
        if k >= 10: tot += time.perf_counter() - t0
    print("expand us/call %.1f nnz %d" % (1e6 * tot / 50, len(ri)))

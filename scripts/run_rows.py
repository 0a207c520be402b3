#!/usr/bin/env python3
"""Run one of the 'next rows' workloads of bench.py a few times (for ncu):
    python scripts/run_rows.py pool | mckp | batch16"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dip_b200 import capi, instances as I  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402

which = sys.argv[1]
torch.cuda.set_device(0)
flush = lambda: None  # noqa: E731
if which == "pool":
    print(bench.bench_pool(torch, 0, 6542.4, "n/a", flush, 0))
elif which == "mckp":
    print(bench.bench_mckp(torch, 0, 37000.0, flush))
elif which == "bigdp":
    print(bench.bench_bigdp(torch, 0, 37000.0, flush, m=int(sys.argv[2]) if len(sys.argv) > 2 else 60))
elif which == "intknap":
    print(bench.bench_intknap(torch, 0, 37000.0, flush))
else:
    model, duals = bench.workload()
    with GpuPricer(0) as p:
        p.setModel(model)
        p.setDualSupport(model.dual_support)
        U = np.stack(duals)
        for _ in range(3):
            res = p.priceRoundsBatch(U)
        print(sum(r.cells for r in res), sum(r.cellsEvaluated for r in res))

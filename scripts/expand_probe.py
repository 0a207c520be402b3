#!/usr/bin/env python3
"""dipgpu_expand_columns after a GAP-E round, L2 flushed: wall clock around the raw C-ABI call (bench.py's
`expand_columns` leg without the CPU comparison)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import bench  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

model, duals = bench.workload()
wb = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rb = torch.ones(64 << 20, dtype=torch.float32, device="cuda")
cap_nnz = 1 << 20
col_start = np.zeros(model.m + 1, np.int64)
row_ind = np.zeros(cap_nnz, np.int32)
val = np.zeros(cap_nnz, np.float64)
hsh = np.zeros(model.m, np.uint64)
mc = capi.MCols()
mc.capCols, mc.capNnz = model.m, cap_nnz
mc.colStart, mc.rowInd, mc.val, mc.hash = capi.ptr(col_start), capi.ptr(row_ind), capi.ptr(val), capi.ptr(hsh)
lib = capi.lib()
with GpuPricer(0) as p:
    p.setModel(model)
    res = RoundResult(model.m, model.m, model.N)
    tot = 0.0
    for k in range(60):
        p.priceRound(duals[k % 16], result=res)
        wb.fill_(1)
        rb.sum()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.dipgpu_expand_columns(p._ctx, 1e-6, C.byref(mc))
        assert rc == 0
        if k >= 10:
            tot += time.perf_counter() - t0
    print("expand us/call %.1f nnz %d" % (1e6 * tot / 50, mc.nNnz))

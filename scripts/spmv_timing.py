#!/usr/bin/env python3
"""How the L2 flush method changes the event-timed duration of the reduced-cost sweep (MILPBlock shape):
(a) 256 MiB written before the sweep (the L2 is then full of DIRTY lines whose write-back competes with the
sweep's reads), (b) written then 256 MiB read (L2 full of clean lines), (c) an empty kernel after (a)/(b):
the fixed cost of the event pair itself."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dip_b200 import instances as I  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
wbuf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rbuf = torch.ones(64 << 20, dtype=torch.float32, device=dev)
tiny = torch.zeros(32, device=dev)

model, u = I.milpblock_model()
p = GpuPricer(0)
p.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
p.setObjective(model.objective)
ud = torch.from_numpy(u).to(dev)
by = model.spmv_bytes()


def flush(kind):
    if kind in ("write", "write+read"):
        wbuf.fill_(1)
    if kind in ("read", "write+read"):
        rbuf.sum()


def run(kind, what):
    ts = []
    for it in range(23):
        flush(kind)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        if what == "spmv":
            p.calcRedCostDev(ud.data_ptr(), model.u_len, stream=stream.cuda_stream)
        else:
            tiny.add_(1)
        b.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    print(f"flush={kind:11s} {what:6s} {ms*1e3:7.1f} us  min {min(ts)*1e3:7.1f}  "
          + (f"{by/ms/1e6:8.1f} GB/s" if what == "spmv" else ""), flush=True)


for kind in ("write", "read", "write+read", "none"):
    run(kind, "spmv")
    run(kind, "empty")

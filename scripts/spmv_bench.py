#!/usr/bin/env python3
"""Time the reduced-cost SpMV (device-resident duals, CUDA events, L2 flushed) on the MILPBlock
and GAP-E shapes: sweep of the stream kernel's tile entry budget, ring depth and warps per CTA,
plus the lane-cooperative fallback."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rflush = torch.ones(64 << 20, dtype=torch.float32, device=dev)


def timeit(p, ud, model, reps=13):
    ts = []
    for it in range(reps):
        flush.fill_(1)   # 256 MiB written ...
        rflush.sum()     # ... then 256 MiB read: the L2 holds clean lines, no write-back under the timed sweep
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        p.calcRedCostDev(ud.data_ptr(), model.u_len, stream=stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(min(ts))


def run(name, model, u, configs):
    with GpuPricer(0) as p:
        p.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
        p.setObjective(model.objective)
        ud = torch.from_numpy(u).to(dev)
        by = model.spmv_bytes()
        for cfg in configs:
            lanes = cfg.get("lanes", 0)
            p.set_option("spmv_lanes", lanes)
            if lanes == 0:
                for opt, dflt in (("chunk", 0), ("window", 0), ("stages", 0), ("warps", 0), ("u_smem", -1), ("max_cols", 0)):
                    p.set_option("spmv_" + opt, cfg.get(opt, dflt))
            ms, mn = timeit(p, ud, model)
            print(f"{name:10s} {str(cfg):60s} {ms*1e3:8.1f} us  min {mn*1e3:8.1f} us  "
                  f"{by/ms/1e6:8.1f} GB/s ({by/1e6:.1f} MB)", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "milp"):
    model, u = I.milpblock_model()
    cfgs = [{}]
    for window in (32, 1024, 4096, 16384):
        for chunk in (256, 384):
            for warps in (16, 32):
                cfgs.append({"window": window, "chunk": chunk, "warps": warps})
    cfgs += [{"window": 4096, "chunk": 192}, {"window": 4096, "chunk": 256, "stages": 3}, {"window": 4096, "chunk": 512}]
    cfgs.append({"lanes": 8})
    run("milpblock", model, u, cfgs)
if which in ("all", "gape"):
    g = I.gap_pricing_model(I.gap_class_e())
    cfgs = [{}, {"warps": 4}, {"stages": 3}, {"chunk": 192}, {"max_cols": 64}, {"lanes": 1}]
    run("gap-e", g, I.gap_duals(g, np.random.default_rng(1)), cfgs)

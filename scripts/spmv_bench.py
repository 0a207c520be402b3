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


def timeit(p, ud, model, reps=13):
    ts = []
    for it in range(reps):
        flush.fill_(1)
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
        last_chunk = None
        for cfg in configs:
            lanes = cfg.get("lanes", 0)
            p.set_option("spmv_lanes", lanes)
            if lanes == 0:
                if cfg.get("chunk", 0) != last_chunk:
                    p.set_option("spmv_chunk", cfg.get("chunk", 0))
                    last_chunk = cfg.get("chunk", 0)
                p.set_option("spmv_stages", cfg.get("stages", 0))
                p.set_option("spmv_warps", cfg.get("warps", 0))
                p.set_option("spmv_u_smem", cfg.get("usmem", -1))
            ms, mn = timeit(p, ud, model)
            print(f"{name:10s} {str(cfg):60s} {ms*1e3:8.1f} us  min {mn*1e3:8.1f} us  "
                  f"{by/ms/1e6:8.1f} GB/s ({by/1e6:.1f} MB)", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "milp"):
    model, u = I.milpblock_model()
    cfgs = [{}]
    for chunk in (256, 320, 384, 512):
        for stages in (2, 3):
            for warps in (12, 16, 32):
                cfgs.append({"chunk": chunk, "stages": stages, "warps": warps})
    cfgs.append({"lanes": 8})
    run("milpblock", model, u, cfgs)
if which in ("all", "gape"):
    g = I.gap_pricing_model(I.gap_class_e())
    cfgs = [{}, {"stages": 2, "warps": 4}, {"stages": 2, "warps": 8}, {"stages": 3, "warps": 8}, {"chunk": 192},
            {"chunk": 96}, {"lanes": 1}, {"lanes": 4}]
    run("gap-e", g, I.gap_duals(g, np.random.default_rng(1)), cfgs)

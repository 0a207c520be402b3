#!/usr/bin/env python3
"""Time the reduced-cost SpMV (device-resident duals, CUDA events, L2 flushed) on the MILPBlock
and GAP-E shapes for the stream kernel and the lane-cooperative fallback."""
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


def run(name, model, u, lanes_opts, extra=()):
    with GpuPricer(0) as p:
        p.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
        p.setObjective(model.objective)
        ud = torch.from_numpy(u).to(dev)
        for lanes in list(lanes_opts) + list(extra):
            if isinstance(lanes, tuple):
                p.set_option("spmv_lanes", 0)
                p.set_option("spmv_warps", lanes[0])
                p.set_option("spmv_u_smem", lanes[1])
                tag = f"w{lanes[0]}u{lanes[1]}"
                lanes = 0
            else:
                p.set_option("spmv_lanes", lanes)
                tag = ""
            ts = []
            for it in range(13):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                p.calcRedCostDev(ud.data_ptr(), model.u_len, stream=stream.cuda_stream)
                b.record(stream)
                torch.cuda.synchronize()
                if it >= 3:
                    ts.append(a.elapsed_time(b))
            by = model.spmv_bytes()
            ms = float(np.mean(ts))
            print(f"{name:10s} lanes={lanes:2d} {tag:6s} {ms*1e3:8.1f} us  min {min(ts)*1e3:8.1f} us  "
                  f"{by/ms/1e6:8.1f} GB/s ({by/1e6:.1f} MB)", flush=True)


model, u = I.milpblock_model()
run("milpblock", model, u, [0, 8], extra=[(12, 1), (24, 1), (8, 2), (12, 2), (16, 2), (20, 2), (24, 2), (16, 0)])
g = I.gap_pricing_model(I.gap_class_e())
run("gap-e", g, I.gap_duals(g, np.random.default_rng(1)), [0, 1, 4])

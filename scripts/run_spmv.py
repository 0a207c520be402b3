#!/usr/bin/env python3
"""Run the reduced-cost SpMV sweep on the MILPBlock shape a few times (for ncu)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dip_b200 import instances as I  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402

model, u = I.milpblock_model()
with GpuPricer(0) as p:
    p.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
    p.setObjective(model.objective)
    if len(sys.argv) > 1:
        p.set_option("spmv_lanes", int(sys.argv[1]))
    for _ in range(3):
        rc = p.calcRedCost(u)
    print(float(rc.sum()))

#!/usr/bin/env python3
"""Batched GAP-E rounds through the host-buffer call: wall clock per call with the in-kernel dual pull of ticketed batches on / off.
    python scripts/batch_probe.py [devices, e.g. 0,1,2,3] [batch sizes ...]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench as B  # noqa: E402
from dip_b200 import capi  # noqa: E402
from dip_b200.pricer import GpuPricer  # noqa: E402


def main():
    import torch
    devs = [int(d) for d in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
    sizes = [int(a) for a in sys.argv[2:]] or [16, 64, 256]
    model, duals = B.workload()
    p = GpuPricer(devs if len(devs) > 1 else devs[0])
    p.setModel(model)
    p.set_option("batch_pull_min", 0)
    p.setDualSupport(model.dual_support)
    big = max(sizes)
    pin = capi.PinnedArray((big, model.u_len), np.float64)
    for k in range(big):
        pin.array[k] = duals[k % len(duals)]
    fl = B.Flusher(torch, devs)
    res, arr = p._batch_results(big)
    for n in sizes:
        for pipe in (0, 1, 0, 1):
            p.set_option("batch_pull", pipe)
            for _ in range(3):
                p.priceRoundsBatchRaw(n, pin.array.ctypes.data, model.u_len, capi.PHASE_PRICE2, 1e-4, arr)
            t = 0.0
            reps = 15
            for _ in range(reps):
                fl(sync=True)
                t0 = time.perf_counter()
                p.priceRoundsBatchRaw(n, pin.array.ctypes.data, model.u_len, capi.PHASE_PRICE2, 1e-4, arr)
                t += time.perf_counter() - t0
            print(f"devices {devs} batch {n:4d} batch_pull={pipe}: {1e6 * t / reps:9.1f} us/call  {n * reps / t:10.0f} rounds/s", flush=True)
    p.close()


if __name__ == "__main__":
    main()

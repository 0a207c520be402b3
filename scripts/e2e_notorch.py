#!/usr/bin/env python3
"""The end-to-end GAP-E round (host-buffer C-ABI call) timed in a process WITHOUT torch: the L2 flush between rounds
goes through libcudart directly (256 MiB device-to-device copy).  Tells the library's own host cost (launch call,
completion wait, unpack) from whatever a torch process adds.  Options on the command line: name=value ..."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dip_b200 import capi, instances as I  # noqa: E402
from dip_b200.pricer import GpuPricer, RoundResult  # noqa: E402

assert "torch" not in sys.modules
rt = C.CDLL("libcudart.so.12")
opts = dict(a.split("=") for a in sys.argv[1:])
model = I.gap_pricing_model(I.gap_class_e())
rng = np.random.default_rng(801600)   # bench.workload(), without importing bench (it imports nothing heavy, but stays apart)
branch_rows = I.gap_branch_rows(model, rng)
duals = [I.gap_duals(model, rng, branch_rows=branch_rows) for _ in range(16)]
model.dual_support = I.gap_dual_support(model, branch_rows)
u_pin = capi.PinnedArray((16, model.u_len), np.float64)
for k in range(16):
    u_pin.array[k] = duals[k]
ptrs = [u_pin.array[k].ctypes.data for k in range(16)]
res = RoundResult(model.m, model.m, model.N)
with GpuPricer(0) as p:
    for k, v in opts.items():
        p.set_option(k, int(v))
    p.setModel(model)
    p.setDualSupport(model.dual_support)
    a, b = C.c_void_p(), C.c_void_p()
    nbytes = 256 << 20
    assert rt.cudaMalloc(C.byref(a), C.c_size_t(nbytes)) == 0 and rt.cudaMalloc(C.byref(b), C.c_size_t(nbytes)) == 0
    rt.cudaMemset(a, 1, C.c_size_t(nbytes))
    p.set_option("host_timing", 1)
    for flush in (1, 0):
        ts, hs = [], []
        for k in range(310):
            if flush:
                rt.cudaMemcpyAsync(b, a, C.c_size_t(nbytes), 3, None)
            rt.cudaDeviceSynchronize()
            t0 = time.perf_counter()
            p.priceRoundRaw(ptrs[k % 16], model.u_len, capi.PHASE_PRICE2, 1e-4, res)
            if k >= 10:
                ts.append(time.perf_counter() - t0)
                h = np.zeros(4)
                p.set_option("host_timing", h.ctypes.data)
                hs.append(h)
        h = np.mean(hs, axis=0)
        print(f"L2 {'flushed' if flush else 'warm'}: call mean {1e6 * np.mean(ts):.2f} us median {1e6 * np.median(ts):.2f} | "
              f"DP launch call begins {h[3]:.2f}  enqueued {h[0]:.2f}  result seen {h[1]:.2f}  unpacked {h[2]:.2f}", flush=True)

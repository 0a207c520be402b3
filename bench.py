#!/usr/bin/env python3
"""bench.py -- pricing rounds/s on GAP-E 80x1600 (BASELINE.json `metric`), with the knapsack-DP
GCUPS (CSP batch) and the reduced-cost SpMV GB/s (MILPBlock shape) reported beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One STEP = one pricing round of DIP's price-and-cut loop (DecompAlgo::generateVars,
Dip/src/DecompAlgo.cpp:4622-5260) for ONE master dual vector over all 80 blocks:
reduced costs c - A''^T u  ->  80 knapsack DPs  ->  backtrack  ->  negative-reduced-cost filter.

  value : rounds/s with the dual vector already in HBM (device-resident entry point), device time
          by CUDA events per step on the launching stream, L2 flushed between steps.
  e2e   : the same metric through the reference-facing C-ABI call dipgpu_price_round with HOST
          buffers (pinned u in, packed columns out), H2D + D2H inside the timed region.
  N > 1 : one process per GPU (torchrun) for the plumbing, ONE caller for the work: DIP is one host
          process, so rank 0 prices its stream of dual vectors through ONE multi-device context
          (dipgpu_create_multi over the N GPUs) while the other ranks wait.  `value` / `e2e` stay one
          dual vector per step, its 80 blocks partitioned over the GPUs (strong scaling of a
          latency-bound round: no faster than one GPU, and it says so); `batch16` / `batch64` /
          `ladder8` are the calls that do scale (dual vectors partitioned), each beside the same call on
          one GPU in the same run.  `replicas` (every rank pricing its own rounds) is an extra only.

--impl reference runs the CPU pricing round (the oracle's restatement of the reference path,
OpenMP over blocks like DecompAlgo.cpp:4815) on the host cores for the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dip_b200 import instances as I  # noqa: E402

METRIC = "pricing_rounds_per_s_gap_e_80x1600"
UNIT = "rounds/s"
N_DUALS = 16          # distinct master dual vectors cycled through the steps
N_BATCH = 64          # ... and priced per call by the batch legs
N_BIG = 512           # multi-GPU runs: a batch that leaves every one of 8 GPUs 64 vectors
RC_EPS = 1e-4         # DecompParam::RedCostEpsilon (Dip/src/DecompParam.h:691)


def load_peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(key, scale=1.0):
    """dram__bytes_read + dram__bytes_write per launch from the committed ncu capture of this kernel
    (profiles/ncu_traffic.json, written by scripts/summarize_ncu.py), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]
        return (d["dram_bytes_read"] + d["dram_bytes_write"]) * scale
    except Exception:
        return None


def workload():
    """GAP-E 80x1600 and 16 master dual vectors of ONE node of the tree: the same 1 % of the branch rows
    carry a bound (and a dual) in all of them, every other branch row is free (zero dual)."""
    model = I.gap_pricing_model(I.gap_class_e())
    rng = np.random.default_rng(801600)
    branch_rows = I.gap_branch_rows(model, rng)
    duals = [I.gap_duals(model, rng, branch_rows=branch_rows) for _ in range(N_BATCH)]
    model.dual_support = I.gap_dual_support(model, branch_rows)
    return model, duals


def workload_config(model):
    """`config` of BOTH arms (the GPU path and --impl reference): the same dict, key for key."""
    return {"workload": "GAP-E 80x1600 pricing round (SpMV c-A''^T u + 80 knapsack DPs + backtrack + filter), "
                        "one master dual vector per step, 16 dual vectors of one node cycled",
            "blocks": int(model.m), "cols": int(model.N), "rows": int(model.R), "nnz": int(model.nnz),
            "red_cost_eps": RC_EPS}


# ------------------------------------------------------------------------ clocks sampler


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names.setdefault(v, k.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bit & (bit - 1) == 0 and r & bit and nm not in ("GpuIdle", "None", "All"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------- CPU baseline


def cpu_round_fn(model, threads, solver="dp"):
    """One whole pricing round on the host cores (oracle/dip_oracle.c): SpMV single-threaded as DIP does it, the blocks
    under OpenMP schedule(dynamic,1) like Dip/src/DecompAlgo.cpp:4815.  solver "dp": the restatement of the reference's
    knapsack01 DP; "hs": the reference's own KnapsackOptimizeHS compiled from Dip/src/UtilKnapsack.cpp (oracle/_ref)."""
    from oracle import dip_oracle as orc
    orc.build()
    mode = 0
    if solver == "hs":
        if not orc.enable_ref_hs():
            raise RuntimeError("oracle/_ref is not built")
        mode = orc.PROFIT_REF_HS

    def run(u):
        return orc.price_round(model.R, model.N, model.row_start, model.col_ind, model.val, model.objective, u,
                               model.n_base_rows, model.m, model.block_col_start, model.weight, model.capacity,
                               profit_mode=mode, red_cost_eps=RC_EPS, n_threads=threads, want_ind=True)
    return run


def _time_rounds(run, duals, budget_s, max_n=2000):
    run(duals[0])
    t0 = time.perf_counter()
    n = 0
    while True:
        run(duals[n % len(duals)])
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= max_n:
            return n, el


def reference_solver():
    """The CPU arm's knapsack solver: the reference's compiled Horowitz-Sahni B&B when oracle/_ref exists, else the
    DP port."""
    from oracle import dip_oracle as orc
    orc.build()
    return ("hs", "reference") if orc.have_ref() else ("dp", "port")


def cpu_baseline(model, duals, budget_s=2.0):
    """CPU pricing timed on the host cores: whole GAP-E rounds with the blocks priced (i) by the
    reference's own Horowitz-Sahni B&B (KnapsackOptimizeHS, Dip/src/UtilKnapsack.cpp:246-502, compiled from the
    reference into oracle/_ref) and (ii) by the DP port (knapsack01), each at 1 thread, at 4 (DIP's default
    NumConcurrentThreadsSubProb, Dip/src/DecompParam.h:636) and at all host cores; and the reduced-cost sweep alone,
    single-threaded row-major as DIP does it and column-partitioned over 8 threads.  `value` is the reference arm's
    configuration (HS at all cores when oracle/_ref exists)."""
    from oracle import dip_oracle as orc
    cores = os.cpu_count() or 1
    solver, kind = reference_solver()
    variants = {}
    for sv in (["hs"] if orc.have_ref() else []) + ["dp"]:
        variants[sv] = {}
        for th in sorted({1, min(4, cores), cores}):
            n, el = _time_rounds(cpu_round_fn(model, th, sv), duals, budget_s)
            variants[sv][str(th)] = n / el
    # how far the B&B's optimum is from the exact DP's on this workload (it prunes with tolerances)
    hs_note = None
    if orc.have_ref():
        a = cpu_round_fn(model, cores, "hs")(duals[0])
        b = cpu_round_fn(model, cores, "dp")(duals[0])
        rel = np.abs(a.z - b.z) / np.maximum(np.abs(b.z), 1.0)
        hs_note = {"blocks_whose_optimum_differs_by_more_than_1e-9_rel": int(np.sum(rel > 1e-9)), "max_rel_diff": float(rel.max()),
                   "note": "the B&B adds the profits in ratio order (last-bit differences) and prunes with tolerances"}
    ua = orc.adjust_duals(duals[0], model.n_base_rows, model.m)
    cs, ri, va = orc.csc_descending(model.R, model.N, model.row_start, model.col_ind, model.val)
    spmv = {}
    t0 = time.perf_counter()
    for k in range(40):
        orc.calc_redcost(model.R, model.N, model.row_start, model.col_ind, model.val, ua, model.objective)
    spmv["row_major_1_thread_ms"] = (time.perf_counter() - t0) / 40 * 1e3
    for th in sorted({1, min(8, cores)}):
        t0 = time.perf_counter()
        for k in range(40):
            orc.calc_redcost_csc(model.N, cs, ri, va, ua, model.objective, n_threads=th)
        spmv[f"column_ordered_{th}_threads_ms"] = (time.perf_counter() - t0) / 40 * 1e3
    val = variants[solver][str(cores)]
    return {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"whole GAP-E 80x1600 pricing rounds for {budget_s:.0f} s per variant (16 dual vectors cycled); value = "
                      + ("KnapsackOptimizeHS from oracle/_ref" if solver == "hs" else "the DP port")
                      + f" under OpenMP dynamic,1 over blocks on {cores} threads, SpMV single-thread as DIP",
            "rounds_per_s": {"reference_hs_bnb": variants.get("hs"), "dp_port_knapsack01": variants["dp"]},
            "hs_vs_exact_dp": hs_note,
            "spmv_gap_e": spmv,
            "note": "DIP's GAP example prices with Pisinger's combo (un-vendored, cannot be built here) and MILP "
                    "subproblems with Cbc (absent): neither can be timed; HS is the exact solver the reference ships"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    model, duals = workload()
    cores = os.cpu_count() or 1
    solver, kind = reference_solver()
    run = cpu_round_fn(model, cores, solver)
    for k in range(max(args.warmup, 1)):
        run(duals[k % N_DUALS])
    # a CPU round takes milliseconds: bound the run to a few minutes
    t0 = time.perf_counter()
    for k in range(args.steps):
        run(duals[k % N_DUALS])
    el = time.perf_counter() - t0
    val = args.steps / el
    what = ("the reference's KnapsackOptimizeHS (Dip/src/UtilKnapsack.cpp, compiled into oracle/_ref) under the OpenMP block loop "
            "of DecompAlgo.cpp:4815" if solver == "hs" else
            "the oracle's restatement of knapsack01 + generateVars (oracle/_ref absent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(model),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} whole rounds on {cores} host threads: {what}; DIP's GAP example itself prices "
                                   "with Pisinger's combo.c (not vendored)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------- GPU arm


_REAL_STDOUT = None


def _claim_stdout():
    """Everything libraries print (NCCL banner, warnings) goes to stderr; stdout carries exactly the
    one JSON line the driver parses."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


class Flusher:
    """L2 flush of every device a leg runs on: 256 MiB written, then 256 MiB read, so that the timed kernels
    start with an L2 full of CLEAN foreign lines (after a write-only flush the timed kernel also pays the
    write-back of ~126 MB of the flush's dirty lines -- scripts/spmv_timing.py: 37.0 vs 30.8 us)."""

    def __init__(self, torch, devices):
        self.torch, self.devices = torch, list(dict.fromkeys(devices))
        self.bufs = []
        for d in self.devices:
            dev = torch.device("cuda", d)
            self.bufs.append((torch.empty(256 << 20, dtype=torch.uint8, device=dev),
                              torch.ones(64 << 20, dtype=torch.float32, device=dev)))

    def __call__(self, sync=False):
        for d, (w, r) in zip(self.devices, self.bufs):
            with self.torch.cuda.device(d):
                w.fill_(1)
                r.sum()
        if sync or len(self.devices) > 1:
            for d in self.devices:
                self.torch.cuda.synchronize(d)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-extras", action="store_true", help="skip the CSP GCUPS / SpMV / CPU legs")
    ap.add_argument("--csp-blocks", type=int, default=10_000)
    ap.add_argument("--devices", default="", help="single process, several GPUs: comma-separated ordinals for the "
                                                  "multi-device legs (what torchrun's rank 0 does with --gpus N)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the pricing path is CUDA-only (no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # waiting ranks must not spin an NCCL kernel on a GPU rank 0 drives

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    devices = [int(d) for d in args.devices.split(",") if d != ""] if args.devices else list(range(world))
    if world == 1 and len(devices) <= 1:
        line = single_gpu_line(args, torch, dev, local_rank)
        _emit(line)
        return

    # ---- N > 1.  Extra first: every rank prices its own rounds (independent replicas, no exchange).
    model, duals = workload()
    replicas = replica_rounds(args, torch, dev, local_rank, model, duals, world, barrier, max_over_ranks) if world > 1 else None
    barrier()
    if rank == 0:
        line = multi_gpu_line(args, torch, devices, model, duals, replicas)
    if world > 1:
        dist.barrier(group=cpu_group)   # the other ranks wait on the host while rank 0 drives every GPU
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def replica_rounds(args, torch, dev, local_rank, model, duals, world, barrier, max_over_ranks):
    """EXTRA at N > 1: every rank prices its own GAP-E rounds on its own GPU (device-resident, as `value` at
    N = 1).  Independent replicas with no exchange: an upper bound on what N GPUs deliver, not a scaling result."""
    from dip_b200.pricer import GpuPricer
    K, W = min(args.steps, 100), args.warmup
    pr = GpuPricer(local_rank)
    pr.setModel(model)
    stream = torch.cuda.Stream(device=dev)
    flush = Flusher(torch, [local_rank])
    u_dev = [torch.from_numpy(u).to(dev) for u in duals[:N_DUALS]]
    with torch.cuda.stream(stream):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(W):
            pr.priceRoundDev(u_dev[k % N_DUALS].data_ptr(), model.u_len, redCostEps=RC_EPS, stream=stream.cuda_stream)
        barrier()
        for k in range(K):
            flush()
            stream.wait_stream(torch.cuda.default_stream(dev))
            ev[k][0].record(stream)
            pr.priceRoundDev(u_dev[k % N_DUALS].data_ptr(), model.u_len, redCostEps=RC_EPS, stream=stream.cuda_stream)
            ev[k][1].record(stream)
        barrier()
        stream.synchronize()
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    pr.close()
    return {"rounds_per_s": world * K / (ms * 1e-3), "ms_per_round_per_gpu": ms / K,
            "note": "every rank prices its own rounds, no exchange: an upper bound, not a scaling result"}


def single_gpu_line(args, torch, dev, local_rank):
    from dip_b200.pricer import GpuPricer, RoundResult
    from dip_b200 import capi
    hbm_peak, peak_src = load_peaks()
    model, duals = workload()
    K, W = args.steps, args.warmup
    rank, world = 0, 1

    def barrier():
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return x

    pr = GpuPricer(local_rank)
    pr.setModel(model)
    # a non-default stream: the library treats a NULL stream as "use the context's own"
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    assert sptr != 0
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    flush_rd = torch.ones(64 << 20, dtype=torch.float32, device=dev)   # 256 MiB, read after the write

    def flush_l2():
        # 256 MiB written, then 256 MiB read: every line of the timed kernels' inputs is evicted, and the
        # L2 is left full of CLEAN lines -- after a write-only flush the timed kernel also pays the
        # write-back of ~126 MB of the flush's dirty lines (scripts/spmv_timing.py: 37.0 vs 30.8 us)
        flush_buf.fill_(1)
        flush_rd.sum()

    u_dev = [torch.from_numpy(u).to(dev) for u in duals[:N_DUALS]]
    u_pin = capi.PinnedArray((N_BATCH, model.u_len), np.float64)
    for k, u in enumerate(duals):
        u_pin.array[k] = u
    res = RoundResult(model.m, model.m, model.N)

    # ---- parity spot check of what is about to be timed (device-resident vs host-buffer entry points)
    pr.priceRoundDev(u_dev[0].data_ptr(), model.u_len, redCostEps=RC_EPS, stream=sptr)
    torch.cuda.synchronize()
    r0 = pr.priceRound(duals[0], redCostEps=RC_EPS, result=res)
    cells_per_round, cells_eval_per_round = [], []
    for k in range(N_DUALS):
        rk = pr.priceRound(duals[k], redCostEps=RC_EPS, result=res)
        cells_per_round.append(rk.cells)
        cells_eval_per_round.append(rk.cellsEvaluated)
    kept0 = r0.nCols

    # ---- device-resident rounds (value).  The rows that can carry a dual at this node are declared once
    # (dipgpu_set_dual_support: free branch rows have zero duals), which is how the integration drives the library
    # (INTEGRATION.md); `no_dual_support` repeats the measurement without the declaration.
    def value_loop(steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(W):
            flush_l2()
            pr.priceRoundDev(u_dev[k % N_DUALS].data_ptr(), model.u_len, redCostEps=RC_EPS, stream=sptr)
        barrier()
        a = pr.counters()
        t0 = time.perf_counter()
        for k in range(steps):
            flush_l2()
            ev[k][0].record(stream)
            pr.priceRoundDev(u_dev[k % N_DUALS].data_ptr(), model.u_len, redCostEps=RC_EPS, stream=sptr)
            ev[k][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
        b = pr.counters()
        return sum(x.elapsed_time(y) for x, y in ev), b["kernel_launches"] - a["kernel_launches"], wall

    ptrs = [u_pin.array[k].ctypes.data for k in range(N_DUALS)]

    # end-to-end rounds through the host-buffer C-ABI call (e2e): pinned u in, packed columns out
    def e2e_loop(steps):
        for k in range(W):
            pr.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        barrier()
        c0 = pr.counters()
        tot = 0.0
        for k in range(steps):
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pr.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
            tot += time.perf_counter() - t0
        c1 = pr.counters()
        return tot, (c1["h2d_bytes"] - c0["h2d_bytes"]) // steps, (c1["d2h_bytes"] - c0["d2h_bytes"]) // steps

    pr.setDualSupport(model.dual_support)
    with ClockSampler(local_rank) as clk:
        dev_ms, launches, wall_dev = value_loop(K)
    value = K / (dev_ms * 1e-3)
    e2e_s, h2d_b, d2h_b = e2e_loop(K)
    # the same call on PAGEABLE dual vectors (what an LP solver hands over): u[support] is gathered on the host into
    # mapped pinned staging, the kernel pulls it from there in one contiguous piece -- an extra
    u_page = [np.array(duals[k], copy=True) for k in range(N_DUALS)]
    page_ptrs = [u.ctypes.data for u in u_page]
    page_s = 0.0
    for k in range(W + min(K, 100)):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr.priceRoundRaw(page_ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        if k >= W:
            page_s += time.perf_counter() - t0
    # the same call without the L2 flush between rounds (what a column-generation loop sees: the GPU idles while
    # the master LP is solved, its L2 keeps the model) -- an extra, the headline numbers are the flushed ones
    warm_s = 0.0
    for k in range(W + K):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        if k >= W:
            warm_s += time.perf_counter() - t0
    pr.setDualSupport(None)
    Kd = min(K, 100)
    dense_dev_ms, _, _ = value_loop(Kd)
    dense_s, dense_h2d, _ = e2e_loop(Kd)
    pr.setDualSupport(model.dual_support)
    e2e = {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
           "ms_per_step": 1e3 * e2e_s / K,
           "dual_support_rows": int(len(model.dual_support)),
           "pageable_u": {"value": min(K, 100) / page_s, "ms_per_step": 1e3 * page_s / min(K, 100),
                          "note": "same call on pageable dual vectors: host gather of u[support] into mapped staging (extra)"},
           "warm_l2": {"value": K / warm_s, "ms_per_step": 1e3 * warm_s / K,
                       "note": "same call, no L2 flush between rounds (extra; not the headline)"},
           "no_dual_support": {"value": Kd / dense_s, "ms_per_step": 1e3 * dense_s / Kd, "h2d_bytes_per_step": dense_h2d},
           "timing": "host wall clock around the synchronous dipgpu_price_round call (pinned u in, columns out); "
                     "`value`: the rows that can carry a dual declared once, only they are uploaded; "
                     "`no_dual_support`: the whole dual vector uploaded every round"}

    # ---- per-stage device times of the timed workload (CUDA events on the library's stream)
    pr.set_option("stage_timing", 1)
    stage = {}
    for k in range(min(K, 100)):
        flush_l2()
        torch.cuda.synchronize()
        pr.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        for nm, ms in pr.last_stage_ms().items():
            stage.setdefault(nm, []).append(ms)
    pr.set_option("stage_timing", 0)
    stage_ms = {nm: float(np.mean(v)) for nm, v in stage.items()}

    # ---- the step after the round: master columns + fingerprints of the kept columns
    expand = None
    if not args.skip_extras:
        expand = bench_expand(pr, model, duals, ptrs, res, flush_l2, torch, rank)

    extras = {}
    smem_peak = None
    if not args.skip_extras:
        smem_peak = pr.measureSmemGBs()
        extras["batched"] = bench_batched(pr, model, duals, u_pin, flush_l2, torch, smem_peak)
        extras["pool"] = bench_pool(torch, local_rank, hbm_peak, peak_src, flush_l2, rank)
        extras["mckp"] = bench_mckp(torch, local_rank, smem_peak, flush_l2)
        extras["bigdp"] = bench_bigdp(torch, local_rank, smem_peak, flush_l2)
        extras["intknap"] = bench_intknap(torch, local_rank, smem_peak, flush_l2)
        extras["other_shapes"] = bench_other_shapes(torch, dev, local_rank, flush_l2)
        extras["real_duals"] = bench_real_duals(torch, dev, local_rank, flush_l2)
    cells_mean = float(np.mean(cells_per_round))
    cells_eval_mean = float(np.mean(cells_eval_per_round))
    # the dominant kernel of the step IS the step: one fused kernel per device-resident round (gpu_launches == steps),
    # its average launch duration = the CUDA-event time of the timed region / steps
    dp_ms = dev_ms / K if launches == K else stage_ms["knapsack_dp"]
    roofline = dp_roofline(cells_mean, cells_eval_mean, dp_ms, smem_peak)
    roofline["avg_launch_ms_source"] = ("CUDA events of the timed region (value loop), one kernel per step" if launches == K
                                        else "library stage events of the host-buffer call")

    if not args.skip_extras:
        extras["dp_csp"] = bench_csp(torch, [local_rank], args, smem_peak, flush_l2)
        extras["spmv_milpblock"] = bench_spmv(torch, dev, local_rank, hbm_peak, peak_src, flush_l2, world, rank,
                                              max_over_ranks, barrier)
    cpu = None
    if not args.skip_extras:
        cpu = cpu_baseline(model, duals)

    cfg = workload_config(model)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "run": {"cells_per_round": cells_mean, "cells_evaluated_per_round": cells_eval_mean,
                "kept_columns_round0": kept0, "parallelism": "1 GPU",
                "l2": "flushed between steps (256 MiB written then 256 MiB read, outside the per-step CUDA-event pairs)"},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches,
        "wall_s_device_loop": wall_dev, "stage_ms": stage_ms, "roofline": roofline,
        "value_no_dual_support": {"value": Kd / (dense_dev_ms * 1e-3), "ms_per_step": dense_dev_ms / Kd},
    }
    if "dp_csp" in extras:
        line["roofline_dp_csp"] = extras["dp_csp"]
        line["dp_gcups_csp"] = extras["dp_csp"]["gcups"]
        line["dp_csp_frac_smem_16B"] = extras["dp_csp"]["frac_16B"]
    if "spmv_milpblock" in extras:
        line["roofline_spmv"] = extras["spmv_milpblock"]
        line["spmv_gbs"] = extras["spmv_milpblock"]["achieved"]
        line["spmv_frac"] = extras["spmv_milpblock"]["frac"]
    if "batched" in extras:
        line["batched_rounds"] = extras["batched"]
        for nm in ("batch16", "batch64", "ladder8"):
            if nm in extras["batched"]:
                line[nm + "_rounds_per_s"] = extras["batched"][nm]["e2e_rounds_per_s"]
    if "pool" in extras:
        line["roofline_pool"] = extras["pool"]
    if "mckp" in extras:
        line["roofline_mckp"] = extras["mckp"]
    if "bigdp" in extras:
        line["roofline_dp_big"] = extras["bigdp"]
        line["dp_gcups_cap1e5"] = extras["bigdp"]["gcups"]
    if "intknap" in extras:
        line["roofline_intknap"] = extras["intknap"]
    if "other_shapes" in extras:
        line["other_shapes"] = extras["other_shapes"]
        line["gap_d_e2e_rounds_per_s"] = extras["other_shapes"]["gap_d_20x200"]["e2e_rounds_per_s"]
        line["gap12_e2e_rounds_per_s"] = extras["other_shapes"]["gap12_class_c_10x60"]["e2e_rounds_per_s"]
    if extras.get("real_duals"):
        line["real_duals"] = extras["real_duals"]
        for nm, r in extras["real_duals"].items():
            line["real_duals_" + nm + "_e2e_rounds_per_s"] = r["e2e_rounds_per_s"]
    if expand is not None:
        line["expand_columns"] = expand
    if cpu is not None:
        line["cpu_baseline"] = cpu
    pr.close()
    return line


def dp_roofline(cells_mean, cells_eval_mean, dp_ms, smem_peak):
    """Dominant kernel of the step = knap_dp_kernel.  The fused kernel first drops the items no optimal solution
    can contain and evaluates `cells_evaluated` of the reference DP's cells: `frac` is the fraction of the
    shared-memory roofline the cells it TOUCHES account for (24 B each) -- the honest figure for a
    kernel that is a chain of latencies; `algorithmic_*` count the reference's cells against the same time."""
    ok = dp_ms > 0 and smem_peak
    ach = cells_eval_mean * 24.0 / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None
    alg = cells_mean * 24.0 / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None
    return {
        "kernel": "knap_dp_kernel", "bound": "smem", "achieved": ach, "peak": smem_peak, "unit": "GB/s",
        "frac": ach / smem_peak if ok else None,
        "cells_evaluated_per_launch": cells_eval_mean, "cells_per_launch": cells_mean,
        "algorithmic_achieved": alg, "algorithmic_frac": alg / smem_peak if ok else None,
        "algorithmic_gcups": cells_mean / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None,
        "traffic": ncu_traffic("prof_round_gape_dev"),
        "traffic_note": "DRAM bytes per launch, ncu --set full capture of this kernel on this shape "
                        "(profiles/ncu_traffic.json); the bound is shared memory, HBM only sees inputs",
        "peak_source": "in-library shared-memory copy microbenchmark (dipgpu_measure_smem_gbs), this run",
        "note": "80 six-warp CTAs on 148 SMs: the GAP-E round is LATENCY-bound (staging, reduction pre-pass, ~70 DP "
                "steps, backtrack, in-kernel filter), not roofline-bound; roofline_dp_csp is the bandwidth-bound "
                "shape of the same kernel",
        "avg_launch_ms": dp_ms,
    }


def multi_gpu_line(args, torch, devices, model, duals, replicas):
    """Rank 0 at N > 1 (or --devices): ONE caller, N GPUs, through one multi-device context."""
    from dip_b200.pricer import GpuPricer, RoundResult, unpack_results
    from dip_b200 import capi
    N = len(devices)
    K, W = args.steps, args.warmup
    dev0 = torch.device("cuda", devices[0])
    flush = Flusher(torch, devices)
    flush1 = Flusher(torch, devices[:1])
    pm = GpuPricer(devices)
    pm.setModel(model)
    pm.setDualSupport(model.dual_support)
    p1 = GpuPricer(devices[0])          # the same calls on ONE GPU, same run, same clock: the scaling baseline
    p1.setModel(model)
    p1.setDualSupport(model.dual_support)
    pone = GpuPricer(devices[:1])       # the multi-device code path over one device: same timing method
    pone.setModel(model)
    pone.setDualSupport(model.dual_support)
    pf = GpuPricer(devices)             # the blocks FORCED onto all devices (the default policy keeps a round whose
    pf.set_option("shard_rounds", N)    # blocks are co-resident on one GPU on the primary device)
    pf.setModel(model)
    pf.setDualSupport(model.dual_support)
    n_dev, peer = pm.deviceCount()
    ranges = pm.shardRanges()
    ranges_forced = pf.shardRanges()
    u_dev = [torch.from_numpy(u).to(dev0) for u in duals[:N_DUALS]]
    u_pin = capi.PinnedArray((N_BATCH, model.u_len), np.float64)
    for k, u in enumerate(duals):
        u_pin.array[k] = u
    ptrs = [u_pin.array[k].ctypes.data for k in range(N_DUALS)]
    res = RoundResult(model.m, model.m, model.N)
    sup_bytes = 8 * len(model.dual_support)

    class _DevBytes:
        def __init__(self, p, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p, False), "version": 2}

    # ---- value: ONE dual vector per step, device-resident (duals in the primary's HBM, results gathered in the
    # primary's HBM), the blocks partitioned; every device brackets its part with CUDA events on its own stream,
    # the step costs the slowest device.
    def dev_loop(p, fl, steps, warm):
        p.set_option("round_timing", 1)
        tot, per_dev = 0.0, None
        for k in range(warm + steps):
            fl(sync=True)
            p.priceRoundDev(u_dev[k % N_DUALS].data_ptr(), model.u_len, redCostEps=RC_EPS)
            p.sync()
            ms, mx = p.lastDeviceMs()
            if k >= warm:
                tot += mx
                per_dev = np.array(ms) if per_dev is None else per_dev + np.array(ms)
        p.set_option("round_timing", 0)
        return tot, (per_dev / steps).tolist()

    c0 = pm.counters()
    with ClockSampler(devices[0]) as clk:
        dev_ms, per_dev_ms = dev_loop(pm, flush, K, W)
    c1 = pm.counters()
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    one_ms, _ = dev_loop(pone, flush1, min(K, 100), W)
    one_ms = one_ms / min(K, 100)
    forced_ms, forced_per_dev = dev_loop(pf, flush, min(K, 100), W)
    forced_ms = forced_ms / min(K, 100)
    # merged == unsharded, on the last dual vector priced (default policy and forced partition)
    full = None
    same_all = []
    for p, k_last in ((pm, (W + K - 1) % N_DUALS), (pf, (W + min(K, 100) - 1) % N_DUALS)):
        ptr, nbytes = p.resultDev()
        off, siz = p.resultShards()
        host = torch.as_tensor(_DevBytes(ptr, nbytes), device=dev0).cpu().numpy()
        merged = unpack_results(host, off, siz, model.m, model.N)
        full = p1.priceRound(duals[k_last], redCostEps=RC_EPS, result=RoundResult(model.m, model.m, model.N))
        same_all.append(bool(merged.nCols == full.nCols and np.array_equal(merged.blockObj, full.blockObj)
                             and np.array_equal(merged.blockRedCost, full.blockRedCost)
                             and merged.mostNegReducedCost == full.mostNegReducedCost
                             and np.array_equal(merged.colInd[:merged.c.nInd], full.colInd[:full.c.nInd])))
    same = all(same_all)
    result_bytes = int(sum(int(x) for x in siz[1:]))   # (forced partition) upper bound: the shards' whole packed buffers
    kept_bytes = int(merged.c.nInd) * 4 + model.m * 44
    value = K / (dev_ms * 1e-3)

    # ---- e2e: the same round through the host-buffer call (pinned u in, merged columns out), wall clock
    def e2e_loop(p, fl, steps, warm):
        for k in range(warm):
            p.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        a = p.counters()
        tot = 0.0
        for k in range(steps):
            fl(sync=True)
            t0 = time.perf_counter()
            p.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
            tot += time.perf_counter() - t0
        b = p.counters()
        return tot, (b["h2d_bytes"] - a["h2d_bytes"]) // steps, (b["d2h_bytes"] - a["d2h_bytes"]) // steps

    e2e_s, h2d_b, d2h_b = e2e_loop(pm, flush, K, W)
    e2e1_s, _, _ = e2e_loop(p1, flush1, min(K, 100), W)
    e2ef_s, h2df_b, d2hf_b = e2e_loop(pf, flush, min(K, 100), W)
    e2e = {"value": K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
           "ms_per_step": 1e3 * e2e_s / K, "one_gpu_ms_per_step": 1e3 * e2e1_s / min(K, 100),
           "dual_support_rows": int(len(model.dual_support)),
           "timing": "host wall clock around the synchronous dipgpu_price_round call on the multi-device context: "
                     "every GPU reads u[support] out of the caller's pinned vector (PCIe) and mirrors its part of the "
                     "packed result into mapped host memory; h2d / d2h bytes are summed over the GPUs"}

    # ---- per-stage device times (slowest device per stage) and the DP roofline per device
    pm.set_option("stage_timing", 1)
    stage, cells, cells_ev = {}, [], []
    for k in range(min(K, 60)):
        flush(sync=True)
        pm.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        for nm, ms in pm.last_stage_ms().items():
            stage.setdefault(nm, []).append(ms)
        cells.append(res.cells)
        cells_ev.append(res.cellsEvaluated)
    pm.set_option("stage_timing", 0)
    stage_ms = {nm: float(np.mean(v)) for nm, v in stage.items()}
    smem_peak = p1.measureSmemGBs() if not args.skip_extras else None
    roofline = dp_roofline(float(np.mean(cells)), float(np.mean(cells_ev)), stage_ms["knapsack_dp"], smem_peak * N if smem_peak else None)
    roofline["note"] = f"{N} GPUs, each with 1/{N} of the 80 blocks: the slowest device's DP stage against {N} x the shared-memory peak; " + roofline["note"]

    # ---- the calls that scale: several dual vectors per call, partitioned over the GPUs; same call on one GPU beside it
    batched = None
    if not args.skip_extras:
        b_multi = bench_batched(pm, model, duals, u_pin, flush, torch, None, big=N_BIG)
        b_one = bench_batched(p1, model, duals, u_pin, flush1, torch, None, big=N_BIG)
        batched = {}
        for nm in ("batch16", "batch64", f"batch{N_BIG}", "ladder8"):
            batched[nm] = {"rounds_per_call": b_multi[nm]["rounds_per_call"],
                           "rounds_per_s": b_multi[nm]["e2e_rounds_per_s"], "ms_per_call": b_multi[nm]["ms_per_call"],
                           "one_gpu_rounds_per_s": b_one[nm]["e2e_rounds_per_s"], "one_gpu_ms_per_call": b_one[nm]["ms_per_call"],
                           "speedup_vs_one_gpu": b_multi[nm]["e2e_rounds_per_s"] / b_one[nm]["e2e_rounds_per_s"],
                           "efficiency": b_multi[nm]["e2e_rounds_per_s"] / b_one[nm]["e2e_rounds_per_s"] / N,
                           "stage_ms_slowest_device": b_multi[nm]["stage_ms"]}
        batched["timing"] = (f"host wall clock around the synchronous C-ABI call ({N_BIG} / 64 / 16 pinned dual vectors, or one dualRM "
                             "and an 8-step smoothing ladder), duals in and columns out inside, L2 of every GPU flushed before "
                             "each call; efficiency = speed-up over the same call on one GPU / number of GPUs")
    csp = None
    if not args.skip_extras:
        csp = bench_csp(torch, devices, args, smem_peak, flush)
    for p in (pm, p1, pone, pf):
        p.close()
    cfg = workload_config(model)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "run": {"parallelism": f"ONE caller (rank 0), one multi-device context over {N} GPUs.  value / e2e: ONE dual vector per "
                               "step under the library's default policy -- the 80 blocks of a GAP-E round are 80 one-CTA latency "
                               "chains that run side by side on one GPU's 148 SMs, so the round stays on the primary device "
                               "(`sharded`: the same round with the blocks forced onto all GPUs); the calls that do scale "
                               "(several dual vectors per call, partitioned over the GPUs) are `batched_rounds`",
                "blocks_per_gpu": [c for _, c in ranges],
                "l2": "every GPU flushed between steps (256 MiB written then 256 MiB read, outside the per-device CUDA-event pairs)",
                "timing": "per step: each GPU's CUDA-event pair on its own stream around its part, the step costs the slowest GPU"},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "stage_ms": stage_ms, "roofline": roofline,
        "sharded": {"default_policy_ms_per_round": dev_ms / K, "default_policy_blocks_per_gpu": [c for _, c in ranges],
                    "ms_per_round": forced_ms, "ms_per_round_per_gpu": forced_per_dev,
                    "blocks_per_gpu": [c for _, c in ranges_forced],
                    "e2e_ms_per_round": 1e3 * e2ef_s / min(K, 100),
                    "one_gpu_same_method_ms": one_ms, "vs_one_gpu": one_ms / forced_ms,
                    "default_policy_vs_one_gpu": one_ms / (dev_ms / K),
                    "merged_equals_unsharded": same, "peer_access": bool(peer),
                    "exchange": "duals: every GPU pulls u[support] out of the primary's HBM (peer loads over NVLink); results: "
                                "every fused DP kernel mirrors its packed result into the primary's gather buffer (peer stores)",
                    "nvlink_bytes_per_round": {"duals": (N - 1) * sup_bytes, "results_max": result_bytes, "results_kept_about": kept_bytes},
                    "pcie_bytes_per_round_e2e": {"h2d": int(h2df_b), "d2h": int(d2hf_b)},
                    "note": "forced partition (option shard_rounds = N): duals pulled from the primary's HBM over NVLink, "
                            "results mirrored into the primary's gather buffer by peer stores; a remote device pays the "
                            "NVLink round trips on top of the same per-block latency chain, which is why the default "
                            "policy does not partition co-resident rounds"},
        "sharded_ms_per_round": forced_ms, "sharded_vs_one_gpu": one_ms / forced_ms,
        "round_default_policy_ms": dev_ms / K, "round_default_policy_vs_one_gpu": one_ms / (dev_ms / K),
    }
    if batched is not None:
        line["batched_rounds"] = batched
        for nm in ("batch16", "batch64", f"batch{N_BIG}", "ladder8"):
            line[nm + "_rounds_per_s"] = batched[nm]["rounds_per_s"]
            line[nm + "_speedup_vs_one_gpu"] = batched[nm]["speedup_vs_one_gpu"]
    if csp is not None:
        line["roofline_dp_csp"] = csp
        line["dp_gcups_csp"] = csp["gcups"]
        line["dp_csp_frac_smem_16B"] = csp["frac_16B"]
    if replicas is not None:
        line["replicas"] = replicas
    return line


def bench_batched(pr, model, duals, u_pin, flush_l2, torch, smem_peak, big=0):
    """Several dual vectors per submission: dipgpu_price_rounds_batch on 16 and on 64 pinned dual
    vectors, and the Wentges ladder of 8 smoothing parameters from ONE uploaded dualRM
    (dipgpu_price_round_stab_ladder).  Host wall clock around the synchronous C-ABI calls, H2D and D2H
    inside, L2 flushed before every call; stage times from the library's CUDA events (a multi-device context
    reports the slowest device per stage)."""
    from dip_b200 import capi
    out = {}
    res, arr = pr._batch_results(max(N_BATCH, big))
    legs = [("batch16", 16), ("batch64", N_BATCH), ("ladder8", 8)]
    u_big = None
    if big > N_BATCH:
        # a batch large enough that every one of 8 GPUs still gets tens of vectors: the N_BATCH dual vectors repeated
        u_big = capi.PinnedArray((big, model.u_len), np.float64)
        for k in range(big):
            u_big.array[k] = u_pin.array[k % N_BATCH]
        legs.append((f"batch{big}", big))
    for name, steps in legs:
        reps = 30 if steps <= N_BATCH else 8
        src = u_big if steps > N_BATCH else u_pin

        def call():
            if name == "ladder8":
                pr.priceRoundStabLadderRaw(u_pin.array[1].ctypes.data, model.u_len, 0.1, 0.90, 8, capi.PHASE_PRICE2,
                                           RC_EPS, arr)
            else:
                pr.priceRoundsBatchRaw(steps, src.array.ctypes.data, model.u_len, capi.PHASE_PRICE2, RC_EPS, arr)
        if name == "ladder8":
            pr.stabSetCenter(duals[0])
        for _ in range(3):
            call()
        pr.set_option("stage_timing", 1)
        t, stage = 0.0, {}
        for _ in range(reps):
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            call()
            t += time.perf_counter() - t0
            for nm, ms in pr.last_stage_ms().items():
                stage.setdefault(nm, []).append(ms)
        pr.set_option("stage_timing", 0)
        st = {nm: float(np.mean(v)) for nm, v in stage.items()}
        cells = float(sum(arr[k].cells for k in range(steps)))
        cells_ev = float(sum(arr[k].cellsEvaluated for k in range(steps)))
        dp_ms = st["knapsack_dp"]
        ach = cells_ev * 24.0 / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None
        out[name] = {
            "rounds_per_call": steps, "e2e_rounds_per_s": steps * reps / t, "ms_per_call": 1e3 * t / reps,
            "stage_ms": st, "dp_cells_per_call": cells, "dp_cells_evaluated_per_call": cells_ev,
            "dp_roofline": {"kernel": "knap_dp_kernel (fused, gridDim.y = vectors)", "bound": "smem", "achieved": ach,
                            "peak": smem_peak, "unit": "GB/s", "frac": ach / smem_peak if (ach and smem_peak) else None,
                            "note": "achieved / frac count the cells the kernel evaluates after its reduction pre-pass (24 B "
                                    "each); algorithmic_gcups counts the reference DP's cells against the same time",
                            "algorithmic_gcups": cells / (dp_ms * 1e-3) / 1e9 if dp_ms > 0 else None},
        }
    out["note"] = ("value / e2e above stay ONE dual vector per step (what DecompAlgo::generateVars prices per call); "
                   "these are the same rounds submitted several at a time")
    return out


def _time_shape(torch, dev, local_rank, model, dual_list, support, flush_l2, steps, warm=5):
    """Rounds of `model` under the dual vectors `dual_list` (cycled): device-resident (CUDA events on the launching
    stream, like `value`) and through the host-buffer call with pinned duals (wall clock, like `e2e`)."""
    from dip_b200.pricer import GpuPricer, RoundResult
    from dip_b200 import capi
    pr = GpuPricer(local_rank)
    pr.setModel(model)
    if support is not None:
        pr.setDualSupport(support)
    stream = torch.cuda.current_stream(dev)
    nd = len(dual_list)
    u_dev = [torch.from_numpy(np.ascontiguousarray(u)).to(dev) for u in dual_list]
    u_pin = capi.PinnedArray((nd, model.u_len), np.float64)
    for k, u in enumerate(dual_list):
        u_pin.array[k] = u
    ptrs = [u_pin.array[k].ctypes.data for k in range(nd)]
    res = RoundResult(model.m, model.m, model.N)
    kept, cells, cells_ev = [], [], []
    for k in range(nd):
        r = pr.priceRound(dual_list[k], redCostEps=RC_EPS, result=res)
        kept.append(r.nCols)
        cells.append(r.cells)
        cells_ev.append(r.cellsEvaluated)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(warm + steps):
        flush_l2()
        if k >= warm:
            ev[k - warm][0].record(stream)
        pr.priceRoundDev(u_dev[k % nd].data_ptr(), model.u_len, redCostEps=RC_EPS, stream=stream.cuda_stream)
        if k >= warm:
            ev[k - warm][1].record(stream)
    torch.cuda.synchronize()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    tot = 0.0
    for k in range(warm + steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pr.priceRoundRaw(ptrs[k % nd], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        if k >= warm:
            tot += time.perf_counter() - t0
    geom = pr.dpGeometry() if hasattr(pr, "dpGeometry") else None
    pr.close()
    u_pin.free()
    return {"blocks": int(model.m), "cols": int(model.N), "rows": int(model.R), "nnz": int(model.nnz), "dual_vectors": nd,
            "device_ms_per_round": dev_ms, "device_rounds_per_s": 1e3 / dev_ms,
            "e2e_ms_per_round": 1e3 * tot / steps, "e2e_rounds_per_s": steps / tot,
            "kept_columns_mean": float(np.mean(kept)), "cells_per_round": float(np.mean(cells)),
            "cells_evaluated_per_round": float(np.mean(cells_ev))}


def bench_other_shapes(torch, dev, local_rank, flush_l2):
    """The other GAP shapes BASELINE.json names, same method as value / e2e: GAP class D 20x200 (configs[1]) and the
    gap12-sized class C 10x60 (configs[0]; the OR-Library file is not in the reference tree, so a class-C instance of
    that size is generated), each under 16 synthetic dual vectors of one node (1 % of the branch rows bounded)."""
    out = {}
    for name, g in (("gap_d_20x200", I.gap_class_d()), ("gap12_class_c_10x60", I.gap_class_c())):
        model = I.gap_pricing_model(g)
        rng = np.random.default_rng(20200 if "d_" in name else 1060)
        br = I.gap_branch_rows(model, rng)
        duals = [I.gap_duals(model, rng, branch_rows=br) for _ in range(16)]
        out[name] = _time_shape(torch, dev, local_rank, model, duals, I.gap_dual_support(model, br), flush_l2, 100)
    return out


def bench_real_duals(torch, dev, local_rank, flush_l2):
    """Replay of RECORDED master duals beside the synthetic ones: the dual vectors of real column-generation
    runs (scripts/record_dual_trace.py: examples/gap_colgen.py's loop, restricted master in HiGHS) -- the whole
    root-node run of GAP-D 20x200 (every 4th round) and the first rounds of GAP-E 80x1600 -- scattered into DIP's
    master layout (root node: every branch row free, zero duals) and priced one vector per call."""
    out = {}
    for name, g in (("gap_d_20x200", I.gap_class_d()), ("gap_e_80x1600", I.gap_class_e())):
        path = os.path.join(ROOT, "tests", "golden", "dual_trace_" + ("gap_d" if "gap_d" in name else "gap_e") + ".npz")
        if not os.path.exists(path):
            continue
        tr = np.load(path)
        model = I.gap_pricing_model(g)
        n, m = int(tr["n"]), int(tr["m"])
        assert n == model.meta["n"] and m == model.m
        U = []
        for row in tr["duals"]:
            u = np.zeros(model.u_len)
            u[:n] = row[:n]
            u[model.n_base_rows:model.n_base_rows + m] = row[n:n + m]
            U.append(u)
        support = np.concatenate([np.arange(n), model.n_base_rows + np.arange(m)]).astype(np.int32)
        r = _time_shape(torch, dev, local_rank, model, U, support, flush_l2, min(100, 4 * len(U)))
        r["trace"] = {"rounds_recorded": int(tr["rounds"][-1]) + 1, "vectors_replayed": len(U),
                      "run_converged": bool(tr["converged"]), "z_rmp_last": float(tr["z_rmp"][-1]),
                      "lower_bound_last": float(tr["lower_bound"][-1])}
        out[name] = r
    return out


def bench_pool(torch, local_rank, hbm_peak, peak_src, flush_l2, rank, pool_kernel=None, options=None):
    """Waiting-column re-pricing (DecompVarPool::setReducedCosts) on a synthetic pool:
    2*10^5 master columns of 20..80 entries over the GAP-E master rows.  Kernel time from the library's
    CUDA events (L2 flushed), algorithmic bytes 12*nnz + 24*cols + 8*uLen."""
    from dip_b200.pricer import GpuPricer
    rng = np.random.default_rng(77)
    # master columns shaped like GAP-E's: for the ~16 jobs a machine takes, the assignment row j, the two
    # branch rows of x[b,j] (x <= u, x >= l) and the block's convexity row -- ascending master rows
    m, n = 80, 1600
    N = m * n
    n_base = n + 2 * N
    u_len = n_base + m
    n_cols = 200_000
    per = 16
    blk = rng.integers(0, m, size=n_cols)
    jobs = np.sort(rng.integers(0, n, size=(n_cols, per)), axis=1)
    x = blk[:, None] * n + jobs
    rows = np.concatenate([jobs, n + x, n + N + x, (n_base + blk)[:, None]], axis=1)
    ri = rows.astype(np.int32).ravel()
    lens = np.full(n_cols, 3 * per + 1)
    cs = np.zeros(n_cols + 1, np.int64)
    np.cumsum(lens, out=cs[1:])
    nnz = int(cs[-1])
    va = np.ones(nnz)
    oc = rng.uniform(0, 1000, size=n_cols)
    # duals of one node: zero on the free branch rows (99 % of them)
    bounded = np.sort(rng.choice(2 * N, size=2 * N // 100, replace=False))
    support = np.concatenate([np.arange(n), n + bounded, np.arange(n_base, u_len)]).astype(np.int32)
    u = np.zeros(u_len)
    u[support] = rng.uniform(0, 50, size=len(support))
    pr = GpuPricer(local_rank)
    if pool_kernel is not None:
        pr.set_option("pool_kernel", pool_kernel)
    for k_, v_ in (options or {}).items():
        pr.set_option(k_, v_)
    pr.setModelCore(u_len - 1, 1, np.zeros(u_len, np.int64), np.zeros(0, np.int32), np.zeros(0), u_len - 1, 1)
    pr.poolAppend(cs, ri, va, oc)
    by = 12 * nnz + 24 * n_cols + 8 * u_len

    def timed():
        pr.poolSetReducedCosts(u)
        pr.set_option("stage_timing", 1)
        ts = []
        for _ in range(12):
            flush_l2()
            torch.cuda.synchronize()
            pr.poolSetReducedCosts(None)
            ts.append(pr.last_stage_ms()["redcost"])
        pr.set_option("stage_timing", 0)
        return float(np.mean(ts[2:]))

    ms_plain = timed()
    pr.setDualSupport(support)       # the kernel then skips the gathers of the rows that cannot carry a dual
    ms = timed()
    pr.close()
    ach = by / (ms * 1e-3) / 1e9
    # with the support declared the pass reads a compacted copy of the pool (entries of rows that can carry a dual):
    # what it really streams
    in_support = np.zeros(u_len, bool)
    in_support[support] = True
    nnz_eval = int(in_support[ri].sum())
    by_eval = 12 * nnz_eval + 24 * n_cols + 8 * len(support)
    return {"entries_evaluated": nnz_eval, "evaluated_bytes": by_eval, "frac_evaluated": by_eval / (ms * 1e-3) / 1e9 / hbm_peak,
            "kernel": "pool_redcost_kernel<BITMAP>" if pool_kernel == 0 else "pool_redcost_persist_kernel<0, 256> over the compacted copy",
            "workload": f"{n_cols} waiting columns shaped like GAP-E master columns, {nnz} entries, {u_len} master rows, "
                        f"{len(support)} of them with a dual (one node: 1 % of the branch rows bounded)",
            "note": "every stored entry gathers one dual (8 B of a 32 B sector, L2-resident): the sector traffic, not the "
                    "12 B/entry stream, bounds this pass; with the dual support declared the gathers of rows without a "
                    "dual are left out altogether: the pass streams a compacted copy of the pool (they add +-0 to a sum that is "
                    "never -0: same bits); `achieved` / `frac` count the whole pool's 12 B/entry against the time, "
                    "`frac_evaluated` only what the kernel streams; persistent CTAs, the window ring streams across tile borders",
            "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic("prof_pool_bitmap"),
            "avg_launch_ms": ms, "without_dual_support": {"avg_launch_ms": ms_plain, "achieved": by / (ms_plain * 1e-3) / 1e9,
                                                          "frac": by / (ms_plain * 1e-3) / 1e9 / hbm_peak},
            "algorithmic_bytes": by, "peak_source": peak_src}


def bench_mckp(torch, local_rank, smem_peak, flush_l2):
    """Multiple-choice knapsack blocks (the MMKP example's MCKP0 relaxation): 2 000 blocks of
    25 groups x 10 columns, capacity 1 000 -- cells = columns x (capacity + 1); shared-memory bytes per cell:
    8 read (the previous group's row) + 8/K written (the new row, once per group of K columns)."""
    from dip_b200.pricer import GpuPricer
    rng = np.random.default_rng(88)
    m, G, K, cap = 2000, 25, 10, 1000
    gcs = (np.arange(m * G + 1) * K).astype(np.int32)
    bgs = (np.arange(m + 1) * G).astype(np.int32)
    N = m * G * K
    w = rng.integers(1, 81, size=N).astype(np.int32)
    rc = rng.uniform(-50, 50, size=N)
    pr = GpuPricer(local_rank)
    pr.setObjective(rng.uniform(1, 100, size=N))
    pr.setMckpBlocks(bgs, gcs, w, np.full(m, cap, np.int32))
    alpha = np.zeros(m)
    res = pr.solveBlocks(rc, alpha, redCostEps=-np.inf)
    pr.set_option("stage_timing", 1)
    ts = []
    for _ in range(6):
        flush_l2()
        torch.cuda.synchronize()
        pr.solveBlocks(rc, alpha, redCostEps=-np.inf, result=res)
        ts.append(pr.last_stage_ms()["knapsack_dp"])
    pr.close()
    ms = float(np.mean(ts[1:]))
    cells = float(res.cells)
    by_cell = 8.0 + 8.0 / K
    ach = cells * by_cell / (ms * 1e-3) / 1e9
    return {"kernel": "mckp_dp_kernel (+ mckp_backtrack_kernel)", "workload": f"{m} blocks x {G} groups x {K} columns, capacity {cap}",
            "bound": "smem", "cells": cells, "ms": ms, "gcups": cells / (ms * 1e-3) / 1e9, "achieved": ach,
            "peak": smem_peak, "unit": "GB/s", "frac": ach / smem_peak if smem_peak else None,
            "algorithmic_bytes_per_cell": by_cell, "kept_columns": int(res.nCols)}


def bench_bigdp(torch, local_rank, smem_peak, flush_l2, m=60, n=500, cap=100_000):
    """Knapsacks whose DP row exceeds one CTA (VERDICT r1 item 5 / GAP_KnapPisinger.h:243-249: any capacity):
    m knapsacks, n = 500, capacity 10^5, weights U{1..cap/10}.  One 8-CTA cluster per knapsack holds the row in
    distributed shared memory; 15 such clusters are resident on a B200 (ncu: launch__cluster_max_active), m = 60 is
    four full waves."""
    from dip_b200.pricer import GpuPricer
    b = I.csp_batch(m=m, n=n, cap=cap, seed=77, wmax=cap // 10)
    pr = GpuPricer(local_rank)
    pr.setObjective(b.orig_cost)
    pr.setKnapsackBlocks(b.block_col_start, b.weight, b.capacity)
    geo = pr.dpGeometry()
    pr.set_option("stage_timing", 1)
    alpha = np.zeros(m)
    res = pr.solveBlocks(b.red_cost, alpha, redCostEps=-np.inf)
    ts = []
    for _ in range(3):
        flush_l2()
        torch.cuda.synchronize()
        pr.solveBlocks(b.red_cost, alpha, redCostEps=-np.inf, result=res)
        ts.append(pr.last_stage_ms())
    pr.close()
    best = min(ts, key=lambda d: d["knapsack_dp"])
    cells = float(m) * n * (cap + 1)
    assert res.cells == cells, (res.cells, cells)
    dp_ms = best["knapsack_dp"]
    ach = cells * 24.0 / (dp_ms * 1e-3) / 1e9
    return {"kernel": "knap_dp_cluster_kernel<%d>" % geo["cells_per_thread"] if geo["cells_per_thread"] > 0 else "knap_dp_big_kernel",
            "workload": f"{m} knapsacks x n={n} x cap={cap}", "bound": "smem", "gcups": cells / (dp_ms * 1e-3) / 1e9,
            "dp_ms": dp_ms, "backtrack_ms": best["backtrack_emit"], "achieved": ach, "peak": smem_peak, "unit": "GB/s",
            "frac": ach / smem_peak if smem_peak else None, "algorithmic_bytes_per_cell": 24, "cells": cells, "geometry": geo}


def bench_intknap(torch, local_rank, smem_peak, flush_l2, m=2000, n=60, cap=2000):
    """Integer (unbounded) knapsack blocks of cutting-stock pricing (kp of
    Dip/src/dippy/tests/test_cutting_stock.py:154-179): m patterns' pricing problems, n item lengths each, roll
    length `cap`; cells = n x (cap + 1) table entries per block, 16 B each (value + item/source)."""
    from dip_b200.pricer import GpuPricer
    rng = np.random.default_rng(99)
    bcs = (np.arange(m + 1) * n).astype(np.int32)
    w = rng.integers(cap // 40, cap // 3, size=m * n).astype(np.int32)
    rc = -rng.uniform(0.01, 1.0, size=m * n) * w / cap
    pr = GpuPricer(local_rank)
    pr.setObjective(np.zeros(m * n))
    pr.setIntKnapBlocks(bcs, w, np.full(m, cap, np.int32))
    pr.set_option("stage_timing", 1)
    alpha = np.zeros(m)
    res = pr.solveBlocks(rc, alpha, redCostEps=-np.inf)
    ts = []
    for _ in range(5):
        flush_l2()
        torch.cuda.synchronize()
        pr.solveBlocks(rc, alpha, redCostEps=-np.inf, result=res)
        ts.append(pr.last_stage_ms()["knapsack_dp"])
    pr.close()
    ms = float(min(ts[1:]))
    cells = float(m) * n * (cap + 1)
    ach = cells * 16.0 / (ms * 1e-3) / 1e9
    return {"kernel": "intknap_kernel", "workload": f"{m} integer knapsacks x n={n} x cap={cap}", "bound": "smem",
            "cells": cells, "ms": ms, "gcups": cells / (ms * 1e-3) / 1e9, "achieved": ach, "peak": smem_peak, "unit": "GB/s",
            "frac": ach / smem_peak if smem_peak else None, "algorithmic_bytes_per_cell": 16,
            "note": "capacity-sequential, item-parallel table (the recurrence's cell F[j] needs every F[j - w_i]): a latency "
                    "chain of cap + 1 steps per block, many blocks side by side", "kept_columns": int(res.nCols)}


def bench_expand(pr, model, duals, ptrs, res, flush_l2, torch, rank):
    """dipgpu_expand_columns after each of 50 rounds (wall clock around the C-ABI call, D2H inside),
    beside the CPU restatement of DecompAlgo::addVarsToPool's per-column work (single thread, as DIP)."""
    from dip_b200 import capi
    import ctypes as C
    cap_nnz = 1 << 20
    col_start = np.zeros(model.m + 1, np.int64)
    row_ind = np.zeros(cap_nnz, np.int32)
    val = np.zeros(cap_nnz, np.float64)
    hsh = np.zeros(model.m, np.uint64)
    mc = capi.MCols()
    mc.capCols, mc.capNnz = model.m, cap_nnz
    mc.colStart, mc.rowInd, mc.val, mc.hash = capi.ptr(col_start), capi.ptr(row_ind), capi.ptr(val), capi.ptr(hsh)
    lib = capi.lib()
    t_gpu = 0.0
    reps = 50
    for k in range(reps + 3):
        pr.priceRoundRaw(ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, res)
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.dipgpu_expand_columns(pr._ctx, 1e-6, C.byref(mc))
        dt = time.perf_counter() - t0
        assert rc == 0, rc
        if k >= 3:
            t_gpu += dt
    # the round and its expansion in one submission (dipgpu_price_round_expand), same clock
    t_fused = 0.0
    for k in range(reps + 3):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.dipgpu_price_round_expand(pr._ctx, ptrs[k % N_DUALS], model.u_len, capi.PHASE_PRICE2, RC_EPS, 1e-6,
                                           C.byref(res.c), C.byref(mc))
        dt = time.perf_counter() - t0
        assert rc == 0, rc
        if k >= 3:
            t_fused += dt
    out = {"gpu_ms_per_round": 1e3 * t_gpu / reps, "columns": int(mc.nCols), "master_nnz": int(mc.nNnz),
           "timing": "host wall clock around dipgpu_expand_columns (kernel + D2H), L2 flushed",
           "round_plus_expand_fused_ms": 1e3 * t_fused / reps,
           "round_plus_expand_note": "dipgpu_price_round_expand: round and expansion in one submission, one wait; compare "
                                     "with e2e.ms_per_step + gpu_ms_per_round for the two separate calls"}
    if rank == 0:
        from oracle import dip_oracle as orc
        cols = [res.column(k).copy() for k in range(res.nCols)]
        blocks = [int(res.colBlock[k]) for k in range(res.nCols)]
        t0 = time.perf_counter()
        for ind, b in zip(cols, blocks):
            orc.expand_column(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m,
                              ind, b)
        out["cpu_ms_per_round"] = 1e3 * (time.perf_counter() - t0)
        out["cpu"] = "oracle restatement of addVarsToPool per column (fillDenseArr + M->times + dense scan), 1 thread as DIP"
    return out


def bench_csp(torch, devices, args, smem_peak, flush_l2):
    """Knapsack DP GCUPS on the cutting-stock batch (BASELINE.json configs[3]): 10^4 knapsacks, n = 500,
    capacity 10^4.  Several GPUs: ONE dipgpu_solve_blocks call on the multi-device context, the knapsacks
    partitioned over the GPUs (no exchange but the result); DP time = the slowest device's DP stage."""
    from dip_b200.pricer import GpuPricer
    m_total = args.csp_blocks
    b = I.csp_batch(m=m_total)
    n = 500
    ndev = len(devices)
    pr = GpuPricer(devices if ndev > 1 else devices[0])
    pr.setObjective(b.orig_cost)
    pr.setKnapsackBlocks((np.arange(m_total + 1) * n).astype(np.int32), b.weight, b.capacity)
    pr.set_option("stage_timing", 1)
    rc = b.red_cost.copy()
    alpha = np.zeros(m_total)
    res = pr.solveBlocks(rc, alpha, redCostEps=-np.inf)
    times = []
    for rep in range(4):
        flush_l2()
        torch.cuda.synchronize()
        pr.solveBlocks(rc, alpha, redCostEps=-np.inf, result=res)
        times.append(pr.last_stage_ms())
    best = min(times[1:], key=lambda d: d["knapsack_dp"])
    dp_ms, bt_ms = best["knapsack_dp"], best["backtrack_emit"]
    cells = m_total * n * (10_000 + 1)
    assert res.cells == cells, (res.cells, cells)
    pr.close()
    ach = cells * 24.0 / (dp_ms * 1e-3) / 1e9
    peak = (smem_peak * ndev) if smem_peak else None
    return {"kernel": "knap_dp_kernel", "workload": f"CSP batch {m_total} x n=500 x cap=10^4, {ndev} GPU(s), one call",
            "bound": "smem", "gcups": cells / (dp_ms * 1e-3) / 1e9, "dp_ms": dp_ms, "backtrack_ms": bt_ms,
            "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak if peak else None,
            "frac_16B": (cells * 16.0 / (dp_ms * 1e-3) / 1e9 / peak) if peak else None,
            "note": "achieved / frac at the 24 B/cell SURVEY 8d counts as algorithmic (row-in-smem formulation); the kernel keeps "
                    "the cells in registers and moves 16 B/cell through shared memory: frac_16B is the share of the crossbar it uses",
            "algorithmic_bytes_per_cell": 24, "cells": cells,
            # ncu capture is of 1184 knapsacks (8 waves): scaled to one device's block count
            "traffic": ncu_traffic("prof_dp_csp", scale=(m_total / ndev) / 1184.0),
            "peak_source": "in-library shared-memory copy microbenchmark x n_gpus"}


def bench_spmv(torch, dev, local_rank, hbm_peak, peak_src, flush_l2, world, rank, max_over_ranks, barrier):
    """Reduced-cost SpMV sweep on the MILPBlock shape (BASELINE.json configs[4]): 256 blocks,
    10^7-nnz linking rows; every rank sweeps the whole matrix (replicas), rank 0 reports."""
    from dip_b200.pricer import GpuPricer
    model, u = I.milpblock_model()
    pr = GpuPricer(local_rank)
    pr.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
    pr.setObjective(model.objective)
    ud = torch.from_numpy(u).to(dev)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    for _ in range(3):
        pr.calcRedCostDev(ud.data_ptr(), model.u_len, stream=sptr)
    times = []
    for _ in range(20):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        pr.calcRedCostDev(ud.data_ptr(), model.u_len, stream=sptr)
        b.record(stream)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = float(np.mean(times))
    # ---- 25 cut rows appended in place (DecompConstraintSet::appendRow, DecompAlgo.cpp:6074): the rows stay a
    # delta (O(new entries)); beside it the same append with everything re-transposed and re-tiled
    append = None
    if not getattr(bench_spmv, "skip_append", False):
        rng = np.random.default_rng(25)
        k, nn = 25, 2000
        r2 = np.sort(rng.integers(0, k, size=nn))
        rs2 = np.zeros(k + 1, np.int64)
        np.cumsum(np.bincount(r2, minlength=k), out=rs2[1:])
        ci2 = rng.integers(0, model.N, size=nn).astype(np.int32)
        v2 = rng.uniform(-5, 5, size=nn)
        u2 = np.concatenate([u, rng.standard_normal(k)])   # master layout: [base rows | convexity | cut rows], new cuts last
        out = {}
        for label, min_nnz in (("delta", 0), ("full_rebuild", 1 << 40)):
            q = GpuPricer(local_rank)
            q.set_option("append_delta_min_nnz", min_nnz)
            q.setModelCore(model.R, model.N, model.row_start, model.col_ind, model.val, model.n_base_rows, model.m)
            q.setObjective(model.objective)
            q.calcRedCost(u)                      # tiles built, as in a running column generation
            t0 = time.perf_counter()
            q.appendCoreRows(rs2, ci2, v2)
            t_app = time.perf_counter() - t0
            t0 = time.perf_counter()
            rc_first = q.calcRedCost(u2)          # the first sweep after the append (the rebuild re-tiles here)
            t_first = time.perf_counter() - t0
            t0 = time.perf_counter()
            q.calcRedCost(u2)                     # steady state of the same host call (8 MB of reduced costs back)
            t_second = time.perf_counter() - t0
            u2d = torch.from_numpy(u2).to(dev)
            ts = []
            for _ in range(8):
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                q.calcRedCostDev(u2d.data_ptr(), len(u2), stream=sptr)
                b.record(stream)
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out[label] = {"append_call_ms": 1e3 * t_app, "first_sweep_call_ms": 1e3 * t_first, "next_sweep_call_ms": 1e3 * t_second, "sweep_ms": float(np.mean(ts[1:])), "_rc": rc_first}
            q.close()
        same = bool(np.array_equal(out["delta"].pop("_rc"), out["full_rebuild"].pop("_rc")))
        append = {"rows": k, "new_entries": nn, "bit_equal_to_full_rebuild": same, **out}
    pr.close()
    by = model.spmv_bytes()
    ach = by / (ms * 1e-3) / 1e9
    return {"kernel": "redcost_stream_kernel", "workload": model.name, "bound": "hbm", "achieved": ach, "append_cut_rows": append,
            "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": ncu_traffic("prof_spmv_milp"),
            "avg_launch_ms": ms,
            "min_launch_ms": float(np.min(times)), "algorithmic_bytes": by, "peak_source": peak_src,
            "l2": "flushed before every sweep (256 MiB written then 256 MiB read; inputs 141 MB vs 126 MB L2)"}


if __name__ == "__main__":
    main()

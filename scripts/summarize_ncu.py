#!/usr/bin/env python3
"""Turn gpurun_out/*.ncu-rep and launches.csv into the tracked summaries under profiles/.
    python scripts/summarize_ncu.py r01
Reads with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def raw(rep):
    if rep.endswith(".csv"):   # exported on the GPU box (scripts/gpu_bench_profile.sh)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    lines = [f"# ncu summaries, round {tag[1:]} (from gpurun_out/*.ncu-rep; `--set full --clock-control none`)\n"]
    for name in sorted(os.listdir(OUT)):
        if not (name.endswith(".ncu-rep") or name.endswith(".raw.csv")):
            continue
        recs, units = raw(os.path.join(OUT, name))
        for r in recs:
            lines.append(f"## {name}: {r.get('Kernel Name', '?')}\n")
            lines.append("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in r and r[k] != "":
                    lines.append(f"| {k} | {r[k]} | {units.get(k, '')} |")
            stalls = sorted(((float(v), k) for k, v in r.items()
                             if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v),
                            reverse=True)[:6]
            tot = sum(s for s, _ in stalls) or 1.0
            lines.append("")
            lines.append("top stall reasons (pc samples): " + ", ".join(
                f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {int(s)}" for s, k in stalls))
            lines.append("")
    open(os.path.join(PROF, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
    # DRAM traffic per launch of the captured kernels (bench.py reports it as roofline.traffic)
    import json
    traffic = {}
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for name in sorted(os.listdir(OUT)):
        if not (name.endswith(".ncu-rep") or name.endswith(".raw.csv")):
            continue
        recs, units = raw(os.path.join(OUT, name))
        for r in recs:
            try:
                rd = float(r["dram__bytes_read.sum"]) * unit[units["dram__bytes_read.sum"]]
                wr = float(r["dram__bytes_write.sum"]) * unit[units["dram__bytes_write.sum"]]
            except Exception:
                continue
            traffic[name.replace(".ncu-rep", "").replace(".raw.csv", "")] = {
                "kernel": r.get("Kernel Name", "?"), "dram_bytes_read": rd, "dram_bytes_write": wr,
                "grid": r.get("launch__grid_size"), "duration": r.get("gpu__time_duration.sum"),
                "duration_unit": units.get("gpu__time_duration.sum")}
    json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    # launch list
    lc = os.path.join(OUT, "launches.csv")
    if os.path.exists(lc):
        txt = [l for l in open(lc) if l.startswith('"')]
        rows = list(csv.DictReader(io.StringIO("".join(txt))))
        agg = {}
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"])
            u = r["Metric Unit"]
            us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
            k = r["Kernel Name"].split("(")[0]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += us
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(PROF, f"{tag}_launches_summary.md"), "w") as f:
            f.write(f"# Launch list of `python bench.py --steps 5 --warmup 3 --skip-extras` under ncu "
                    f"(`--metrics gpu__time_duration.sum --clock-control none`), round {tag[1:]}\n\n"
                    "Times are cold-cache and serialised: compare SHARES, not absolutes.  The `at::` kernels are the "
                    "bench's L2 flush (256 MiB written, 256 MiB read) between steps, OUTSIDE the per-step CUDA-event pairs; "
                    "the last column is the share among the library's own kernels.  With the dual support declared a round "
                    "is ONE `knap_dp_kernel` launch (the `value` and `e2e` loops); the `redcost_stream_kernel` launches belong "
                    "to the `no_dual_support` legs and the parity check in front of the timed loops.\n\n"
                    "| kernel | launches | total us | mean us | share | share of dipgpu kernels |\n|---|---|---|---|---|---|\n")
            own = sum(a[1] for k, a in agg.items() if "dipgpu::" in k) or 1.0
            for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                mine = f"{100 * us / own:.1f}%" if "dipgpu::" in k else "-"
                f.write(f"| {k} | {n} | {us:.1f} | {us / n:.2f} | {100 * us / tot:.1f}% | {mine} |\n")
        import shutil
        shutil.copy(lc, os.path.join(PROF, f"{tag}_launches.csv"))
    for fn in ("bench_1gpu.json", "bench_ref.json"):
        p = os.path.join(OUT, fn)
        if os.path.exists(p):
            import shutil
            shutil.copy(p, os.path.join(PROF, f"{tag}_{fn}"))
    print("wrote", sorted(os.listdir(PROF)))


if __name__ == "__main__":
    main()

# Round-end measurement recipe (run under gpurun): bench line, reference arm, ncu launch list of
# the bench step, and one --set full capture per hot kernel.  Every ncu run follows a plain run
# of the same command that exited 0.  TAG names the round (r02): files land in gpurun_out/, the
# tracked summaries are made from them by scripts/summarize_ncu.py TAG.
TAG=${TAG:-r02}
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 300 --warmup 5 > gpurun_out/bench_ref.json 2>&1
python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
cap() { # name, kernel regex, skip, command...
  name=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/p_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/n_$name.log 2>&1; echo "ncu $name rc=$?"
  # gpurun brings back at most 64 MiB: keep the raw page (all metrics) as CSV, and the per-line source page of
  # the top kernel; the .ncu-rep itself stays on the box
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/prof_$name.raw.csv 2>/dev/null
  if [ "$name" = round_gape_dev ] || [ "$name" = pool_bitmap ]; then ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv > gpurun_out/prof_$name.source.csv 2>/dev/null; fi
  rm -f gpurun_out/prof_$name.ncu-rep
}
cap round_gape_dev knap_dp 2 python scripts/run_round.py dev
cap round_gape_host knap_dp 2 python scripts/run_round.py host
cap dp_csp knap_dp 1 python scripts/run_dp.py csp_small
cap spmv_milp redcost_stream 1 python scripts/run_spmv.py
cap pool_bitmap pool_redcost 16 python scripts/run_rows.py pool
cap mckp mckp_dp 1 python scripts/run_rows.py mckp
cap dp_batch16 knap_dp 1 python scripts/run_rows.py batch16
cap bigdp knap_dp_big 1 python scripts/run_rows.py bigdp
cap intknap intknap 1 python scripts/run_rows.py intknap

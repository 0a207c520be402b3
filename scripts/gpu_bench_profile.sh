# Round-end measurement recipe (run under gpurun): bench line, reference arm, ncu launch list of
# the bench step, and one --set full capture per hot kernel.  Every ncu run follows a plain run
# of the same command that exited 0.
mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 300 --warmup 5 > gpurun_out/bench_ref.json 2>&1
python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
python scripts/run_dp.py gape > gpurun_out/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_gape python scripts/run_dp.py gape > gpurun_out/n1.log 2>&1; echo "ncu dp gape rc=$?"
python scripts/run_dp.py csp_small > gpurun_out/p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_csp python scripts/run_dp.py csp_small > gpurun_out/n2.log 2>&1; echo "ncu dp csp rc=$?"
python scripts/run_spmv.py > gpurun_out/p3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:redcost -s 1 -c 1 -f -o gpurun_out/prof_spmv_milp python scripts/run_spmv.py > gpurun_out/n3.log 2>&1; echo "ncu spmv rc=$?"
python scripts/run_rows.py pool > gpurun_out/p4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_redcost -s 16 -c 1 -f -o gpurun_out/prof_pool_bitmap python scripts/run_rows.py pool > gpurun_out/n4.log 2>&1; echo "ncu pool rc=$?"
python scripts/run_rows.py mckp > gpurun_out/p5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mckp_dp -s 1 -c 1 -f -o gpurun_out/prof_mckp python scripts/run_rows.py mckp > gpurun_out/n5.log 2>&1; echo "ncu mckp rc=$?"
python scripts/run_rows.py batch16 > gpurun_out/p6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_batch16 python scripts/run_rows.py batch16 > gpurun_out/n6.log 2>&1; echo "ncu batch16 rc=$?"

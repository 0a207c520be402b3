mkdir -p gpurun_out
python scripts/run_dp.py gape 0 0 0 > gpurun_out/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_gape_fused python scripts/run_dp.py gape 0 0 0 > gpurun_out/n1.log 2>&1; echo rc=$?
python scripts/run_spmv.py > gpurun_out/p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:redcost -s 1 -c 1 -f -o gpurun_out/prof_spmv_milp python scripts/run_spmv.py > gpurun_out/n2.log 2>&1; echo rc=$?
cat gpurun_out/p1.log gpurun_out/p2.log

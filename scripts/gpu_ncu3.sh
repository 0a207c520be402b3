mkdir -p gpurun_out
python scripts/run_spmv.py > gpurun_out/p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:redcost -s 1 -c 1 -f -o gpurun_out/prof_spmv_milp python scripts/run_spmv.py > gpurun_out/n2.log 2>&1; echo rc=$?
cat gpurun_out/p2.log

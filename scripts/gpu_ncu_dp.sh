mkdir -p gpurun_out
python scripts/run_dp.py gape 1 1 0 > gpurun_out/p1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_gape_s1 python scripts/run_dp.py gape 1 1 0 > gpurun_out/n1.log 2>&1; echo rc=$?
python scripts/run_dp.py gape 1 2 0 > gpurun_out/p2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_gape_s1i2 python scripts/run_dp.py gape 1 2 0 > gpurun_out/n2.log 2>&1; echo rc=$?
python scripts/run_dp.py csp_small 21 2 0 > gpurun_out/p3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_csp_s21i2 python scripts/run_dp.py csp_small 21 2 0 > gpurun_out/n3.log 2>&1; echo rc=$?
cat gpurun_out/p1.log gpurun_out/p2.log gpurun_out/p3.log

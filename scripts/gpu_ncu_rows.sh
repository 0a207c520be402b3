mkdir -p gpurun_out
python scripts/run_rows.py pool > gpurun_out/p4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pool_redcost -s 16 -c 1 -f -o gpurun_out/prof_pool_bitmap python scripts/run_rows.py pool > gpurun_out/n4.log 2>&1; echo "ncu pool rc=$?"
python scripts/run_rows.py mckp > gpurun_out/p5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:mckp_dp -s 1 -c 1 -f -o gpurun_out/prof_mckp python scripts/run_rows.py mckp > gpurun_out/n5.log 2>&1; echo "ncu mckp rc=$?"
python scripts/run_rows.py batch16 > gpurun_out/p6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 1 -c 1 -f -o gpurun_out/prof_dp_batch16 python scripts/run_rows.py batch16 > gpurun_out/n6.log 2>&1; echo "ncu batch16 rc=$?"
for f in gpurun_out/p4.log gpurun_out/p5.log gpurun_out/p6.log; do tail -n 2 $f; done

mkdir -p gpurun_out
set -x
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
python bench.py --impl reference --steps 300 --warmup 5 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu1.log 2>&1; echo "ncu1 rc=$?"
python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:knap_dp -s 8 -c 2 -f -o gpurun_out/prof_dp_gape python bench.py --steps 5 --warmup 3 --skip-extras > gpurun_out/ncu2.log 2>&1; echo "ncu2 rc=$?"

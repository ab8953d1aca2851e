set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
tail -c 3500 gpurun_out/bench_r1d.json
tail -3 gpurun_out/bench_r1d.err
python bench.py --workload newton --steps 10 --warmup 3 --cpu-budget 15 > gpurun_out/bench_newton_r1d.json 2> gpurun_out/bench_newton_r1d.err
tail -c 2500 gpurun_out/bench_newton_r1d.json
tail -3 gpurun_out/bench_newton_r1d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/ncu_d.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg_res -s 6 -c 1 -o gpurun_out/prof_pcg_res_r1d python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_d.log 2>&1
tail -2 gpurun_out/ncu_full_d.log | cut -c1-300
ls -la gpurun_out | tail -6

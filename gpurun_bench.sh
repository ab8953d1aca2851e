set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
tail -c 2500 gpurun_out/bench_r1c.json
tail -3 gpurun_out/bench_r1c.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err
tail -c 1200 gpurun_out/bench_ref_r1c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 6 -c 1 -o gpurun_out/prof_pcg_r1c python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_c.log 2>&1
tail -2 gpurun_out/ncu_full_c.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 4 -c 1 -o gpurun_out/prof_pcg_big_r1c python bench.py --workload partitioned --size 400x400x20 --steps 2 --warmup 3 > gpurun_out/ncu_full_big.log 2>&1
tail -2 gpurun_out/ncu_full_big.log | cut -c1-300
ls -la gpurun_out | tail -8

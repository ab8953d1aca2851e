set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -c 3000 gpurun_out/bench1.json
tail -5 gpurun_out/bench1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 6 -c 2 -o gpurun_out/prof_pcg_r1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out

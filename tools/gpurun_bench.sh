set -x
TAG=${TAG:-r1h}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t_$TAG.log; cat gpurun_out/t_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3500 gpurun_out/bench_$TAG.json
tail -3 gpurun_out/bench_$TAG.err
python bench.py --workload coupled --steps 20 --warmup 3 --cpu-budget 15 > gpurun_out/bench_coupled_$TAG.json 2> gpurun_out/bench_coupled_$TAG.err
tail -c 1500 gpurun_out/bench_coupled_$TAG.json
python bench.py --workload newton --steps 10 --warmup 3 --cpu-budget 15 > gpurun_out/bench_newton_$TAG.json 2> gpurun_out/bench_newton_$TAG.err
tail -c 1500 gpurun_out/bench_newton_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_coupled_$TAG.csv python bench.py --workload coupled --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_coupled_$TAG.log 2>&1
tail -2 gpurun_out/ncu_coupled_$TAG.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
tail -c 800 gpurun_out/bench_ref_$TAG.json
ls -la gpurun_out | tail -6

TAG=${TAG:-r1k}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_$TAG.log; cat gpurun_out/t_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json
python bench.py --workload coupled --steps 20 --warmup 3 --cpu-budget 10 > gpurun_out/bench_coupled_$TAG.json 2> gpurun_out/bench_coupled_$TAG.err
tail -c 300 gpurun_out/bench_coupled_$TAG.json

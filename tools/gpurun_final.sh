# End-of-campaign check on one B200: smoke, the GPU test suite, headline bench (default flags, as the driver runs it) and the coupled workload.
TAG=${TAG:-r2b}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
tail -c 600 gpurun_out/${TAG}_bench_1gpu.json; tail -3 gpurun_out/${TAG}_bench_1gpu.err
python bench.py --workload coupled --steps 20 --warmup 3 --cpu-budget 10 > gpurun_out/${TAG}_bench_coupled_1gpu.json 2> gpurun_out/${TAG}_bench_coupled_1gpu.err
tail -c 300 gpurun_out/${TAG}_bench_coupled_1gpu.json

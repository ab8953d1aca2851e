set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload enkf --members 256 --steps 2 --warmup 1 --concurrent 4 > gpurun_out/enkf256_8gpu_r1e.json 2> gpurun_out/enkf256_8gpu_r1e.err
tail -2 gpurun_out/enkf256_8gpu_r1e.err | cut -c1-300; cat gpurun_out/enkf256_8gpu_r1e.json | cut -c1-1500
timeout 300 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_8gpu_r1e.json 2> gpurun_out/bench_8gpu_r1e.err
tail -2 gpurun_out/bench_8gpu_r1e.err | cut -c1-300; cat gpurun_out/bench_8gpu_r1e.json | cut -c1-1500

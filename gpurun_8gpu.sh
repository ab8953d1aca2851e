set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --workload partitioned --size 1000x1000x30 --steps 3 --warmup 3 > gpurun_out/part_full_8gpu.json 2> gpurun_out/part_full_8gpu.err
tail -2 gpurun_out/part_full_8gpu.err | cut -c1-300; cat gpurun_out/part_full_8gpu.json | cut -c1-1500
timeout 400 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --workload partitioned --size 1000x1000x30 --steps 3 --warmup 3 > gpurun_out/part_full_4gpu.json 2> gpurun_out/part_full_4gpu.err
tail -2 gpurun_out/part_full_4gpu.err | cut -c1-300; cat gpurun_out/part_full_4gpu.json | cut -c1-1500
timeout 400 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload enkf --members 256 --steps 1 --warmup 1 --concurrent 4 > gpurun_out/enkf256_8gpu.json 2> gpurun_out/enkf256_8gpu.err
tail -2 gpurun_out/enkf256_8gpu.err | cut -c1-300; cat gpurun_out/enkf256_8gpu.json | cut -c1-1500
timeout 300 $TR --nproc-per-node 8 --master-port 29524 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -2 gpurun_out/bench_8gpu.err | cut -c1-300; cat gpurun_out/bench_8gpu.json | cut -c1-1500

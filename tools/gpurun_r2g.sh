# Round-2 closing run on one B200: smoke, the whole GPU suite, the driver's default bench, the coupled and pre-processor
# workloads (both arms for the latter), the launch list and ncu full captures of the pre-processor kernels.
TAG=${TAG:-r2g}
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
tail -c 400 gpurun_out/${TAG}_bench_1gpu.json; tail -2 gpurun_out/${TAG}_bench_1gpu.err
python bench.py --workload coupled --steps 20 --warmup 3 --cpu-budget 10 > gpurun_out/${TAG}_bench_coupled_1gpu.json 2> gpurun_out/${TAG}_bench_coupled_1gpu.err
tail -c 300 gpurun_out/${TAG}_bench_coupled_1gpu.json
python bench.py --workload prepro --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_prepro.json 2> gpurun_out/${TAG}_bench_prepro.err
python bench.py --workload prepro --impl reference --steps 1 > gpurun_out/${TAG}_bench_prepro_reference_arm.json 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_prepro.json').read().strip().splitlines()[-1])
print('prepro: device %.1f ms, e2e %.2f s, value %.4g e2e %.4g cpu %s stages %s' % (d['ms_per_step'], d['e2e']['seconds'], d['value'], d['e2e']['value'], d.get('cpu_baseline'), d['stage_ms']))"
tail -c 400 gpurun_out/${TAG}_bench_prepro_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches_prepro.csv python tools/gpurun_prepro_probe.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pp_sweep|k_pp_depit|k_pp_qsplit|k_pp_qsmall|k_pp_local|k_pp_smean" -c 8 -o gpurun_out/${TAG}_prof_prepro python tools/gpurun_prepro_probe.py > /dev/null 2> gpurun_out/${TAG}_ncu_prepro.err
ncu -i gpurun_out/${TAG}_prof_prepro.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_prepro.raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_prof_prepro.ncu-rep
ls -la gpurun_out | grep ${TAG}

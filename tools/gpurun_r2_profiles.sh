# Round-2 evidence run (one B200): launch lists + ncu --set full captures of the dominant kernels; outputs land in gpurun_out/.
set -x
TAG=${TAG:-r2a}
B="python bench.py --no-cpu --quick --sharded off"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2.csv $B --steps 2 --warmup 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_coupled_steps2.csv $B --workload coupled --steps 2 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pcg_res2|k_assemble_a|k_curves|k_rhs_lhs|k_norms" -s 12 -c 6 -o gpurun_out/${TAG}_prof_picard $B --steps 2 --warmup 3 > /dev/null 2> gpurun_out/${TAG}_ncu1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bicgstab_res|k_permute_cols|k_bres_sym|k_assemble_newton|k_route" -s 10 -c 6 -o gpurun_out/${TAG}_prof_coupled $B --workload coupled --steps 2 --warmup 3 > /dev/null 2> gpurun_out/${TAG}_ncu2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pcg_tma|k_spmv" -s 2 -c 3 -o gpurun_out/${TAG}_prof_tma python bench.py --workload partitioned --size 400x400x20 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/${TAG}_ncu3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_enkf_crosscov|k_enkf_update" -s 2 -c 2 -o gpurun_out/${TAG}_prof_enkf python tools/gpurun_enkf_probe.py > /dev/null 2> gpurun_out/${TAG}_ncu4.err
# the .ncu-rep files together exceed what gpurun brings back (64 MiB): keep their raw pages as CSV
for r in picard coupled tma enkf; do ncu -i gpurun_out/${TAG}_prof_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_$r.raw.csv 2>/dev/null; rm -f gpurun_out/${TAG}_prof_$r.ncu-rep; done
python tools/gpurun_enkf_probe.py 2>&1 | tail -1 > gpurun_out/${TAG}_enkf_probe.log
NE=32 python tools/gpurun_enkf_probe.py 2>&1 | tail -1 >> gpurun_out/${TAG}_enkf_probe.log
cat gpurun_out/${TAG}_enkf_probe.log
ls -la gpurun_out | grep ${TAG}

timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pcg|k_assemble|k_spmv|k_curves" -s 8 -c 8 -o gpurun_out/prof_r1_final python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1
tail -3 gpurun_out/ncu_full2.log | cut -c1-300

python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench3.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench3.json'))
print('value %.4g e2e %.4g ms/step %.3f dev_ms/step %.3f pcg_frac %.3f pcg_share %.3f spmv_frac %.3f iters %d solves %d launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['device_ms_per_step'], d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline']['spmv_only']['frac'], d['config']['pcg_iters'], d['config']['pcg_solves'], d['gpu_launches']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1

python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/t_res.log; cat gpurun_out/t_res.log
timeout 900 python gpurun_pcgvar.py > gpurun_out/pcgvar7.log 2>&1; cat gpurun_out/pcgvar7.log
python bench.py --steps 20 --warmup 3 --no-cpu 2>gpurun_out/bench4.err | tail -1 > gpurun_out/bench4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench4.json'))
print('value %.4g e2e %.4g ms/step %.3f dev_ms/step %.3f pcg_frac %.3f pcg_share %.3f spmv_frac %.3f iters %d solves %d launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['device_ms_per_step'], d['roofline']['frac'], d['roofline']['share_of_step'], d['roofline']['spmv_only']['frac'], d['config']['pcg_iters'], d['config']['pcg_solves'], d['gpu_launches']))
PY

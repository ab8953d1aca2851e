timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "newton or ponding" 2>&1 | tail -5
python bench.py --workload newton --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('newton: ms/step %.3f value %.4g us/it %.2f share %.3f its %d nl %d' % (d['ms_per_step'], d['value'], d['roofline']['us_per_pcg_iter'], d['roofline']['share_of_step'], d['config']['pcg_iters'], d['config']['nonlinear_its']))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_newton_r1f.csv python bench.py --workload newton --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<'PY'
import sys
sys.path.insert(0,'tools')
import summarise_profiles as sp
sp.launch_shares('gpurun_out/launches_newton_r1f.csv','gpurun_out/newton_shares_r1f.md','x')
print(''.join(open('gpurun_out/newton_shares_r1f.md').readlines()[5:14]))
PY

timeout 800 python -m pytest tests/test_gpu_parity.py -x -q -k "newton or ponding" 2>&1 | tail -6
for wl in newton coupled; do python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl: ms/step %.3f value %.4g us/it %.2f share %.3f its %d nl %d' % (d['ms_per_step'], d['value'], d['roofline']['us_per_pcg_iter'], d['roofline']['share_of_step'], d['config']['pcg_iters'], d['config']['nonlinear_its']))"; done

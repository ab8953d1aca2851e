python - <<'PY'
import torch
p=torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size/2**20, 'MB')
PY
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "newton" 2>&1 | tail -30
run() { python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$2 $1 ms/step %.3f value %.4g e2e %.4g us/it %.2f share %.3f its %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['us_per_pcg_iter'], d['roofline']['share_of_step'], d['config']['pcg_iters']))"; }
for w in coupled newton; do
CATHY_L2_PERSIST=0 run $w persist=0
run $w persist=max,reset
CATHY_L2_RESET=0 run $w persist=max,noreset
CATHY_L2_PERSIST=64 run $w persist=64MB,reset
done

timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "newton or derived or curves" 2>&1 | tail -30
for w in newton coupled; do
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w ms/step %.3f value %.4g e2e %.4g launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches']))"
CATHY_PLAN_STORED=1 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$w STORED ms/step %.3f value %.4g e2e %.4g launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches']))"
done

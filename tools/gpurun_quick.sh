timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "hillslope or storm or ponding or routing or coupled or aux" 2>&1 | tail -12
python bench.py --workload coupled --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('coupled ms/step %.3f value %.4g e2e %.4g launches %d share %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['share_of_step']))"

timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "curves or zones or newton" 2>&1 | tail -60
python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms/step %.3f value %.4g e2e %.4g share_pcg %.3f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['share_of_step'], d['gpu_launches']))"

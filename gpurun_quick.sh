for pf in 0 1; do CATHY_PCG_PREFETCH=$pf python bench.py --workload coupled --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('prefetch $pf: ms/step %.3f value %.4g us/it %.2f frac %.3f share %.3f its %d' % (d['ms_per_step'], d['value'], d['roofline']['us_per_pcg_iter'], d['roofline']['frac'], d['roofline']['share_of_step'], d['config']['pcg_iters']))"; done

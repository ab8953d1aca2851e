for g in 0 148; do if [ $g = 0 ]; then unset CATHY_PCG_GRID; else export CATHY_PCG_GRID=$g; fi; python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('grid env=$g: ms/step %.3f dev %.3f value %.4g e2e %.4g share_pcg %.3f us/it %.2f' % (d['ms_per_step'], d['device_ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['share_of_step'], d['roofline']['us_per_pcg_iter']))"; done

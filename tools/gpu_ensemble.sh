python bench.py --steps ${STEPS:-200} --warmup 21 --workload ensemble_graphene 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload']); print('  value %.4g  ms per ensemble step %.4f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches']))
    else: print(l.rstrip())
"

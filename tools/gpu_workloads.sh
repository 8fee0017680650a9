for w in lj_fluid ab_gas graphene_cu; do
python bench.py --steps ${STEPS:-200} --warmup 21 --no-cpu-baseline --no-e2e --workload $w 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload']); print('  value %.4g ms/step %.4f launches/step %.1f'%(d['value'],d['ms_per_step'],d['gpu_launches']/d['steps'])); print('  ',{k:round(v,4) for k,v in d['kernels_ms_per_step'].items()})
    else: print(l.rstrip())
"
done

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 21 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g ms/step %.4f'%(d['value'],d['ms_per_step'])); print(d['kernels_ms_per_step']); print(d['roofline']['frac'], d['clocks'])
    else: print(l.rstrip())
"

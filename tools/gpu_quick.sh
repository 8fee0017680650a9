# quick GPU check: parity tests + one bench line (run under gpurun)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps ${STEPS:-100} --warmup 21 --no-cpu-baseline --no-e2e ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g ms/step %.4f launches %d'%(d['value'],d['ms_per_step'],d['gpu_launches'])); print({k:round(v,4) for k,v in d['kernels_ms_per_step'].items()}); print('frac',d['roofline']['frac'], d['clocks'])
    else: print(l.rstrip())
"

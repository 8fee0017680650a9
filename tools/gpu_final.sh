# final single-GPU evidence of the round
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; tail -c 300 gpurun_out/final_n1.err
python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
for w in lj_fluid ab_gas graphene_cu ensemble_graphene; do python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 200 > gpurun_out/final_$w.json 2>/dev/null; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/final_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY

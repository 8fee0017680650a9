mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2q_tests.log
tail -3 gpurun_out/r2q_tests.log
timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > "gpurun_out/r2q_ensemble.json" 2>> gpurun_out/r2q.err
timeout 300 python bench.py --workload graphene_cu --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > "gpurun_out/r2q_graphene_cu.json" 2>> gpurun_out/r2q.err
tail -c 300 gpurun_out/r2q.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2q_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"])
PY

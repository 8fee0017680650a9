mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2o_tests.log
tail -3 gpurun_out/r2o_tests.log
python bench.py --steps 100 --warmup 21 --no-cpu-baseline --no-variants > gpurun_out/r2o_bench_n1.json 2>> gpurun_out/r2o_bench.err
for wl in ab_gas graphene_cu lj_deposition lj_fluid; do
  timeout 300 python bench.py --workload $wl --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2o_${wl}.json 2>> gpurun_out/r2o_bench.err
done
timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > gpurun_out/r2o_ensemble_1gpu.json 2>> gpurun_out/r2o_bench.err
tail -c 600 gpurun_out/r2o_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "launches", d.get("gpu_launches"), d.get("kernels_ms_per_step"))
PY

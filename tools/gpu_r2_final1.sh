# final 1-GPU call of round 2: whole suite, smoke, driver-style bench lines, workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -16 > gpurun_out/r2v_tests.log
tail -3 gpurun_out/r2v_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -1 gpurun_out/r2v_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2v_bench_ref.json 2>> gpurun_out/r2v.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_n1_k20.json 2>> gpurun_out/r2v.err
python bench.py > gpurun_out/r2v_bench_n1.json 2>> gpurun_out/r2v.err
for wl in ab_gas graphene_cu lj_fluid; do
  timeout 300 python bench.py --workload $wl --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2v_${wl}.json 2>> gpurun_out/r2v.err
done
timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > gpurun_out/r2v_ensemble_1gpu.json 2>> gpurun_out/r2v.err
tail -c 400 gpurun_out/r2v.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2v_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            r = d.get("roofline") or {}
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "frac", r.get("frac"), d.get("kernels_ms_per_step"))
PY

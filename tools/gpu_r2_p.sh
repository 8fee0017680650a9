mkdir -p gpurun_out
for v in "" "PFMDS_SMALL_FORK=0"; do
  env $v timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > "gpurun_out/r2p_ensemble_${v}.json" 2>> gpurun_out/r2p.err
  env $v timeout 300 python bench.py --workload graphene_cu --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > "gpurun_out/r2p_graphene_cu_${v}.json" 2>> gpurun_out/r2p.err
  env $v timeout 300 python bench.py --workload ab_gas --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > "gpurun_out/r2p_ab_gas_${v}.json" 2>> gpurun_out/r2p.err
done
timeout 300 python bench.py --workload lj_fluid --steps 200 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2p_lj_fluid.json 2>> gpurun_out/r2p.err
timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > "gpurun_out/r2p_ensemble_again.json" 2>> gpurun_out/r2p.err
tail -c 300 gpurun_out/r2p.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], d["clocks"])
PY

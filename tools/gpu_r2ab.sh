#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ab_tests.log 2>&1
tail -4 gpurun_out/r2ab_tests.log
for wl in ab_gas graphene_cu; do
  for r in 1 0; do
    PFMDS_GRAPH_REBUILDS=$r timeout 200 python bench.py --workload $wl --steps 2000 --warmup 21 > gpurun_out/r2ab_${wl}_r$r.json 2> gpurun_out/r2ab_${wl}_r$r.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2ab_${wl}_r$r.json").read().strip().splitlines()[-1])
    print("$wl graph_rebuilds=$r", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"])
except Exception as e:
    print("$wl r=$r failed", e); print(open("gpurun_out/r2ab_${wl}_r$r.err").read()[-1500:])
P
  done
done
timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > gpurun_out/r2ab_ensemble_1gpu.json 2> gpurun_out/r2ab_ensemble_1gpu.err; tail -c 700 gpurun_out/r2ab_ensemble_1gpu.json

#!/bin/bash
# round 2, late: full GPU suite, then the small workloads (hoisted loads in k_sum_kick_ke, populated-block grid, four steps per graph launch) and the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_tests.log 2>&1
tail -6 gpurun_out/r2aa_tests.log
for wl in ab_gas graphene_cu; do
  for g in 4 1 8; do
    PFMDS_GRAPH_STEPS=$g timeout 200 python bench.py --workload $wl --steps 2000 --warmup 21 > gpurun_out/r2aa_${wl}_g$g.json 2> gpurun_out/r2aa_${wl}_g$g.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2aa_${wl}_g$g.json").read().strip().splitlines()[-1])
    print("$wl graph_steps=$g", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"])
except Exception as e:
    print("$wl g=$g failed", e); print(open("gpurun_out/r2aa_${wl}_g$g.err").read()[-1500:])
P
  done
done
timeout 400 python bench.py > gpurun_out/r2aa_bench_n1.json 2> gpurun_out/r2aa_bench_n1.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2aa_bench_n1.json").read().strip().splitlines()[-1])
print("cu_fcc", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
P

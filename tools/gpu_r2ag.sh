#!/bin/bash
# survivor-list capacity of the cell-tiled list build (shared memory per warp against flushes): PFMDS_NL_LCAP sweep on the bench workload
mkdir -p gpurun_out
for l in 42 58 74 90 106 122 154; do
  PFMDS_NL_LCAP=$l timeout 200 python bench.py --steps 100 --warmup 21 --no-variants --no-cpu-baseline --no-e2e > gpurun_out/r2ag_lcap$l.json 2> gpurun_out/r2ag_lcap$l.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2ag_lcap$l.json").read().strip().splitlines()[-1])
    print("lcap $l", "ms/step %.4f"%d["ms_per_step"], "nl_build %.4f"%d["kernels_ms_per_step"]["nl_build"])
except Exception as e:
    print("lcap $l failed", e)
P
done
timeout 300 python -m pytest tests/test_zz_variants_gpu.py -q -x 2>&1 | tail -3

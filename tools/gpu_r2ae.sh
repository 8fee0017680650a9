#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_tests.log 2>&1
tail -4 gpurun_out/r2ae_tests.log
timeout 400 python bench.py > gpurun_out/r2ae_bench_n1.json 2> gpurun_out/r2ae_bench_n1.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2ae_bench_n1.json").read().strip().splitlines()[-1])
print("cu_fcc", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e", "%.4g"%d["e2e"]["value"], "frac", "%.3f"%d["roofline"]["frac"], d["kernels_ms_per_step"])
P
timeout 200 python bench.py --workload lj_fluid --steps 400 --warmup 21 --no-variants > gpurun_out/r2ae_lj_fluid.json 2> gpurun_out/r2ae_lj_fluid.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2ae_lj_fluid.json").read().strip().splitlines()[-1])
print("lj_fluid", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["kernels_ms_per_step"])
P

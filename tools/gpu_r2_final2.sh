#!/bin/bash
# last call of round 2: what the driver runs at round end (GPU suite, smoke, both bench arms), on the final tree
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2fin_tests.log 2>&1
tail -4 gpurun_out/r2fin_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2fin_smoke.log 2>&1; tail -2 gpurun_out/r2fin_smoke.log
timeout 300 python bench.py --impl reference > gpurun_out/r2fin_bench_ref.json 2> gpurun_out/r2fin_bench_ref.err
timeout 500 python bench.py > gpurun_out/r2fin_bench_n1.json 2> gpurun_out/r2fin_bench_n1.err
python - <<P
import json
r=json.loads(open("gpurun_out/r2fin_bench_ref.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/r2fin_bench_n1.json").read().strip().splitlines()[-1])
print("reference arm", "%.4g"%r["value"], r["cpu_baseline"]["sample"][:80])
print("cu_fcc", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "e2e %.4g"%d["e2e"]["value"], "frac %.3f"%d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
print(d["kernels_ms_per_step"])
P
for wl in ab_gas graphene_cu; do
  timeout 200 python bench.py --workload $wl --steps 2000 --warmup 21 > gpurun_out/r2fin_$wl.json 2> gpurun_out/r2fin_$wl.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2fin_$wl.json').read().strip().splitlines()[-1]);print('$wl', '%.4g'%d['value'], 'ms/step %.4f'%d['ms_per_step'])"
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2ah_tests.log 2>&1
tail -4 gpurun_out/r2ah_tests.log
for t in 1 0; do
  PFMDS_TB_SUM=$t timeout 200 python bench.py --workload graphene_cu --steps 2000 --warmup 21 > gpurun_out/r2ah_graphene_cu_tb$t.json 2> gpurun_out/r2ah_graphene_cu_tb$t.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2ah_graphene_cu_tb$t.json").read().strip().splitlines()[-1])
    print("graphene_cu tb_sum=$t", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["kernels_ms_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2ah_graphene_cu_tb$t.err").read()[-1500:])
P
done
timeout 200 python bench.py --workload lj_fluid --steps 400 --warmup 21 --no-variants > gpurun_out/r2ah_lj_fluid.json 2> gpurun_out/r2ah_lj_fluid.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2ah_lj_fluid.json").read().strip().splitlines()[-1])
print("lj_fluid", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["kernels_ms_per_step"], "frac", d["roofline"]["frac"])
P
timeout 300 python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > gpurun_out/r2ah_ensemble_1gpu.json 2> gpurun_out/r2ah_ensemble_1gpu.err; python -c "
import json;d=json.loads(open('gpurun_out/r2ah_ensemble_1gpu.json').read().strip().splitlines()[-1]);print('ensemble 1 GPU', '%.4g'%d['value'])"

#!/bin/bash
# round 2, persistent step kernel: bit-equality tests, then the small workloads with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_zz_persist_gpu.py -x -q > gpurun_out/r2y_persist_tests.log 2>&1
tail -15 gpurun_out/r2y_persist_tests.log
(timeout 100 python tools/stamps_probe.py persist ab_gas; timeout 100 python tools/stamps_probe.py persist graphene_cu; timeout 100 python tools/stamps_probe.py run ab_gas) > gpurun_out/r2y_stamps.txt 2>&1
grep -v Warning gpurun_out/r2y_stamps.txt | tail -40
for wl in ab_gas graphene_cu; do
  for p in 1 0; do
    PFMDS_PERSIST=$p timeout 200 python bench.py --workload $wl --steps 2000 --warmup 21 > gpurun_out/r2y_${wl}_persist$p.json 2> gpurun_out/r2y_${wl}_persist$p.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2y_${wl}_persist$p.json").read().strip().splitlines()[-1])
    print("$wl persist=$p", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "launches", d.get("gpu_launches"))
except Exception as e:
    print("$wl persist=$p failed", e); print(open("gpurun_out/r2y_${wl}_persist$p.err").read()[-1500:])
P
  done
done

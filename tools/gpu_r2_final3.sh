#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 2 --steps 200 --warmup 21 > gpurun_out/r2fin_scale_2gpu.json 2> gpurun_out/r2fin_scale_2gpu.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2fin_scale_2gpu.json").read().strip().splitlines()[-1])
    print("2 GPUs", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["kernels_ms_per_step"], "check", d.get("check",{}).get("ok"), "e2e", d.get("e2e",{}).get("value"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2fin_scale_2gpu.err").read()[-2500:])
P

#!/bin/bash
# 4 GPUs: the slab list build takes the cell-tiled kernel by the rank's share of the cells (was: k_build_mask from 4 ranks up)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 4 --steps 200 --warmup 21 > gpurun_out/r2ac_scale_4gpu.json 2> gpurun_out/r2ac_scale_4gpu.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r2ac_scale_4gpu.json").read().strip().splitlines()[-1])
    print("4 GPUs", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["kernels_ms_per_step"], "check", d.get("check",{}).get("ok"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2ac_scale_4gpu.err").read()[-2500:])
P
timeout 300 python -m pytest tests/test_slab_gpu.py -q 2>&1 | tail -3

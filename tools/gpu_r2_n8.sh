# 8 GPUs (gpurun --gpus 8): default slab line, BASELINE.json configs[3] at 1.0165e8 atoms, configs[4] ensemble of 64 runs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 600 $TR 29655 bench.py --gpus 8 --steps 200 --warmup 21 > gpurun_out/r2u_scale_8gpu.json 2> gpurun_out/r2u_scale_8gpu.err
tail -c 600 gpurun_out/r2u_scale_8gpu.err
timeout 900 $TR 29656 bench.py --gpus 8 --workload cu_fcc_1e8 --steps 100 --warmup 21 > gpurun_out/r2u_cu_fcc_1e8_8gpu.json 2> gpurun_out/r2u_cu_fcc_1e8_8gpu.err
tail -c 600 gpurun_out/r2u_cu_fcc_1e8_8gpu.err
timeout 600 $TR 29657 bench.py --gpus 8 --workload ensemble_graphene --steps 200 --warmup 21 > gpurun_out/r2u_ensemble_8gpu.json 2> gpurun_out/r2u_ensemble_8gpu.err
tail -c 600 gpurun_out/r2u_ensemble_8gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2u_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", d.get("e2e") and d["e2e"].get("value"), d.get("check"))
            print("   ", d.get("kernels_ms_per_step"))
PY

mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 600 $TR 29657 bench.py --gpus 8 --workload ensemble_graphene --steps 400 --warmup 21 > gpurun_out/r2w_ensemble_8gpu.json 2> gpurun_out/r2w_ensemble_8gpu.err
timeout 600 $TR 29655 bench.py --gpus 8 --steps 200 --warmup 21 > gpurun_out/r2w_scale_8gpu.json 2> gpurun_out/r2w_scale_8gpu.err
timeout 600 $TR 29658 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2w_scale_8gpu_k20.json 2>> gpurun_out/r2w_scale_8gpu.err
tail -c 300 gpurun_out/r2w_scale_8gpu.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2w_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", d.get("e2e") and d["e2e"].get("value"), (d.get("check") or {}).get("ok"))
            print("   ", d.get("kernels_ms_per_step"))
PY

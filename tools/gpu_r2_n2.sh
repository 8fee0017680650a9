# 2 GPUs: the slab tests (driver's box cannot run them), then the self-checking bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2j_slab_tests.log
tail -3 gpurun_out/r2j_slab_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 100 --warmup 21 > gpurun_out/r2j_scale_2gpu.json 2> gpurun_out/r2j_scale_2gpu.err
tail -c 800 gpurun_out/r2j_scale_2gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r2j_scale_2gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["check"])
        print(d["kernels_ms_per_step"])
PY
PFMDS_SLAB_LEAN=0 PFMDS_SLAB_KE_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 100 --warmup 21 --no-e2e > gpurun_out/r2j_scale_2gpu_old_halo.json 2>> gpurun_out/r2j_scale_2gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r2j_scale_2gpu_old_halo.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("old halo + NCCL KE:", d["value"], d["ms_per_step"], d["check"]["ok"])
        print(d["kernels_ms_per_step"])
PY

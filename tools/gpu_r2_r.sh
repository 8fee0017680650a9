# 2 GPUs: halo variants A/B (lean = one kernel per exchange + consumers wait; mailboxes for the thermostat KE)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2r_slab_tests.log
cat gpurun_out/r2r_slab_tests.log
i=0
for v in "" "PFMDS_SLAB_LEAN=0" "PFMDS_SLAB_KE_NCCL=1" "PFMDS_SLAB_LEAN=0 PFMDS_SLAB_KE_NCCL=1"; do
  i=$((i+1))
  env $v timeout 600 $TR $((29660+i)) bench.py --gpus 2 --steps 200 --warmup 21 --no-e2e > "gpurun_out/r2r_scale_2gpu_$i.json" 2>> gpurun_out/r2r.err
  python - "$v" "gpurun_out/r2r_scale_2gpu_$i.json" <<'PY'
import json, sys
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l)
        print(repr(sys.argv[1]), d["value"], d["ms_per_step"], d["check"]["ok"], {k: round(v, 4) for k, v in d["kernels_ms_per_step"].items()})
PY
done
tail -c 300 gpurun_out/r2r.err

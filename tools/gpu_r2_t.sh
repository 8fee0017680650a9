mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 600 python -m pytest tests/test_slab_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2t_slab_tests.log
cat gpurun_out/r2t_slab_tests.log
timeout 600 $TR 29671 bench.py --gpus 2 --steps 200 --warmup 21 > gpurun_out/r2t_scale_2gpu.json 2> gpurun_out/r2t.err
python - <<'PY'
import json
for l in open("gpurun_out/r2t_scale_2gpu.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["check"]["ok"], d["e2e"]["value"], d["e2e"]["phases_ms"], {k: round(v, 4) for k, v in d["kernels_ms_per_step"].items()})
PY
tail -c 300 gpurun_out/r2t.err
for v in "PFMDS_NL_LCAP=154" "PFMDS_NL_LCAP=186"; do
  env $v python bench.py --steps 100 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > "gpurun_out/r2t_bench_n1_${v}.json" 2>> gpurun_out/r2t.err
  python - "gpurun_out/r2t_bench_n1_${v}.json" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l); print(sys.argv[1], d["ms_per_step"], d["kernels_ms_per_step"]["nl_build"])
PY
done

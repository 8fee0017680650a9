N=${1:-4}
timeout -k 5 ${TMO:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps ${STEPS:-200} --warmup 21 ${BENCH_ARGS} > gpurun_out/scale_$N.log 2>&1
grep '^{' gpurun_out/scale_$N.log > gpurun_out/scale_$N.json
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
print('value %.4g ms/step %.4f e2e %s'%(d['value'],d['ms_per_step'],d['e2e'] and '%.4g'%d['e2e']['value'])); print({k:round(v,4) for k,v in d['kernels_ms_per_step'].items()})
for r in d["per_rank"]: print(r["rank"], round(r["ms"],1), r["kernels_ms_per_step"].get("rjl_force"), r["kernels_ms_per_step"].get("other"), r["clocks"]["sm_mhz"], r["clocks"]["reasons"])
PY

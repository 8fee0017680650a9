# slab parity test + N-GPU bench with hard time limits (run under gpurun --gpus N)
N=${1:-2}
timeout -k 5 200 python -m pytest tests/test_slab_gpu.py -x -q 2>&1 | tail -4
timeout -k 5 ${TMO:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps ${STEPS:-100} --warmup 21 ${BENCH_ARGS} > gpurun_out/scale_$N.log 2>&1
grep '^{' gpurun_out/scale_$N.log > gpurun_out/scale_$N.json
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_$N.json').read().strip().splitlines()[-1])
    print(d['config']['workload'],'|',d['config']['parallelism']); print('  value %.4g ms/step %.4f e2e %s'%(d['value'],d['ms_per_step'],d['e2e'] and '%.4g'%d['e2e']['value'])); print('  ',{k:round(v,4) for k,v in d['kernels_ms_per_step'].items()})
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/scale_$N.log').read()[-1500:])
PY

# N-GPU bench under torchrun (run under gpurun --gpus N):  N=$1
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps ${STEPS:-100} --warmup 21 ${BENCH_ARGS} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'], '|', d['config']['parallelism']); print('  value %.4g ms/step %.4f  e2e %s'%(d['value'],d['ms_per_step'], d['e2e'] and '%.4g'%d['e2e']['value'])); print('  ',{k:round(v,4) for k,v in d['kernels_ms_per_step'].items()}); print('  ', d['clocks'])
    elif 'Error' in l or 'error' in l or 'Traceback' in l or 'assert' in l: print(l.rstrip())
"

N=${1:-2}
PFMDS_SLAB_DEBUG=1 timeout -k 5 ${TMO:-90} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --no-e2e > gpurun_out/dbg.log 2>&1
grep -n -i "slab \|pfmds error\|PfmdsError\|\"value\|assert" gpurun_out/dbg.log | tail -14 | cut -c1-300

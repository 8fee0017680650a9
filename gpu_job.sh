set -x
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 100 --warmup 21 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; tail -c 3000 gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err

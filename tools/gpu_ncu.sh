# ncu capture of kernels matching $1 into gpurun_out/$2.ncu-rep (run under gpurun)
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${SKIP:-42} -c ${COUNT:-2} -o gpurun_out/$2 python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log

# Round 2, second GPU call: third-generation rjl kernels (node-table exponentials): parity, A/B timing, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_zz_variants_gpu.py tests/test_zz_random_systems.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2b_tests.log
tail -3 gpurun_out/r2b_tests.log
python bench.py --steps 100 --warmup 21 --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 600 gpurun_out/r2b_bench_n1.err
ncu --set full --clock-control none --import-source on -k regex:'k_rjl_(force|density)' -s 44 -c 2 -o gpurun_out/r2b_rjl python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r2b_ncu_rjl.log 2>&1
PFMDS_RJL_MINB=8 ncu --set full --clock-control none --import-source on -k regex:'k_rjl_force' -s 22 -c 1 -o gpurun_out/r2b_rjl_mb8 python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r2b_ncu_rjl8.log 2>&1
ls -la gpurun_out | grep r2b

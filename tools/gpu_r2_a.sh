# Round 2, first GPU call (one B200):  gpurun --timeout 1500 -- 'bash tools/gpu_r2_a.sh'
# whole GPU suite (now also the large-path parity tests), default bench + reference arm, launch list, ncu --set full of the default kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 > gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1_k20.json 2> gpurun_out/r2a_bench_n1_k20.err
python bench.py --no-variants > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 1500 gpurun_out/r2a_bench_n1_k20.json
# launch list: 2 steady steps + one rebuild step
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 25 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r2a_ncu_list.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:'k_rjl_(force|density)' -s 44 -c 2 -o gpurun_out/r2a_rjl python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r2a_ncu_rjl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_build' -s 1 -c 1 -o gpurun_out/r2a_build python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r2a_ncu_build.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rjl_force_e|k_kick|k_nhc' -s 60 -c 4 -o gpurun_out/r2a_misc python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-variants > gpurun_out/r2a_ncu_misc.log 2>&1
ls -la gpurun_out | tail -20

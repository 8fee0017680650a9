# final round-1 evidence: launch list of the bench command + full captures of the top kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 160 --csv --log-file gpurun_out/r1_launches_final.csv python bench.py --steps 25 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rjl|k_build|k_kick' -s 3 -c 6 -o gpurun_out/r1e_top python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
python bench.py --impl reference > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
tail -c 600 gpurun_out/bench_r1_n1.json; tail -c 400 gpurun_out/bench_r1_ref.json

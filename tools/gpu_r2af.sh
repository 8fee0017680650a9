#!/bin/bash
# round 2, last profiles: the kernels as they ship (density pass at 9 blocks per SM, integrator loads issued with the mask, k_kick_ke with prefetch)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 100 --warmup 21 --no-variants --no-cpu-baseline > gpurun_out/r2af_bench_n1.json 2> gpurun_out/r2af_bench_n1.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2af_bench_n1.json").read().strip().splitlines()[-1])
print("cu_fcc", "%.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["kernels_ms_per_step"])
P
B="python bench.py --steps 25 --warmup 21 --no-cpu-baseline --no-e2e --no-variants"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 260 --csv --log-file gpurun_out/r2af_launches.csv $B > gpurun_out/r2af_ncu_list.log 2>&1
B3="python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e --no-variants"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rjl_(force|density)' -s 44 -c 2 -o gpurun_out/r2af_rjl $B3 > gpurun_out/r2af_ncu_rjl.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_build' -s 1 -c 1 -o gpurun_out/r2af_build $B3 > gpurun_out/r2af_ncu_build.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_kick|k_nhc' -s 60 -c 4 -o gpurun_out/r2af_misc $B3 > gpurun_out/r2af_ncu_misc.log 2>&1
for f in rjl build misc; do ncu -i gpurun_out/r2af_$f.ncu-rep --page raw --csv > gpurun_out/r2af_$f.raw.csv 2>/dev/null; python profiles/ncu_raw.py gpurun_out/r2af_$f.raw.csv > gpurun_out/r2af_${f}_raw_summary.txt 2>&1; done
ls -la gpurun_out | grep r2af
head -12 gpurun_out/r2af_misc_raw_summary.txt | cut -c1-200

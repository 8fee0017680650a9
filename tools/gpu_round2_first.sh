# First GPU call of the next round (run under gpurun, one B200): everything that was written after the round-1 GPU budget ran out.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
# 1. the whole GPU suite (new: golden fixtures of deposition / rebosc, fitting, restart cases, the lj1g variant)
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r2_tests.log
# 2. bench lines: default workload, LJ fluid with and without the pipelined lj1g kernel, the two new workloads
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 60 > gpurun_out/r2_bench_ref.json 2>/dev/null
python bench.py --workload lj_fluid --no-cpu-baseline --no-e2e > gpurun_out/r2_lj_fluid_default.json 2>/dev/null
PFMDS_LJ1G_PIPE=1 python bench.py --workload lj_fluid --no-cpu-baseline --no-e2e > gpurun_out/r2_lj_fluid_pipe.json 2>/dev/null
python bench.py --workload graphene_rebosc --steps 100 --no-cpu-baseline --no-e2e > gpurun_out/r2_graphene_rebosc.json 2>/dev/null
python bench.py --workload lj_deposition --steps 400 --no-cpu-baseline --no-e2e > gpurun_out/r2_lj_deposition.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); r = d.get("roofline") or {}
            print(f, "%.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], r.get("kernel"), r.get("frac"), {k: round(v, 4) for k, v in list((d.get("kernels_ms_per_step") or {}).items())[:4]})
PY
# 3. memcheck / racecheck of the new kernels on small systems (the reference has no sanitizer story)
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rebosc_gpu.py tests/test_deposition_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_memcheck.log
# 4. (separate call, 8 GPUs)  BASELINE.json configs[3] at its stated size, 1.0165e8 atoms:
#   gpurun --gpus 8 --timeout 900 -- 'BENCH_ARGS="--workload cu_fcc_1e8 --steps 100" bash tools/gpu_scale.sh 8'
# 5. (second call) launch list + full capture of the second-generation rjl kernels, to replace profiles/r1e_* :
#   gpurun --timeout 900 -- 'ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 160 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 25 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_list.log 2>&1; bash tools/gpu_ncu.sh "k_rjl" r2_rjl_gen2'
#   then here:  python profiles/ncu_raw.py gpurun_out/r2_rjl_gen2.ncu-rep  /  python profiles/sass_mix.py ...   (see profiles/README.md)

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_zz_variants_gpu.py tests/test_zz_random_systems.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2s2_tests.log
tail -2 gpurun_out/r2s2_tests.log
for v in "" "PFMDS_NL_LCAP=90" "PFMDS_NL_CELL=0"; do
  env $v python bench.py --steps 100 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > "gpurun_out/r2s2_bench_n1_${v}.json" 2>> gpurun_out/r2s2_bench.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s2_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], d.get("kernels_ms_per_step"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_build' -s 1 -c 1 -o gpurun_out/r2s2_build python bench.py --steps 3 --warmup 21 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/r2s2_ncu_build.log 2>&1

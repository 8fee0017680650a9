# 1 GPU: small-system graph depth (pre-open, lj branches, normals in branch): parity + A/B;  then compute-sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2m_tests.log
tail -3 gpurun_out/r2m_tests.log
for wl in ab_gas graphene_cu; do
  python bench.py --workload $wl --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2m_${wl}.json 2>> gpurun_out/r2m_bench.err
  PFMDS_PRE_OPEN=0 python bench.py --workload $wl --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2m_${wl}_nopreopen.json 2>> gpurun_out/r2m_bench.err
done
python bench.py --workload ensemble_graphene --steps 400 --warmup 21 > gpurun_out/r2m_ensemble_1gpu.json 2>> gpurun_out/r2m_bench.err
tail -c 600 gpurun_out/r2m_bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "launches", d.get("gpu_launches"))
PY
bash tools/gpu_r2_sanitize.sh

mkdir -p gpurun_out
for wl in ab_gas graphene_cu; do
PFMDS_GRAPHS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 80 --csv --log-file gpurun_out/r2n_launches_$wl.csv python bench.py --workload $wl --steps 30 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2n_ncu_$wl.log 2>&1
done
PFMDS_GRAPHS=0 timeout 200 python bench.py --workload ab_gas --steps 2000 --warmup 21 --no-cpu-baseline --no-e2e > gpurun_out/r2n_ab_gas_nographs.json 2>gpurun_out/r2n.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, "%.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "launches", d.get("gpu_launches"))
PY

# compile-time variants of the rjl kernels, one bench line each (run under gpurun; nvcc is on the box)
for v in "-DRJL_MINB=4" "-DRJL_MINB=5" "-DRJL_MINB=7" "-DRJL_MINB=8"; do
  PFMDS_NVCC_EXTRA="$v" python -m pfmds_b200.build --force > /dev/null 2>&1
  echo "== $v"; python bench.py --steps 60 --warmup 21 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernels_ms_per_step']; print('  ms/step %.4f force %.4f density %.4f'%(d['ms_per_step'],k['rjl_force'],k['rjl_density']))
"
done

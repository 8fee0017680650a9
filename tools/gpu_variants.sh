# compile-time variants, one short bench line each (run under gpurun; nvcc is on the box); the default build comes last
timeout 120 python -m pytest tests -m gpu -x -q -k "store_instead or full_size" 2>&1 | tail -2
for v in "-DMX_NOINLINE_SWITCH" "-DRJL_MINB_D=8" "-DRJL_MINB_D=6" ""; do
  PFMDS_NVCC_EXTRA="$v" python -m pfmds_b200.build --force > /dev/null 2>&1
  echo "== variant [$v]"; python bench.py --steps 60 --warmup 21 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernels_ms_per_step']; print('  ms/step %.4f force %.4f density %.4f zero %s'%(d['ms_per_step'],k['rjl_force'],k['rjl_density'],k.get('zero_forces')))
"
done

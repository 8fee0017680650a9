# compute-sanitizer over a parity subset that launches every kernel family on BOTH sides of the size switches;
# one B200.  memcheck: out-of-bounds / misaligned global + shared accesses, leaks of device errors;  racecheck: shared-memory hazards
# (block_sum, scans, the staged list build);  initcheck: reads of uninitialised device memory.
mkdir -p gpurun_out
SUB='tests/test_parity_gpu.py::test_step0_lists_forces_energies tests/test_parity_gpu.py::test_energies_from_the_force_pass tests/test_deposition_gpu.py::test_deposition_matches_the_oracle tests/test_rebosc_gpu.py::test_rebosc_energy_and_numerical_forces'
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest $SUB "tests/test_parity_gpu.py::test_trajectory_22_steps" -k "not nvms and not nve" -m gpu -q -x -p no:cacheprovider > gpurun_out/r2s_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2s_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2s_$tool.log | tail -3 >> gpurun_out/r2s_summary.txt
done
# the 10^5-atom kernels as the bench launches them (thread per atom, cell-tiled list build with bulk copies, fused kick), a few steps
cat > /tmp/san_big.py <<'PY'
import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests")
from pfmds_b200 import inputs
from pfmds_b200.engine import configure
import numpy as np
case = inputs.cu_fcc(ncell=30, jitter=0.03, period=2)
e = configure(case)
e.advance("nvt", 2.0, 0, 4, with_energy=True)
rows = e.advance_logged("nvt", 2.0, 4, 3, log_period=1)
p, v, f = e.download()
assert np.isfinite(f).all() and np.abs(f).max() > 0.05
print("big ok", e.energies()[0], e.launch_count())
PY
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python /tmp/san_big.py > gpurun_out/r2s_big_$tool.log 2>&1
  echo "big $tool rc=$?" >> gpurun_out/r2s_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|big ok" gpurun_out/r2s_big_$tool.log | tail -3 >> gpurun_out/r2s_summary.txt
done
cat gpurun_out/r2s_summary.txt

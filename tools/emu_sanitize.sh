# Host replay of the device library (tests/emu) under AddressSanitizer + UBSan: every cudaMalloc is a malloc there, so a slot number,
# list row or partial-sum index out of range in a kernel or in the C-ABI orchestration is a heap-buffer-overflow report.  ~8 min on CPU.
#   bash tools/emu_sanitize.sh [pytest args]          e.g.  -k serial   (53 tests, 8 min)   or   -k 'lockstep and (step0 or golden or anchors or parity_misc)'
# Last runs: serial flavour 53 passed, lock-step subset 21 passed, slab decomposition with ranks as threads
# (pytest tests/test_slab_lockstep.py under the same environment) 4 passed; no sanitizer report in any.
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so)"   # libstdc++ first-loaded: ASan intercepts __cxa_throw
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 PFMDS_EMU_SANITIZE=1
python -m pytest tests/test_emulated_library.py -q -p no:cacheprovider "$@" 2>&1 | tee /tmp/pfmds_emu_sanitize.log | tail -5
echo "sanitizer reports: $(grep -c -E 'runtime error|ERROR: AddressSanitizer' /tmp/pfmds_emu_sanitize.log)"

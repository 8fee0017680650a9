"""Where does a small system's step go?  Debug build of the library with per-step time stamps (-DPFMDS_STAMPS, common.cuh):
globaltimer at entry / exit of kick+drift, the lj / lj1g branches and the phases of k_sum_kick_ke, for every step of a run.
Not the product: builds tools/_stamps/libpfmds_b200_stamps.so (git-ignored), prints the mean gaps between the stamps.

  python tools/stamps_probe.py build          (here, no GPU)
  python tools/stamps_probe.py run [workload] (GPU box)
"""
import ctypes as C
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_stamps")
LIB = os.path.join(OUT, "libpfmds_b200_stamps.so")
SLOTS, STEPS = 16, 4096
NAMES = {0: "kick_drift entry (min)", 1: "kick_drift exit (max)", 2: "force branches entry (min)", 3: "force branches exit (max)",
         4: "sum_kick entry (min)", 5: "sum_kick after atom loop (max)", 6: "sum_kick partials stored (max)", 7: "last block knows (max)",
         8: "last block summed partials (max)", 9: "chain update done (max)"}


def build():
    from pfmds_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    def one(s):
        o = os.path.join(OUT, s[:-3] + ".o")
        r = subprocess.run([b.NVCC] + b.NVFLAGS + ["-DPFMDS_STAMPS", "-I" + os.path.join(ROOT, "include"), "-c", os.path.join(b.CSRC, s), "-o", o],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            raise RuntimeError(r.stdout)
        return o
    with ThreadPoolExecutor(6) as ex:
        objs = list(ex.map(one, b.SOURCES))
    subprocess.check_call([b.NVCC] + b.ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"])
    print(LIB)


def run(workload="ab_gas", steps=1000):
    import bench
    from pfmds_b200 import engine
    case, integrator, _ = bench.build_case(workload, 1, 1000)
    dt = case["integrators"][0][1]
    eng = engine.configure(case, lib_path=LIB)
    lib = eng._lib
    eng.advance(integrator, dt, 0, 41)
    eng.synchronize()
    assert lib.pfmds_debug_stamps_begin() == 0
    eng.advance(integrator, dt, 41, steps)
    eng.synchronize()
    buf = np.zeros(1 + STEPS * 2 * SLOTS, np.uint64)
    lib.pfmds_debug_stamps_read(buf.ctypes.data_as(C.c_void_p))
    n = int(buf[0])
    st = buf[1:].reshape(STEPS, 2, SLOTS)[:n].astype(np.int64)
    mins, maxs = st[:, 0, :], st[:, 1, :]
    t = np.zeros((n, 10), np.float64)
    for s in range(10):
        t[:, s] = (mins if s in (0, 2, 4) else maxs)[:, s]
    ok = np.all(t > 0, axis=1) & np.all(mins[:, [0, 2, 4]] < (1 << 62), axis=1)
    # a step: kick_drift(s) ... chain done(s); the next step's kick_drift entry gives the launch gap
    print("workload %s: %d steps stamped, %d complete" % (workload, n, int(ok.sum())))
    step_len = np.diff(t[:, 0])
    good = ok[:-1] & ok[1:] & (step_len < 5e5)
    print("step length (kick_drift entry to the next one): mean %.2f us  median %.2f us" % (step_len[good].mean() / 1e3, np.median(step_len[good]) / 1e3))
    print("distinct values of (t mod 1024 ns) in slot 0: %d  (timer resolution check)" % len(set((t[:, 0] % 1024).astype(int).tolist())))
    prev = 0
    for s in range(1, 10):
        d = (t[:, s] - t[:, prev])[ok]
        print("  %-40s -> %-40s mean %7.2f us  median %7.2f us" % (NAMES[prev], NAMES[s], d.mean() / 1e3, np.median(d) / 1e3))
        prev = s
    d = (t[1:, 0] - t[:-1, 9])[good]
    print("  %-40s -> %-40s mean %7.2f us  median %7.2f us" % (NAMES[9], "next step's kick_drift entry", d.mean() / 1e3, np.median(d) / 1e3))
    eng.close()


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        run(*(sys.argv[2:3] or ["ab_gas"]))

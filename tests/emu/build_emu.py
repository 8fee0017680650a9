"""TEST SUPPORT ONLY.  Builds the HOST REPLAY of the device library: the very .cu sources of pfmds_b200/csrc compiled by g++
(-x c++ -DPFMDS_EMU_LIB -DSMALL_N=0) with tests/emu/host_emu.hpp standing in for the CUDA language extensions and runtime,
plus the run_md_simulation / run_gr_moire_fitting hosts linked against it.  Kernels run as serial loops over (block, thread).
This exists so that the C-ABI orchestration (pfmds_advance's step sequence, the fused NVT path, deposition, checkpoints ...) and
the kernels' arithmetic can be checked against the oracle in the CPU test suite; it is never built, loaded or shipped by the
product (pfmds_b200/build.py compiles with nvcc only, pfmds_b200/engine.py loads libpfmds_b200.so only)."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pfmds_b200", "csrc")
HOST = os.path.join(ROOT, "pfmds_b200", "host")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
# PFMDS_EMU_SANITIZE=1: the same replay under AddressSanitizer + UBSan in its own directory (cudaMalloc maps onto malloc, so an
# out-of-range slot number, list row or partial-sum index of a kernel is a heap-buffer-overflow report).  Run as
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 PFMDS_EMU_SANITIZE=1 pytest tests/test_emulated_library.py
SAN = os.environ.get("PFMDS_EMU_SANITIZE") == "1"
SANFLAGS = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g"] if SAN else []
SOURCES = ["capi.cu", "forces.cu", "nl.cu", "integrate.cu", "rebosc.cu"]


class Build:
    """One flavour of the host replay.  warp=False: the SERIAL replay (threads of a block one after the other; the library takes
    its thread-per-atom kernels, -DSMALL_N=0).  warp=True: the LOCK-STEP replay (-DPFMDS_EMU_WARP, host_emu.hpp): every thread of
    a block is a fiber, warp collectives and block barriers exchange values between them, so the library takes the very kernels
    the GPU launches for small systems (8 lanes per atom, warp-per-atom list build, scans, shared-memory reductions)."""

    def __init__(self, warp=False):
        self.warp = warp
        self.out = os.path.join(HERE, "_build" + ("_warp" if warp else "") + ("_san" if SAN else ""))
        self.flags = ["-x", "c++", "-std=c++17", "-O1" if SAN else "-O2", "-ffp-contract=off", "-DPFMDS_EMU_LIB", "-fPIC", "-I", HERE] + \
                     (["-DPFMDS_EMU_WARP"] if warp else ["-DSMALL_N=0"]) + SANFLAGS
        self.lib = os.path.join(self.out, "libpfmds_b200_emu.so")
        self.exe = os.path.join(self.out, "run_md_simulation_emu")
        self.exe_fit = os.path.join(self.out, "run_gr_moire_fitting_emu")

    def build(self):
        os.makedirs(self.out, exist_ok=True)
        hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".hpp"))] + [os.path.join(ROOT, "include", "pfmds_b200.h")] + \
               [os.path.join(HERE, h) for h in ("host_emu.hpp", "nccl_emu.hpp")]
        jobs, objs = [], []
        for s in SOURCES + (["slab.cu"] if self.warp else []):   # the slab decomposition needs the lock-step flavour (scans, spin waits)
            obj = os.path.join(self.out, s[:-3] + ".o")
            objs.append(obj)
            if _newer(obj, [os.path.join(CSRC, s)] + hdrs):
                jobs.append([CXX] + self.flags + ["-c", os.path.join(CSRC, s), "-o", obj])
        stub = os.path.join(self.out, "slab_stub.o")
        if not self.warp:
            objs.append(stub)
        if not self.warp and _newer(stub, [os.path.join(HERE, "slab_stub.cpp")] + hdrs):
            jobs.append([CXX, "-std=c++17", "-O2", "-fPIC", "-I", HERE] + SANFLAGS + ["-c", os.path.join(HERE, "slab_stub.cpp"), "-o", stub])
        with ThreadPoolExecutor(4) as ex:
            list(ex.map(_run, jobs))
        if jobs or _newer(self.lib, objs):
            _run([CXX, "-shared"] + SANFLAGS + ["-o", self.lib] + objs + ["-ldl", "-pthread"])
        host_deps = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".hpp", ".cpp"))]
        for exe, src in ((self.exe, "run_md_simulation.cpp"), (self.exe_fit, "run_gr_moire_fitting.cpp")):
            if _newer(exe, host_deps + [self.lib]):
                _run([CXX, "-O2", "-std=c++17"] + SANFLAGS + ["-o", exe, os.path.join(HOST, src), "-pthread", "-L" + self.out, "-lpfmds_b200_emu", "-Wl,-rpath,$ORIGIN"])
        return self.lib


def _newer(target, deps):
    return not os.path.exists(target) or any(os.path.getmtime(d) > os.path.getmtime(target) for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulated build failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))


SERIAL, LOCKSTEP = Build(False), Build(True)
# the serial replay under its historical names
OUT, LIB, EXE, EXE_FIT = SERIAL.out, SERIAL.lib, SERIAL.exe, SERIAL.exe_fit


def build_emu():
    return SERIAL.build()


if __name__ == "__main__":
    import sys
    print((LOCKSTEP if "--lockstep" in sys.argv else SERIAL).build())

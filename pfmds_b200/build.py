"""In-tree build of the CUDA library (sm_100a) and the run_md_simulation host.  nvcc cross-compiles
without a GPU; the products (libpfmds_b200.so, run_md_simulation) stay next to their sources so they
travel to the GPU box with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", CXX] + os.environ.get("PFMDS_NVCC_EXTRA", "").split()
SOURCES = ["nl.cu", "forces.cu", "rebosc.cu", "integrate.cu", "capi.cu", "slab.cu"]
LIB = os.path.join(CSRC, "libpfmds_b200.so")
EXE = os.path.join(HOST, "run_md_simulation")
EXE_FIT = os.path.join(HOST, "run_gr_moire_fitting")
EXE_ANALYSIS = os.path.join(HOST, "run_gr_analysis")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith((".cuh", ".hpp"))] + [os.path.join(HERE, "..", "include", "pfmds_b200.h")]
    objs = []
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(CSRC, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append([NVCC] + NVFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(4) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        print("\n".join(outs))
    if force or jobs or _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-ccbin", CXX, "-o", LIB] + objs + ["-ldl"])
    host_deps = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".hpp", ".cpp"))]
    for exe, src in ((EXE, "run_md_simulation.cpp"), (EXE_FIT, "run_gr_moire_fitting.cpp")):
        if force or _newer(exe, host_deps + [LIB]):
            _run([CXX, "-O2", "-std=c++17", "-o", exe, os.path.join(HOST, src), "-pthread", "-L" + CSRC, "-lpfmds_b200", "-Wl,-rpath,$ORIGIN/../csrc"])
    if force or _newer(EXE_ANALYSIS, host_deps):
        _run([CXX, "-O2", "-std=c++17", "-o", EXE_ANALYSIS, os.path.join(HOST, "run_gr_analysis.cpp")])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

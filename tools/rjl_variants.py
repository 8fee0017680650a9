"""A/B of compile-time variants of the rjl pair kernels (block size, resident blocks per SM of the density / force pass) on the bench
workload.  Each variant is forces.cu rebuilt with other -D flags and linked with the product's other objects into
tools/_variants/lib_<name>.so (git-ignored; not the product).

  python tools/rjl_variants.py build   (here, no GPU)
  python tools/rjl_variants.py run     (GPU box)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_variants")
VARIANTS = {
    "base": [],
    "dens8": ["-DRJL_MINB_D=8"],
    "dens9": ["-DRJL_MINB_D=9"],
    "dens6": ["-DRJL_MINB_D=6"],
    "ft64": ["-DFT=64", "-DRJL_MINB=14", "-DRJL_MINB_D=14"],
    "ft64_d16": ["-DFT=64", "-DRJL_MINB=14", "-DRJL_MINB_D=16"],
    "ft256": ["-DFT=256", "-DRJL_MINB=3", "-DRJL_MINB_D=4"],
    "ft96": ["-DFT=96", "-DRJL_MINB=9", "-DRJL_MINB_D=10"],
}


def build():
    from pfmds_b200 import build as b
    b.build()
    os.makedirs(OUT, exist_ok=True)
    others = [os.path.join(b.CSRC, s[:-3] + ".o") for s in b.SOURCES if s != "forces.cu"]
    for name, flags in VARIANTS.items():
        o = os.path.join(OUT, "forces_%s.o" % name)
        r = subprocess.run([b.NVCC] + b.NVFLAGS + flags + ["-Xptxas", "-v", "-c", os.path.join(b.CSRC, "forces.cu"), "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            raise RuntimeError(r.stdout[-3000:])
        lines = r.stdout.splitlines()
        for i, l in enumerate(lines):
            if "Compiling entry function" in l and ("k_rjl_forceI4RjlFE" in l or "k_rjl_densityILb0E4RjlDE" in l):
                print(name, l.split("'")[1][:40], lines[i + 2].strip(), lines[i + 3].strip()[:40])
        subprocess.check_call([b.NVCC] + b.ARCH + ["-shared", "-o", os.path.join(OUT, "lib_%s.so" % name), o] + others + ["-ldl"])


def run():
    import bench
    from pfmds_b200.engine import configure
    case, integ, _ = bench.build_case("cu_fcc", 2, 200)
    for rep in range(2):
        for name in VARIANTS:
            lib = os.path.join(OUT, "lib_%s.so" % name)
            if not os.path.exists(lib):
                continue
            eng = configure(case, lib_path=lib)
            eng.advance(integ, 2.0, 0, 21)
            eng.synchronize()
            eng.timer_start()
            eng.advance(integ, 2.0, 21, 100)
            ms = eng.timer_stop()
            eng.set_profiling(True)
            eng.advance(integ, 2.0, 121, 40)
            kt = eng.kernel_times()
            eng.close()
            k = {n: round(v[0] / max(v[1], 1), 4) if isinstance(v, (tuple, list)) else v for n, v in kt.items()} if isinstance(kt, dict) else kt
            print("%-9s %.4f ms/step   %s" % (name, ms / 100, {n: k[n] for n in k if n in ("rjl_force", "rjl_density")} if isinstance(k, dict) else k), flush=True)


if __name__ == "__main__":
    build() if sys.argv[1] == "build" else run()

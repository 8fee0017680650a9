#!/usr/bin/env python
"""bench.py — atom-steps/s of the PFMDS MD inner loop on B200 (BASELINE.json metric).

Workload at N=1: BASELINE.json configs[1] — Cu fcc crystal, rjl (Rosato-Guillope-Legrand), NVT at
300 K, 63^3 cells = 1 000 188 atoms, dt 2 fs, neighbour-list rebuild every 20 steps, synthetic
lattice with Maxwell velocities (seed 2).  One "step" is one MD step of md()'s loop
(md_simulation.f90:138-186) through the C ABI (pfmds_advance), including the rebuilds that fall
into the timed region.  N>1 (torchrun, one rank per GPU), weak scaling at 1 000 188 atoms per GPU:
  --decomp slab (default)   ONE crystal of N x 63 x 63 x 63 cells split into x-slabs (BASELINE.json configs[3]):
                            ghost positions and ghost 1/Eb exchanged with the two neighbours every step,
                            KE all-reduced for the thermostat, atoms migrate at list rebuilds (NCCL inside
                            the library, on the context's stream);
  --decomp ensemble         the reference's MPI mode: every rank integrates its own replica, no collective.
`value` is the atoms of all ranks x steps divided by the slowest rank's device time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cu_fcc|cu_fcc_1e8|lj_fluid|ab_gas|graphene_cu|ensemble_graphene|graphene_rebosc|lj_deposition]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic (source-level) FP64 operation counts per listed directed pair and HBM bytes per list-owner atom and launch
# (4n for the int32 row + per-atom records): frozen next to the kernel source in pfmds_b200/csrc/roofline.json
# (DESIGN.md section 4 / SURVEY.md 8d: FMA = 2, every exp/sqrt/div/sin/cos = 1, distance + min-image = 30).
_RM = json.load(open(os.path.join(ROOT, "pfmds_b200", "csrc", "roofline.json")))["kernels"]
FLOPS_PER_PAIR = {k: v["flop_per_pair"] for k, v in _RM.items()}
_RMX = json.load(open(os.path.join(ROOT, "pfmds_b200", "csrc", "roofline.json"))).get("executed_fp64_per_pair_ncu", {})
_RML = json.load(open(os.path.join(ROOT, "pfmds_b200", "csrc", "roofline.json"))).get("l1_gather_model")
BYTES_PER_ATOM = {k: (lambda n, c=v["bytes_per_atom_const"]: 4 * n + c) for k, v in _RM.items() if v.get("bytes_per_atom_const") is not None}


BIG_CELLS_PER_RANK = (37, 293, 293)   # x 8 ranks = 296 x 293 x 293 cells = 101 645 216 atoms (BASELINE.json configs[3])


def build_case(workload, seed, steps, nx=1):
    from pfmds_b200 import inputs
    if workload == "cu_fcc":
        if nx > 1:
            return (inputs.cu_fcc(cells=(63 * nx, 63, 63), seed=seed, steps=steps), "nvt",
                    "Cu fcc %dx63x63 cells = %d atoms in x-slabs, rjl, NVT 300 K, dt 2 fs, r_cut 6.5, rebuild/20" % (63 * nx, 1000188 * nx))
        return inputs.cu_fcc(ncell=63, seed=seed, steps=steps), "nvt", "Cu fcc 63^3x4 = 1000188 atoms, rjl, NVT 300 K, dt 2 fs, r_cut 6.5, rebuild/20"
    if workload == "cu_fcc_1e8":   # one GPU's share of the 10^8-atom crystal (the N>1 path generates per rank, see main)
        cx, cy, cz = BIG_CELLS_PER_RANK
        return (inputs.cu_fcc(cells=BIG_CELLS_PER_RANK, seed=seed, steps=steps), "nvt",
                "Cu fcc %dx%dx%d cells = %d atoms, rjl, NVT 300 K, dt 2 fs, r_cut 6.5, rebuild/20" % (cx, cy, cz, 4 * cx * cy * cz))
    if workload == "lj_fluid":
        return inputs.lj_fluid(n_side=128, seed=seed, steps=steps), "nve", "LJ fluid (lj1g) 128^3 = 2097152 atoms, NVE, dt 0.5 fs, r_cut 7.5, rebuild/20"
    if workload == "ab_gas":
        return inputs.ab_gas(seed=seed), "nvt", "A/B LJ gas 22^3 = 10648 atoms, lj + 2 x lj1g, NVT 100 K, dt 0.5 fs"
    if workload == "graphene_rebosc":   # numerical forces, 200 x 116 cells = 92 800 C atoms
        return inputs.graphene_rebosc(cells=(200, 116), seed=seed, steps=steps), "nve", "graphene sheet 200x116 cells = 92800 atoms, rebosc (numerical forces), NVE, dt 0.5 fs"
    if workload == "lj_deposition":     # a growing group (small system, launch bound)
        return inputs.lj_deposition(n_side=40, n_layers=6, n_deposit=400, steps=steps, ts2=10 ** 9), "nvt", "LJ substrate 40x40x6 + 400 deposited atoms (change_group_num), NVT 80 K, dt 1 fs"
    if workload == "graphene_cu":
        return inputs.graphene_on_cu(seed=seed), "nvt", "graphene on Cu(111) moire, 11028 atoms, tb + ljc + rjl, NVT 300 K, dt 1 fs"
    raise SystemExit("unknown workload " + workload)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    (profiles/r2_traffic.json, same workload); None when no capture exists for it."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(kernel)
    except Exception:
        return None


def cpu_baseline_sample(steps, threads=None, warmup=None):
    """The CPU oracle (oracle/, the reference's algorithm restated in C++/OpenMP; kind "port") timed on the
    host cores on a bounded sample of the same workload: Cu fcc 20^3 cells = 32 000 atoms (the reference's
    O(N^2) rebuild makes 10^6 atoms infeasible: ~10^12 pair tests per rebuild)."""
    from pfmds_b200 import inputs
    from pfmds_b200.engine import configure, load_library
    lib = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(lib):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, stdout=subprocess.DEVNULL)
    # the timed CPU arm uses the -march=native flavour, compiled on THIS box; the portable one if that fails
    native = os.path.join(ROOT, "oracle", "_build", "native", "liboracle.so")
    tag = os.path.join(ROOT, "oracle", "_build", "native", "host")
    try:
        cpu_id = " ".join(l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags")))[:20000]
    except Exception:
        cpu_id = "unknown"
    if not (os.path.exists(native) and os.path.exists(tag) and open(tag).read() == cpu_id):   # a library built for another CPU does not count
        try:
            if os.path.exists(native):
                os.remove(native)
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "native"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
            open(tag, "w").write(cpu_id)
        except Exception:
            native = None
    if native and os.path.exists(native):
        lib = native
    cores = threads or os.cpu_count() or 1
    L = load_library(lib, "oracle_")
    L.oracle_set_threads.restype = int
    cores = L.oracle_set_threads(int(cores))
    case = inputs.cu_fcc(ncell=20, steps=steps + (warmup or 0))
    n = len(case["mass"])
    e = configure(case, lib_path=lib, prefix="oracle_")
    if warmup is not None:   # reference arm: step 0 and `warmup` steps untimed, then exactly `steps` steps (with the rebuilds that fall among them)
        e.advance("nvt", 2.0, 0, 1 + warmup)
        t0 = time.perf_counter()
        e.advance("nvt", 2.0, 1 + warmup, steps)
        return n, steps, time.perf_counter() - t0, cores
    t0 = time.perf_counter()
    e.advance("nvt", 2.0, 0, 1)
    e.advance("nvt", 2.0, 1, steps)
    dt = time.perf_counter() - t0
    return n, steps, dt, cores


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (the C++/OpenMP restatement in oracle/; the
    Fortran original cannot be compiled in this image) on all host cores, on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    k = max(1, min(args.steps, 200))    # bounded: ~0.1 s per step on 16 cores, 200 steps = 10 O(N^2) rebuilds
    w = max(0, min(args.warmup, 21))
    n, steps, dt, cores = cpu_baseline_sample(k, warmup=w)
    value = n * steps / dt
    sample = "Cu fcc 20^3x4 = %d atoms (bounded sample of the 1000188-atom workload), %d NVT steps timed after step 0 + %d warm-up steps, rebuild/20 by the reference's O(N^2) search" % (n, steps, w)
    line = {
        "impl": "reference", "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": w,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "Cu fcc, rjl, NVT 300 K, dt 2 fs (BASELINE.json configs[1]); CPU arm runs a bounded sample", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ensemble(args, rank, world, local, dist, K, W):
    """BASELINE.json configs[4]: 64 independent graphene-on-Cu runs (run_md_simulation_mpi list mode), sharded over the
    ranks by the reference's rule mod(i-1,n)==rank-1 and executed concurrently on each GPU, one context (CUDA stream)
    per run, driven from one host thread through the asynchronous pfmds_advance."""
    import torch
    from pfmds_b200 import inputs
    from pfmds_b200.engine import configure
    from pfmds_b200.ensemble import shard
    runs = shard(64, world, rank)
    ctxs = []
    for i in runs:
        case = inputs.graphene_on_cu(seed=i)
        ctxs.append(configure(case, device=local))
    n_atoms = len(case["mass"])
    chunk = 10
    for e in ctxs:
        e.advance("nvt", 1.0, 0, W)
    for e in ctxs:
        e.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = sum(e.launch_count() for e in ctxs)
    with ClockSampler(local) as cs:
        for e in ctxs:
            e.timer_start()
        for s in range(W, W + K, chunk):
            for e in ctxs:
                e.advance("nvt", 1.0, s, min(chunk, W + K - s))
        ms = max(e.timer_stop() for e in ctxs)
        torch.cuda.synchronize()
    launches = sum(e.launch_count() for e in ctxs) - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        line = {"metric": "atom-steps/s", "value": 64 * n_atoms * K / (ms * 1e-3), "unit": "atom-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "ensemble of 64 graphene-on-Cu(111) runs (tb + ljc + rjl, NVT 300 K, 11028 atoms each), %d per GPU on separate streams" % len(runs),
                           "runs_per_gpu": len(runs), "l2": "each replica is L2 resident; replicas differ (seeds 1..64)"},
                "clocks": cs.summary(), "e2e": None, "gpu_launches": launches, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def time_variant(case, integrator, dt, local, env, W, K):
    """ms/step and per-kernel times of the same workload in a fresh context created under `env` (kernel variants are chosen
    from the environment at pfmds_create).  Not part of `value`: reported beside it under "variants"."""
    from pfmds_b200.engine import configure
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        eng = configure(case, device=local)
        eng.advance(integrator, dt, 0, W)
        eng.synchronize()
        eng.set_profiling(True)
        eng.timer_start()
        eng.advance(integrator, dt, W, K)
        ms = eng.timer_stop()
        kt = eng.kernel_times()
        eng.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return {"ms_per_step": ms / K, "steps": K, "atom_steps_per_s": len(case["mass"]) * K / (ms * 1e-3),
            "kernels_ms_per_step": {k: round(v[0] / K, 5) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])[:6]}}


def e2e_steps(eng, integrator, dt, ke2e, stepwise=False):
    """Steps 1..ke2e of the e2e leg with the energies of EVERY step brought to the host (period_log = 1).  Default: what the
    run_md_simulation host does (md_driver.hpp) — one pfmds_advance_logged call, every step's energies produced by its force pass
    and logged on the device, one D2H copy of the rows (in slab mode every rank gets the all-reduced rows).  stepwise
    (PFMDS_BENCH_STEPWISE_E2E=1, or if the logged call fails): one pfmds_advance_with_energy(1) + pfmds_energies round trip per step.  Returns (D2H energy bytes per step, description)."""
    if not stepwise:
        try:
            rows = eng.advance_logged(integrator, dt, 1, ke2e, log_period=1)
            assert rows[0].shape[0] == ke2e and np.isfinite(rows[0]).all() and np.isfinite(rows[1]).all()
            return 8 * (rows[0].shape[1] + 1 + rows[3].shape[1] * 3 * 3), "pfmds_advance_logged(%d steps, log_period 1: energies of every step, one D2H of the log)" % ke2e
        except Exception as ex:
            print("bench: pfmds_advance_logged failed (%r), e2e falls back to one call per step" % (ex,), file=sys.stderr)
    e_bytes = 0
    for s in range(1, ke2e + 1):
        eng.advance(integrator, dt, s, 1, with_energy=True)
        e = eng.energies()
        e_bytes = 8 * (len(e[0]) + 1 + len(e[3]) * (3 * 3 + 2)) + 16
    return e_bytes, "%d x [pfmds_advance_with_energy(1) + pfmds_energies (D2H)]" % ke2e


def run_variants(local, small=False):
    """--variants-only (child of the default run): ms/step of the rjl and lj1g kernel variants on this GPU, one cumulative JSON line
    per finished experiment.  small: miniature systems and a few steps (the CPU suite runs this function on the host replay)."""
    from pfmds_b200 import inputs
    W, K = (2, 3) if small else (21, 100)
    out = {"note": "child process, fresh contexts after the headline measurement, %d warm-up + %d timed steps each; `value` is the default configuration" % (W, K)}
    if small:
        case, integrator = inputs.cu_fcc(ncell=4, jitter=0.05, period=5), "nvt"
        ljc = inputs.lj_fluid(n_side=8, seed=2, steps=20, period=5)
    else:
        from pfmds_b200.build import build
        build()
        case, integrator, _ = build_case("cu_fcc", seed=2, steps=200)
        ljc = inputs.lj_fluid(n_side=96, seed=2, steps=200)
    dt = case["integrators"][0][1]
    if not small:   # elementary functions of mathx.cuh on this device (the bounds tests/test_parity_gpu.py asserts): [exp, switch, rsqrt, seed] / short forms
        try:
            import ctypes as C
            from pfmds_b200.engine import load_library
            lib, err = load_library(), (C.c_double * 4)()
            for fn in ("pfmds_selftest_math", "pfmds_selftest_math2"):
                f = getattr(lib, fn)
                f.argtypes, f.restype = [C.c_int, C.POINTER(C.c_double)], C.c_int
                rc = f(local, err)
                out[fn] = {"rc": rc, "err": [float(x) for x in err]}
        except Exception as ex:
            out["device_math"] = {"error": repr(ex)[:300]}
    for name, cs, integ, h, env in (
            ("cu_fcc defaults", case, integrator, dt, {}),
            ("cu_fcc, rjl third generation (PFMDS_RJL_GEN=3: node-table exponentials, fewer FP64 instructions but twice the L1 wavefronts)", case, integrator, dt, {"PFMDS_RJL_GEN": "3"}),
            ("cu_fcc, rjl first generation (PFMDS_RJL_GEN=1, the round-1 kernels)", case, integrator, dt, {"PFMDS_RJL_GEN": "1"}),
            ("cu_fcc, list build with the exact test inside the candidate loop (PFMDS_NL_MASK=0, k_build): compare nl_build", case, integrator, dt, {"PFMDS_NL_MASK": "0"}),
            ("lj_fluid 96^3 lj1g defaults (pipelined kernel)", ljc, "nve", ljc["integrators"][0][1], {}),
            ("lj_fluid 96^3 lj1g plain kernel (PFMDS_LJ1G_PIPE=0)", ljc, "nve", ljc["integrators"][0][1], {"PFMDS_LJ1G_PIPE": "0"})):
        try:
            out[name] = time_variant(cs, integ, h, local, env, W, K)
        except Exception as ex:
            out[name] = {"error": repr(ex)[:300]}
        print(json.dumps(out), flush=True)   # cumulative: the parent keeps the last complete line, also when it has to stop this process
    # the other device rows of SURVEY 8(f) and the small-system configurations, default kernels (per-kernel events on: small systems
    # run without their CUDA graphs here, so these are upper bounds of their ms/step; `bench.py --workload X` is the clean line)
    for wl, k in (("graphene_rebosc", 50), ("lj_deposition", 200), ("ab_gas", 200), ("graphene_cu", 200)):
        try:
            if small:
                k = K
                cs, integ, desc = {"graphene_rebosc": lambda: (inputs.graphene_rebosc(), "nve", "miniature"),
                                   "lj_deposition": lambda: (inputs.lj_deposition(), "nvt", "miniature"),
                                   "ab_gas": lambda: (inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=5), "nvt", "miniature"),
                                   "graphene_cu": lambda: (inputs.graphene_on_cu_small(interface="ljc", period=5), "nvt", "miniature")}[wl]()
            else:
                cs, integ, desc = build_case(wl, seed=2, steps=21 + k)
            out["workload " + wl + ": " + desc] = time_variant(cs, integ, cs["integrators"][0][1], local, {}, W, k)
        except Exception as ex:
            out["workload " + wl] = {"error": repr(ex)[:300]}
        print(json.dumps(out), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=21)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="cu_fcc")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the A/B timing of alternative kernel variants after the headline measurement")
    ap.add_argument("--decomp", default="slab", choices=["slab", "ensemble"])
    ap.add_argument("--variants-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--device", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.variants_only:
        return run_variants(args.device)

    import torch
    from pfmds_b200.build import build
    from pfmds_b200.engine import configure, measure_peaks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (pfmds_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build()

    K, W = args.steps, max(3, args.warmup)
    if args.workload == "ensemble_graphene":
        return run_ensemble(args, rank, world, local, dist, K, W)
    slab = world > 1 and args.decomp == "slab" and args.workload in ("cu_fcc", "cu_fcc_1e8")
    n_global = None
    if slab and args.workload == "cu_fcc_1e8":
        # BASELINE.json configs[3] at its stated size: 8 x (37 x 293 x 293) cells = 1.0165e8 atoms on 8 GPUs (1.27e7 per GPU at any N);
        # every rank generates only its own slab, the velocity initialisation's global sums go through one all-reduce
        from pfmds_b200.slab import broadcast_unique_id, configure_slab, cu_fcc_slab_inputs

        def allsum(v):
            t = torch.from_numpy(np.ascontiguousarray(v, np.float64)).cuda()
            dist.all_reduce(t)
            return t.cpu().numpy()
        case, loc = cu_fcc_slab_inputs(rank, world, BIG_CELLS_PER_RANK, allsum, seed=2, steps=K + W)
        integrator, n_global = "nvt", loc["n_global"]
        desc = "Cu fcc %dx293x293 cells = %d atoms in x-slabs (generated per rank), rjl, NVT 300 K, dt 2 fs, r_cut 6.5, rebuild/20" % (37 * world, n_global)
        uid = broadcast_unique_id(dist, torch.device("cuda", local))
        eng = configure_slab(case, rank, world, local, uid, capacity_factor=1.15, local=loc)
        n_atoms = eng.n_local0
        del loc
    elif slab:
        from pfmds_b200.slab import broadcast_unique_id, configure_slab
        case, integrator, desc = build_case(args.workload, seed=2, steps=K + W, nx=world)   # every rank builds the same crystal
        uid = broadcast_unique_id(dist, torch.device("cuda", local))
        eng = configure_slab(case, rank, world, local, uid, capacity_factor=1.25)
        n_atoms = eng.n_local0
        n_global = len(case["mass"])
    else:
        case, integrator, desc = build_case(args.workload, seed=2 + rank, steps=K + W)
        eng = configure(case, device=local)
        n_atoms = len(case["mass"])
    dt = case["integrators"][0][1]
    n_total = n_global if slab else n_atoms * world

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-GPU lines check themselves (the driver's test box has one GPU: the 2-GPU pytest never runs there) ----
    check = None
    if slab:
        check = {}
        try:
            from pfmds_b200.slab import slab_parity_check
            check["slab_parity_detail"] = slab_parity_check(dist, rank, world, local)
            check["slab_parity"] = "ok"
        except AssertionError as ex:
            check["slab_parity"] = "FAILED: " + repr(ex)[:300]

    # ---- warm-up: step 0 (lists + forces) and W-1 steps ----
    eng.advance(integrator, dt, 0, 1)
    if slab:   # potential energy per atom of the perfect lattice at step 0 against a single-context crystal (size independent)
        e0 = eng.energies()                      # all-reduced inside the library: every rank calls
        check["pe_per_atom_step0"] = float(np.sum(e0[0]) / n_total)
        if rank == 0:
            from pfmds_b200 import inputs as _inp
            small = configure(_inp.cu_fcc(ncell=4), device=local)
            small.advance("nvt", dt, 0, 1)
            pe1 = float(np.sum(small.energies()[0]) / 256)
            small.close()
            check["pe_per_atom_single_gpu"] = pe1
            check["pe_rel_diff"] = abs(check["pe_per_atom_step0"] - pe1) / abs(pe1)
    eng.advance(integrator, dt, 1, W - 1)
    eng.synchronize()
    pairs = eng.pair_count(0, 0) // (world if slab else 1)   # the library sums over ranks in slab mode
    pairs_within_cache = {}

    def eng_pairs_within(r):
        return pairs_within_cache[r]
    try:   # counted now (the engine is closed before the variants child runs)
        it0 = case["interactions"][0]
        r2_of = {"rjl": 6, "lj1g": 3, "lj": 3}.get(it0["name"])
        if r2_of is not None:
            pairs_within_cache[it0["params"][r2_of]] = eng.pair_count_within(0, 0, it0["params"][r2_of]) // (world if slab else 1)
    except Exception as ex:
        print("bench: pair_count_within failed: %r" % (ex,), file=sys.stderr)

    # ---- timed region: exactly K steps, device events on the library's stream, clocks sampled meanwhile ----
    # Per-kernel CUDA events bracket every launch inside the timed region.  Small systems (< 2e5 atoms) replay their
    # steady-state steps from CUDA graphs, which events would break up: they are timed clean and profiled in a second pass.
    launches0 = eng.launch_count()
    prof_in_region = n_atoms >= 200000
    eng.set_profiling(prof_in_region)
    with ClockSampler(local) as cs:
        barrier()   # after the sampler is up: in the lock-stepped slab mode a rank that starts late makes every other rank wait
        eng.timer_start()
        eng.advance(integrator, dt, W, K)
        ms = eng.timer_stop()
        barrier()
    launches = eng.launch_count() - launches0
    ms_prof = ms
    if not prof_in_region:
        eng.set_profiling(True)
        eng.timer_start()
        eng.advance(integrator, dt, W + K, K)
        ms_prof = eng.timer_stop()
    ktimes = eng.kernel_times()
    eng.set_profiling(False)
    if slab:   # after the timed steps (rebuilds with migration among them): every atom still has exactly one owner, sum F = 0
        fs = eng.diagnostics()[0]                # all-reduced inside the library
        cnt = torch.tensor([eng.slab_counts()[0]], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        check["atoms"] = int(cnt.item())
        check["atoms_expected"] = int(n_total)
        check["force_sum_max_abs"] = float(np.abs(fs).max())
        check["ok"] = bool(check["atoms"] == check["atoms_expected"] and check.get("slab_parity") == "ok" and check["force_sum_max_abs"] < 1e-6 and
                           (rank != 0 or check.get("pe_rel_diff", 1.0) < 1e-12))
    ms_t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = n_total * K / (ms_max * 1e-3)

    # ---- e2e: through the C ABI with host buffers.  Per run: H2D of positions+velocities from pinned host
    # memory, then ke2e steps with the energies of every step read back (what md() logs with period_log = 1,
    # see e2e_steps), and a final D2H of positions, velocities and forces.
    e2e = None

    def e2e_leg():
        ke2e = K   # the same K steps as the device-timed region, now from host buffers to host buffers
        if slab:
            gid0, p0, v0, _ = eng.download(forces=False)
            n_here = len(gid0)
        else:
            p0, v0, n_here = case["pos"], case["vel"], n_atoms
        hp = torch.empty((n_here, 3), dtype=torch.float64).pin_memory()
        hv = torch.empty((n_here, 3), dtype=torch.float64).pin_memory()
        hp.numpy()[:] = p0
        hv.numpy()[:] = v0
        # results land in pinned host memory too (single GPU; the slab path returns this rank's atoms in fresh arrays)
        ho = None if slab else [torch.empty((n_here, 3), dtype=torch.float64).pin_memory().numpy() for _ in range(3)]
        def one_pass():
            barrier()
            t0 = time.perf_counter()
            if slab:
                eng.upload_local(hp.numpy(), hv.numpy())
            else:
                eng.upload_ptr(hp.data_ptr(), hv.data_ptr())
            eng.synchronize()
            t1 = time.perf_counter()
            eng.advance(integrator, dt, 0, 1)
            eng.synchronize()
            t2 = time.perf_counter()
            e_bytes, how = e2e_steps(eng, integrator, dt, ke2e, stepwise=bool(os.environ.get("PFMDS_BENCH_STEPWISE_E2E")))
            t3 = time.perf_counter()
            out = eng.download() if slab else eng.download(out=ho)
            barrier()
            t_e2e = time.perf_counter() - t0
            phases = {"upload": (t1 - t0) * 1e3, "step0_lists_forces": (t2 - t1) * 1e3, "steps_with_energies": (t3 - t2) * 1e3, "download": (t0 + t_e2e - t3) * 1e3}
            te = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            del out
            return float(te.item()), {k: round(v, 3) for k, v in phases.items()}, e_bytes, how

        # The same leg twice, both complete (H2D of the inputs, step 0, K steps with every step's energies read back, D2H of the
        # state).  The first pass is the first time this process runs pfmds_upload / pfmds_advance_logged / pfmds_download: it also
        # pays the one-off lazy loading of those kernels' code by the CUDA runtime.  `value` is the second pass (what every later
        # call of a run costs); the first is reported beside it.
        t_cold, ph_cold, e_bytes, how = one_pass()
        t_warm, ph_warm, e_bytes, how = one_pass()
        return {"value": n_total * ke2e / t_warm, "unit": "atom-steps/s", "h2d_bytes_per_step": int(2 * 24 * n_atoms / ke2e),
                "d2h_bytes_per_step": int(e_bytes + ((3 * 32 + 4) if slab else 3 * 24) * n_atoms / ke2e), "steps": ke2e,
                "what": "pfmds_upload(H2D pinned) + " + how + " + pfmds_download(D2H); second of two identical passes, first_pass_value = the first (pays one-off kernel loading); "
                        "phases_ms [upload, step0_lists_forces, steps_with_energies, download]: second pass %s, first pass %s"
                        % (list(ph_warm.values()), list(ph_cold.values())),
                "first_pass_value": n_total * ke2e / t_cold, "phases_ms": ph_warm, "first_pass_phases_ms": ph_cold}

    if not args.no_e2e:
        if dist is not None:
            e2e = e2e_leg()   # ranks are in lock step: an exception on one of them has to end the job
        else:
            try:
                e2e = e2e_leg()
            except Exception as ex:   # the device-timed headline must survive a failing e2e leg; the failure is reported in its place
                print("bench: e2e leg failed: %r" % (ex,), file=sys.stderr)
                e2e = {"value": None, "unit": "atom-steps/s", "error": repr(ex)[:300]}

    # per-rank view (explains stragglers in the lock-stepped slab mode): own kernel times and clocks
    per_rank = None
    if dist is not None:
        mine = {"rank": rank, "ms": ms, "kernels_ms_per_step": {k: round(v[0] / K, 4) for k, v in ktimes.items()}, "clocks": cs.summary()}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_file):
        try:
            hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    try:
        dfma_tf, copy_gbs = measure_peaks(local)
    except Exception as ex:   # roofline.frac is then null and says why; the headline line is still printed
        print("bench: pfmds_measure_peaks failed: %r" % (ex,), file=sys.stderr)
        dfma_tf, copy_gbs = None, None
    top = max(ktimes.items(), key=lambda kv: kv[1][0]) if ktimes else (None, (0, 0))
    roofline = None
    if top[0] is not None:
        name, (tot_ms, cnt) = top
        avg_ms = tot_ms / max(cnt, 1)
        n_per_atom = pairs / n_atoms
        # pairs of a row between the potential's R2 and the list's r_cut leave the pair routine after the distance test:
        # they are charged the distance + minimum image (30 flop) only
        pairs_in = pairs
        try:
            it0 = case["interactions"][0]
            r2_of = {"rjl": 6, "lj1g": 3, "lj": 3}.get(it0["name"])
            if r2_of is not None and name in ("rjl_force", "rjl_density", "lj1g", "lj"):
                pairs_in = eng_pairs_within(it0["params"][r2_of])
        except Exception as ex:
            print("bench: pair_count_within failed: %r" % (ex,), file=sys.stderr)
        flops = pairs_in * FLOPS_PER_PAIR.get(name, 0) + (pairs - pairs_in) * 30
        byts = n_atoms * BYTES_PER_ATOM.get(name, lambda n: 0)(n_per_atom)
        ach_tf = flops / (avg_ms * 1e-3) * 1e-12
        ach_gbs = byts / (avg_ms * 1e-3) * 1e-9
        roofline = {
            "kernel": name, "bound": "fp64", "achieved": ach_tf, "peak": dfma_tf, "unit": "TFLOP/s", "frac": ach_tf / dfma_tf if dfma_tf else None,
            "peak_source": "pfmds_measure_peaks: DFMA micro-benchmark run live in this bench (MEASURED_PEAKS.json carries no FP64 figure)",
            "peak_theoretical": 148 * 64 * 2 * (cs.summary().get("sm_max_mhz") or 1965.0) * 1e6 * 1e-12,
            "peak_theoretical_how": "148 SMs x 64 FP64 FMA lanes x 2 flop x sm_max_mhz",
            "avg_launch_ms": avg_ms, "launches": cnt, "share_of_step": tot_ms / ms_prof, "flop_per_pair": FLOPS_PER_PAIR.get(name, 0), "pairs_per_launch": pairs, "pairs_within_R2": pairs_in, "flop_per_pair_beyond_R2": 30,
            "executed_fp64_per_pair": _RMX.get(name),
            "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": hbm_src, "copy_gbs_live": copy_gbs},
            "traffic": _ncu_traffic(name),
            "traffic_source": "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture of the default (second-generation) kernels, profiles/r2a_rjl_raw_summary.txt",
            "l1_gather": _RML if name in ("rjl_force", "rjl_density") else None,
        }
        if n_atoms < 200000:
            roofline["note"] = "system of %d atoms: every kernel is launch/latency bound (grid smaller than one wave), the FP64 fraction is not the figure of merit" % n_atoms
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            # same protocol as `--impl reference` (step 0 and the warm-up untimed, then K steps with the rebuilds that fall among them)
            n, steps, tcpu, cores = cpu_baseline_sample(max(1, min(K, 200)), warmup=min(W, 21))
            cpu = {"value": n * steps / tcpu, "unit": "atom-steps/s", "cores": cores, "kind": "port",
                   "sample": "Cu fcc 20^3x4 = %d atoms (bounded sample of the workload: the reference's O(N^2) rebuild), %d NVT steps timed after step 0 + %d warm-up steps, rebuild/20, "
                             "C++/OpenMP restatement of the reference built -O3 -march=native on this host, %.1f s" % (n, steps, min(W, 21), tcpu)}
        except Exception as ex:
            print("bench: cpu_baseline failed: %r" % (ex,), file=sys.stderr)
            cpu = {"value": None, "unit": "atom-steps/s", "cores": 0, "kind": "port", "sample": "failed: " + repr(ex)[:200]}
    # ---- kernel variants, same box, same run (A/B evidence for the defaults; never part of `value`) ----
    # Measured by a CHILD process after this one has released its context: an experiment that faults (sticky CUDA error,
    # abort) or hangs cannot take the headline line with it.
    variants = None
    if not args.no_variants and world == 1 and args.workload == "cu_fcc":
        def last_row(text):
            if isinstance(text, bytes):
                text = text.decode("utf-8", "replace")
            for l in reversed((text or "").splitlines()):
                if l.startswith("{") and l.rstrip().endswith("}"):
                    try:
                        return json.loads(l)
                    except Exception:
                        continue
            return None
        try:
            eng.close()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--variants-only", "--device", str(local)], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True, timeout=240)
            variants = last_row(r.stdout) or {"error": "child exit %d: %s" % (r.returncode, r.stderr[-300:])}
            if r.returncode != 0:
                variants["child_exit"] = r.returncode
        except subprocess.TimeoutExpired as ex:   # keep what the child had finished
            variants = last_row(ex.stdout) or {}
            variants["error"] = "child stopped after 240 s"
        except Exception as ex:  # the headline line must survive a failing experiment
            variants = {"error": repr(ex)[:300]}
    line = {
        "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "atoms_per_gpu": n_atoms, "parallelism": ("x-slab decomposition over %d GPUs, halo by direct NVLink stores into the neighbours' ghost slots (IPC peer memory), NCCL for migration and all-reduce" % world if slab else "ensemble x%d (independent replicas, reference MPI mode)" % world) if world > 1 else "single GPU",
                   "l2": "working set (lists %.0f MB + state) exceeds the 126 MB L2" % (pairs * 4 / 1e6), "ns_per_day": K / (ms_max * 1e-3) * dt * 86400e-6},
        "clocks": cs.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        "kernels_ms_per_step": {k: v[0] / K for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])},
        "per_rank": per_rank, "variants": variants, "check": check,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

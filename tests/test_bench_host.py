"""bench.py helpers that can run without a GPU: the A/B variant timer through the host replay of the library (kernel variants
are read from the environment at pfmds_create and the environment is restored), the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest

from pfmds_b200 import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def test_time_variant_on_the_host_replay(oracle_lib, monkeypatch):
    import build_emu as B
    import pfmds_b200.engine as E
    import bench
    B.build_emu()
    orig = E.configure
    monkeypatch.setattr(E, "configure", lambda case, device=0, **kw: orig(case, lib_path=B.LIB))
    case = inputs.cu_fcc(ncell=4, jitter=0.05, period=5)
    monkeypatch.delenv("PFMDS_RJL_GEN", raising=False)
    for gen in ("1", "2"):
        r = bench.time_variant(case, "nvt", 2.0, 0, {"PFMDS_RJL_GEN": gen}, 2, 3)
        assert r["steps"] == 3 and set(r["kernels_ms_per_step"]) >= {"rjl_force", "rjl_density"}
        assert "PFMDS_RJL_GEN" not in os.environ


def test_mask_list_build_on_the_host_replay(oracle_lib, monkeypatch):
    """k_build_mask (nl.cu, the default; PFMDS_NL_MASK=0 selects k_build) on the serial host replay, which takes the thread-per-atom build for every size: rows and
    forces identical to k_build, bit for bit, through partial rebuilds (lists of different periods, no re-sort) too."""
    import numpy as np
    import build_emu as B
    from pfmds_b200.engine import configure
    from util import list_ids, neighbours
    B.build_emu()
    gr = inputs.graphene_on_cu_small(interface="ljc", period=5)
    gr["interactions"][0]["lists"] = [(1, 1, 12, 2.6, 3)]
    gr["interactions"][2]["lists"] = [(2, 2, 100, 6.5, 7)]
    for case, integ, dt in ((inputs.cu_fcc(ncell=4, jitter=0.05, period=5), "nvt", 2.0), (gr, "nvt", 1.0),
                            (inputs.ab_gas(n_side=6, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=5), "nvt", 0.5)):
        monkeypatch.setenv("PFMDS_NL_MASK", "0")     # k_build
        a = configure(case, lib_path=B.LIB)
        monkeypatch.setenv("PFMDS_NL_MASK", "1")     # k_build_mask (the default)
        b = configure(case, lib_path=B.LIB)
        monkeypatch.delenv("PFMDS_NL_MASK")
        for e in (a, b):
            e.advance(integ, dt, 0, 8)
        for k, j in list_ids(case):
            x, y = neighbours(a, case, k, j), neighbours(b, case, k, j)
            assert all(np.array_equal(u, v) for u, v in zip(x, y)) and x[1].max() > 0
        assert all(np.array_equal(u, v) for u, v in zip(a.download(), b.download()))
        a.close()
        b.close()


def test_variants_child_on_the_host_replay(oracle_lib, monkeypatch, capsys):
    """bench.run_variants (the child process of the default bench run) in miniature: every experiment and every extra workload
    yields a timing, one cumulative JSON line per finished experiment."""
    import build_emu as B
    import pfmds_b200.engine as E
    import bench
    B.build_emu()
    orig = E.configure
    monkeypatch.setattr(E, "configure", lambda case, device=0, **kw: orig(case, lib_path=B.LIB))
    assert bench.run_variants(0, small=True) == 0
    rows = [json.loads(l) for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(rows) == 10 and all(len(b) == len(a) + 1 for a, b in zip(rows, rows[1:]))
    last = rows[-1]
    bad = {k: v for k, v in last.items() if k != "note" and "error" in v}
    assert not bad, bad
    assert sum(k.startswith("workload ") for k in last) == 4 and all(v["atom_steps_per_s"] > 0 for k, v in last.items() if k != "note")


def test_e2e_steps_logged_and_stepwise_agree(oracle_lib):
    """The e2e leg's two ways of reading every step's energies (one pfmds_advance_logged call / one round trip per step) leave the
    same state and count their D2H bytes."""
    import build_emu as B
    import bench
    from pfmds_b200.engine import configure
    import numpy as np
    B.build_emu()
    case = inputs.cu_fcc(ncell=5, jitter=0.05, period=5)
    a, b = configure(case, lib_path=B.LIB), configure(case, lib_path=B.LIB)
    for e in (a, b):
        e.advance("nvt", 2.0, 0, 1)
    ba, ha = bench.e2e_steps(a, "nvt", 2.0, 7)
    bb, hb = bench.e2e_steps(b, "nvt", 2.0, 7, stepwise=True)
    assert "advance_logged" in ha and "pfmds_energies" in hb and ba == 8 * (1 + 1 + 9) and bb > ba
    assert np.array_equal(a.download()[0], b.download()[0]) and np.array_equal(a.download()[2], b.download()[2])


def test_work_model_is_the_frozen_one():
    """bench.py takes its algorithmic flop / byte counts from pfmds_b200/csrc/roofline.json (frozen next to the kernels)."""
    import bench
    assert bench.FLOPS_PER_PAIR["rjl_force"] == 75 and bench.FLOPS_PER_PAIR["rjl_density"] == 55 and bench.FLOPS_PER_PAIR["lj1g"] == 71
    assert bench.BYTES_PER_ATOM["rjl_force"](86) == 4 * 86 + 140 and bench.BYTES_PER_ATOM["rjl_density"](86) == 4 * 86 + 76
    cx, cy, cz = bench.BIG_CELLS_PER_RANK
    assert 8 * 4 * cx * cy * cz == 101645216


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert out.returncode == 0, out.stderr[-400:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "atom-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_main_dry_run_on_the_host_replay(oracle_lib, monkeypatch, capsys):
    """bench.py's default arm end to end WITHOUT a GPU: main() runs unchanged on the host replay of the library (a small crystal
    in place of the 10^6-atom one, torch.cuda stubbed) and must print one complete JSON line.  No number of it means anything;
    the point is that no code path of the bench script is first executed on the GPU box."""
    import numpy as np
    import torch
    import build_emu as B
    import pfmds_b200.engine as E
    import bench
    B.build_emu()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"}))
    orig = E.configure
    monkeypatch.setattr(E, "configure", lambda case, device=0, **kw: orig(case, lib_path=B.LIB))
    monkeypatch.setattr(E, "measure_peaks", lambda device=0: (34.8, 6400.0))          # the device micro-benchmarks do not exist on the host
    small = inputs.cu_fcc(ncell=5, jitter=0.02, period=20, steps=40)
    monkeypatch.setattr(bench, "build_case", lambda workload, seed, steps, nx=1: (small, "nvt", "dry run: Cu fcc 5^3 cells"))
    monkeypatch.setattr(bench, "cpu_baseline_sample", lambda steps, threads=None: (500, steps, 1.0, 1))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "8", "--warmup", "3", "--no-variants"])
    assert bench.main() == 0
    line = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["steps"] == 8 and line["gpu_launches"] > 0 and line["config"]["workload"].startswith("dry run")
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and "pfmds_advance_logged" in line["e2e"]["what"]
    assert line["cpu_baseline"]["kind"] == "port"


@pytest.mark.parametrize("child", ["ok", "timeout", "crash"])
def test_bench_main_keeps_its_line_whatever_the_variants_child_does(oracle_lib, monkeypatch, capsys, child):
    """The parent side of the variants leg (never reached with --no-variants): the child's last complete cumulative line is kept when
    it finishes, when it is stopped at the time limit and when it dies; the headline line is printed in every case."""
    import torch
    import build_emu as B
    import pfmds_b200.engine as E
    import bench
    B.build_emu()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"}))
    orig = E.configure
    monkeypatch.setattr(E, "configure", lambda case, device=0, **kw: orig(case, lib_path=B.LIB))
    monkeypatch.setattr(E, "measure_peaks", lambda device=0: (34.8, 6400.0))
    small = inputs.cu_fcc(ncell=4, jitter=0.02, period=20, steps=40)
    monkeypatch.setattr(bench, "build_case", lambda workload, seed, steps, nx=1: (small, "nvt", "dry run: Cu fcc 4^3 cells"))
    monkeypatch.setattr(bench, "cpu_baseline_sample", lambda steps, threads=None: (500, steps, 1.0, 1))
    partial = json.dumps({"note": "n", "a": {"ms_per_step": 1.0}}) + "\n" + json.dumps({"note": "n", "a": {"ms_per_step": 1.0}, "b": {"ms_per_step": 2.0}}) + "\n"
    real_run = subprocess.run

    def fake_run(cmd, **kw):
        if "--variants-only" not in cmd:
            return real_run(cmd, **kw)
        if child == "timeout":
            raise subprocess.TimeoutExpired(cmd, kw.get("timeout"), output=partial + '{"note": "n", "a": {"ms_per', stderr="")
        if child == "crash":
            return subprocess.CompletedProcess(cmd, -11, stdout=partial + '{"trunc', stderr="Segmentation fault")
        return subprocess.CompletedProcess(cmd, 0, stdout="noise\n" + partial, stderr="")
    monkeypatch.setattr(bench.subprocess, "run", fake_run)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "6", "--warmup", "3"])
    assert bench.main() == 0
    line = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1])
    v = line["variants"]
    assert line["value"] > 0 and v["b"]["ms_per_step"] == 2.0 and v["a"]["ms_per_step"] == 1.0
    assert ("error" in v) == (child == "timeout") and v.get("child_exit") == (-11 if child == "crash" else None)


def test_bench_main_survives_failing_auxiliary_legs(oracle_lib, monkeypatch, capsys):
    """A failing e2e leg, peak micro-benchmark or CPU baseline is reported in its own field; the device-timed line is still printed."""
    import torch
    import build_emu as B
    import pfmds_b200.engine as E
    import bench
    B.build_emu()
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda data, **kw: real_tensor(data, **{k: v for k, v in kw.items() if k != "device"}))
    orig = E.configure
    monkeypatch.setattr(E, "configure", lambda case, device=0, **kw: orig(case, lib_path=B.LIB))

    def boom(*a, **kw):
        raise RuntimeError("injected")
    monkeypatch.setattr(E, "measure_peaks", boom)
    monkeypatch.setattr(bench, "e2e_steps", boom)
    monkeypatch.setattr(bench, "cpu_baseline_sample", boom)
    small = inputs.cu_fcc(ncell=4, jitter=0.02, period=20, steps=40)
    monkeypatch.setattr(bench, "build_case", lambda workload, seed, steps, nx=1: (small, "nvt", "dry run: Cu fcc 4^3 cells"))
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "6", "--warmup", "3", "--no-variants"])
    assert bench.main() == 0
    line = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1])
    assert line["value"] > 0 and line["gpu_launches"] > 0 and line["kernels_ms_per_step"]
    assert "injected" in line["e2e"]["error"] and line["e2e"]["value"] is None
    assert line["roofline"]["frac"] is None and line["roofline"]["peak"] is None and line["roofline"]["achieved"] >= 0
    assert line["cpu_baseline"]["value"] is None and "injected" in line["cpu_baseline"]["sample"]


def test_smoke_dry_run_on_the_lockstep_replay(oracle_lib, monkeypatch, capsys):
    """__graft_entry__.smoke() as the driver calls it, with the engine factory pointed at the lock-step host replay (the small-system
    kernels the GPU would launch): the graphene-on-Cu case, its oracle comparison and its assertions run unchanged."""
    import build_emu as B
    import pfmds_b200.engine as E
    import __graft_entry__ as G
    B.LOCKSTEP.build()
    orig = E.configure

    def configure(case, device=0, lib_path=None, prefix="pfmds_"):
        return orig(case, lib_path=lib_path or B.LOCKSTEP.lib, prefix=prefix)
    monkeypatch.setattr(E, "configure", configure)
    G.smoke()
    assert "smoke ok" in capsys.readouterr().out

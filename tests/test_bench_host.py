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
    case = inputs.cu_fcc(ncell=5, jitter=0.05, period=5)
    monkeypatch.delenv("PFMDS_RJL_GEN", raising=False)
    for gen in ("1", "2"):
        r = bench.time_variant(case, "nvt", 2.0, 0, {"PFMDS_RJL_GEN": gen}, 3, 5)
        assert r["steps"] == 5 and set(r["kernels_ms_per_step"]) >= {"rjl_force", "rjl_density"}
        assert "PFMDS_RJL_GEN" not in os.environ


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, OMP_NUM_THREADS="4"))
    assert out.returncode == 0, out.stderr[-400:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "atom-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0

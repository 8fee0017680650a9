"""SURVEY.md 8(f) row 4 — exact restart: `-checkpoint_period n` writes <prefix>checkpoint_NNNNNN.chk (full-precision state,
thermostat chains, group sizes, the driver's counters), `-restart file` continues the interrupted run.  CPU side: the shared
driver on the oracle engine, one thread (the reference's OpenMP partial sums are not reproducible with more)."""
import os
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from conftest import ORACLE_EXE


def log_rows(path):
    return [l for l in open(path).read().splitlines() if l[:6].strip() in ("nvt", "nve", "nvms")]


def run(exe, d, prefix, *extra):
    r = subprocess.run([exe, "-ipath", d, "-p", d + prefix, "-op", "10", "-omp_n", "1", *extra], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout


def restart_cases():
    ab = inputs.ab_gas(n_side=5, cap_aa=125, cap_ab=125, cap_ba=125, cap_bb=125, period=5, period_log=5, steps=(30, 20, 10))
    ab["integrators"] = [(n, dt, ln, 10 ** 9, 5) for (n, dt, ln, _, _) in ab["integrators"]]
    gr = inputs.graphene_on_cu_small(interface="morsec", period=5, steps=(30, 20))
    gr["integrators"] = [(n, dt, ln, 10 ** 9, 5) for (n, dt, ln, _, _) in gr["integrators"]]
    dep = inputs.lj_deposition(steps=50)
    return {"ab_gas": ab, "graphene": gr, "deposition": dep}


def check_restart(exe, tmp_path, case, at=20):
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    run(exe, d, "full_", "-checkpoint_period", "10")
    assert os.path.exists(d + "full_checkpoint_%06d.chk" % at)
    out = run(exe, d, "part_", "-restart", d + "full_checkpoint_%06d.chk" % at)
    assert "restarted from" in out
    # the continued run writes the same final state, digit for digit (7f27.16 rows)
    assert open(d + "full_final_init.xyz").read() == open(d + "part_final_init.xyz").read()
    full, part = log_rows(d + "full_md_run.log"), log_rows(d + "part_md_run.log")
    later = [l for l in full if int(l[6:15]) > at]
    assert part == later and len(later) >= 3


@pytest.mark.parametrize("name", ["ab_gas", "graphene", "deposition"])
def test_restart_reproduces_the_interrupted_run(tmp_path, oracle_lib, name):
    check_restart(ORACLE_EXE, tmp_path, restart_cases()[name], at=40 if name == "ab_gas" else 20)


def test_checkpoints_only_on_rebuild_steps(tmp_path, oracle_lib):
    case = inputs.cu_fcc(ncell=3, steps=30, period=4)
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    run(ORACLE_EXE, d, "x_", "-checkpoint_period", "10")
    assert sorted(f for f in os.listdir(d) if f.endswith(".chk")) == ["x_checkpoint_000020.chk"]     # 10 and 30 are not multiples of 4


def test_restart_refuses_a_foreign_checkpoint(tmp_path, oracle_lib):
    a, b = str(tmp_path / "a") + os.sep, str(tmp_path / "b") + os.sep
    inputs.write_case(a, inputs.cu_fcc(ncell=3, steps=10, period=5))
    inputs.write_case(b, inputs.cu_fcc(ncell=4, steps=10, period=5))
    run(ORACLE_EXE, a, "x_", "-checkpoint_period", "10")
    r = subprocess.run([ORACLE_EXE, "-ipath", b, "-p", b + "y_", "-restart", a + "x_checkpoint_000010.chk"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode != 0 and "does not belong" in r.stdout

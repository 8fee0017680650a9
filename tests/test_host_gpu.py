"""The drop-in host on the GPU: pfmds_b200/host/run_md_simulation against the oracle binary on the same input
files — same log rows (to the printed digits where the physics allows), same final xyz, same stdout blocks."""
import os
import re
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.build import EXE
from pfmds_b200.host_io import read_xyz
from conftest import ORACLE_EXE

pytestmark = pytest.mark.gpu


def _rows(path):
    out = []
    for l in open(path).read().splitlines():
        if l[:6].strip() in ("nvt", "nve", "nvms"):
            out.append((l[:6].strip(), int(l[6:15]), [float(x) for x in l[15:].split()]))
    return out


@pytest.mark.parametrize("which", ["ab_gas", "graphene"])
def test_same_outputs_as_the_cpu_reference_port(tmp_path, cuda_lib, oracle_lib, which):
    if which == "ab_gas":
        case = inputs.ab_gas(n_side=6, cap_aa=216, cap_ab=216, cap_ba=216, cap_bb=216, period=5, period_log=10, steps=(40, 40, 60))
        case["integrators"] = [(n, dt, ln, 40, 10) for (n, dt, ln, _, _) in case["integrators"]]
    else:
        case = inputs.graphene_on_cu_small(interface="morsec", period=5, steps=(40, 40))
        case["integrators"] = [(n, dt, ln, 40, 10) for (n, dt, ln, _, _) in case["integrators"]]
    outs = {}
    for tag, exe in (("gpu", EXE), ("cpu", ORACLE_EXE)):
        d = str(tmp_path / tag) + os.sep
        inputs.write_case(d, case)
        r = subprocess.run([exe, "-ipath", d, "-p", d + "x_", "-op", "20", "-omp_n", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        outs[tag] = (d, r.stdout)
    (dg, og), (dc, oc) = outs["gpu"], outs["cpu"]
    rg, rc = _rows(dg + "x_md_run.log"), _rows(dc + "x_md_run.log")
    assert [(a, b) for a, b, _ in rg] == [(a, b) for a, b, _ in rc]          # same phases, same logged steps, same exit step
    for (_, step, a), (_, _, b) in zip(rg, rc):
        assert np.allclose(a, b, rtol=1e-7, atol=2e-6), (step, a, b)          # f24.6 / f20.9 columns
    fg, fc = read_xyz(dg + "x_final_init.xyz"), read_xyz(dc + "x_final_init.xyz")
    assert np.abs(fg["pos"] - fc["pos"]).max() < 1e-7 and fg["names"] == fc["names"]
    assert os.path.exists(dg + "x_snapshot_000040.xyz")
    strip = lambda s: [l for l in s.splitlines() if l.startswith(("step =", "neib", " steps number")) or re.match(r"^ +(lj|lj1g|ljc|morsec|tb|rjl) +\d+ +\d+ /", l)]
    assert [l for l in strip(og) if "exe time" not in l][:40] == [l for l in strip(oc) if "exe time" not in l][:40]
    pd = lambda s: [float(x) for x in re.findall(r"potential energy difference:\s+(\S+)", s)]
    assert np.allclose(pd(og), pd(oc), rtol=1e-6, atol=1e-12)                # 17 printed digits: equal to rounding noise only
    assert re.search(r"steps number:\s+\d+", og).group(0) == re.search(r"steps number:\s+\d+", oc).group(0)


def test_gpu_ensemble_ranks(tmp_path, cuda_lib):
    """run_md_simulation_mpi semantics on the GPU host: two ranks share the list, one context each."""
    case = inputs.cu_fcc(ncell=5, steps=10, period=5)
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    open(d + "list.txt", "w").write("3\nmd_run_settings.txt a_\nmd_run_settings.txt b_\nmd_run_settings.txt c_\n")
    procs = [subprocess.Popen([EXE, "-node", str(r + 1), "-nodes", "2", "-gpu", "0", "-ipath", d, "-ilist", "list.txt", "-opath", d, "-op", "100"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=d) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    a, c = read_xyz(d + "0001-a_final_init.xyz"), read_xyz(d + "0001-c_final_init.xyz")
    b = read_xyz(d + "0002-b_final_init.xyz")
    assert np.array_equal(a["pos"], c["pos"]) and np.array_equal(a["pos"], b["pos"])  # same input, deterministic kernels


def test_queued_log_rows_are_the_stepwise_log(tmp_path, cuda_lib):
    """period_log = 1 under a sparse stdout period: the host queues the steps between hard events (pfmds_advance_logged, one
    copy of the log rows per queue) — every output file is identical to the run that reads the energies step by step."""
    case = inputs.cu_fcc(ncell=5, steps=45, period=5, jitter=0.05)
    case["integrators"] = [(n, dt, ln, 20, 1) for (n, dt, ln, _, _) in case["integrators"]]
    case["roles"]["period_traj"] = 15
    outs = []
    for tag, env in (("queued", {}), ("stepwise", {"PFMDS_HOST_STEPWISE_LOG": "1"})):
        d = str(tmp_path / tag) + os.sep
        inputs.write_case(d, case)
        r = subprocess.run([EXE, "-ipath", d, "-p", d + "x_", "-op", "25"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout[-2000:]
        outs.append(d)
    a, b = outs
    log = open(a + "x_md_run.log").read()
    assert len(_rows(a + "x_md_run.log")) == 46 and log == open(b + "x_md_run.log").read()
    for f in ("x_final_init.xyz", "x_snapshot_000020.xyz", "x_snapshot_000040.xyz"):   # (the trajectory group of this case is empty)
        assert open(a + f).read() == open(b + f).read(), f

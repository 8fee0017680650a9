"""The run_md_simulation host logic (settings grammar, phase control, log / xyz formats), exercised on CPU with the
oracle binary, which shares pfmds_b200/host/md_driver.hpp with the GPU host."""
import os
import re
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.ensemble import node_prefix, shard
from conftest import ORACLE_EXE


@pytest.fixture(scope="module", autouse=True)
def _lib(oracle_lib):
    return None


def _tiny_case(steps=(20, 20, 30)):
    c = inputs.ab_gas(n_side=5, cap_aa=125, cap_ab=125, cap_ba=125, cap_bb=125, period=5, period_log=10, steps=steps)
    c["integrators"] = [(n, dt, ln, 20, 10) for (n, dt, ln, _, _) in c["integrators"]]
    c["roles"]["period_traj"] = 25
    return c


def _run(tmp, case, extra=(), **kw):
    d = str(tmp) + os.sep
    inputs.write_case(d, case, **kw)
    r = subprocess.run([ORACLE_EXE, "-ipath", d, "-p", d + "t_", "-op", "20", "-omp_n", "2"] + list(extra), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return d, r


def test_full_run_outputs(tmp_path):
    case = _tiny_case()
    d, r = _run(tmp_path, case)
    out = r.stdout
    assert r.returncode == 0, out
    # settings echo, md_simulation.f90:48-93
    assert "settings_filename: md_run_settings.txt" in out
    assert re.search(r"^box_size:\s+20\.000000\s+20\.000000\s+20\.000000$", out, re.M)
    assert re.search(r"^particles_num:\s+125$", out, re.M)
    assert "RUNNING ON " in out and "OPENMP THREADS" in out
    assert "PERFOMANCE:" in out and "TIME STEPS PER HOUR:" in out
    assert "neib lists load:" in out and re.search(r"^      lj  1\s+\d+ /\s+125$", out, re.M)
    # log file: '(A6,i9,7f24.6)' + 3 x f20.9 + 1 x f20.9 per logged step, then the final '(A32,i9,5f20.9)' line
    log = open(d + "t_md_run.log").read().splitlines()
    rows = [l for l in log if l[:6].strip() in ("nvt", "nve", "nvms")]
    assert len(rows[0]) == 6 + 9 + 7 * 24 + 3 * 20 + 20
    steps_logged = [int(l[6:15]) for l in rows]
    assert steps_logged[:3] == [0, 10, 20] and [l[:6].strip() for l in rows[:3]] == ["nvt", "nvt", "nvt"]
    assert any(l[:6].strip() == "nve" for l in rows) and any(l[:6].strip() == "nvms" for l in rows)
    # phase switch at md_step-1 == cumulative length (:121): step 21 is the first nve step, logged at 30
    assert [l[:6].strip() for l in rows if int(l[6:15]) == 30] == ["nve"]
    # snapshot / trajectory / final files, md_read_write.f90:65-107
    assert os.path.exists(d + "t_snapshot_000020.xyz") and os.path.exists(d + "t_final_init.xyz") and os.path.exists(d + "t_traj_03.xyz")
    fin = open(d + "t_final_init.xyz").read().splitlines()
    assert int(fin[0]) == 125 and fin[1].startswith('Lattice="') and "Properties=pos:R:3:vel:R:3:mass:R:1:species:S:1" in fin[1]
    assert len(fin[2]) == 7 * 27 + 4 + 32
    traj = open(d + "t_traj_03.xyz").read().splitlines()
    assert traj[1].startswith("time_step:         0    Lattice=")
    # NVE phase conserves the logged conserved energy
    nve = np.array([[float(l[15 + 24 * k:15 + 24 * (k + 1)]) for k in range(7)] for l in rows if l[:6].strip() == "nve"])
    assert np.ptp(nve[:, 1]) < 1e-4 * abs(nve[0, 5]) + 1e-5


def test_final_xyz_restarts_exactly(tmp_path):
    """final_*.xyz carries full FP64 state (7f27.16): a restart from it continues the NVE trajectory."""
    case = _tiny_case(steps=(0, 40, 0))
    case["integrators"] = [("nve", 0.5, 40, 1000, 10)]
    case["nhc"] = []
    d, r = _run(tmp_path / "a", case)
    assert r.returncode == 0, r.stdout
    half = dict(case, integrators=[("nve", 0.5, 20, 1000, 10)])
    d1, r1 = _run(tmp_path / "b", half)
    from pfmds_b200.host_io import read_xyz  # noqa
    st = read_xyz(d1 + "t_final_init.xyz")
    cont = dict(half, pos=st["pos"], vel=st["vel"])
    d2, r2 = _run(tmp_path / "c", cont)
    a, b = read_xyz(d + "t_final_init.xyz"), read_xyz(d2 + "t_final_init.xyz")
    assert np.abs(a["pos"] - b["pos"]).max() < 1e-9


def test_nvms_exit_rule(tmp_path):
    """an nvms phase entered right after a non-logged step exits at once (PE == PE_prev)."""
    case = _tiny_case(steps=(0, 13, 50))
    case["integrators"] = [("nve", 0.5, 13, 1000, 10), ("nvms", 0.5, 50, 1000, 10)]
    d, r = _run(tmp_path, case)
    assert "potential energy diffrence is small enough" in r.stdout
    assert re.search(r"steps number:\s+13$", r.stdout, re.M)


def test_no_more_integrators(tmp_path):
    case = _tiny_case(steps=(5, 0, 0))
    case["integrators"] = [("nve", 0.5, 5, 1000, 10)]
    d, r = _run(tmp_path, case, md_step_limit=50)
    assert "no more integrators" in r.stdout and re.search(r"steps number:\s+5$", r.stdout, re.M)


def test_reference_error_messages(tmp_path):
    case = _tiny_case()
    case["interactions"][1]["lists"] = [(2, 2, 3, 7.5, 5)]
    d, r = _run(tmp_path, case)
    assert r.returncode != 0 and "too many neighbours" in r.stdout


def test_list_mode_and_ensemble_sharding(tmp_path):
    case = _tiny_case(steps=(5, 0, 0))
    case["integrators"] = [("nve", 0.5, 5, 1000, 5)]
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    open(d + "list.txt", "w").write("3\nmd_run_settings.txt a_\nmd_run_settings.txt b_\nmd_run_settings.txt c_\n")
    r = subprocess.run([ORACLE_EXE, "-ipath", d, "-ilist", "list.txt", "-opath", d, "-op", "100"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    allout = open(d + "all_out.txt").read().splitlines()
    assert len([l for l in allout if l.strip()]) == 3 and os.path.exists(d + "c_final_init.xyz")
    # run_md_simulation_mpi.f90:74-85
    assert shard(3, 2, 0) == [1, 3] and shard(3, 2, 1) == [2] and node_prefix(1) == "0002-"
    with pytest.raises(ValueError):
        shard(1, 2, 0)
    for rank in (0, 1):
        rr = subprocess.run([ORACLE_EXE, "-node", str(rank + 1), "-nodes", "2", "-ipath", d, "-ilist", "list.txt", "-opath", d, "-op", "100"],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=d)   # NNNN-out.txt goes to the working directory
        assert rr.returncode == 0
    assert os.path.exists(d + "0001-a_final_init.xyz") and os.path.exists(d + "0002-b_final_init.xyz") and os.path.exists(d + "0001-c_final_init.xyz")
    assert not os.path.exists(d + "0002-a_final_init.xyz")
    assert "RUNNING ON NODE      1 OUT OF     2 NODES" in open(d + "0001-out.txt").read()


def test_new_velocities_are_seeded_and_order_independent(tmp_path):
    """new_velocities T (set_new_temperature, md_general.f90:114-159,328-340) with the host's counter-based generator:
    exact target temperature, no centre-of-mass motion, Maxwell widths, and the same velocities for the same seed."""
    from pfmds_b200.host_io import read_xyz
    case = inputs.cu_fcc(ncell=6, steps=0, period=5)
    case["vel"][:] = 0.0
    case["integrators"] = [("nve", 2.0, 1, 1, 1)]
    case["initial_temperature"] = 250.0
    outs = []
    for k in range(2):
        d, r = _run(tmp_path / ("r%d" % k), case, new_velocities=True, md_step_limit=1)
        assert r.returncode == 0, r.stdout[-2000:]
        outs.append((d, r.stdout))
    a, b = read_xyz(outs[0][0] + "t_snapshot_000001.xyz"), read_xyz(outs[1][0] + "t_snapshot_000001.xyz")
    assert np.array_equal(a["vel"], b["vel"])                                      # same rand_seed (1): same draw for every atom
    log = [l for l in open(outs[0][0] + "t_md_run.log").read().splitlines() if l[:6].strip() == "nve"]
    t0 = float(log[0].split()[8])
    assert abs(t0 - 250.0) < 1e-5                                                   # rescaled to the requested temperature (f24.6 column)
    v = a["vel"]
    m = case["mass"]
    assert np.abs((m[:, None] * v).sum(0)).max() < 1e-9 * m.sum()                   # momentum removed (one short step later)
    sig = np.sqrt(inputs.COEF * 250.0 / m[0])
    assert abs(v.std() / sig - 1.0) < 0.05 and abs(np.mean(np.abs(v) < sig) - 0.6827) < 0.03   # Maxwell: 68 % inside one sigma
    # ensemble ranks seed with their rank number (run_md_simulation_mpi.f90:62): different velocities
    d2 = str(tmp_path / "e") + os.sep
    inputs.write_case(d2, case, new_velocities=True, md_step_limit=1)
    rr = subprocess.run([ORACLE_EXE, "-node", "2", "-nodes", "2", "-ipath", d2, "-opath", d2, "-op", "100"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=d2)
    assert rr.returncode == 0
    c = read_xyz(d2 + "0002-snapshot_000001.xyz")
    assert not np.allclose(c["vel"], a["vel"])


def test_queued_log_rows_equal_the_stepwise_log(tmp_path):
    """period_log = 1 with a sparse stdout period: the driver queues the steps between hard events and reads their log rows in
    one go (advance_logged); PFMDS_HOST_STEPWISE_LOG=1 is the reference's loop, one energy read per step.  Same files."""
    case = _tiny_case()
    case["integrators"] = [(n, dt, ln, 20, 1 if n != "nve" else 3) for (n, dt, ln, _, _) in case["integrators"]]
    d1, r1 = _run(tmp_path / "queued", case)
    assert r1.returncode == 0, r1.stdout
    os.environ["PFMDS_HOST_STEPWISE_LOG"] = "1"
    try:
        d2, r2 = _run(tmp_path / "stepwise", case)
    finally:
        del os.environ["PFMDS_HOST_STEPWISE_LOG"]
    assert r2.returncode == 0, r2.stdout
    a, b = open(d1 + "t_md_run.log").read(), open(d2 + "t_md_run.log").read()
    assert len(a.splitlines()) > 45 and a == b
    from pfmds_b200.host_io import read_xyz
    for f in ("t_final_init.xyz", "t_snapshot_000020.xyz"):   # the oracle's OpenMP sums differ in the last bits from run to run
        x, y = read_xyz(d1 + f), read_xyz(d2 + f)
        assert np.allclose(x["pos"], y["pos"], rtol=0, atol=1e-10) and np.allclose(x["vel"], y["vel"], rtol=0, atol=1e-12)
    assert len(open(d1 + "t_traj_03.xyz").read().splitlines()) == len(open(d2 + "t_traj_03.xyz").read().splitlines())

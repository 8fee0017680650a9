"""SURVEY.md 8(f) row 1 on the device: deposition (`change_group_num`) through the C ABI against the CPU oracle — group sizes,
neighbour sets (bit-exact), forces, energies and trajectories with a group that grows while the run is resident on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.build import EXE
from pfmds_b200.host_io import read_xyz
from conftest import ORACLE_EXE
from util import RTOL, gpu, oracle, rel_err, neighbours

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("thermostat", [True, False])
def test_deposition_matches_the_oracle(thermostat):
    case = inputs.lj_deposition(thermostat=thermostat)
    g, o = gpu(case), oracle(case)
    kind = "nvt" if thermostat else "nve"
    s = 0
    for n in (1, 3, 1, 1, 5, 11, 7, 9, 6):      # release steps fall first, in the middle and last in a call; CUDA-graph replays in between
        g.advance(kind, 1.0, s, n)
        o.advance(kind, 1.0, s, n)
        s += n
        assert g.group_size(3) == o.group_size(3) and g.group_size(2) == o.group_size(2) == 108
        pg, vg, fg = g.download()
        po, vo, fo = o.download()
        assert np.abs(pg - po).max() < 1e-9 and np.abs(vg - vo).max() < 1e-11
        assert rel_err(fg, fo) < RTOL
        a, b = neighbours(g, case, 0, 0), neighbours(o, case, 0, 0)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)          # nlist, nnum, lessnnum: rows beyond group%N stay empty, partners come from the first N
        eg, eo = g.energies(), o.energies()
        assert rel_err(eg[0], eo[0]) < RTOL and abs(eg[1] - eo[1]) < RTOL * abs(eo[1]) and abs(eg[2] - eo[2]) < RTOL * abs(eo[2])
        if thermostat:
            assert np.allclose(eg[3], eo[3], rtol=1e-8, atol=1e-12)
        dg, do = g.diagnostics(), o.diagnostics()
        assert np.allclose(dg[1], do[1], rtol=1e-10) and np.allclose(dg[2], do[2], rtol=1e-9, atol=1e-16)   # c.o.m. over the growing all_atoms group
    assert g.group_size(3) == 116            # every parked atom has been released by step 43


def test_forces_outside_all_atoms_accumulate_like_the_reference():
    """zero_forces only touches the all_atoms group (md_integrators.f90:147-163): an atom with a list row that is not in
    all_atoms keeps summing its forces, step after step — in the reference and here."""
    case = inputs.cu_fcc(ncell=4, jitter=0.05, period=5)
    n = len(case["mass"])
    case["names"] = ["CU"] * (n - 7) + ["CUX"] * 7
    case["groups"] = [["CU", "CUX"], ["CU", "#"], ["#", "#"]]          # 1 everything (lists), 2 all_atoms / moving without the last 7
    case["roles"] = dict(all_moving=2, xyz_moving=2, z_moving=3, all_atoms=2, traj_group=3, period_traj=10 ** 9)
    case["nhc"] = [(2, 300.0, 3, case["nhc"][0][3])]
    g, o = gpu(case), oracle(case)
    for e in (g, o):
        e.advance("nvt", 2.0, 0, 6)
    fg, fo = g.download()[2], o.download()[2]
    assert rel_err(fg, fo) < RTOL
    assert np.abs(fo[-7:]).max() > 2 * np.abs(fo[:-7]).max()   # six steps of force piled up on the outsiders


def test_host_deposition_run_matches_the_cpu_port(tmp_path, cuda_lib, oracle_lib):
    case = inputs.lj_deposition(steps=30)
    outs = {}
    for tag, exe in (("gpu", EXE), ("cpu", ORACLE_EXE)):
        d = str(tmp_path / tag) + os.sep
        inputs.write_case(d, case)
        r = subprocess.run([exe, "-ipath", d, "-p", d + "x_", "-op", "10", "-omp_n", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
        outs[tag] = d
    for name in ("x_snapshot_000020.xyz", "x_final_init.xyz"):
        a, b = read_xyz(outs["gpu"] + name), read_xyz(outs["cpu"] + name)
        assert len(a["mass"]) == len(b["mass"]) and a["names"] == b["names"]
        assert np.abs(a["pos"] - b["pos"]).max() < 1e-8
    assert open(outs["gpu"] + "x_traj_03.xyz").read().count("time_step:") == open(outs["cpu"] + "x_traj_03.xyz").read().count("time_step:") == 2


def test_deposition_matches_the_golden_fixture():
    """Against the committed fixture (tests/golden/lj_deposition.npz, frozen oracle outputs)."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lj_deposition.npz"))
    case = inputs.lj_deposition()
    e = gpu(case)
    e.advance("nvt", 1.0, 0, 1)
    assert rel_err(e.download()[2], g["frc0"]) < RTOL and np.allclose(e.energies()[0], g["e0"], rtol=RTOL, atol=0)
    nl = neighbours(e, case, 0, 0)
    assert np.array_equal(nl[1], g["nnum_0_0"]) and np.array_equal(nl[0], g["nlist_0_0"])
    e.advance("nvt", 1.0, 1, 10)
    p, v, f = e.download()
    assert np.abs(p - g["pos10"]).max() < 1e-10 and rel_err(f, g["frc10"]) < 1e-8
    assert [e.group_size(k + 1) for k in range(4)] == g["group_n10"].tolist()
    assert np.array_equal(neighbours(e, case, 0, 0)[1], g["nnum10_0_0"])

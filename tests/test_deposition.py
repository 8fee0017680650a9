"""SURVEY.md 8(f) row 1 — deposition through `change_group_num` (change_particle_group_N, md_general.f90:82-94; applied at the top
of every md step, md_simulation.f90:116-119).  CPU side: the oracle's restatement against the closed form of the recurrence,
what a growing group means for the integrator / lists / writers, and the drop-in host on the oracle engine."""
import os
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.host_io import read_xyz
from conftest import ORACLE_EXE
from util import oracle, neighbours


def expected_n(step, n_from, cap, ts1, ts2, frec):
    """Closed form of the recurrence for a constant source group."""
    if step < ts1:
        return min(n_from, cap)
    last = min(step, ts2 - 1)
    extra = (last - ts1) // frec if last > ts1 else 0
    return min(cap, n_from + 1 + extra)


@pytest.mark.parametrize("ts1,ts2,frec", [(3, 40, 4), (0, 10, 1), (5, 6, 3), (2, 1000, 7)])
def test_group_size_follows_the_reference_recurrence(oracle_lib, ts1, ts2, frec):
    case = inputs.lj_deposition(ts1=ts1, ts2=ts2, frec=frec, thermostat=False)
    e = oracle(case)
    ns, cap = 108, 116
    for s in range(0, 45):
        e.advance("nve", 1.0, s, 1)
        assert e.group_size(3) == expected_n(s, ns, cap, ts1, ts2, frec), s
        assert e.group_size(2) == ns and e.group_size(1) == cap


def test_parked_atoms_wait_for_their_turn(oracle_lib):
    case = inputs.lj_deposition(thermostat=False)
    e = oracle(case)
    p0 = case["pos"].copy()
    ns = 108
    for s in range(0, 24):
        e.advance("nve", 1.0, s, 1)
        n = e.group_size(3)
        p, v, f = e.download()
        assert np.array_equal(p[n:], p0[n:])                       # not in the moving group yet: never integrated
        if n > ns and s > 4:
            assert (p[ns:n, 2] < p0[ns:n, 2]).all()                # released atoms fly down
        nl, nn, less = neighbours(e, case, 0, 0)
        assert (nn[n:] == 0).all()                                 # rows beyond group%N are never built
        assert nl.max() <= n                                       # partners come from the first N entries of the group only
    # an atom released between two rebuilds has no row (and is nobody's partner) until the next rebuild
    e2 = oracle(case)
    for s in range(0, 4):
        e2.advance("nve", 1.0, s, 1)
    assert e2.group_size(3) == ns + 1
    _, nn, _ = neighbours(e2, case, 0, 0)
    assert nn[ns] == 0
    e2.advance("nve", 1.0, 4, 2)                                   # step 5 rebuilds
    _, nn, _ = neighbours(e2, case, 0, 0)
    assert nn[ns] > 0


def test_temperature_uses_the_current_group_size(oracle_lib):
    case = inputs.lj_deposition()
    e = oracle(case)
    e.advance("nvt", 1.0, 0, 12)
    n = e.group_size(3)
    _, ke, temp, _ = e.energies()
    assert n == 111 and np.isclose(temp, 2 * ke / inputs.KB / (3 * n), rtol=1e-14)


def test_host_writers_follow_group_size(tmp_path, oracle_lib):
    """snapshot / final / trajectory frames hold group%N atoms at the time of writing (md_read_write.f90:65-107)."""
    case = inputs.lj_deposition(steps=30)
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    r = subprocess.run([ORACLE_EXE, "-ipath", d, "-p", d + "x_", "-op", "10", "-omp_n", "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "group_change_from_to:" in r.stdout and "change_ts1_ts2_freq:" in r.stdout
    snap = read_xyz(d + "x_snapshot_000020.xyz")
    assert len(snap["mass"]) == expected_n(20, 108, 116, 3, 40, 4) == 113
    fin = read_xyz(d + "x_final_init.xyz")
    assert len(fin["mass"]) == expected_n(30, 108, 116, 3, 40, 4)
    frames = [int(l) for l in open(d + "x_traj_03.xyz").read().splitlines() if l.strip().isdigit()]
    assert frames == [expected_n(s, 108, 116, 3, 40, 4) for s in (0, 20)]

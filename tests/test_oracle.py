"""CPU tests of the oracle (C++ restatement of the reference; parity unpinned — see oracle/oracle.hpp):
physics invariants the reference itself prints warnings about (md_simulation.f90:212-227), and the frozen
golden fixtures."""
import os

import numpy as np
import pytest

from pfmds_b200 import inputs
from util import list_ids, neighbours, oracle, rel_err, small_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _lib(oracle_lib):
    return None


def _only(case, names):
    c = dict(case)
    c["interactions"] = [i for i in case["interactions"] if i["name"] in names]
    return c


def _fd_error(case, atoms, h=1e-5):
    e = oracle(case)
    e.advance("nve", 1.0, 0, 1)
    F = e.download()[2]
    worst = 0.0
    for a in atoms:
        for k in range(3):
            E = []
            for s in (+1, -1):
                c = dict(case)
                c["pos"] = case["pos"].copy()
                c["pos"][a, k] += s * h
                ee = oracle(c)
                ee.advance("nve", 1.0, 0, 1)
                E.append(ee.energies()[0].sum())
            worst = max(worst, abs(-(E[0] - E[1]) / (2 * h) - F[a, k]) / np.abs(F).max())
    return worst, F


@pytest.mark.parametrize("names,iface", [(("tb",), "ljc"), (("ljc",), "ljc"), (("morsec",), "morsec"), (("rjl",), "ljc")])
def test_forces_are_minus_grad_energy(names, iface):
    """Each analytic force equals -grad E of the reference's own energy expression (finite differences).
    ljc/morsec alone exercise the own-search path of the nearest-3 list (md_interactions.f90:168-171)."""
    case = _only(inputs.graphene_on_cu_small(interface=iface, jitter=0.05), names)
    err, F = _fd_error(case, [0, 10, 95, 96, 300])
    assert err < 5e-8
    assert np.abs(F.sum(0)).max() < 1e-12  # Newton's third law, md_simulation.f90:224


def test_lj_forces_are_minus_grad_energy():
    case = _only(small_cases()["ab_gas"], ("lj",))
    err, F = _fd_error(case, [0, 100, 500])
    assert err < 5e-8 and np.abs(F.sum(0)).max() < 1e-13


def test_lj1g_force_matches_energy_only_when_switch_width_is_one():
    """f_dfr_cut lacks 1/(R2-R1) (cut_off_poly.f90:41); with R2-R1 = 1 forces are consistent."""
    case = _only(small_cases()["ab_gas"], ("lj1g",))
    err, _ = _fd_error(case, [0, 100, 500])
    assert err < 5e-8
    bad = _only(small_cases()["ab_gas"], ("lj1g",))
    for it in bad["interactions"]:
        it["params"] = [it["params"][0], it["params"][1], 5.0, 7.0]
    err2, _ = _fd_error(bad, [0, 100, 500])
    assert err2 > 1e-4  # the reference's bug, reproduced


def test_converse_list_is_the_transpose():
    case = small_cases()["ab_gas"]
    o = oracle(case)
    o.advance("nve", 0.5, 0, 1)
    a, na, _ = neighbours(o, case, 0, 0)
    b, nb, _ = neighbours(o, case, 0, 1)
    pairs_a = {(i, a[i, p] - 1) for i in range(len(na)) for p in range(na[i])}
    pairs_b = {(b[j, p] - 1, j) for j in range(len(nb)) for p in range(nb[j])}
    assert pairs_a == pairs_b and len(pairs_a) > 0


def _conserved_deviation(name, integrator, dt, steps):
    case = small_cases()[name]
    o = oracle(case)
    o.advance(integrator, dt, 0, 1)
    e0 = o.energies()
    c0 = e0[0].sum() + e0[1] + e0[3].sum()
    dev = []
    for k in range(5):
        o.advance(integrator, dt, 1 + k * (steps // 5), steps // 5)
        e = o.energies()
        dev.append(abs(e[0].sum() + e[1] + e[3].sum() - c0))
    return max(dev), e0[1]


@pytest.mark.parametrize("name,integrator", [("ab_gas", "nve"), ("cu_fcc", "nve"), ("cu_fcc", "nvt"), ("gr_cu_ljc", "nve")])
def test_conserved_energy(name, integrator):
    """NVE total energy / NVT extended energy (PE+KE+NHC, md_simulation.f90:199-201): the deviation is bounded and is
    pure velocity-Verlet discretisation error — it falls by ~4 when dt is halved (forces are consistent with energies)."""
    dt = small_cases()[name]["integrators"][0][1]
    d1, ke = _conserved_deviation(name, integrator, dt, 100)
    d2, _ = _conserved_deviation(name, integrator, dt / 2, 200)
    assert d1 < 5e-3 * ke
    assert d2 < 0.4 * d1 + 1e-9 * ke


def test_nvms_quench_lowers_potential_energy():
    case = small_cases()["cu_fcc"]
    o = oracle(case)
    o.advance("nvms", 2.0, 0, 1)
    e0 = o.energies()[0].sum()
    o.advance("nvms", 2.0, 1, 60)
    e1 = o.energies()
    assert e1[0].sum() < e0 and e1[1] < 0.5 * oracle(case).energies()[1] + 1e9


@pytest.mark.parametrize("name", list(small_cases()))
def test_oracle_matches_golden(name):
    """The restatement reproduces its frozen outputs (tests/golden/make_golden.py)."""
    case = small_cases()[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert np.array_equal(g["pos0"], case["pos"])  # the seeded generators are stable
    o = oracle(case)
    integ, dt = case["integrators"][0][0], case["integrators"][0][1]
    o.advance(integ, dt, 0, 1)
    assert rel_err(o.download()[2], g["frc0"]) < 1e-12
    assert np.allclose(o.energies()[0], g["e0"], rtol=1e-12)
    for k, j in list_ids(case):
        nl = neighbours(o, case, k, j)
        assert np.array_equal(nl[1], g["nnum_%d_%d" % (k, j)]) and np.array_equal(nl[0], g["nlist_%d_%d" % (k, j)])
    o.advance(integ, dt, 1, 10)
    p, v, f = o.download()
    assert np.abs(p - g["pos10"]).max() < 1e-11 and rel_err(f, g["frc10"]) < 1e-9


@pytest.mark.parametrize("name", ["lj_deposition", "graphene_rebosc"])
def test_oracle_matches_golden_next_rows(name):
    """Fixtures of the SURVEY.md 8(f) rows.  rebosc forces are central differences: they repeat to the rounding noise of
    E(-dx) - E(+dx) only (thread partial sums), hence the looser force bound."""
    from util import next_row_cases
    case = next_row_cases()[name]
    g = np.load(os.path.join(GOLD, name + ".npz"))
    assert np.array_equal(g["pos0"], case["pos"])
    o = oracle(case)
    integ, dt = case["integrators"][0][0], case["integrators"][0][1]
    ftol = 2e-7 if name == "graphene_rebosc" else 1e-12
    o.advance(integ, dt, 0, 1)
    assert rel_err(o.download()[2], g["frc0"]) < ftol
    assert np.allclose(o.energies()[0], g["e0"], rtol=1e-12)
    nl = neighbours(o, case, 0, 0)
    assert np.array_equal(nl[1], g["nnum_0_0"]) and np.array_equal(nl[0], g["nlist_0_0"])
    o.advance(integ, dt, 1, 10)
    p, v, f = o.download()
    assert np.abs(p - g["pos10"]).max() < 1e-9 and rel_err(f, g["frc10"]) < max(ftol, 1e-9)
    if "group_n10" in g:
        assert [o.group_size(k + 1) for k in range(len(case["groups"]))] == g["group_n10"].tolist()
        assert np.array_equal(neighbours(o, case, 0, 0)[1], g["nnum10_0_0"])


def test_half_list_rule_drops_pairs_for_non_monotone_groups():
    """lessnnum counts the entries before the first one whose GLOBAL index exceeds the owner's
    (md_neighbours.f90:78).  For a group whose indexes are not ascending (type columns 'B A' with file order A.., B..)
    lj1g then silently loses pairs.  The oracle reproduces that literally; the CUDA path refuses such input."""
    base = small_cases()["ab_gas"]
    mono = dict(base, groups=[["A", "B"], ["A", "#"], ["B", "#"], ["#", "#"]],
                interactions=[dict(name="lj1g", params=[0.0103, 3.405, 6.0, 7.0], lists=[(1, 1, 200, 7.5, 5)])])
    swapped = dict(mono, groups=[["B", "A"], ["A", "#"], ["B", "#"], ["#", "#"]])
    e = []
    for case in (mono, swapped):
        o = oracle(case)
        o.advance("nve", 0.5, 0, 1)
        nl = neighbours(o, case, 0, 0)
        e.append((o.energies()[0][0], int(nl[2].sum()), int(nl[1].sum())))
    (e_mono, half_mono, full_mono), (e_swap, half_swap, full_swap) = e
    assert full_mono == full_swap and half_mono * 2 == full_mono      # monotone: the half list holds every pair once
    assert half_swap != half_mono and abs(e_swap - e_mono) > 1e-6 * abs(e_mono)   # swapped columns: pairs are lost


def test_advance_logged_is_the_stepwise_sequence(oracle_lib):
    """oracle_advance_logged (checker twin of pfmds_advance_logged): rows = energies() after each logged step."""
    from util import oracle, small_cases
    case = small_cases()["cu_fcc"]
    a, b = oracle(case), oracle(case)
    a.advance("nvt", 2.0, 0, 1)
    rows = a.advance_logged("nvt", 2.0, 1, 9, log_period=2)
    b.advance("nvt", 2.0, 0, 1)
    got = []
    for s in range(1, 10):
        b.advance("nvt", 2.0, s, 1)
        if s % 2 == 0:
            got.append(b.energies())
    assert len(got) == 4 and rows[0].shape == (4, 1)
    for r, g in enumerate(got):   # OpenMP partial sums: two oracle runs agree to rounding, not bit for bit
        for k in range(4):
            assert np.allclose(rows[k][r], g[k], rtol=1e-11, atol=1e-12)

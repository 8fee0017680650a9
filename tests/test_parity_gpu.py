"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on identical inputs.
Bar (BASELINE.json north_star): neighbour sets bit-exact; forces and energies within 1e-9 relative."""
import numpy as np
import pytest

from pfmds_b200.engine import PfmdsError
from util import RTOL, gpu, list_ids, neighbours, oracle, rel_err, small_cases

pytestmark = pytest.mark.gpu
CASES = small_cases()


@pytest.fixture(scope="module", autouse=True)
def _libs(cuda_lib, oracle_lib):
    return None


@pytest.mark.parametrize("name", list(CASES))
def test_step0_lists_forces_energies(name):
    case = CASES[name]
    g, o = gpu(case), oracle(case)
    g.advance("nve", 1.0, 0, 1)
    o.advance("nve", 1.0, 0, 1)
    for k, j in list_ids(case):
        a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
        assert np.array_equal(a[1], b[1]), "nnum differs in list %d/%d" % (k, j)
        assert np.array_equal(a[0], b[0]), "nlist differs in list %d/%d" % (k, j)
        if not (j == 1 and case["interactions"][k]["name"] in ("lj", "ljc", "morsec")):  # the converse list has no lessnnum
            if j != 2:
                assert np.array_equal(a[2], b[2]), "lessnnum differs in list %d/%d" % (k, j)
    pg, vg, fg = g.download()
    po, vo, fo = o.download()
    assert rel_err(fg, fo) < RTOL
    eg, eo = g.energies(), o.energies()
    assert np.allclose(eg[0], eo[0], rtol=RTOL, atol=0)
    assert abs(eg[1] - eo[1]) <= RTOL * abs(eo[1])
    assert abs(eg[2] - eo[2]) <= RTOL * abs(eo[2])
    dg, do = g.diagnostics(), o.diagnostics()
    assert np.allclose(dg[1], do[1], rtol=1e-12)  # centre of mass
    assert abs(dg[3] - do[3]) <= 1e-12 * do[3]    # max velocity
    assert np.array_equal(dg[4], do[4])           # neighbour-list load
    for k, it in enumerate(case["interactions"]):
        if it["name"] in ("ljc", "morsec"):
            assert np.abs(g.normals(k) - o.normals(k)).max() < 1e-12


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("integrator", ["nve", "nvt", "nvms"])
def test_trajectory_22_steps(name, integrator):
    """Steps 0..21 with rebuilds at 0,5,10,15,20: state, lists and energies stay within tolerance."""
    case = CASES[name]
    dt = case["integrators"][0][1]
    g, o = gpu(case), oracle(case)
    for e in (g, o):
        e.advance(integrator, dt, 0, 1)
        e.advance(integrator, dt, 1, 21)
    pg, vg, fg = g.download()
    po, vo, fo = o.download()
    assert np.abs(pg - po).max() < 1e-9
    assert rel_err(vg, vo) < 1e-8
    assert rel_err(fg, fo) < 1e-7  # forces amplify the accumulated 1e-12 position differences
    for k, j in list_ids(case):
        a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    eg, eo = g.energies(), o.energies()
    assert np.allclose(eg[0], eo[0], rtol=1e-8, atol=1e-9)
    assert abs(eg[1] - eo[1]) <= 1e-8 * abs(eo[1]) + 1e-12
    if case["nhc"]:
        assert np.allclose(eg[3], eo[3], rtol=1e-7, atol=1e-9)
        for k in range(len(case["nhc"])):
            assert np.allclose(g.get_nhc(k)[1], o.get_nhc(k)[1], rtol=1e-7, atol=1e-15)


def test_single_step_forces_after_move():
    """One full step (kick, drift, refresh without rebuild, forces): forces still match to 1e-9."""
    case = small_cases()["cu_fcc"]
    g, o = gpu(case), oracle(case)
    for e in (g, o):
        e.advance("nve", 2.0, 0, 2)
    assert rel_err(g.download()[2], o.download()[2]) < RTOL


def test_zero_momentum_and_invert_z():
    case = small_cases()["ab_gas"]
    case = dict(case, zero_momentum_period=3, invert_z_vel=True)
    z = case["pos"][:, 2] / case["box"][2]
    assert ((z > 0.8) & (z < 0.9)).sum() > 10  # some atoms sit inside the elastic-wall slab
    g, o = gpu(case), oracle(case)
    for e in (g, o):
        e.advance("nve", 0.5, 0, 8)
    pg, vg, _ = g.download()
    po, vo, _ = o.download()
    assert np.abs(pg - po).max() < 1e-10
    assert rel_err(vg, vo) < 1e-9


def test_too_many_neighbours_is_reported():
    case = small_cases()["cu_fcc"]
    case["interactions"][0]["lists"] = [(1, 1, 20, 6.5, 5)]
    g = gpu(case)
    g.advance("nve", 1.0, 0, 1)
    with pytest.raises(PfmdsError) as ei:
        g.synchronize()
    assert ei.value.code == 11 and "too many neighbours" in str(ei.value)


def test_particle_out_of_cell_is_reported():
    case = small_cases()["cu_fcc"]
    case["pos"] = case["pos"].copy()
    case["pos"][7, 1] = case["box"][1] + 0.5
    g = gpu(case)
    g.advance("nve", 1.0, 0, 1)
    with pytest.raises(PfmdsError) as ei:
        g.synchronize()
    assert ei.value.code == 10 and "particle out of cell" in str(ei.value) and " 8 " in str(ei.value)


def test_not_enough_graphene_neighbours_is_reported():
    case = small_cases()["gr_cu_ljc"]
    case["interactions"][1]["lists"][2] = (1, 1, 3, 1.2, 5)  # r_cut_nn below the C-C bond length
    g = gpu(case)
    g.advance("nve", 1.0, 0, 1)
    with pytest.raises(PfmdsError) as ei:
        g.synchronize()
    assert ei.value.code == 12


def test_refuses_what_the_reference_gets_wrong_silently():
    case = small_cases()["ab_gas"]
    case["groups"] = [["A", "B"], ["B", "A"], ["B", "#"], ["#", "#"]]  # lj1g on a non-monotone group
    case["interactions"] = [dict(name="lj1g", params=[0.0103, 3.405, 6.0, 7.0], lists=[(2, 2, 120, 7.5, 5)])]
    g = gpu(case)
    with pytest.raises(PfmdsError) as ei:
        g.advance("nve", 1.0, 0, 1)
    assert ei.value.code == 20

"""(Named zz: on the GPU box it runs after the parity tests.)  Seeded random systems against the oracle: non-cubic boxes with a different number of cells per axis (down to two), sparse and
dense regions (empty cells, crowded cells), atoms exactly on the box faces, random cut-offs and switch radii — neighbour lists
bit-exact, forces and energies to 1e-9, a few steps of trajectory.  Runs on the serial host replay (thread-per-atom kernels, the
large-system path), on the lock-step replay (warp-per-atom list build, 8 lanes per atom: the small-system path) and on the GPU."""
import os
import sys

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.engine import configure
from util import gpu, list_ids, neighbours, oracle, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))


def _serial(case):
    import build_emu as B
    B.SERIAL.build()
    return configure(case, lib_path=B.SERIAL.lib)


def _lockstep(case):
    import build_emu as B
    B.LOCKSTEP.build()
    return configure(case, lib_path=B.LOCKSTEP.lib)


def _gpu_large(case):
    """The GPU with the size switches at zero: the thread-per-atom pipelined pair kernels and the thread-per-atom list build,
    i.e. the kernels the 10^6-atom bench times (util.PATHS)."""
    return gpu(case, "large")


ENGINES = [pytest.param(_serial, id="serial-replay"), pytest.param(_lockstep, id="lockstep-replay"), pytest.param(gpu, id="gpu", marks=pytest.mark.gpu),
           pytest.param(_gpu_large, id="gpu-large-path", marks=pytest.mark.gpu)]


@pytest.fixture(scope="module", autouse=True)
def _lib(oracle_lib):
    return None


def random_case(seed, kind):
    rng = np.random.default_rng(seed)
    box = rng.uniform(13.5, 34.0, 3)
    if seed % 3 == 0:
        box[rng.integers(3)] = rng.uniform(13.2, 14.9)             # one axis only two cells wide
    n_try = int(rng.integers(150, 420))
    pos = rng.uniform(0, 1, (n_try, 3)) * box
    if seed % 2 == 0:                                              # a crowded blob and a void
        pos[: n_try // 3] = (box * rng.uniform(0.2, 0.8, 3) + rng.normal(0, 2.5, (n_try // 3, 3))) % box
    pos[0] = [0.0, 0.0, 0.0]                                       # exactly on the faces (x = L is "out of cell" in the reference)
    pos[1] = [0.0, box[1] * (1 - 2.0 ** -53), box[2] / 2]
    pos[2] = [box[0] * (1 - 2.0 ** -53), box[1] / 2, box[2] / 3]
    keep = []                                                      # no pair closer than 2.1 A (minimum image)
    for i in range(n_try):
        d = pos[keep] - pos[i]
        d -= box * np.round(d / box)
        if not keep or (d * d).sum(1).min() > 2.1 ** 2:
            keep.append(i)
    pos = pos[keep]
    n = len(pos)
    rcut = float(rng.uniform(5.0, min(6.6, 0.49 * box.min())))
    R2 = float(rng.uniform(rcut - 1.2, rcut - 0.05))
    R1 = R2 - 1.0 if kind == "lj1g" else float(rng.uniform(R2 - 1.0, R2 - 0.2))       # lj1g: the switch width the reference's derivative assumes
    mass = np.full(n, 63.546)
    vel = inputs.maxwell(rng, mass, 300.0)
    if kind == "lj1g":
        inter = dict(name="lj1g", file="p.txt", params=[0.0103, 3.405, R1, R2], lists=[(1, 1, 200, rcut, 3)])
    else:
        inter = dict(name="rjl", file="p.txt", params=list(inputs.RJL_CU[:5]) + [R1, R2], lists=[(1, 1, 200, rcut, 3)])
    return dict(title="random", box=box, pos=pos, vel=vel, mass=mass, names=["A"] * n, groups=[["A"], ["#"]],
                roles=dict(all_moving=1, xyz_moving=1, z_moving=2, all_atoms=1, traj_group=2, period_traj=10 ** 9),
                integrators=[("nve", 0.5, 10, 10 ** 9, 10 ** 9)], ms_de=1e-8, nhc=[], zero_momentum_period=10 ** 9, invert_z_vel=False,
                initial_temperature=300.0, interactions=[inter])


@pytest.mark.parametrize("kind", ["lj1g", "rjl"])
@pytest.mark.parametrize("make", ENGINES)
def test_random_systems_match_the_oracle(make, kind):
    for seed in range(1, 4 if make is _lockstep else 7):              # the lock-step replay is ~4x slower per launch
        case = random_case(seed, kind)
        g, o = make(case), oracle(case)
        for e in (g, o):
            e.advance("nve", 0.5, 0, 1)
        a, b = neighbours(g, case, 0, 0), neighbours(o, case, 0, 0)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), (seed, "lists")
        assert a[1].max() > 3, seed
        fo = o.download()[2]
        assert rel_err(g.download()[2], fo) < 1e-9, (seed, "forces")
        assert abs(g.energies()[0][0] - o.energies()[0][0]) < 1e-9 * abs(o.energies()[0][0]) + 1e-12, (seed, "energy")
        for e in (g, o):
            e.advance("nve", 0.5, 1, 4)                              # a rebuild at step 3
        assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-9, (seed, "trajectory")
        g.close()
        o.close()


def random_graphene_case(seed, interface):
    """Rippled, jittered graphene on Cu(111) (tb + ljc|morsec + rjl) with random interface parameters and switch radii; a
    seed in three uses the `simplified` normals (n = z)."""
    rng = np.random.default_rng(100 + seed)
    case = inputs.graphene_on_cu_small(interface=interface, seed=seed, jitter=float(rng.uniform(0.02, 0.06)), period=3, simplified=(seed % 3 == 0))
    box, pos = case["box"], case["pos"]
    nc = case["names"].count("C")
    amp, ph = rng.uniform(0.1, 0.4), rng.uniform(0, 2 * np.pi, 2)
    pos[:nc, 2] += amp * np.sin(2 * np.pi * pos[:nc, 0] / box[0] + ph[0]) * np.cos(2 * np.pi * pos[:nc, 1] / box[1] + ph[1])
    R2 = float(rng.uniform(5.6, 7.2))
    R1 = R2 - float(rng.uniform(0.5, 1.2))
    it = case["interactions"][1]
    simp = it["params"][-1]
    if interface == "ljc":
        it["params"] = [float(rng.uniform(0.01, 0.04)), float(rng.uniform(2.6, 3.3)), float(rng.uniform(1.0, 3.0)), R1, R2, simp]
    else:
        it["params"] = [float(rng.uniform(0.01, 0.05)), float(rng.uniform(2.8, 3.5)), float(rng.uniform(0.8, 1.6)), float(rng.uniform(1.0, 3.0)), R1, R2, simp]
    return case


@pytest.mark.parametrize("interface", ["ljc", "morsec"])
@pytest.mark.parametrize("make", ENGINES)
def test_random_graphene_on_cu_matches_the_oracle(make, interface):
    for seed in range(1, 3 if make is _lockstep else 4):
        case = random_graphene_case(seed, interface)
        g, o = make(case), oracle(case)
        for e in (g, o):
            e.advance("nvt", 1.0, 0, 1)
        for k, j in list_ids(case):
            a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), (seed, "list", k, j)
        assert rel_err(g.normals(1), o.normals(1)) < 1e-12, (seed, "normals")
        assert rel_err(g.download()[2], o.download()[2]) < 1e-9, (seed, "forces")
        eg, eo = g.energies(), o.energies()
        assert np.allclose(eg[0], eo[0], rtol=1e-9, atol=1e-12), (seed, "energies", eg[0], eo[0])
        for e in (g, o):
            e.advance("nvt", 1.0, 1, 4)                              # a rebuild at step 3
        assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-9, (seed, "trajectory")
        assert np.allclose(g.get_nhc(0)[1], o.get_nhc(0)[1], rtol=1e-8, atol=1e-15), (seed, "thermostat")
        g.close()
        o.close()


def random_ab_gas_case(seed):
    """Two-species lj gas (list + converse list) plus lj1g within each species: random composition, box and cut-offs."""
    rng = np.random.default_rng(200 + seed)
    n_side = int(rng.integers(6, 9))
    case = inputs.ab_gas(n_side=n_side, spacing=float(rng.uniform(3.5, 4.3)), seed=seed, frac_b=float(rng.uniform(0.1, 0.4)), cap_aa=120, cap_ab=120, cap_ba=120, cap_bb=120, period=3)
    return case


@pytest.mark.parametrize("make", ENGINES)
def test_random_ab_gas_matches_the_oracle(make):
    for seed in range(1, 3 if make is _lockstep else 4):
        case = random_ab_gas_case(seed)
        integ, dt = case["integrators"][0][0], case["integrators"][0][1]
        g, o = make(case), oracle(case)
        for e in (g, o):
            e.advance(integ, dt, 0, 1)
        for k, j in list_ids(case):
            a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), (seed, "list", k, j)
            if not (j == 1 and case["interactions"][k]["name"] == "lj"):   # the converse list has no lessnnum (md_neighbours.f90:128-160)
                assert np.array_equal(a[2], b[2]), (seed, "lessnnum", k, j)
        assert rel_err(g.download()[2], o.download()[2]) < 1e-9, (seed, "forces")
        assert np.allclose(g.energies()[0], o.energies()[0], rtol=1e-9, atol=1e-12), (seed, "energies")
        for e in (g, o):
            e.advance(integ, dt, 1, 4)
        assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-9, (seed, "trajectory")
        g.close()
        o.close()


@pytest.mark.parametrize("n_b", [0, 1])
@pytest.mark.parametrize("make", ENGINES)
def test_empty_and_single_atom_groups(make, n_b):
    """Edge of the two-group interactions: species B absent (the lj list, its converse and the B-B lj1g list have no rows) or
    present with one atom (a one-row converse list, an lj1g list whose only row is empty)."""
    n = 6 ** 3
    case = inputs.ab_gas(n_side=6, frac_b=(n_b + 0.25) / n, cap_aa=120, cap_ab=120, cap_ba=240, cap_bb=8, period=3)
    assert case["names"].count("B") == n_b
    g, o = make(case), oracle(case)
    for e in (g, o):
        e.advance("nvt", 0.5, 0, 1)
    for k, j in list_ids(case):
        a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
        assert a[0].shape[0] == (n_b if (k, j) in ((0, 1), (2, 0)) else n - n_b)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), ("list", k, j)
    assert rel_err(g.download()[2], o.download()[2]) < 1e-9
    eg, eo = g.energies(), o.energies()
    assert np.allclose(eg[0], eo[0], rtol=1e-9, atol=1e-12) and eg[0][2] == 0.0 and (n_b or eg[0][0] == 0.0)
    for e in (g, o):
        e.advance("nvt", 0.5, 1, 7)                                  # rebuilds at steps 3 and 6
    assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-9
    assert abs(g.energies()[2] - o.energies()[2]) < 1e-7            # temperature
    g.close()
    o.close()


@pytest.mark.parametrize("make", ENGINES)
def test_list_capacity_exactly_reached(make):
    """neighb_num_max equal to the fullest row is enough, one less stops with the reference's message (md_neighbours.f90:82,
    :148 for the converse list) — on the oracle and on the device alike.  Perfect fcc within 6.5 A: 86 neighbours everywhere."""
    from pfmds_b200.engine import PfmdsError
    for cap, ok in ((86, True), (85, False)):
        case = inputs.cu_fcc(ncell=4, period=5)
        case["interactions"][0]["lists"] = [(1, 1, cap, 6.5, 5)]
        for which, e in enumerate((make(case), oracle(case))):
            if ok:
                e.advance("nve", 1.0, 0, 1)
                e.synchronize()
                assert e.diagnostics()[4][0] == 86
            else:
                with pytest.raises(PfmdsError) as ei:
                    e.advance("nve", 1.0, 0, 1)
                    e.synchronize()
                assert "too many neighbours" in str(ei.value) and (which == 1 or ei.value.code == 11)   # the oracle has one code for every stop
            e.close()
    # the converse list of lj has its own capacity (settings line 2)
    n = 6 ** 3
    base = inputs.ab_gas(n_side=6, frac_b=0.2, cap_aa=120, cap_ab=120, cap_ba=240, cap_bb=120, period=3)
    o = oracle(base)
    o.advance("nvt", 0.5, 0, 1)
    full = int(neighbours(o, base, 0, 1)[1].max())
    for cap, ok in ((full, True), (full - 1, False)):
        case = inputs.ab_gas(n_side=6, frac_b=0.2, cap_aa=120, cap_ab=120, cap_ba=cap, cap_bb=120, period=3)
        for which, e in enumerate((make(case), oracle(case))):
            if ok:
                e.advance("nvt", 0.5, 0, 1)
                e.synchronize()
                assert int(neighbours(e, case, 0, 1)[1].max()) == full
            else:
                with pytest.raises(PfmdsError) as ei:
                    e.advance("nvt", 0.5, 0, 1)
                    e.synchronize()
                assert "too many neighbours" in str(ei.value) and (which == 1 or ei.value.code == 11)   # the oracle has one code for every stop
            e.close()


@pytest.mark.parametrize("make", ENGINES)
def test_z_only_movers_and_fixed_atoms(make):
    """The z_moving group (integrate_verlet_z_velocities / _z_positions, md_integrators.f90:34-56,79-97; its nvms quench
    molecular_static_1D_velocities, :125-145) next to xyz movers and fixed atoms, through nvt, nve and nvms steps."""
    case = inputs.cu_fcc(ncell=4, jitter=0.06, period=3)
    n = len(case["mass"])
    rng = np.random.default_rng(5)
    kind = rng.permutation(n) % 4                                     # 0,1: xyz movers, 2: z movers, 3: fixed
    order = np.argsort(kind, kind="stable")                           # file order by type keeps the multi-type groups index-monotone
    for key in ("pos", "vel"):
        case[key] = case[key][order]
    kind = kind[order]
    case["names"] = [("CU", "CU", "CUZ", "CUF")[k] for k in kind]
    case["vel"][kind == 3] = 0.0
    case["groups"] = [["CU", "CUZ", "CUF"], ["CU", "CUZ", "#"], ["CU", "#", "#"], ["CUZ", "#", "#"]]
    case["roles"] = dict(all_moving=2, xyz_moving=3, z_moving=4, all_atoms=1, traj_group=4, period_traj=10 ** 9)
    case["nhc"] = [(2, 300.0, 3, case["nhc"][0][3])]
    p0 = case["pos"].copy()
    g, o = make(case), oracle(case)
    step = 0
    for integ, k in (("nvt", 7), ("nve", 4), ("nvms", 5)):
        for e in (g, o):
            e.advance(integ, 2.0, step, k)
        step += k
        (pg, vg, fg), (po, vo, fo) = g.download(), o.download()
        assert np.abs(pg - po).max() < 1e-9 and rel_err(vg, vo) < 1e-8 and rel_err(fg, fo) < 1e-9, integ
        assert np.array_equal(pg[kind == 3], p0[kind == 3]) and np.array_equal(pg[kind == 2][:, :2], p0[kind == 2][:, :2])
        assert np.abs(pg[kind == 2][:, 2] - p0[kind == 2][:, 2]).max() > 1e-4
        eg, eo = g.energies(), o.energies()
        assert np.allclose(eg[0], eo[0], rtol=1e-9) and abs(eg[1] - eo[1]) < 1e-8 * abs(eo[1]) + 1e-12
    vz = g.download()[1][kind == 2]
    assert np.all(vz[:, :2] == vz[:, :2]) and (np.abs(vz[:, 2]) == 0).any() and (np.abs(vz[:, 2]) > 0).any()   # the 1D quench zeroed some, kept others
    g.close()
    o.close()


@pytest.mark.parametrize("make", ENGINES)
def test_several_thermostats_of_different_chain_lengths(make):
    """nhc_num > 1 (md_simulation.f90:75-84, :150-154, :176-180): thermostats on disjoint groups with M = 2 and M = 1, and a third of
    M = 4 over both (its atoms are scaled twice per half step, in file order of the thermostats)."""
    from pfmds_b200.inputs import KB
    case = inputs.graphene_on_cu_small(interface="morsec", period=4)
    case["groups"] = case["groups"] + [["CU", "#", "#"]]             # 6: the moving copper
    nC, nCu = case["names"].count("C"), case["names"].count("CU")
    q = lambda n, T: 3 * n * KB * T * 100.0 ** 2
    case["nhc"] = [(1, 350.0, 2, q(nC, 350.0)), (6, 250.0, 1, q(nCu, 250.0)), (3, 300.0, 4, q(nC + nCu, 300.0))]
    g, o = make(case), oracle(case)
    for e in (g, o):
        e.advance("nvt", 1.0, 0, 13)                                  # rebuilds at steps 4, 8, 12
    (pg, vg, fg), (po, vo, fo) = g.download(), o.download()
    assert np.abs(pg - po).max() < 1e-9 and rel_err(vg, vo) < 1e-8 and rel_err(fg, fo) < 1e-9
    eg, eo = g.energies(), o.energies()
    assert np.allclose(eg[0], eo[0], rtol=1e-9) and abs(eg[1] - eo[1]) < 1e-8 * eo[1] and np.allclose(eg[3], eo[3], rtol=1e-7, atol=1e-12)
    for k, (_, _, M, _) in enumerate(case["nhc"]):
        (xg, wg), (xo, wo) = g.get_nhc(k), o.get_nhc(k)
        assert len(xg) == M and np.allclose(xg, xo, rtol=1e-7, atol=1e-15) and np.allclose(wg, wo, rtol=1e-7, atol=1e-15), k
        assert np.abs(wo).max() > 0
    g2 = make(case)                                                   # the same steps one at a time, energies logged on the device
    g2.advance("nvt", 1.0, 0, 1)
    rows = g2.advance_logged("nvt", 1.0, 1, 12, log_period=1)
    assert rows[0].shape == (12, 3) and rows[3].shape == (12, 3)
    assert np.allclose(rows[0][-1], eo[0], rtol=1e-9) and np.allclose(rows[3][-1], eo[3], rtol=1e-7, atol=1e-12)
    assert np.abs(g2.download()[0] - po).max() < 1e-9
    for e in (g, g2, o):
        e.close()


@pytest.mark.parametrize("make", ENGINES)
def test_momentum_removal_reaches_the_fixed_atoms_too(make):
    """zero_momentum works on the all_atoms group (md_simulation.f90:162, md_general.f90:236-253): atoms outside the moving groups
    receive the velocity shift as well (and keep their positions).  With invert_z_vel on, a layer of fast atoms inside the
    reflecting slab [0.8 Lz, 0.9 Lz] (md_general.f90:382-398)."""
    case = inputs.graphene_on_cu_small(interface="ljc", period=4, lz=40.0)
    case = dict(case, zero_momentum_period=2, invert_z_vel=True)
    fixed = np.array([nm == "CU_fixed" for nm in case["names"]])
    nc = case["names"].count("C")
    case["pos"] = case["pos"].copy(); case["vel"] = case["vel"].copy()
    case["pos"][:nc, 2] += 0.85 * 40.0 - case["pos"][:nc, 2].mean()      # lift the sheet into the reflecting slab (away from the metal)
    case["vel"][:nc:2, 2] = 0.01
    case["vel"][1:nc:2, 2] = -0.01
    p0 = case["pos"].copy()
    g, o = make(case), oracle(case)
    for e in (g, o):
        e.advance("nvt", 1.0, 0, 7)
    (pg, vg, fg), (po, vo, fo) = g.download(), o.download()
    assert np.abs(pg - po).max() < 1e-9 and np.abs(vg - vo).max() < 1e-9 * np.abs(vo).max()
    assert np.array_equal(pg[fixed], p0[fixed]) and np.abs(vo[fixed]).max() > 0 and np.abs(vg[fixed] - vo[fixed]).max() < 1e-12 * np.abs(vo).max() + 1e-18
    dg, do = g.diagnostics(), o.diagnostics()
    assert np.allclose(dg[2], do[2], rtol=1e-6, atol=1e-14) and abs(dg[3] - do[3]) < 1e-9 * do[3]
    g.close()
    o.close()


@pytest.mark.parametrize("make", ENGINES)
def test_lists_with_different_update_periods(make):
    """Every list follows its own mod(md_step, update_period) (md_neighbours.f90:41): tb every 3, the interface lists every 5, rjl
    every 7 steps.  Only steps on which all of them rebuild re-sort the atoms on the device; the partial rebuilds in between must
    give the reference's lists all the same."""
    case = inputs.graphene_on_cu_small(interface="ljc", period=5, jitter=0.05)
    case["interactions"][0]["lists"] = [(1, 1, 12, 2.6, 3)]
    case["interactions"][1]["lists"] = [(1, 2, 64, 7.5, 5), (2, 1, 96, 7.5, 5), (1, 1, 3, 1.9, 3)]
    case["interactions"][2]["lists"] = [(2, 2, 100, 6.5, 7)]
    case["vel"] = case["vel"] * 3.0                                   # hotter: list membership changes within a few steps
    g, o = make(case), oracle(case)
    done = 0
    for upto in (4, 6, 8, 11, 15, 16):                                # stops after rebuilds of one, two or all lists
        for e in (g, o):
            e.advance("nvt", 1.0, done, upto - done)
        done = upto
        for k, j in list_ids(case):
            a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), (upto, "list", k, j)
        (pg, vg, fg), (po, vo, fo) = g.download(), o.download()
        assert np.abs(pg - po).max() < 1e-9 and rel_err(fg, fo) < 1e-9, upto
        assert np.allclose(g.energies()[0], o.energies()[0], rtol=1e-9), upto
    g.close()
    o.close()


@pytest.mark.parametrize("make", ENGINES)
def test_random_rebosc_sheets_match_the_oracle(make):
    """rebosc with its numerical forces (REBOsolidcarbon.f90:27-91, md_interactions.f90:273-311) on rippled, jittered sheets of
    different sizes, alone and with a tb interaction behind it: energies 1e-9, forces within 2e-7 of max|F| (the noise floor of the
    reference's own central differences), a short trajectory."""
    for seed in range(1, 3 if make is _lockstep else 4):
        rng = np.random.default_rng(300 + seed)
        cells = [(4, 3), (5, 3), (6, 4)][seed % 3]
        case = inputs.graphene_rebosc(cells=cells, seed=seed, jitter=float(rng.uniform(0.02, 0.07)), period=3, with_tb=(seed % 2 == 0))
        box, pos = case["box"], case["pos"]
        pos[:, 2] += rng.uniform(0.1, 0.4) * np.sin(2 * np.pi * pos[:, 0] / box[0] + rng.uniform(0, 6.28)) * np.cos(2 * np.pi * pos[:, 1] / box[1])
        g, o = make(case), oracle(case)
        for e in (g, o):
            e.advance("nve", 0.5, 0, 1)
        for k, j in list_ids(case):
            a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), (seed, "list", k, j)
        fo = o.download()[2]
        assert rel_err(g.download()[2], fo) < 2e-7 and np.abs(fo).max() > 0.1, (seed, "forces")
        assert np.allclose(g.energies()[0], o.energies()[0], rtol=1e-9, atol=1e-12), (seed, "energies")
        for e in (g, o):
            e.advance("nve", 0.5, 1, 4)
        assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-8, (seed, "trajectory")
        g.close()
        o.close()

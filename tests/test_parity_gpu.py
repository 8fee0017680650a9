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


PATH_IDS = ["small", "large", "large_build"]   # util.PATHS: which side of the size switches the context takes


@pytest.mark.parametrize("path", PATH_IDS)
@pytest.mark.parametrize("name", list(CASES))
def test_step0_lists_forces_energies(name, path):
    case = CASES[name]
    g, o = gpu(case, path), oracle(case)
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


@pytest.mark.parametrize("path", ["small", "large"])
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("integrator", ["nve", "nvt", "nvms"])
def test_trajectory_22_steps(name, integrator, path):
    """Steps 0..21 with rebuilds at 0,5,10,15,20: state, lists and energies stay within tolerance."""
    case = CASES[name]
    dt = case["integrators"][0][1]
    g, o = gpu(case, path), oracle(case)
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


def test_device_math_functions():
    """mathx.cuh on the device (constant-bank exp, select-free cosine switch, MUFU-seeded rsqrt) against the CUDA math library."""
    import ctypes as C
    from pfmds_b200.engine import load_library
    lib = load_library()
    err = (C.c_double * 4)()
    lib.pfmds_selftest_math.argtypes = [C.c_int, C.POINTER(C.c_double)]
    assert lib.pfmds_selftest_math(0, err) == 0
    print("device math errors: exp %.2e switch %.2e rsqrt %.2e seed %.2e" % tuple(err))
    assert err[0] < 3e-14 and err[1] < 1e-15 and err[2] < 1e-15
    lib.pfmds_selftest_math2.argtypes = [C.c_int, C.POINTER(C.c_double)]
    assert lib.pfmds_selftest_math2(0, err) == 0
    print("short forms: exp %.2e switch %.2e rsqrt_q %.2e exp(wide) %.2e" % tuple(err))
    assert err[0] < 1e-14 and err[1] < 2e-14 and err[2] < 3e-12 and err[3] < 4e-14


@pytest.mark.parametrize("path", ["small", "large"])
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_matches_golden_fixtures(name, path):
    """Against the committed fixtures (tests/golden/, frozen oracle outputs): lists bit-exact, forces/energies 1e-9."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    case = CASES[name]
    e = gpu(case, path)
    integ, dt = case["integrators"][0][0], case["integrators"][0][1]
    e.advance(integ, dt, 0, 1)
    assert rel_err(e.download()[2], g["frc0"]) < RTOL
    assert np.allclose(e.energies()[0], g["e0"], rtol=RTOL, atol=0)
    for k, j in list_ids(case):
        nl = neighbours(e, case, k, j)
        assert np.array_equal(nl[1], g["nnum_%d_%d" % (k, j)]) and np.array_equal(nl[0], g["nlist_%d_%d" % (k, j)])
    e.advance(integ, dt, 1, 10)
    p, v, f = e.download()
    assert np.abs(p - g["pos10"]).max() < 1e-10 and rel_err(f, g["frc10"]) < 1e-8


def test_upload_restarts_a_context():
    case = small_cases()["cu_fcc"]
    a, b = gpu(case), gpu(case)
    a.advance("nve", 2.0, 0, 11)
    p, v, _ = a.download()
    b.advance("nve", 2.0, 0, 1)
    b.upload(p, v)
    b.advance("nve", 2.0, 0, 1)   # step 0: lists + forces on the uploaded state
    a.advance("nve", 2.0, 0, 1)
    assert rel_err(b.download()[2], a.download()[2]) < 1e-12


def test_multi_step_advance_equals_single_steps():
    case = small_cases()["gr_cu_morsec"]
    a, b = gpu(case), gpu(case)
    a.advance("nvt", 1.0, 0, 13)
    for s in range(13):
        b.advance("nvt", 1.0, s, 1)
    pa, va, fa = a.download()
    pb, vb, fb = b.download()
    assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)  # bitwise: fixed-order reductions


def test_two_contexts_share_a_gpu():
    """Ensemble mode: independent contexts (own streams) on one device do not interfere."""
    c1, c2 = small_cases()["cu_fcc"], small_cases()["ab_gas"]
    a, b, ref = gpu(c1), gpu(c2), gpu(c1)
    for s in range(6):
        a.advance("nvt", 2.0, s, 1)
        b.advance("nve", 0.5, s, 1)
    ref.advance("nvt", 2.0, 0, 6)
    assert np.array_equal(a.download()[0], ref.download()[0])


def test_full_size_properties_cu_fcc():
    """BASELINE.json configs[1] at full size (1 000 188 atoms, no oracle at this size): size-independent properties.
    sum F = 0 (Newton 3), perfect-lattice neighbour count 86, extended energy conserved, momentum stays zero."""
    from pfmds_b200 import inputs
    case = inputs.cu_fcc(ncell=63)
    e = gpu(case)
    e.advance("nvt", 2.0, 0, 1)
    fs, mc, mcv, vmax, load = e.diagnostics()
    f = e.download()[2]
    assert np.abs(fs).max() < 1e-9 and np.abs(f).max() < 1e-9      # perfect fcc lattice: every force vanishes
    assert load[0] == 86 and e.pair_count(0, 0) == 86 * e.n        # shells within 6.5 A: 12+6+24+12+24+8
    en0 = e.energies()
    c0 = en0[0].sum() + en0[1] + en0[3].sum()
    assert abs(en0[2] - 300.0) < 1e-6                               # generator rescales to exactly 300 K
    small = inputs.cu_fcc(ncell=4)                                  # perfect lattice: energy per atom is size independent
    o = oracle(small)
    o.advance("nvt", 2.0, 0, 1)
    assert abs(en0[0][0] / e.n - o.energies()[0][0] / len(small["mass"])) < 1e-10 * abs(en0[0][0] / e.n)
    e.advance("nvt", 2.0, 1, 40)
    en1 = e.energies()
    c1 = en1[0].sum() + en1[1] + en1[3].sum()
    assert abs(c1 - c0) < 2e-3 * en0[1]
    fs, mc, mcv, vmax, load = e.diagnostics()
    assert np.abs(fs).max() < 1e-7 and np.linalg.norm(mcv) < 1e-12


@pytest.mark.parametrize("name", ["ab_gas", "cu_fcc"])
def test_nve_energy_drift_over_10k_steps(name):
    """north_star: NVE energy drift matches over 10^4 steps.  Trajectories of a chaotic system separate after ~10^3
    steps, so positions are not compared; the conserved energy is: its excursion stays bounded and equal in size on
    both sides, and the final total energies agree to the level set by that bounded fluctuation."""
    case = CASES[name]
    dt = case["integrators"][0][1] / 2
    g, o = gpu(case), oracle(case)
    tr = {}
    for tag, e in (("gpu", g), ("cpu", o)):
        e.advance("nve", dt, 0, 1)
        en = e.energies()
        e0, ke0 = en[0].sum() + en[1], en[1]
        dev = []
        for k in range(10):
            e.advance("nve", dt, 1 + 1000 * k, 1000)
            en = e.energies()
            dev.append(en[0].sum() + en[1] - e0)
        tr[tag] = (np.array(dev), ke0, e0)
    dg, ke0, e0 = tr["gpu"]
    dc = tr["cpu"][0]
    assert abs(tr["gpu"][2] - tr["cpu"][2]) <= 1e-9 * abs(e0)                  # same starting energy
    assert np.abs(dg).max() < 2e-3 * ke0 and np.abs(dc).max() < 2e-3 * ke0     # bounded: no drift on either side
    assert abs(np.abs(dg).max() - np.abs(dc).max()) < 1e-3 * ke0               # same size of the Verlet fluctuation
    assert abs(dg[:1] - dc[:1]).max() < 1e-6 * ke0                             # still the same trajectory after 10^3 steps


@pytest.mark.parametrize("path", ["small", "large"])
@pytest.mark.parametrize("name", list(CASES))
def test_energies_from_the_force_pass(name, path):
    """pfmds_advance_with_energy: the potential energies produced inside the last step's force pass equal the separate sweep
    (and, large path, the oracle's: k_rjl_force_e / k_lj1g_pipe<E> / k_lj<.,E,1> on hardware)."""
    case = CASES[name]
    integ, dt = case["integrators"][0][0], case["integrators"][0][1]
    a, b = gpu(case, path), gpu(case, path)
    o = oracle(case)
    o.advance(integ, dt, 0, 1)
    c = gpu(case, path)
    c.advance(integ, dt, 0, 1, with_energy=True)
    assert np.allclose(c.energies()[0], o.energies()[0], rtol=RTOL, atol=0)
    a.advance(integ, dt, 0, 7, with_energy=True)
    b.advance(integ, dt, 0, 7)
    ea, eb = a.energies(), b.energies()
    # (rjl takes the in-step energy from its force pass, the sweep from its density pass: two formulas, measured 2e-16 apart)
    assert np.allclose(ea[0], eb[0], rtol=1e-12, atol=0) and ea[1] == eb[1]
    assert np.array_equal(a.download()[2], b.download()[2])      # and the forces are the same bits
    a.advance(integ, dt, 7, 3)                                     # a later plain step invalidates the cached energies
    b.advance(integ, dt, 7, 3)
    assert np.allclose(a.energies()[0], b.energies()[0], rtol=1e-12, atol=0)


def test_store_instead_of_zero_plus_accumulate():
    """When rjl is the first interaction and owns every atom its force kernel stores the force (no zero pass).  The same
    system with a no-op interaction in front (lj1g on an empty group) takes the zero + accumulate path: identical bits."""
    from pfmds_b200 import inputs
    a = inputs.cu_fcc(ncell=30, jitter=0.03, period=5)          # 108 000 atoms: the large-system kernels
    b = dict(a, interactions=[dict(name="lj1g", params=[0.01, 3.0, 6.0, 7.0], lists=[(2, 2, 8, 6.5, 5)])] + a["interactions"])   # same r_cut: same cell grid, same atom order
    ea, eb = gpu(a), gpu(b)
    for e in (ea, eb):
        e.advance("nvt", 2.0, 0, 12)
    pa, va, fa = ea.download()
    pb, vb, fb = eb.download()
    assert np.array_equal(fa, fb) and np.array_equal(pa, pb) and np.array_equal(va, vb)
    assert np.abs(fa).max() > 0.1


@pytest.mark.parametrize("path", ["small", "large"])
@pytest.mark.parametrize("name", ["ab_gas", "cu_fcc", "gr_cu_ljc"])
def test_advance_logged_rows(name, path):
    """pfmds_advance_logged: the device-resident energy log returns, bit for bit, what advance_with_energy(1) + energies() give
    step by step, and leaves the same state; against the oracle the rows agree to 1e-9."""
    case = CASES[name]
    integ, dt = case["integrators"][0][0], case["integrators"][0][1]
    a, b, o = gpu(case, path), gpu(case, path), oracle(case)
    for e in (a, b, o):
        e.advance(integ, dt, 0, 1)
    rows = a.advance_logged(integ, dt, 1, 12, log_period=3)
    ref = o.advance_logged(integ, dt, 1, 12, log_period=3)
    got = []
    for s in range(1, 13):
        b.advance(integ, dt, s, 1, with_energy=(s % 3 == 0))
        if s % 3 == 0:
            got.append(b.energies())
    assert rows[0].shape[0] == 4
    for r, g in enumerate(got):
        assert np.array_equal(rows[0][r], g[0]) and rows[1][r] == g[1] and rows[2][r] == g[2] and np.array_equal(rows[3][r], g[3])
    pa, va, fa = a.download()
    pb, vb, fb = b.download()
    assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)
    for k in range(4):
        assert np.allclose(rows[k], ref[k], rtol=1e-7, atol=1e-9 * max(1.0, np.abs(ref[k]).max()))   # 12 steps of a trajectory, not step 0
    assert np.allclose(rows[0][0], ref[0][0], rtol=RTOL * 100, atol=0)


def test_download_into_caller_buffers():
    """pfmds_download gathers into file order on the device and copies straight into the caller's arrays; arrays that are
    not asked for (NULL) are skipped."""
    case = CASES["gr_cu_ljc"]
    e = gpu(case)
    e.advance("nvt", 1.0, 0, 7)                     # past a rebuild: slots are in cell order, not file order
    p, v, f = e.download()
    n = len(case["mass"])
    bp, bf = np.full((n, 3), np.nan), np.full((n, 3), np.nan)
    q = e.download(out=(bp, None, bf))
    assert q[0] is bp and q[1] is None and np.array_equal(bp, p) and np.array_equal(bf, f)
    p2, v2, f2 = e.download(forces=False)
    assert f2 is None and np.array_equal(p2, p) and np.array_equal(v2, v)
    with pytest.raises(PfmdsError):
        e.download(out=(np.zeros((n, 2)), None, None))


def test_rjl_in_a_box_narrower_than_twice_R2():
    """The second-generation rjl routines decide the minimum image after r^2, which needs R2 <= half the shortest box edge; a
    narrower box (10.8 A against R2 = 6.0) takes the first-generation routines.  Same lists, forces and trajectory as the oracle."""
    from pfmds_b200 import inputs
    case = inputs.cu_fcc(ncell=3, jitter=0.05, period=5)
    g, o = gpu(case), oracle(case)
    for e in (g, o):
        e.advance("nvt", 2.0, 0, 1)
    a, b = neighbours(g, case, 0, 0), neighbours(o, case, 0, 0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert rel_err(g.download()[2], o.download()[2]) < 1e-12 and abs(g.energies()[0][0] / o.energies()[0][0] - 1) < 1e-13   # libm-grade routines
    for e in (g, o):
        e.advance("nvt", 2.0, 1, 12)
    assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-12


def _check_steps_0_1_20_21(case, path, integ, dt, lists_every_time=True):
    """parity at steps 0, 1, 20, 21 (update_period 20: rebuilds at 0 and 20).  Neighbour sets bit-exact, per-atom
    forces and energies within 1e-9 relative at each of those steps."""
    g, o = gpu(case, path), oracle(case)
    at = 0
    for upto in (0, 1, 20, 21):
        for e in (g, o):
            e.advance(integ, dt, at, upto + 1 - at)
        at = upto + 1
        fg, fo = g.download()[2], o.download()[2]
        assert np.abs(fo).max() > 1e-3
        assert rel_err(fg, fo) < RTOL, "forces at step %d" % upto
        eg, eo = g.energies(), o.energies()
        assert np.allclose(eg[0], eo[0], rtol=RTOL, atol=0), "energies at step %d" % upto
        assert abs(eg[1] - eo[1]) <= RTOL * abs(eo[1])
        if lists_every_time or upto in (0, 20):
            for k, j in list_ids(case):
                a, b = neighbours(g, case, k, j), neighbours(o, case, k, j)
                assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]), "list %d/%d at step %d" % (k, j, upto)
                if j == 0:
                    assert np.array_equal(a[2], b[2])
    assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-9
    g.close()


@pytest.mark.parametrize("path", ["small", "large"])
def test_config0_ab_gas_at_its_stated_size(path):
    """BASELINE.json configs[0] at its stated 22^3 = 10 648 atoms: lj + 2 x lj1g, nvt."""
    from pfmds_b200 import inputs
    case = inputs.ab_gas()
    assert len(case["mass"]) == 10648
    _check_steps_0_1_20_21(case, path, "nvt", case["integrators"][0][1])


@pytest.mark.parametrize("path", ["small", "large"])
def test_config2_graphene_on_cu_at_its_stated_size(path):
    """BASELINE.json configs[2] at its stated 11 028 atoms: tb + ljc + rjl, nvt."""
    from pfmds_b200 import inputs
    case = inputs.graphene_on_cu()
    assert len(case["mass"]) == 11028
    period = case["interactions"][0]["lists"][0][4]
    assert 20 % period == 0
    _check_steps_0_1_20_21(case, path, "nvt", case["integrators"][0][1])


def test_config1_kernels_on_a_108000_atom_crystal():
    """The kernels BASELINE.json configs[1] is timed on (k_rjl_density / k_rjl_force / k_rjl_force_e, k_build_mask, thread per atom,
    zero_forces fused away, no CUDA graphs... exactly what a 10^6-atom run launches) on the largest jittered crystal the O(N^2)
    oracle handles in tens of seconds: 30^3 cells = 108 000 atoms, default switches (108 000 >= small_n; the list build is
    forced to the thread-per-atom kernel, which the default takes from 200 000 atoms up)."""
    from pfmds_b200 import inputs
    case = inputs.cu_fcc(ncell=30, jitter=0.03, period=20)
    assert len(case["mass"]) == 108000
    _check_steps_0_1_20_21(case, {"PFMDS_NL_WARP_N": "0"}, "nvt", 2.0, lists_every_time=False)

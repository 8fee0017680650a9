"""The GPU parity tests re-run on the CPU against the HOST REPLAY of the device library (tests/emu/build_emu.py: the same .cu
sources compiled by g++ with tests/emu/host_emu.hpp, kernels as serial loops over (block, thread)).  What this checks
without a GPU: the C-ABI orchestration of libpfmds_b200 (step sequence of pfmds_advance, fused NVT path, deposition, rebosc,
checkpoint / restore, error reporting, the hosts) and the arithmetic and indexing of every thread-per-atom kernel, against the
oracle and the golden fixtures, with the GPU tests' own assertions; in the lock-step flavour (fibers, see below) also the
warp-cooperative kernels the GPU launches for small systems (8 lanes per atom, warp-per-atom list build, scans, shared-memory
reductions).  What it cannot check: the compiled SASS, CUDA graphs, streams, NVLink — the `-m gpu` run remains the gate.
The product never loads this library (tests/test_cabi.py::test_no_cpu_fallback)."""
import importlib
import os
import sys

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.engine import configure

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu as BE  # noqa: E402

B = BE.SERIAL   # the flavour the current test runs against (set per test by the `flavour` fixture)

# Every test exists in two flavours: `serial` (threads of a block one after the other, thread-per-atom kernels) and `lockstep`
# (fibers; the kernels the GPU launches for systems under 10^5 atoms: 8 lanes per atom, warp-per-atom list build, block scans,
# shuffle / shared-memory reductions).  The lock-step replay costs ~3x more per launch: tests/conftest.py keeps a representative
# subset of its tests in the default run (LOCKSTEP_KEEP); PFMDS_LOCKSTEP_TESTS=all runs all of them (12 min).
def pytest_generate_tests(metafunc):
    if "flavour" in metafunc.fixturenames:
        metafunc.parametrize("flavour", ["serial", "lockstep"])


@pytest.fixture(autouse=True)
def _flavour(flavour, oracle_lib):
    global B
    B = BE.LOCKSTEP if flavour == "lockstep" else BE.SERIAL
    B.build()
    yield
    B = BE.SERIAL


def emu_gpu(case, path="small", **kw):
    """(the replay has one path per flavour: `path` of the GPU tests, util.PATHS, does not apply)"""
    return configure(case, lib_path=B.lib)


def replay(monkeypatch, module, **attrs):
    """Import a GPU test module with its `gpu` engine factory (and executables) pointed at the host replay."""
    m = importlib.import_module(module)
    if hasattr(m, "gpu"):
        monkeypatch.setattr(m, "gpu", emu_gpu)
    for k, v in attrs.items():
        monkeypatch.setattr(m, k, v)
    return m


SMALL = ["ab_gas", "cu_fcc", "gr_cu_ljc", "gr_cu_morsec", "gr_cu_ljc_simplified"]


@pytest.mark.parametrize("name", SMALL)
def test_step0_lists_forces_energies(monkeypatch, name):
    replay(monkeypatch, "test_parity_gpu").test_step0_lists_forces_energies(name, "small")


@pytest.mark.parametrize("name", ["ab_gas", "cu_fcc", "gr_cu_morsec"])
@pytest.mark.parametrize("integrator", ["nve", "nvt", "nvms"])
def test_trajectory_22_steps(monkeypatch, name, integrator):
    replay(monkeypatch, "test_parity_gpu").test_trajectory_22_steps(name, integrator, "small")


@pytest.mark.parametrize("fn", ["test_single_step_forces_after_move", "test_zero_momentum_and_invert_z", "test_too_many_neighbours_is_reported",
                                "test_particle_out_of_cell_is_reported", "test_not_enough_graphene_neighbours_is_reported",
                                "test_refuses_what_the_reference_gets_wrong_silently", "test_upload_restarts_a_context",
                                "test_multi_step_advance_equals_single_steps", "test_two_contexts_share_a_gpu",
                                "test_download_into_caller_buffers", "test_rjl_in_a_box_narrower_than_twice_R2"])
def test_parity_misc(monkeypatch, fn):
    getattr(replay(monkeypatch, "test_parity_gpu"), fn)()


@pytest.mark.parametrize("name", ["ab_gas", "cu_fcc", "gr_cu_ljc"])
def test_advance_logged_rows(monkeypatch, name):
    replay(monkeypatch, "test_parity_gpu").test_advance_logged_rows(name, "small")


@pytest.mark.parametrize("name", SMALL)
def test_golden_fixtures_and_in_step_energies(monkeypatch, name):
    m = replay(monkeypatch, "test_parity_gpu")
    m.test_gpu_matches_golden_fixtures(name, "small")
    m.test_energies_from_the_force_pass(name, "small")


@pytest.mark.parametrize("thermostat", [True, False])
def test_deposition(monkeypatch, thermostat):
    replay(monkeypatch, "test_deposition_gpu").test_deposition_matches_the_oracle(thermostat)


def test_deposition_misc(monkeypatch, tmp_path, oracle_lib):
    m = replay(monkeypatch, "test_deposition_gpu", EXE=B.exe)
    m.test_forces_outside_all_atoms_accumulate_like_the_reference()
    m.test_deposition_matches_the_golden_fixture()
    m.test_host_deposition_run_matches_the_cpu_port(tmp_path, None, oracle_lib)


def test_rebosc(monkeypatch, tmp_path, oracle_lib):
    m = replay(monkeypatch, "test_rebosc_gpu", EXE=B.exe)
    m.test_rebosc_energy_and_numerical_forces(0.04)
    m.test_rebosc_trajectory_and_energy_conservation()
    m.test_rebosc_feeds_the_graphene_normals_of_ljc()
    m.test_rebosc_matches_the_golden_fixture()
    m.test_host_rebosc_run_matches_the_cpu_port(tmp_path, None, oracle_lib)


@pytest.mark.parametrize("name", ["graphene", "ab_gas", "deposition", "cu_fcc"])
def test_restart_is_bit_identical(monkeypatch, tmp_path, name):
    replay(monkeypatch, "test_zz_restart_gpu", EXE=B.exe).test_restart_reproduces_the_interrupted_run_on_the_gpu(tmp_path, None, name)


def test_save_and_restore_state(monkeypatch):
    replay(monkeypatch, "test_zz_restart_gpu").test_save_and_restore_state_through_the_c_abi()


@pytest.mark.parametrize("which", ["ab_gas", "graphene"])
def test_host_outputs(monkeypatch, tmp_path, oracle_lib, which):
    replay(monkeypatch, "test_host_gpu", EXE=B.exe).test_same_outputs_as_the_cpu_reference_port(tmp_path, None, oracle_lib, which)


def test_host_queued_log(monkeypatch, tmp_path):
    replay(monkeypatch, "test_host_gpu", EXE=B.exe).test_queued_log_rows_are_the_stepwise_log(tmp_path, None)


def test_host_ensemble_ranks(monkeypatch, tmp_path):
    replay(monkeypatch, "test_host_gpu", EXE=B.exe).test_gpu_ensemble_ranks(tmp_path, None)


@pytest.mark.parametrize("extra", [(), ("-pair",)])
def test_fitting_rows(monkeypatch, tmp_path, oracle_lib, extra):
    replay(monkeypatch, "test_zz_fitting_gpu", EXE_FIT=B.exe_fit).test_fit_rows_match_the_cpu_port(tmp_path, None, oracle_lib, extra)


@pytest.mark.parametrize("fn", ["test_rjl_copper_cohesive_energy", "test_tb_graphite_sheet_energy", "test_lj_pair", "ljc", "morsec"])
def test_anchors(fn):
    """Closed forms and published figures (tests/test_zz_anchors.py) on this flavour of the replay."""
    import test_zz_anchors as A
    if fn in ("ljc", "morsec"):
        A.test_cosine_potentials_adatom_over_graphene(emu_gpu, fn)
    else:
        getattr(A, fn)(emu_gpu)


def test_pipelined_lj1g_variant(monkeypatch, flavour):
    """The pipelined lj1g kernel (default) against the plain one (PFMDS_LJ1G_PIPE=0) through the C ABI (the serial replay always takes the thread-per-atom kernels)."""
    if flavour == "lockstep":
        pytest.skip("512 atoms take the 8-lanes-per-atom kernel in the lock-step replay, as on the GPU: the variant is not selected")
    from util import oracle, rel_err
    case = inputs.lj_fluid(n_side=8, period=5)
    monkeypatch.setenv("PFMDS_LJ1G_PIPE", "0")
    a = emu_gpu(case)
    monkeypatch.setenv("PFMDS_LJ1G_PIPE", "1")
    b = emu_gpu(case)
    o = oracle(case)
    for e in (a, b, o):
        e.advance("nve", 0.5, 0, 1)
        e.advance("nve", 0.5, 1, 11)
    fa, fb, fo = a.download()[2], b.download()[2], o.download()[2]
    assert rel_err(fb, fo) < 1e-9 and rel_err(fb, fa) < 1e-11 and not np.array_equal(fa, fb)     # a different kernel did run
    assert np.abs(b.download()[0] - o.download()[0]).max() < 1e-10


@pytest.mark.parametrize("integ", ["nvt", "nve"])
def test_multi_step_graphs(monkeypatch, integ):
    """Steady-state steps replayed one, four (default) and seven per graph launch (capi.cu graph_run_ok; the replay records the captured
    launches as closures): the same bits, also when a logged advance cuts the runs short, and the oracle's trajectory."""
    from util import oracle
    case = inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=20)
    res = []
    for gs in ("1", "4", "7"):
        monkeypatch.setenv("PFMDS_GRAPH_STEPS", gs)
        monkeypatch.setenv("PFMDS_GRAPH_REBUILDS", "0" if gs == "1" else "1")   # rebuild steps kernel by kernel / replayed from their two graphs
        e = emu_gpu(case)
        e.advance(integ, 0.5, 0, 1)
        e.advance(integ, 0.5, 1, 45)
        rows = e.advance_logged(integ, 0.5, 46, 24, log_period=8)
        res.append((e.download(), [e.get_nhc(k) for k in range(len(case["nhc"]))], rows))
        e.close()
    (pa, va, fa), na, ra = res[0]
    for (pb, vb, fb), nb, rb in res[1:]:
        assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)
        assert all(np.array_equal(x, y) for ta, tb in zip(na, nb) for x, y in zip(ta, tb))
        assert all(np.array_equal(x, y) for x, y in zip(ra, rb))
    o = oracle(case)
    o.advance(integ, 0.5, 0, 1)
    o.advance(integ, 0.5, 1, 45)
    o.advance(integ, 0.5, 46, 24)
    assert np.abs(pa - o.download()[0]).max() < 1e-9


def test_replay_identifies_itself():
    import ctypes as C
    lib = C.CDLL(B.lib)
    lib.pfmds_version.restype = C.c_char_p
    assert b"HOST REPLAY" in lib.pfmds_version()
    err = (C.c_double * 4)()
    assert lib.pfmds_selftest_math(0, err) == 20            # no device: the device self-test refuses


def _dep_case(changes, **kw):
    case = inputs.lj_deposition(**kw)
    # 1 all, 2 substrate, 3 growing [S, D], 4 empty, 5 a second growing group fed by group 3
    case["groups"] = [["S", "D"], ["S", "#"], ["S", "D"], ["#", "#"], ["S", "D"]]
    case["changes"] = changes
    return case


@pytest.mark.parametrize("changes", [
    [(2, 3, 3, 40, 4), (3, 5, 6, 30, 2)],        # chained: group 5 follows group 3 until step 6, then grows twice as fast
    [(2, 3, -5, 40, 3)],                         # change_ts1 < 0: the `<= ts1` branch never runs, the group starts full and stays capped
    [(2, 3, 0, 1, 1)],                           # one atom at step 0, window closed at once
    [(2, 3, 2, 10 ** 6, 1), (2, 5, 100, 200, 1)],  # one atom per step; an entry that never leaves its first branch
])
def test_deposition_edge_cases_against_the_oracle(changes):
    from util import oracle, rel_err, neighbours
    case = _dep_case(changes, thermostat=False)
    case["interactions"][0]["lists"] = [(5, 3, 80, 7.5, 5)] if len(changes) > 1 and changes[1][1] == 5 and changes[1][2] == 6 else case["interactions"][0]["lists"]
    if case["interactions"][0]["lists"][0][0] == 5:
        case["interactions"][0] = dict(name="lj", file="parameters_LJ.txt", params=[0.0103, 3.405, 6.0, 7.0], lists=[(5, 3, 80, 7.5, 5), (3, 5, 80, 7.5, 5)])
    g, o = emu_gpu(case), oracle(case)
    s = 0
    for n in (1, 2, 4, 5, 9):
        g.advance("nve", 1.0, s, n)
        o.advance("nve", 1.0, s, n)
        s += n
        assert [g.group_size(k) for k in range(1, 6)] == [o.group_size(k) for k in range(1, 6)]
        assert np.abs(g.download()[0] - o.download()[0]).max() < 1e-10 and rel_err(g.download()[2], o.download()[2]) < 1e-9
        for j in range(len(case["interactions"][0]["lists"])):
            a, b = neighbours(g, case, 0, j), neighbours(o, case, 0, j)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_deposition_zero_frequency_is_an_error_like_the_reference():
    from pfmds_b200.engine import PfmdsError
    from util import oracle
    case = _dep_case([(2, 3, 1, 50, 0)], thermostat=False)
    for eng in (emu_gpu(case), oracle(case)):
        eng.advance("nve", 1.0, 0, 2)                      # steps 0, 1: still in the `<= ts1` branch
        with pytest.raises(PfmdsError):
            eng.advance("nve", 1.0, 2, 1)                  # mod(md_step - ts1, 0)

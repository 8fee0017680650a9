"""Kernel variants selected from the environment must give each other's results (the defaults were flipped to the measured
winners in round 2: pipelined lj1g, mask list build; the former defaults stay selectable).  Every variant is also driven against
the oracle by tests/test_parity_gpu.py (PATHS in util.py).  (Named zz: runs after the parity tests.)"""
import numpy as np
import pytest

from pfmds_b200 import inputs
from util import gpu, rel_err

pytestmark = pytest.mark.gpu


def test_pipelined_lj1g_kernel_matches_the_plain_kernel():
    case = inputs.lj_fluid(n_side=47, period=5)                    # 103 823 atoms: the thread-per-atom kernels
    a = gpu(case, {"PFMDS_LJ1G_PIPE": "0"})
    b = gpu(case)                                                  # default: k_lj1g_pipe
    for e in (a, b):
        e.advance("nve", 0.5, 0, 1, with_energy=True)
    fa, fb = a.download()[2], b.download()[2]
    assert np.abs(fa).max() > 1e-3 and rel_err(fb, fa) < 1e-12 and not np.array_equal(fa, fb)     # a different kernel did run
    assert np.allclose(a.energies()[0], b.energies()[0], rtol=1e-12, atol=0)
    for e in (a, b):
        e.advance("nve", 0.5, 1, 12)
    assert np.abs(a.download()[0] - b.download()[0]).max() < 1e-10
    assert np.allclose(a.energies()[0], b.energies()[0], rtol=1e-11, atol=0)     # energy_interaction keeps the plain kernel


def test_mask_list_build_gives_the_plain_rows():
    """k_build_mask (default) and k_build (PFMDS_NL_MASK=0): same candidates in the same order, so the rows, and with them every
    force bit, are identical."""
    for case, integ, dt in ((inputs.cu_fcc(ncell=37, jitter=0.03, period=5), "nvt", 2.0),            # 202 612 atoms (thread-per-atom build from 200 000 up), class-partitioned rows
                            (inputs.lj_fluid(n_side=60, period=5), "nve", 0.5)):                    # 216 000 atoms, half list (lessnnum)
        a = gpu(case, {"PFMDS_NL_MASK": "0"})
        b = gpu(case)
        for e in (a, b):
            e.advance(integ, dt, 0, 12)
        (pa, va, fa), (pb, vb, fb) = a.download(), b.download()
        assert np.abs(fa).max() > 1e-3 and np.array_equal(fa, fb) and np.array_equal(pa, pb) and np.array_equal(va, vb)
        assert a.pair_count(0, 0) == b.pair_count(0, 0) and np.array_equal(a.diagnostics()[4], b.diagnostics()[4])
        a.close()
        b.close()


def test_size_switches_are_read_from_the_environment():
    """PFMDS_SMALL_N / PFMDS_NL_WARP_N move a small system onto the large-system kernels: different summation order, so the
    forces differ in their last bits but agree to rounding; the neighbour sets are the same."""
    case = inputs.cu_fcc(ncell=6, jitter=0.05, period=5)
    a, b = gpu(case), gpu(case, "large")
    for e in (a, b):
        e.advance("nvt", 2.0, 0, 7)
    fa, fb = a.download()[2], b.download()[2]
    assert rel_err(fb, fa) < 1e-12 and not np.array_equal(fa, fb)
    assert a.pair_count(0, 0) == b.pair_count(0, 0)


def test_small_system_branches_give_the_sequential_results():
    """Small systems run their interactions as parallel branches of the step's CUDA graph (own stream and force buffer per
    interaction, capi.cu compute_forces) when the context has its GPU to itself.  Against PFMDS_SMALL_FORK=0 (one stream, one force
    array): the A/B gas has one addend per atom and list, so the bits are the same; graphene on Cu adds the ljc terms in another
    grouping (rounding only)."""
    for case, integ, dt, exact in ((inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=5), "nvt", 0.5, True),
                                   (inputs.graphene_on_cu_small(interface="ljc", period=5), "nvt", 1.0, False),
                                   (inputs.graphene_on_cu_small(interface="morsec", period=5), "nvms", 1.0, False)):
        res = []
        for env in ({}, {"PFMDS_SMALL_FORK": "0"}):
            import gc
            gc.collect()                           # engines of earlier tests still waiting for the collector would count as neighbours
            e = gpu(case, env)                     # alone on the device: the previous engine is closed
            assert e._lib.pfmds_live_contexts(0) == 1
            e.advance(integ, dt, 0, 1)
            e.advance(integ, dt, 1, 23)            # steady-state steps replay the graph with the forked branches
            res.append((e.download(), e.energies(), e.launch_count()))
            e.close()
        (pa, va, fa), ea, _ = res[0]
        (pb, vb, fb), eb, _ = res[1]
        assert np.abs(fa).max() > 1e-3
        if exact:
            assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)
        else:
            assert np.abs(pa - pb).max() < 1e-12 and rel_err(vb, va) < 1e-11 and rel_err(fb, fa) < 1e-11
        assert np.allclose(ea[0], eb[0], rtol=1e-11, atol=0)


def test_multi_step_graphs_give_the_single_step_results():
    """Runs of steady-state steps of small systems are replayed four steps per graph launch (PFMDS_GRAPH_STEPS, capi.cu
    graph_run_ok) and the steps that rebuild every list are replayed from graphs as well (one per half of the double-buffered state,
    PFMDS_GRAPH_REBUILDS): same kernels in the same order as one graph per step, so every bit must be the same -- positions,
    velocities, forces, thermostat chains and the rows of a logged advance whose log period cuts the runs short."""
    import gc
    for case, integ, dt in ((inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=20), "nvt", 0.5),
                            (inputs.graphene_on_cu_small(interface="ljc", period=10), "nvt", 1.0),
                            (inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=20), "nve", 0.5)):
        res = []
        # (first: one graph per steady-state step and the rebuild steps launched kernel by kernel, as in round 1)
        for env in ({"PFMDS_GRAPH_STEPS": "1", "PFMDS_GRAPH_REBUILDS": "0"}, {}, {"PFMDS_GRAPH_STEPS": "7"}):
            gc.collect()
            e = gpu(case, env)
            e.advance(integ, dt, 0, 1)
            e.advance(integ, dt, 1, 60)
            rows = e.advance_logged(integ, dt, 61, 40, log_period=8)
            res.append((e.download(), [e.get_nhc(k) for k in range(len(case["nhc"]))], rows))
            e.close()
        (pa, va, fa), na, ra = res[0]
        assert np.abs(fa).max() > 1e-3
        for (pb, vb, fb), nb, rb in res[1:]:
            assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)
            assert all(np.array_equal(x, y) for ta, tb in zip(na, nb) for x, y in zip(ta, tb))
            assert all(np.array_equal(x, y) for x, y in zip(ra, rb))


def test_lj1g_force_kernel_stores_instead_of_zero_plus_accumulate():
    """lj1g as the only / first interaction, owning every atom, on the thread-per-atom path: the pipelined kernel stores its result
    (zero_forces fused away, like k_rjl_force).  Against the oracle over rebuilds, and against the plain kernel + memset."""
    from util import oracle
    case = inputs.lj_fluid(n_side=10, period=5)
    g, p, o = gpu(case, "large"), gpu(case, "large_build"), oracle(case)
    for e in (g, p, o):
        e.advance("nve", 0.5, 0, 1)
    fo = o.download()[2]
    assert np.abs(fo).max() > 1e-3 and rel_err(g.download()[2], fo) < 1e-9 and rel_err(p.download()[2], fo) < 1e-9
    for e in (g, p, o):
        e.advance("nve", 0.5, 1, 17)
    (pg, vg, fg), (po, vo, fo) = g.download(), o.download()
    assert np.abs(pg - po).max() < 1e-9 and rel_err(fg, fo) < 1e-8 and rel_err(p.download()[2], fo) < 1e-8
    assert np.allclose(g.energies()[0], o.energies()[0], rtol=1e-9, atol=0)
    for e in (g, p, o):
        e.close()

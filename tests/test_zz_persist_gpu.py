"""Persistent step kernel of small systems (pfmds_b200/csrc/persist.cuh): runs of plain steps -- no list rebuild, no momentum
removal, no energies -- go through ONE cooperative launch whose phases are separated by grid barriers.  Every phase calls the
per-thread body of the kernel it replaces, so the results must be the step-by-step path's bit for bit (PFMDS_PERSIST=0), and
through that path the oracle's (tests/test_parity_gpu.py).  (Named zz: runs after the parity tests.)"""
import gc

import numpy as np
import pytest

from pfmds_b200 import inputs
from util import gpu, oracle, rel_err, RTOL

pytestmark = pytest.mark.gpu

CASES = {
    "ab_gas": (lambda: inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=5), 0.5),
    "gr_cu_ljc": (lambda: inputs.graphene_on_cu_small(interface="ljc", period=5), 1.0),
    "gr_cu_morsec": (lambda: inputs.graphene_on_cu_small(interface="morsec", period=5), 1.0),
    "gr_cu_ljc_simplified": (lambda: inputs.graphene_on_cu_small(interface="ljc", period=5, simplified=True), 1.0),
}


def _run(case, env, plan):
    gc.collect()   # engines of earlier tests still waiting for the collector would count as neighbours on the device
    e = gpu(case, env)
    assert e._lib.pfmds_live_contexts(0) == 1
    out = []
    for integ, dt, first, n in plan:
        e.advance(integ, dt, first, n)
        out.append((e.download(), [e.get_nhc(k) for k in range(len(case["nhc"]))]))
    res = (out, e.energies(), e.launch_count())
    e.close()
    return res


@pytest.mark.parametrize("name", sorted(CASES))
def test_persistent_runs_are_bit_identical_to_the_step_by_step_path(name):
    make, dt = CASES[name]
    case = make()
    # nvt: step 0, then 23 steps with rebuilds at 5, 10, 15, 20 (persistent runs 2-4, 6-9, 11-14, 16-19, 21-23); nve afterwards: the
    # first nve step flushes the pending thermostat scale on the step-by-step path, the rest of the call is persistent again
    plan = [("nvt", dt, 0, 1), ("nvt", dt, 1, 23), ("nve", dt, 24, 17), ("nvt", dt, 41, 9)]
    (a, ea, la), (b, eb, lb) = _run(case, {"PFMDS_PERSIST": "0"}, plan), _run(case, {}, plan)
    assert lb < la, "the persistent kernel did not run (launch counts %d / %d)" % (lb, la)
    for ((pa, va, fa), na), ((pb, vb, fb), nb) in zip(a, b):
        assert np.abs(fa).max() > 1e-3
        assert np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)
        for ta, tb in zip(na, nb):
            assert all(np.array_equal(x, y) for x, y in zip(ta, tb))
    assert np.array_equal(ea[0], eb[0]) and ea[1] == eb[1] and np.array_equal(ea[3], eb[3])


def test_persistent_runs_inside_a_logged_advance():
    """pfmds_advance_logged with log_period 4: the unlogged steps between two rows are persistent runs; rows and state as without."""
    case = inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=10)
    res = []
    for env in ({"PFMDS_PERSIST": "0"}, {}):
        gc.collect()
        e = gpu(case, env)
        e.advance("nvt", 0.5, 0, 1)
        rows = e.advance_logged("nvt", 0.5, 1, 40, log_period=4)
        res.append((np.column_stack([np.asarray(r).reshape(len(rows[1]), -1) for r in rows]), e.download(), e.launch_count()))
        e.close()
    (ra, (pa, va, fa), la), (rb, (pb, vb, fb), lb) = res
    assert lb < la and ra.shape == rb.shape and ra.shape[0] == 10
    assert np.array_equal(ra, rb) and np.array_equal(pa, pb) and np.array_equal(va, vb) and np.array_equal(fa, fb)


def test_persistent_trajectory_against_the_oracle():
    """The same 22-step trajectory the parity tests run, here with the persistent kernel in use, against the CPU oracle."""
    case = inputs.graphene_on_cu_small(interface="ljc", period=5)
    gc.collect()
    g, o = gpu(case), oracle(case)
    for e in (g, o):
        e.advance("nvt", 1.0, 0, 1)
        e.advance("nvt", 1.0, 1, 21)
    (pg, vg, fg), (po, vo, fo) = g.download(), o.download()
    assert np.abs(pg - po).max() < 1e-9 and rel_err(vg, vo) < 1e-8 and rel_err(fg, fo) < 1e-7
    assert np.allclose(g.energies()[0], o.energies()[0], rtol=1e-8, atol=0)
    g.close()
    o.close()


def test_out_of_cell_inside_a_persistent_run_is_reported():
    """The position test of the opening kick+drift runs inside the persistent kernel too: an atom pushed out of the cell is reported
    with the reference's message at the next synchronising call."""
    from pfmds_b200.engine import PfmdsError
    case = inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=50)
    case["vel"] = np.array(case["vel"], float)
    case["vel"][3] = (0.0, 0.0, 9.0e3)      # crosses the box within a few steps (positions are wrapped once per step only)
    gc.collect()
    e = gpu(case)
    e.advance("nve", 0.5, 0, 1)
    with pytest.raises(PfmdsError) as err:
        e.advance("nve", 0.5, 1, 40)
        e.synchronize()
    assert "out of cell" in str(err.value)
    e.close()

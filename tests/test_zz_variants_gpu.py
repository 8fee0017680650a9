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


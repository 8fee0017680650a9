"""Opt-in kernel variants that have not been timed on hardware yet: they must give the default kernels'
results.  (Named zz: runs after the parity tests.)"""
import os

import numpy as np
import pytest

from pfmds_b200 import inputs
from util import gpu, rel_err

pytestmark = pytest.mark.gpu


def test_pipelined_lj1g_kernel_matches_the_default_kernel():
    case = inputs.lj_fluid(n_side=47, period=5)                    # 103 823 atoms: the thread-per-atom kernels
    a = gpu(case)
    os.environ["PFMDS_LJ1G_PIPE"] = "1"
    try:
        b = gpu(case)
    finally:
        del os.environ["PFMDS_LJ1G_PIPE"]
    for e in (a, b):
        e.advance("nve", 0.5, 0, 1, with_energy=True)
    fa, fb = a.download()[2], b.download()[2]
    assert np.abs(fa).max() > 1e-3 and rel_err(fb, fa) < 1e-12
    assert np.allclose(a.energies()[0], b.energies()[0], rtol=1e-12, atol=0)
    for e in (a, b):
        e.advance("nve", 0.5, 1, 12)
    assert np.abs(a.download()[0] - b.download()[0]).max() < 1e-10
    assert np.allclose(a.energies()[0], b.energies()[0], rtol=1e-11, atol=0)     # energy_interaction keeps the default kernel


def test_rjl_force_kernel_at_five_blocks_per_sm_gives_the_same_bits():
    """PFMDS_RJL_MINB=5: the second-generation force kernel compiled for 5 instead of 7 blocks per SM (more registers, no constant
    reloads in the pair loop).  Same source, same operations in the same order: identical forces and trajectory, bit for bit."""
    case = inputs.cu_fcc(ncell=30, jitter=0.03, period=5)              # 108 000 atoms: the thread-per-atom kernels
    a = gpu(case)
    os.environ["PFMDS_RJL_MINB"] = "5"
    try:
        b = gpu(case)
    finally:
        del os.environ["PFMDS_RJL_MINB"]
    for e in (a, b):
        e.advance("nvt", 2.0, 0, 12)
    (pa, va, fa), (pb, vb, fb) = a.download(), b.download()
    assert np.abs(fa).max() > 0.05 and np.array_equal(fa, fb) and np.array_equal(pa, pb) and np.array_equal(va, vb)


def test_mask_list_build_gives_the_default_rows():
    """PFMDS_NL_MASK=1: thread-per-atom list build with the FP32 prefilter and the exact test in separate loops (nl.cu k_build_mask).
    Same candidates in the same order: the rows, and with them every force bit, are those of k_build."""
    for case, integ, dt in ((inputs.cu_fcc(ncell=37, jitter=0.03, period=5), "nvt", 2.0),            # 202 612 atoms (thread-per-atom build from 200 000 up), class-partitioned rows
                            (inputs.lj_fluid(n_side=60, period=5), "nve", 0.5)):                    # 216 000 atoms, half list (lessnnum)
        a = gpu(case)
        os.environ["PFMDS_NL_MASK"] = "1"
        try:
            b = gpu(case)
        finally:
            del os.environ["PFMDS_NL_MASK"]
        for e in (a, b):
            e.advance(integ, dt, 0, 12)
        (pa, va, fa), (pb, vb, fb) = a.download(), b.download()
        assert np.abs(fa).max() > 1e-3 and np.array_equal(fa, fb) and np.array_equal(pa, pb) and np.array_equal(va, vb)
        assert a.pair_count(0, 0) == b.pair_count(0, 0) and np.array_equal(a.diagnostics()[4], b.diagnostics()[4])
        a.close()
        b.close()

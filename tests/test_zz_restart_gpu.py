"""SURVEY.md 8(f) row 4 on the device: a run continued from a checkpoint (pfmds_save_state / pfmds_restore_state behind
`-checkpoint_period` / `-restart`) writes the same final xyz, digit for digit, and the same log rows as the uninterrupted run —
the cell order is keyed by atom identity, so list rows and reductions do not depend on the history of the slots.
(Named zz: it runs after the parity tests.)"""
import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.build import EXE
from test_restart import check_restart, restart_cases
from util import gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["graphene", "ab_gas", "deposition", "cu_fcc"])
def test_restart_reproduces_the_interrupted_run_on_the_gpu(tmp_path, cuda_lib, name):
    if name == "cu_fcc":
        case = inputs.cu_fcc(ncell=5, steps=50, period=5, jitter=0.03)
        case["integrators"] = [("nvt", 2.0, 50, 10 ** 9, 5)]
    else:
        case = restart_cases()[name]
    check_restart(EXE, tmp_path, case, at=40 if name == "ab_gas" else 20)


def test_save_and_restore_state_through_the_c_abi():
    """Same thing without the host program: state of one context moved into a fresh one."""
    import ctypes as C
    case = inputs.cu_fcc(ncell=5, period=5, jitter=0.03)
    a = gpu(case)
    a.advance("nvt", 2.0, 0, 21)                    # steps 0..20; 20 is a rebuild step
    n = C.c_longlong()
    lib = a._lib
    lib.pfmds_state_size.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
    lib.pfmds_save_state.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.pfmds_restore_state.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    assert lib.pfmds_state_size(a._ctx, C.byref(n)) == 0 and n.value == 4 + 2 + 1 + 3 * 3 + 4
    blob = np.zeros(n.value)
    dp = C.POINTER(C.c_double)
    assert lib.pfmds_save_state(a._ctx, blob.ctypes.data_as(dp)) == 0
    pos, vel, frc = a.download()
    b = gpu(case)
    p, v = np.ascontiguousarray(pos.reshape(-1)), np.ascontiguousarray(vel.reshape(-1))
    assert lib.pfmds_restore_state(b._ctx, p.ctypes.data_as(dp), v.ctypes.data_as(dp), blob.ctypes.data_as(dp)) == 0
    pb, vb, fb = b.download()
    assert np.array_equal(pb, pos) and np.array_equal(vb, vel) and np.array_equal(fb, frc)      # forces recomputed: same bits
    a.advance("nvt", 2.0, 21, 14)
    b.advance("nvt", 2.0, 21, 14)
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x, y)
    assert np.array_equal(a.get_nhc(0)[0], b.get_nhc(0)[0]) and np.array_equal(a.get_nhc(0)[1], b.get_nhc(0)[1])

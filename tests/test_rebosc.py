"""SURVEY.md 8(f) row 2 — `rebosc` (REBOsolidcarbon.f90) and the numerical-force engine (md_interactions.f90:273-425).
CPU side: (1) the oracle's line-by-line restatement (truncated lists, +-dx shifts) against independent central differences
of the whole-system energy; (2) the DEVICE algorithm (pfmds_b200/csrc/rebosc_core.cuh: one thread per (atom, axis), only the
pair terms that feel the shifted atom) compiled for the host and run thread by thread on the device's ELL list layout,
against the oracle.  The GPU parity tests proper are in tests/test_rebosc_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from util import oracle, neighbours

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# The reference's own forces carry the rounding noise of E(-dx) - E(+dx): eps * |E_cluster| / (2 dx) ~ 6e-14 / 2e-6 = 3e-8 eV/A.
FD_NOISE = 2e-7


@pytest.fixture(scope="module")
def host_kernels(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("reb") / "librebhost.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"), "-o", out, os.path.join(ROOT, "tests", "rebosc_host.cpp")], check=True)
    L = C.CDLL(out)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.reb_host_run.argtypes = [C.c_int, dp, dp, C.c_size_t, ip, ip, dp, dp, dp, ip]
    return L


def ell_list(pos, box, rcut, maxn, order):
    """The device layout: nlist[p*stride + i] with slot numbers; `order` permutes atoms into slots (the cell re-sort)."""
    n = len(pos)
    p = pos[order]
    d = p[None, :, :] - p[:, None, :]
    d -= box * np.round(d / box)
    r2 = (d ** 2).sum(-1)
    stride = (n + 31) // 32 * 32
    nlist = np.zeros((maxn, stride), np.int32)
    nnum = np.zeros(stride, np.int32)
    rng = np.random.default_rng(0)
    for i in range(n):
        js = np.where((r2[i] < rcut * rcut) & (np.arange(n) != i))[0]
        js = rng.permutation(js)              # row order is arbitrary on the device (cell / class order)
        nnum[i] = len(js)
        nlist[: len(js), i] = js
    return p, nlist, nnum, stride


def run_host_kernels(L, case, seed=1):
    n = len(case["mass"])
    order = np.random.default_rng(seed).permutation(n)
    it = case["interactions"][0]
    p, nlist, nnum, stride = ell_list(case["pos"], case["box"], it["lists"][0][3], it["lists"][0][2], order)
    pos4 = np.zeros((n, 4)); pos4[:, :3] = p
    frc4 = np.zeros((n, 4))
    prm = np.ascontiguousarray(it["params"], np.float64)
    box = np.ascontiguousarray(case["box"], np.float64)
    e = C.c_double()
    err = np.zeros(4, np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.reb_host_run(n, pos4.ctypes.data_as(dp), frc4.ctypes.data_as(dp), stride, nlist.ctypes.data_as(ip), nnum.ctypes.data_as(ip),
                   prm.ctypes.data_as(dp), box.ctypes.data_as(dp), C.byref(e), err.ctypes.data_as(ip))
    assert err[0] == 0
    f = np.zeros((n, 3))
    f[order] = frc4[:, :3]
    return e.value, f


def test_oracle_numerical_forces_are_the_energy_gradient(oracle_lib):
    case = inputs.graphene_rebosc()
    e = oracle(case)
    e.advance("nve", 0.5, 0, 1)
    f = e.download()[2]
    assert np.abs(f.sum(0)).max() < 1e-6 and np.abs(f).max() > 1.0
    h = 1e-5
    for i, k in ((0, 0), (7, 1), (19, 2), (33, 2), (47, 0)):
        es = []
        for sgn in (1, -1):
            c2 = dict(case)
            c2["pos"] = case["pos"].copy()
            c2["pos"][i, k] += sgn * h
            o = oracle(c2)
            o.advance("nve", 0.5, 0, 1)
            es.append(o.energies()[0][0])
        assert abs(f[i, k] + (es[0] - es[1]) / (2 * h)) < 5e-7


def test_oracle_nve_energy_is_conserved_with_numerical_forces(oracle_lib):
    """Total energy only wobbles at the O(dt^2) level of velocity Verlet: halving dt quarters the spread (forces consistent with E)."""
    spread = {}
    for dt in (0.5, 0.25):
        e = oracle(inputs.graphene_rebosc())
        tot = []
        per = int(round(5 / dt))
        for s in range(0, 5):
            e.advance("nve", dt, 0 if s == 0 else (s - 1) * per + 1, 1 if s == 0 else per)
            en = e.energies()
            tot.append(en[0].sum() + en[1])
        spread[dt] = max(tot) - min(tot)
    assert spread[0.5] < 2e-2 and spread[0.25] < 0.4 * spread[0.5]


@pytest.mark.parametrize("jitter,seed", [(0.04, 5), (0.12, 6)])
def test_device_algorithm_on_the_host_matches_the_oracle(oracle_lib, host_kernels, jitter, seed):
    case = inputs.graphene_rebosc(cells=(5, 3), jitter=jitter, seed=seed)
    o = oracle(case)
    o.advance("nve", 0.5, 0, 1)
    fo = o.download()[2]
    eo = o.energies()[0][0]
    eh, fh = run_host_kernels(host_kernels, case)
    assert abs(eh - eo) < 1e-12 * abs(eo)
    assert np.abs(fh - fo).max() < FD_NOISE * np.abs(fo).max()
    # a different slot permutation and row order must not matter beyond rounding
    eh2, fh2 = run_host_kernels(host_kernels, case, seed=2)
    assert abs(eh2 - eh) < 1e-12 * abs(eo) and np.abs(fh2 - fh).max() < FD_NOISE * np.abs(fo).max()


def test_device_algorithm_with_bonds_inside_the_switching_zone(oracle_lib, host_kernels):
    """Stretch the sheet so that bonds sit between R1 and R2 (f_c and its neighbours' f_c all active)."""
    case = inputs.graphene_rebosc(cells=(4, 3), jitter=0.05, seed=9)
    s = 1.27                                   # 1.42 A -> 1.80 A: inside [R1, R2) = [1.7, 2.0)
    case["pos"][:, :2] *= s
    case["box"] = case["box"] * np.array([s, s, 1.0])
    o = oracle(case)
    o.advance("nve", 0.5, 0, 1)
    fo, eo = o.download()[2], o.energies()[0][0]
    eh, fh = run_host_kernels(host_kernels, case)
    assert abs(eh - eo) < 1e-12 * abs(eo)
    assert np.abs(fh - fo).max() < FD_NOISE * max(np.abs(fo).max(), 1.0)


def test_host_reads_the_rebosc_parameter_file(tmp_path, oracle_lib):
    from conftest import ORACLE_EXE
    case = inputs.graphene_rebosc(steps=10)
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    r = subprocess.run([ORACLE_EXE, "-ipath", d, "-p", d + "x_", "-op", "5", "-omp_n", "2"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "rebosc" in r.stdout and os.path.exists(d + "x_final_init.xyz")

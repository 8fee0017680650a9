"""The device kernels of pfmds_b200/csrc/forces.cu compiled for the host and run thread by thread (tests/emu/host_emu.hpp,
tests/forces_host.cpp) on the device's data layout, against the CPU oracle: forces and energies of every interaction kind at
step 0.  Covers the kernels' arithmetic and indexing (pipelined rjl loops, class-free row order, converse lists, tb per-slot
parts, graphene normals and the normal-derivative term) without a GPU; the GPU parity tests (-m gpu) remain the gate for the
compiled SASS.  Only the thread-per-atom variants can be emulated (the 8-lanes-per-atom variants shuffle between lanes)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pfmds_b200 import inputs
from util import RTOL, oracle, rel_err, small_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP, IP = C.POINTER(C.c_double), C.POINTER(C.c_int)


@pytest.fixture(scope="module")
def kernels(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fh") / "libforces_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"), "-o", out, os.path.join(ROOT, "tests", "forces_host.cpp")], check=True)
    return C.CDLL(out)


class DeviceLayout:
    """Slots, masks and ELL lists the way the library holds them (random slot permutation, random row order)."""

    def __init__(self, case, seed=1):
        self.case = case
        self.n = len(case["mass"])
        rng = np.random.default_rng(seed)
        self.order = rng.permutation(self.n)              # slot -> file index
        self.slot_of = np.empty(self.n, int)
        self.slot_of[self.order] = np.arange(self.n)
        self.stride = (self.n + 31) // 32 * 32
        self.pos4 = np.zeros((self.n, 4))
        self.pos4[:, :3] = case["pos"][self.order]
        self.box = np.ascontiguousarray(case["box"], np.float64)
        p = self.pos4[:, :3]
        d = p[None, :, :] - p[:, None, :]
        h = self.box / 2
        d = np.where(d >= h, d - self.box, np.where(d < -h, d + self.box, d))       # min_image of common.cuh
        self.r2 = (d ** 2).sum(-1)
        self.rng = rng

    def members(self, g):
        m = np.zeros(self.n, bool)
        m[self.slot_of[inputs.group_indexes(self.case, g) - 1]] = True
        return m

    def ell(self, g1, g2, maxn, rcut):
        own, par = self.members(g1), self.members(g2)
        nlist = np.zeros((maxn + 2, self.stride), np.int32)   # two spare rows of valid slot numbers, as capi.cu allocates them
        nnum = np.zeros(self.stride, np.int32)
        for i in np.where(own)[0]:
            js = np.where(par & (self.r2[i] < rcut * rcut) & (np.arange(self.n) != i))[0]
            js = self.rng.permutation(js)
            assert len(js) <= maxn
            nnum[i] = len(js)
            nlist[: len(js), i] = js
        return nlist, nnum

    def nearest3(self, src, rc_nn):
        """graphenenorm.f90:8-36 / nl.cu k_nearest3: the entries of the carbon list closer than r_cut_nn, in row order."""
        nlist, nnum = src
        nn = np.zeros((3, self.stride), np.int32)
        cnt = np.zeros(self.stride, np.int32)
        for i in range(self.n):
            k = 0
            for p in range(nnum[i]):
                j = nlist[p, i]
                if np.sqrt(self.r2[i, j]) < rc_nn:
                    nn[k, i] = j
                    k += 1
            assert k in (0, 3)
            cnt[i] = k
        return nn, cnt

    def back(self, a4):
        out = np.zeros((self.n, 3))
        out[self.order] = a4[:, :3]
        return out


def ptr(a, t):
    return a.ctypes.data_as(t)


def run_case(L, case, seed=1, rjl_overwrite_first=False, lj1g_pipe=False, rjl_gen=2, ran=None, rjl_energy_in_force=False):
    lay = DeviceLayout(case, seed)
    frc4 = np.zeros((lay.n, 4))
    energies = []
    carbon_list = None
    e = C.c_double()
    for k, it in enumerate(case["interactions"]):
        prm = np.ascontiguousarray(it["params"], np.float64)
        lists = it["lists"]
        if it["name"] == "lj":
            l0 = lay.ell(lists[0][0], lists[0][1], lists[0][2], lists[0][3])
            l1 = lay.ell(lists[0][1], lists[0][0], lists[1][2], lists[0][3])         # converse: swapped groups, r_cut of line 1, capacity of line 2
            L.fh_lj(lay.n, ptr(lay.pos4, DP), ptr(frc4, DP), C.c_size_t(lay.stride), ptr(l0[0], IP), ptr(l0[1], IP), ptr(l1[0], IP), ptr(l1[1], IP),
                    ptr(prm, DP), ptr(lay.box, DP), C.byref(e))
        elif it["name"] == "lj1g":
            l0 = lay.ell(*lists[0][:4])
            L.fh_lj1g(lay.n, ptr(lay.pos4, DP), ptr(frc4, DP), C.c_size_t(lay.stride), ptr(l0[0], IP), ptr(l0[1], IP), ptr(prm, DP), ptr(lay.box, DP), C.byref(e),
                      int(lj1g_pipe))
        elif it["name"] == "rjl":
            l0 = lay.ell(*lists[0][:4])
            g = L.fh_rjl(lay.n, ptr(lay.pos4, DP), ptr(frc4, DP), C.c_size_t(lay.stride), ptr(l0[0], IP), ptr(l0[1], IP), ptr(prm, DP), ptr(lay.box, DP),
                     int(rjl_overwrite_first and k == 0) + (2 if rjl_energy_in_force else 0), C.byref(e), int(rjl_gen))
            if ran is not None:
                ran.append(g)
        elif it["name"] == "tb":
            l0 = lay.ell(*lists[0][:4])
            carbon_list = l0
            L.fh_tb(lay.n, ptr(lay.pos4, DP), ptr(frc4, DP), C.c_size_t(lay.stride), lists[0][2], ptr(l0[0], IP), ptr(l0[1], IP), ptr(prm, DP), ptr(lay.box, DP), C.byref(e))
        elif it["name"] in ("ljc", "morsec"):
            l0 = lay.ell(lists[0][0], lists[0][1], lists[0][2], lists[0][3])
            l1 = lay.ell(lists[0][1], lists[0][0], lists[1][2], lists[0][3])
            l2 = lay.nearest3(carbon_list, lists[2][3])
            gnorm = np.zeros((lay.stride, 4))
            L.fh_cos(int(it["name"] == "morsec"), lay.n, ptr(lay.pos4, DP), ptr(frc4, DP), C.c_size_t(lay.stride), ptr(l0[0], IP), ptr(l0[1], IP),
                     ptr(l1[0], IP), ptr(l1[1], IP), ptr(l2[0], IP), ptr(l2[1], IP), ptr(prm, DP), ptr(lay.box, DP), ptr(gnorm, DP), C.byref(e))
        else:
            raise AssertionError(it["name"])
        energies.append(e.value)
    return lay.back(frc4), np.array(energies)


@pytest.mark.parametrize("name", list(small_cases()))
def test_emulated_kernels_match_the_oracle(oracle_lib, kernels, name):
    case = small_cases()[name]
    o = oracle(case)
    integ, dt = case["integrators"][0][0], case["integrators"][0][1]
    o.advance(integ, dt, 0, 1)
    fo, eo = o.download()[2], o.energies()[0]
    fh, eh = run_case(kernels, case)
    assert rel_err(fh, fo) < RTOL
    assert rel_err(eh, eo) < RTOL
    fh2, eh2 = run_case(kernels, case, seed=7)               # another slot permutation / row order: only rounding changes
    assert rel_err(fh2, fh) < 1e-12 and rel_err(eh2, eh) < 1e-12


def test_emulated_lj1g_pipelined_variant(oracle_lib, kernels):
    """The opt-in pipelined lj1g kernel (PFMDS_LJ1G_PIPE=1): same sums with 1/r^2 and sqrt from mathx.cuh; oracle parity and
    agreement with the default kernel far inside the 1e-9 bar.  Odd and even row lengths, rows of one entry, the switch zone."""
    for case in (small_cases()["ab_gas"], inputs.lj_fluid(n_side=7, period=5), inputs.lj_deposition()):
        o = oracle(case)
        o.advance(case["integrators"][0][0], case["integrators"][0][1], 0, 1)
        fo, eo = o.download()[2], o.energies()[0]
        if case.get("changes"):
            case = dict(case, groups=[["S", "D"], ["S", "#"], ["S", "#"], ["#", "#"]])     # the emulated layout has no group%N: group 3 = substrate at step 0
        fa, ea = run_case(kernels, case)
        fb, eb = run_case(kernels, case, lj1g_pipe=True)
        assert rel_err(fb, fo) < RTOL and rel_err(eb, eo) < RTOL
        assert rel_err(fb, fa) < 1e-13 and rel_err(eb, ea) < 1e-13


def test_emulated_rjl_store_variant_is_bitwise_the_accumulate_variant(kernels):
    case = small_cases()["cu_fcc"]
    a, ea = run_case(kernels, case)
    b, eb = run_case(kernels, case, rjl_overwrite_first=True)
    assert np.array_equal(a, b) and np.array_equal(ea, eb)


def test_emulated_rjl_generations(oracle_lib, kernels):
    """Second-generation rjl pair routines (default) against the oracle and against the first generation: jittered crystal with
    pairs in all three classes (r < R1, switch zone, beyond R2), boundary atoms whose pairs need the minimum image, rows of odd
    and even length; a parameter set the second generation refuses (R2 beyond the half box) falls back to the first."""
    cases = [small_cases()["cu_fcc"], inputs.cu_fcc(ncell=4, jitter=0.25, seed=5, period=5), inputs.cu_fcc(cells=(5, 4, 6), jitter=0.3, seed=9, period=5)]
    for case in cases:
        o = oracle(case)
        o.advance("nvt", case["integrators"][0][1], 0, 1)
        fo, eo = o.download()[2], o.energies()[0]
        ran = []
        f2, e2 = run_case(kernels, case, ran=ran)
        f1, e1 = run_case(kernels, case, rjl_gen=1, ran=ran)
        f3, e3 = run_case(kernels, case, rjl_gen=3, ran=ran)                                # third generation: node-table exponentials
        f3e, e3e = run_case(kernels, case, rjl_gen=3, ran=ran, rjl_energy_in_force=True)    # ... and its force pass that also yields the energy
        assert ran == [2, 1, 3, 3]
        assert rel_err(f2, fo) < RTOL and rel_err(e2, eo) < RTOL
        assert rel_err(f1, fo) < RTOL and rel_err(e1, eo) < RTOL
        assert rel_err(f3, fo) < 1e-10 and rel_err(e3, eo) < 1e-10
        assert rel_err(f2, f1) < 5e-11 and rel_err(e2, e1) < 5e-11   # rsqrt_q: host stand-in seed off by up to 1.9e-6 -> 5e-12 in r (device 9e-7 -> 1.3e-12), times |exponent| <= 15
        assert rel_err(f3, f2) < 5e-13 and rel_err(e3, e2) < 5e-13   # same r, same switch: only the exponentials differ (table + degree-4 Taylor: 1.6e-14 each)
        assert np.array_equal(f3e, f3) and rel_err(e3e, e3) < 1e-13
    small = inputs.cu_fcc(ncell=3, jitter=0.2, seed=3, period=5)       # box 10.8 A: half box 5.42 < R2 = 6.0
    ran = []
    f, e = run_case(kernels, small, ran=ran)
    assert ran == [1]


# ---- neighbour-list build (nl.cu) ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def nl_kernels(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("nh") / "libnl_host.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "emu"), "-o", out, os.path.join(ROOT, "tests", "nl_host.cpp")], check=True)
    return C.CDLL(out)


def emulated_build(L, case, k, j, partition=None, reorder=True, seed=3):
    """Rows of list j of interaction k as {file index of owner: [file indexes of partners in row order]} plus the raw outputs."""
    it = case["interactions"][k]
    g1, g2, maxn, rcut, _ = it["lists"][j]
    if j == 1 and it["name"] in ("lj", "ljc", "morsec"):          # converse list: swapped groups, r_cut of line 1, capacity of line 2
        g1, g2, rcut = it["lists"][0][1], it["lists"][0][0], it["lists"][0][3]
    n = len(case["mass"])
    rng = np.random.default_rng(seed)
    order = rng.permutation(n)                                     # initial slot -> file index
    pos4 = np.zeros((n, 4)); pos4[:, :3] = case["pos"][order]
    gm = np.zeros(n, np.uint32)
    for g in range(1, len(case["groups"]) + 1):
        idx = inputs.group_indexes(case, g) - 1
        mask = np.zeros(n, bool); mask[idx] = True
        gm[mask[order]] |= np.uint32(1 << (g - 1))
    rc_max = max(l[3] for i2 in case["interactions"] for l in i2["lists"])
    stride = (n + 31) // 32 * 32
    slot_orig = np.zeros(n, np.int32)
    nlist = np.zeros((maxn, stride), np.int32)
    nnum = np.zeros(stride, np.int32)
    err = np.zeros(4, np.int32)
    ncell = np.zeros(3, np.int32)
    r1, r2 = partition if partition else (0.0, 0.0)
    box = np.ascontiguousarray(case["box"], np.float64)
    UP = C.POINTER(C.c_uint)
    L.nh_build.argtypes = [C.c_int, DP, UP, IP, DP, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, IP, IP, IP, IP, IP]
    pf_on = L.nh_build(n, ptr(pos4, DP), gm.ctypes.data_as(UP), ptr(order.astype(np.int32), IP), ptr(box, DP), rc_max, g1, g2, maxn, rcut,
                       int(partition is not None), r1, r2, int(reorder), ptr(slot_orig, IP), ptr(nlist, IP), ptr(nnum, IP), ptr(err, IP), ptr(ncell, IP))
    rows = {}
    for s in range(n):
        if nnum[s] or (gm[np.where(order == slot_orig[s])[0][0]] >> (g1 - 1)) & 1:
            rows[int(slot_orig[s])] = [int(slot_orig[t]) for t in nlist[: nnum[s], s]]
    return rows, err, ncell, pf_on, slot_orig


def oracle_rows(o, case, k, j):
    from util import neighbours, rows_of
    it = case["interactions"][k]
    g1, g2 = it["lists"][j][0], it["lists"][j][1]
    if j == 1 and it["name"] in ("lj", "ljc", "morsec"):
        g1, g2 = it["lists"][0][1], it["lists"][0][0]
    G1, G2 = inputs.group_indexes(case, g1) - 1, inputs.group_indexes(case, g2) - 1
    nlist, nnum, _ = neighbours(o, case, k, j)
    return {int(G1[r]): [int(G2[q - 1]) for q in nlist[r, : nnum[r]]] for r in range(len(G1))}


@pytest.mark.parametrize("name", ["ab_gas", "cu_fcc", "gr_cu_ljc"])
@pytest.mark.parametrize("reorder", [True, False])
def test_emulated_list_build_is_bit_exact(oracle_lib, nl_kernels, name, reorder):
    """Same neighbour SETS as the reference's brute-force search for every list of the case (cell grid, FP32 prefilter with its
    margin, exact FP64 test, group masks), with and without the physical re-sort."""
    case = small_cases()[name]
    o = oracle(case)
    o.advance(case["integrators"][0][0], case["integrators"][0][1], 0, 1)
    for k, it in enumerate(case["interactions"]):
        for j in range(len(it["lists"])):
            if it["name"] in ("ljc", "morsec") and j == 2:
                continue                                            # nearest-three list: derived from the tb list (k_nearest3), not searched
            rows, err, ncell, pf_on, _ = emulated_build(nl_kernels, case, k, j, reorder=reorder)
            want = oracle_rows(o, case, k, j)
            assert err[0] == 0
            assert set(rows) == set(want)
            for a in want:
                assert sorted(rows[a]) == sorted(want[a]), (k, j, a)


def test_emulated_build_partitions_rows_by_class(nl_kernels):
    case = small_cases()["cu_fcc"]
    R1, R2 = case["interactions"][0]["params"][5:7]
    rows, err, _, _, _ = emulated_build(nl_kernels, case, 0, 0, partition=(R1, R2))
    plain, _, _, _, _ = emulated_build(nl_kernels, case, 0, 0)
    p, box = case["pos"], case["box"]
    for a, js in rows.items():
        assert sorted(js) == sorted(plain[a])                       # membership untouched
        d = p[js] - p[a]
        d -= box * np.round(d / box)
        r = np.sqrt((d ** 2).sum(1))
        cls = np.where(r < R1, 0, np.where(r < R2, 1, 2))
        assert (np.diff(cls) >= 0).all()                            # [r < R1 | switch zone | beyond R2]
    assert any(len(set(np.where(np.sqrt((((p[js] - p[a]) - box * np.round((p[js] - p[a]) / box)) ** 2).sum(1)) < R1, 0, 1))) > 1 for a, js in rows.items())


def test_emulated_build_in_a_box_of_two_cells_and_overflow(oracle_lib, nl_kernels):
    """Boxes under three cells wide visit each neighbour cell once; a box too small for the FP32 wrap switches the prefilter off."""
    case = inputs.cu_fcc(ncell=4, jitter=0.05, period=5)            # 14.46 A box, r_cut 6.5: 2 x 2 x 2 cells
    o = oracle(case)
    o.advance("nvt", 2.0, 0, 1)
    rows, err, ncell, pf_on, _ = emulated_build(nl_kernels, case, 0, 0)
    want = oracle_rows(o, case, 0, 0)
    assert list(ncell) == [2, 2, 2] and err[0] == 0
    for a in want:
        assert sorted(rows[a]) == sorted(want[a])
    small = dict(case, interactions=[dict(case["interactions"][0], lists=[(1, 1, 160, 7.2, 5)])])   # half box 7.23 < 1.01 r_cut + margin
    rows2, err2, ncell2, pf_on2, _ = emulated_build(nl_kernels, small, 0, 0)
    o2 = oracle(small)
    o2.advance("nvt", 2.0, 0, 1)
    want2 = oracle_rows(o2, small, 0, 0)
    assert pf_on2 == 0 and all(sorted(rows2[a]) == sorted(want2[a]) for a in want2)
    tight = dict(case, interactions=[dict(case["interactions"][0], lists=[(1, 1, 40, 6.5, 5)])])
    _, err3, _, _, _ = emulated_build(nl_kernels, tight, 0, 0)
    assert err3[0] == 11 and err3[2] > 40                           # E_TOO_MANY with the count found (md_neighbours.f90:80)


def test_emulated_cell_order_is_keyed_by_atom_identity(nl_kernels):
    """After the re-sort the slot -> atom map depends on the positions only, not on the slots the atoms came from: what makes a
    restarted run bit-identical."""
    case = small_cases()["cu_fcc"]
    a = emulated_build(nl_kernels, case, 0, 0, seed=11)[4]
    b = emulated_build(nl_kernels, case, 0, 0, seed=12)[4]
    assert np.array_equal(a, b)

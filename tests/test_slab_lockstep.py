"""The slab decomposition (pfmds_b200/csrc/slab.cu) WITHOUT GPUs: the lock-step host replay of the library with the ranks as
threads of this process — NCCL replaced by an in-process stand-in (tests/emu/nccl_emu.hpp), CUDA IPC handles by plain
pointers, the spin-wait / signal kernels of the peer-memory halo running against each other for real.  Same comparison as
tests/test_slab_gpu.py: the decomposed run against the whole system in one context — atom ownership, positions, velocities,
forces, energies, thermostat and diagnostics after steps that include list rebuilds, migration and ghost re-selection; with 2
ranks (left neighbour = right neighbour) and 3, with the direct halo and with the NCCL halo (PFMDS_SLAB_P2P=0)."""
import os
import sys
import threading

import numpy as np
import pytest

from pfmds_b200 import inputs
from pfmds_b200.engine import configure
from pfmds_b200.slab import configure_slab, make_unique_id

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
import build_emu as BE  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _lib(oracle_lib):
    return BE.LOCKSTEP.build()


def _case(which, world=2):
    if which == "rjl":
        case = inputs.cu_fcc(cells=(max(12, 4 * world), 4, 4), jitter=0.05, period=5, temperature=900.0)    # hot, and shifted so that an atomic plane
        case["pos"] = case["pos"].copy()                                                    # lies on every slab face: atoms cross it
        case["pos"][:, 0] = (case["pos"][:, 0] - 0.25 * 3.615 + 0.02) % case["box"][0]
        return case, "nvt", 2.0
    if which == "lj1g":
        return inputs.lj_fluid(n_side=10, period=5, temperature=300.0), "nve", 1.0
    case = inputs.ab_gas(n_side=9, period=5, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, temperature=300.0)
    case["zero_momentum_period"] = 7
    return case, "nvt", 1.0


def _run_ranks(world, body):
    errors = []

    def wrap(rank):
        try:
            body(rank)
        except BaseException as ex:  # noqa: BLE001 - reported below, with the rank
            import traceback
            errors.append("rank %d: %s" % (rank, traceback.format_exc()))
    threads = [threading.Thread(target=wrap, args=(r,), daemon=True) for r in range(world)]   # daemon: a rank stuck in a collective after another one failed must not keep the process alive
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, "\n".join(errors)
    assert not any(t.is_alive() for t in threads), "a rank did not finish (deadlock in the halo protocol?)"


CONFIGS = [("rjl", 2, "1"), ("rjl", 3, "1"), ("rjl", 2, "0"), ("lj", 3, "1")]
if os.environ.get("PFMDS_SLAB_TESTS") == "all":      # the longer list: 8 ranks as in the scaling run, 4 ranks over the NCCL halo, lj1g
    CONFIGS += [("lj1g", 2, "1"), ("rjl", 8, "1"), ("rjl", 4, "0")]


@pytest.mark.parametrize("which,world,p2p", CONFIGS)
def test_slab_ranks_as_threads_match_the_single_context(monkeypatch, which, world, p2p):
    monkeypatch.setenv("PFMDS_SLAB_P2P", p2p)
    lib = BE.LOCKSTEP.lib
    case, integ, dt = _case(which, world)
    n = len(case["mass"])
    ref = configure(case, lib_path=lib)
    snaps = []
    ref.advance(integ, dt, 0, 1)
    snaps.append((ref.download(), ref.energies(), ref.diagnostics()))
    ref.advance(integ, dt, 1, 12)                       # rebuilds (with migration) at 5 and 10
    snaps.append((ref.download(), ref.energies(), ref.diagnostics()))
    rows_ref = ref.advance_logged(integ, dt, 13, 4, log_period=2)     # the device-resident energy log works across ranks too
    uid = make_unique_id(lib)
    owned = [[None, None] for _ in range(world)]

    def body(rank):
        slab = configure_slab(case, rank, world, 0, uid, lib_path=lib)

        def compare(k, tol_f, tol_x):
            (P, V, F), er, dr = snaps[k]
            gid, p, v, f = slab.download()
            owned[rank][k] = gid.copy()
            assert np.abs(p - P[gid - 1]).max() < tol_x and np.abs(v - V[gid - 1]).max() <= tol_x * np.abs(V).max() * 1e3 + 1e-18
            assert np.abs(f - F[gid - 1]).max() < tol_f * np.abs(F).max()
            es = slab.energies()
            assert np.allclose(es[0], er[0], rtol=max(tol_f, 1e-12), atol=1e-9)
            assert abs(es[1] - er[1]) <= max(tol_f, 1e-12) * abs(er[1]) + 1e-12 and abs(es[2] - er[2]) <= 1e-9 * er[2] + 1e-9
            if case["nhc"]:
                assert np.allclose(es[3], er[3], rtol=1e-7, atol=1e-9)
            ds = slab.diagnostics()
            assert np.allclose(ds[1], dr[1], rtol=1e-11) and abs(ds[3] - dr[3]) <= 1e-12 * dr[3] and np.array_equal(ds[4], dr[4])
        slab.advance(integ, dt, 0, 1)
        compare(0, 1e-11, 1e-12)
        slab.advance(integ, dt, 1, 12)
        compare(1, 1e-8, 1e-9)
        rows = slab.advance_logged(integ, dt, 13, 4, log_period=2)
        assert rows[0].shape[0] == 2
        for k in range(4):
            assert np.allclose(rows[k], rows_ref[k], rtol=1e-7, atol=1e-9)
        slab.close()
    _run_ranks(world, body)
    for k in (0, 1):                                     # every atom has exactly one owner, before and after the migrations
        assert sorted(np.concatenate([owned[r][k] for r in range(world)]).tolist()) == list(range(1, n + 1))
    moved = sum(len(set(owned[r][0].tolist()) ^ set(owned[r][1].tolist())) for r in range(world))
    if which == "rjl":
        assert moved > 0, "the hot crystal should have sent atoms across a slab face"


class _ThreadAllReduce:
    """sum of a float64 vector over the rank threads (what torch.distributed does for bench.py)."""

    def __init__(self, world):
        self.world, self.parts, self.barrier = world, {}, threading.Barrier(world)

    def __call__(self, rank, v):
        self.parts[rank] = np.array(v, np.float64)
        self.barrier.wait()
        out = sum(self.parts[r] for r in range(self.world))
        self.barrier.wait()
        return out


def test_per_rank_generated_crystal_and_the_bench_e2e_sequence(monkeypatch):
    """bench.py --workload cu_fcc_1e8 in miniature: every rank generates only its own slab (cu_fcc_slab_inputs, velocity sums
    all-reduced), hands it to pfmds_create_slab (SlabEngine local=...), and runs the e2e leg of the bench in slab mode (download,
    upload, one energy read per step).  Against one context holding the whole crystal assembled from the ranks' own arrays."""
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    from pfmds_b200.slab import cu_fcc_slab_inputs
    world, lib = 3, BE.LOCKSTEP.lib
    monkeypatch.setenv("PFMDS_SLAB_P2P", "1")
    uid = make_unique_id(lib)
    allsum = _ThreadAllReduce(world)
    gather, out = {}, {}
    ready = threading.Barrier(world)

    def body(rank):
        settings, loc = cu_fcc_slab_inputs(rank, world, (4, 4, 4), lambda v: allsum(rank, v), seed=9, steps=40, period=5)
        gather[rank] = loc
        ready.wait()
        eng = configure_slab(settings, rank, world, 0, uid, lib_path=lib, capacity_factor=1.6, local=loc)
        assert eng.n == 3 * 256 and eng.n_local0 == 256
        eng.advance("nvt", 2.0, 0, 4)
        gid0, p0, v0, _ = eng.download(forces=False)            # the e2e leg of bench.py, slab branch
        eng.upload_local(p0, v0)
        eng.advance("nvt", 2.0, 0, 1)
        e_bytes, how = bench.e2e_steps(eng, "nvt", 2.0, 6, stepwise=True)
        assert e_bytes > 0 and "pfmds_energies" in how
        out[rank] = (eng.download(), eng.energies())
        eng.close()
    _run_ranks(world, body)
    # the same crystal in one context
    n = 3 * 256
    pos, vel, mass = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
    for r in range(world):
        g = gather[r]["gid"] - 1
        pos[g], vel[g], mass[g] = gather[r]["pos"], gather[r]["vel"], gather[r]["mass"]
    whole = inputs.cu_fcc(cells=(12, 4, 4), steps=40, period=5)
    q1 = 3 * n * inputs.KB * 300.0 * 100.0 ** 2
    whole.update(pos=pos, vel=vel, mass=mass, nhc=[(1, 300.0, 3, q1)])
    ref = configure(whole, lib_path=lib)
    ref.advance("nvt", 2.0, 0, 4)
    P, V, _ = ref.download()
    ref.upload(P, V)
    ref.advance("nvt", 2.0, 0, 1)
    for s in range(1, 7):
        ref.advance("nvt", 2.0, s, 1, with_energy=True)
    (P, V, F), er = ref.download(), ref.energies()
    for r in range(world):
        (gid, p, v, f), es = out[r]
        assert np.abs(p - P[gid - 1]).max() < 1e-10 and np.abs(f - F[gid - 1]).max() < 1e-9 * np.abs(F).max()
        assert np.allclose(es[0], er[0], rtol=1e-11) and abs(es[1] / er[1] - 1) < 1e-10 and np.allclose(es[3], er[3], rtol=1e-7, atol=1e-9)

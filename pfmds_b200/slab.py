"""Slab spatial decomposition driver (BASELINE.json configs[3]): one process per GPU, torch.distributed for the
rendezvous, NCCL inside libpfmds_b200.so for the halo exchange (include/pfmds_b200.h, pfmds_create_slab)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import Engine, KIND, LIB_PATH, PfmdsError, _d, _i, load_library  # noqa: F401
from .inputs import group_indexes


def slab_partition(case, rank, world):
    """What rank `rank` of `world` hands to pfmds_create_slab: the 0-based file indexes of its atoms (x in
    [rank*Lx/world, (rank+1)*Lx/world), edge values clamped like the device code), the per-atom group bit masks of
    ALL atoms and the global group sizes."""
    box = np.asarray(case["box"], np.float64)
    n_global = len(case["mass"])
    W = box[0] / world
    owner = np.clip(np.floor(np.asarray(case["pos"])[:, 0] / W).astype(np.int64), 0, world - 1)
    mine = np.where(owner == rank)[0]
    ng = len(case["groups"])
    mask = np.zeros(n_global, np.uint32)
    sizes = np.zeros(ng, np.int64)
    for g in range(1, ng + 1):
        idx = group_indexes(case, g)
        sizes[g - 1] = len(idx)
        mask[idx - 1] |= np.uint32(1 << (g - 1))
    return mine, mask, sizes


def halo_atoms(case, rank, world, width):
    """0-based file indexes of the atoms rank `rank` must hold as ghosts: owned by the left / right neighbour and within
    `width` of the shared face (numpy restatement of slab.cu k_sl_border_flags, for tests)."""
    box = np.asarray(case["box"], np.float64)
    W = box[0] / world
    x = np.asarray(case["pos"])[:, 0]
    owner = np.clip(np.floor(x / W).astype(np.int64), 0, world - 1)
    left, right = (rank - 1) % world, (rank + 1) % world
    from_left = np.where((owner == left) & (((left + 1) * W - x) <= width))[0]      # the left neighbour's right border
    from_right = np.where((owner == right) & ((x - right * W) < width))[0]          # the right neighbour's left border
    return from_left, from_right


def cu_fcc_slab_inputs(rank, world, cells_per_rank, all_reduce_sum, seed=2, temperature=300.0, steps=1000, period=20):
    """BASELINE.json configs[3] at its stated size (10^8 atoms cannot be built whole on every rank): this rank generates only
    its own slab of the crystal (`inputs.cu_fcc_slab`), the two global sums of the velocity initialisation go through
    `all_reduce_sum` (a callable that sums a float64 numpy vector over the ranks: torch.distributed in the bench and tests),
    and the settings (groups, roles, thermostat, rjl list) are those of `inputs.cu_fcc` for the whole crystal.
    Returns (settings_case, local) with local = dict(gid, pos, vel, mass, mask, sizes, box, n_global) as SlabEngine takes it."""
    from . import inputs
    gid, pos, vel, mass, box, sums = inputs.cu_fcc_slab(rank, world, cells_per_rank, seed=seed, temperature=temperature)
    n_global = int(all_reduce_sum(np.array([float(len(gid))]))[0])
    sums = all_reduce_sum(np.asarray(sums, np.float64))
    vel = inputs.finish_velocities(vel, mass, sums, n_global, temperature)
    q1 = 3 * n_global * inputs.KB * temperature * 100.0 ** 2
    settings = inputs.cu_fcc(ncell=1, steps=steps, period=period, temperature=temperature, q1=q1)   # 4 atoms: only its tables are used
    for k in ("pos", "vel", "mass", "names"):
        settings.pop(k)
    settings["box"] = box
    mask = np.full(len(gid), 1, np.uint32)                # groups of cu_fcc: 1 = every CU atom, 2 = empty
    sizes = np.array([n_global, 0], np.int64)
    return settings, dict(gid=gid, pos=pos, vel=vel, mass=mass, mask=mask, sizes=sizes, box=box, n_global=n_global)


class SlabEngine(Engine):
    """This rank's share of a case: atoms with x in [rank*Lx/world, (rank+1)*Lx/world).  `local` (see cu_fcc_slab_inputs)
    replaces the whole-crystal arrays of `case` by this rank's own atoms."""

    def __init__(self, case, rank, world, device, unique_id, capacity_factor=1.6, lib_path=LIB_PATH, local=None):
        self._lib = load_library(lib_path)
        self._p = "pfmds_"
        self._ctx = C.c_void_p()
        L = self._lib
        L.pfmds_create_slab.restype = C.c_int
        L.pfmds_create_slab.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_longlong, C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint), C.c_int,
                                        C.POINTER(C.c_longlong), C.POINTER(C.c_double), C.c_int]
        L.pfmds_slab_download.restype = C.c_int
        L.pfmds_slab_download.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        box = np.ascontiguousarray(case["box"], np.float64)
        ng = len(case["groups"])
        if local is None:
            n_global = len(case["mass"])
            mine, mask, sizes = slab_partition(case, rank, world)
            pos = np.ascontiguousarray(case["pos"][mine], np.float64).reshape(-1)
            vel = np.ascontiguousarray(case["vel"][mine], np.float64).reshape(-1)
            mass = np.ascontiguousarray(case["mass"][mine], np.float64)
            gid = np.ascontiguousarray(mine + 1, np.int32)
            m = np.ascontiguousarray(mask[mine], np.uint32)
            self.groups = {g: group_indexes(case, g) for g in range(1, ng + 1)}
        else:
            n_global = int(local["n_global"])
            if n_global >= 2 ** 31:
                raise PfmdsError(1, "atom numbers are int32 in the C ABI")
            W = box[0] / world
            own = np.clip(np.floor(np.asarray(local["pos"])[:, 0] / W).astype(np.int64), 0, world - 1)
            if not np.all(own == rank):
                raise PfmdsError(1, "local atoms outside this rank's slab")
            mine = local["gid"]
            pos = np.ascontiguousarray(local["pos"], np.float64).reshape(-1)
            vel = np.ascontiguousarray(local["vel"], np.float64).reshape(-1)
            mass = np.ascontiguousarray(local["mass"], np.float64)
            gid = np.ascontiguousarray(local["gid"], np.int32)
            m = np.ascontiguousarray(local["mask"], np.uint32)
            sizes = np.ascontiguousarray(local["sizes"], np.int64)
            self.groups = {}
        self.n = n_global
        self.n_local0 = len(mine)
        self.capacity = int(len(mine) * capacity_factor) + 4096
        self.inter, self.nhc_M = [], []
        self._call("create_slab", C.byref(self._ctx), device, rank, world, unique_id, n_global, len(mine), _i(gid), _d(pos), _d(vel), _d(mass),
                   m.ctypes.data_as(C.POINTER(C.c_uint)), ng, sizes.ctypes.data_as(C.POINTER(C.c_longlong)), _d(box), self.capacity)

    def set_group(self, g, idx1):
        raise PfmdsError(1, "groups of a slab context are given at creation")

    def download(self, forces=True):
        """(global 1-based numbers, pos, vel, frc) of this rank's atoms."""
        n = C.c_int()
        gid = np.zeros(self.capacity, np.int32)
        pos, vel = np.zeros((self.capacity, 3)), np.zeros((self.capacity, 3))
        frc = np.zeros((self.capacity, 3)) if forces else None
        self._call("slab_download", self._ctx, C.byref(n), _i(gid), _d(pos), _d(vel), _d(frc))
        k = n.value
        return gid[:k], pos[:k], vel[:k], (frc[:k] if forces else None)


    def slab_counts(self):
        """(atoms owned by this rank, ghost copies held) right now."""
        self._lib.pfmds_slab_counts.restype = C.c_int
        self._lib.pfmds_slab_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        a, b = C.c_int(), C.c_int()
        self._call("slab_counts", self._ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    def upload_local(self, pos, vel):
        """Overwrite this rank's atoms (order of the last download)."""
        self._lib.pfmds_slab_upload.restype = C.c_int
        self._lib.pfmds_slab_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        pos = np.ascontiguousarray(pos, np.float64)
        vel = np.ascontiguousarray(vel, np.float64)
        self._call("slab_upload", self._ctx, len(pos), _d(pos.reshape(-1)), _d(vel.reshape(-1)))


def make_unique_id(lib_path=LIB_PATH):
    lib = load_library(lib_path)
    buf = C.create_string_buffer(128)
    lib.pfmds_slab_unique_id.argtypes = [C.c_char_p]
    if lib.pfmds_slab_unique_id(buf) != 0:
        raise PfmdsError(2, "pfmds_slab_unique_id failed (is libnccl.so.2 loadable?)")
    return buf.raw


def broadcast_unique_id(dist, device):
    """Rank 0 creates the NCCL id of the library's communicator; torch.distributed carries the 128 bytes."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t = torch.tensor(list(make_unique_id()), dtype=torch.uint8, device=device)
    dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def configure_slab(case, rank, world, device, unique_id, **kw):
    e = SlabEngine(case, rank, world, device, unique_id, **kw)
    r = case["roles"]
    e.set_roles(r["all_moving"], r["xyz_moving"], r["z_moving"], r["all_atoms"])
    for g, t, m, q in case["nhc"]:
        e.add_nhc(g, t, m, q)
    e.set_misc(case["zero_momentum_period"], case["invert_z_vel"])
    for it in case["interactions"]:
        e.add_interaction(it["name"], it["params"], it["lists"])
    return e


def slab_parity_check(dist, rank, world, local_device):
    """Built-in correctness check of the decomposed path for multi-GPU bench lines (the driver's test box has one GPU, so the
    2-GPU pytest cannot run there): a hot 1 600..3 200-atom Cu crystal (atoms cross the slab faces) is run as `world` slabs AND
    whole in a single context on every rank; ownership, step-0 forces (1e-11 of max|F|), energies, and positions / forces after
    23 steps with four rebuilds and migration (1e-9 / 1e-8) must agree.  Returns a dict; raises AssertionError on a mismatch."""
    import torch
    from . import inputs
    from .engine import configure
    nx = 16 if world <= 4 else 4 * world           # slab width >= 14 A: more than the 6.5 A halo (two halos for 2 ranks)
    case = inputs.cu_fcc(cells=(nx, 5, 5), jitter=0.05, period=5, temperature=900.0)
    uid = broadcast_unique_id(dist, torch.device("cuda", local_device))
    slab = configure_slab(case, rank, world, local_device, uid)
    ref = configure(case, device=local_device)
    out = {}

    def compare(tag, tol_f, tol_x):
        gid, p, v, f = slab.download()
        P, V, F = ref.download()
        n = torch.tensor([len(gid)], device="cuda")
        dist.all_reduce(n)
        assert int(n.item()) == len(case["mass"]), (tag, "atoms", int(n.item()))
        ex = float(np.abs(p - P[gid - 1]).max())
        ef = float(np.abs(f - F[gid - 1]).max() / np.abs(F).max())
        assert ex < tol_x and ef < tol_f, (tag, ex, ef)
        es, er = slab.energies(), ref.energies()
        assert np.allclose(es[0], er[0], rtol=max(tol_f, 1e-12), atol=1e-9), (tag, es[0], er[0])
        assert abs(es[1] - er[1]) <= max(tol_f, 1e-12) * abs(er[1]) + 1e-12, (tag, "ke")
        out[tag] = {"pos_abs": ex, "force_rel": ef}
        return len(gid)

    for e in (slab, ref):
        e.advance("nvt", 2.0, 0, 1)
    n0 = compare("step0", 1e-11, 1e-12)
    for e in (slab, ref):
        e.advance("nvt", 2.0, 1, 23)               # rebuilds (with migration) at 5, 10, 15, 20
    n1 = compare("step23", 1e-8, 1e-9)
    moved = torch.tensor([abs(n1 - n0)], device="cuda")
    dist.all_reduce(moved)
    out["atoms"] = len(case["mass"])
    out["owner_changes_net"] = int(moved.item())
    slab.close()
    ref.close()
    return out

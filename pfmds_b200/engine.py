"""ctypes binding of the C ABI in include/pfmds_b200.h.

`Engine` speaks to libpfmds_b200.so (the CUDA product).  The same class can be pointed at another
library exporting the same shapes under a different prefix — the tests use that to drive the CPU
oracle (prefix ``oracle_``) through identical calls; the product itself never loads the oracle.
There is no CPU fallback: if the CUDA library is missing or no GPU is present, construction fails.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

NVE, NVT, NVMS = 0, 1, 2
KIND = {"nve": NVE, "nvt": NVT, "nvms": NVMS}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpfmds_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class PfmdsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pfmds error %d: %s" % (code, msg))
        self.code = code


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


_SIGS = {
    "create": [C.POINTER(C.c_void_p), C.c_int, C.c_int, _dp, _dp, _dp, _dp],
    "set_group": [C.c_void_p, C.c_int, C.c_int, _ip],
    "set_roles": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int],
    "add_nhc": [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double],
    "set_misc": [C.c_void_p, C.c_int, C.c_int],
    "add_group_change": [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int],
    "group_size": [C.c_void_p, C.c_int, _ip],
    "add_interaction": [C.c_void_p, C.c_char_p, C.c_int, _dp, C.c_int, _ip, _ip, _dp, _ip],
    "advance": [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int],
    "advance_with_energy": [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int],
    "advance_logged": [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip],
    "energies": [C.c_void_p, _dp, _dp, _dp, _dp],
    "diagnostics": [C.c_void_p, _dp, _dp, _dp, _dp, _ip],
    "download": [C.c_void_p, _dp, _dp, _dp],
    "neighbours": [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _ip],
    "normals": [C.c_void_p, C.c_int, _dp],
    "get_nhc": [C.c_void_p, C.c_int, _dp, _dp],
    "set_nhc": [C.c_void_p, C.c_int, _dp, _dp],
    "state_size": [C.c_void_p, C.POINTER(C.c_longlong)],
    "save_state": [C.c_void_p, _dp],
    "restore_state": [C.c_void_p, _dp, _dp, _dp],
    "timers": [C.c_void_p, _dp],
    "upload": [C.c_void_p, _dp, _dp],
    "pair_count": [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_longlong)],
    "pair_count_within": [C.c_void_p, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_longlong)],
    "set_profiling": [C.c_void_p, C.c_int],
    "kernel_times": [C.c_void_p, C.c_int, _dp, C.POINTER(C.c_longlong)],
    "timer_start": [C.c_void_p],
    "timer_stop": [C.c_void_p, _dp],
    "create_slab": None,
    "slab_download": None,
    "slab_upload": None,
    "launch_count": [C.c_void_p, C.POINTER(C.c_longlong)],
    "synchronize": [C.c_void_p],
    "destroy": [C.c_void_p],
}


_OPTIONAL = ("state_size", "save_state", "restore_state", "advance_with_energy", "upload", "pair_count", "pair_count_within", "set_profiling", "kernel_times", "timer_start", "timer_stop")  # product-only entry points


def load_library(path=LIB_PATH, prefix="pfmds_"):
    if not os.path.exists(path):
        raise FileNotFoundError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = C.CDLL(path)
    for name, args in _SIGS.items():
        fn = getattr(lib, prefix + name, None)
        if args is None:
            continue
        if fn is None:
            if name in _OPTIONAL:
                continue
            raise AttributeError("%s does not export %s%s" % (path, prefix, name))
        fn.argtypes = args
        fn.restype = C.c_int
    getattr(lib, prefix + "last_error").argtypes = [C.c_void_p]
    getattr(lib, prefix + "last_error").restype = C.c_char_p
    getattr(lib, prefix + "version").restype = C.c_char_p
    return lib


class Engine:
    """One simulation context (one CUDA stream on one device)."""

    def __init__(self, pos, vel, mass, box, device=0, lib_path=LIB_PATH, prefix="pfmds_"):
        self._lib = load_library(lib_path, prefix)
        self._p = prefix
        self._ctx = C.c_void_p()
        self.n = int(len(mass))
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1)
        vel = np.ascontiguousarray(vel, np.float64).reshape(-1)
        mass = np.ascontiguousarray(mass, np.float64)
        box = np.ascontiguousarray(box, np.float64)
        self.groups = {}
        self.inter = []   # (name, [(g1,g2,max,rcut,period)])
        self.nhc_M = []
        self._call("create", C.byref(self._ctx), device, self.n, _d(pos), _d(vel), _d(mass), _d(box))

    def _call(self, name, *args):
        rc = getattr(self._lib, self._p + name)(*args)
        if rc != 0:
            msg = getattr(self._lib, self._p + "last_error")(self._ctx)
            raise PfmdsError(rc, (msg or b"").decode(errors="replace"))

    # ---- setup ----
    def set_group(self, g, idx1):
        idx1 = np.ascontiguousarray(idx1, np.int32)
        self.groups[g] = idx1
        self._call("set_group", self._ctx, g, len(idx1), _i(idx1))

    def set_roles(self, all_moving, xyz_moving, z_moving, all_atoms):
        self.roles = (all_moving, xyz_moving, z_moving, all_atoms)
        self._call("set_roles", self._ctx, all_moving, xyz_moving, z_moving, all_atoms)

    def add_nhc(self, group, temperature, M, q1):
        self.nhc_M.append(M)
        self._call("add_nhc", self._ctx, group, float(temperature), M, float(q1))

    def set_misc(self, zero_momentum_period, invert_z_vel):
        self._call("set_misc", self._ctx, int(zero_momentum_period), int(bool(invert_z_vel)))

    def add_group_change(self, group_from, group_to, ts1, ts2, frec):
        """One change_group_num entry (deposition): md_simulation.f90:63-71, md_general.f90:82-94."""
        self._call("add_group_change", self._ctx, int(group_from), int(group_to), int(ts1), int(ts2), int(frec))

    def group_size(self, g):
        n = C.c_int()
        self._call("group_size", self._ctx, int(g), C.byref(n))
        return n.value

    def add_interaction(self, name, params, lists):
        p = np.ascontiguousarray(params, np.float64)
        gn = np.ascontiguousarray([g for l in lists for g in l[:2]], np.int32)
        mx = np.ascontiguousarray([l[2] for l in lists], np.int32)
        rc = np.ascontiguousarray([l[3] for l in lists], np.float64)
        pe = np.ascontiguousarray([l[4] for l in lists], np.int32)
        self._call("add_interaction", self._ctx, name.encode(), len(p), _d(p), len(lists), _i(gn), _i(mx), _d(rc), _i(pe))
        self.inter.append((name, list(lists)))

    # ---- hot path ----
    def advance(self, integrator, dt, first_md_step, n_steps, with_energy=False):
        """with_energy: the last step's force pass also yields the potential energies (no second sweep in energies())."""
        kind = KIND[integrator] if isinstance(integrator, str) else int(integrator)
        name = "advance_with_energy" if with_energy and hasattr(self._lib, self._p + "advance_with_energy") else "advance"
        self._call(name, self._ctx, kind, float(dt), int(first_md_step), int(n_steps))

    def advance_logged(self, integrator, dt, first_md_step, n_steps, log_period=1):
        """n_steps steps; the energies of every step with step % log_period == 0 are logged on the device and returned with one
        copy: (e_inter[rows, n_inter], ke[rows], temperature[rows], e_nhc[rows, n_nhc])."""
        kind = KIND[integrator] if isinstance(integrator, str) else int(integrator)
        ni, nt = len(self.inter), len(self.nhc_M)
        w = ni + 2 + nt
        rows = np.zeros((max(1, n_steps), w))
        nr = C.c_int()
        self._call("advance_logged", self._ctx, kind, float(dt), int(first_md_step), int(n_steps), int(log_period), _d(rows), w, C.byref(nr))
        rows = rows[: nr.value]
        return rows[:, :ni].copy(), rows[:, ni].copy(), rows[:, ni + 1].copy(), rows[:, ni + 2:].copy()

    def synchronize(self):
        self._call("synchronize", self._ctx)

    # ---- observables ----
    def energies(self):
        e = np.zeros(max(1, len(self.inter)))
        en = np.zeros(max(1, len(self.nhc_M)))
        ke, t = C.c_double(), C.c_double()
        self._call("energies", self._ctx, _d(e), C.byref(ke), C.byref(t), _d(en))
        return e[: len(self.inter)].copy(), ke.value, t.value, en[: len(self.nhc_M)].copy()

    def diagnostics(self):
        fs, mc, mcv = np.zeros(3), np.zeros(3), np.zeros(3)
        vmax = C.c_double()
        nl = np.zeros(max(1, sum(len(l) for _, l in self.inter)), np.int32)
        self._call("diagnostics", self._ctx, _d(fs), _d(mc), _d(mcv), C.byref(vmax), _i(nl))
        return fs, mc, mcv, vmax.value, nl

    def download(self, forces=True, out=None):
        """(pos, vel, frc) in file order, [n, 3] each.  `out`: three preallocated C-contiguous float64 [n, 3] arrays to fill
        (e.g. views of pinned host memory; None entries are skipped) instead of fresh ones."""
        if out is not None:
            pos, vel, frc = out
            for a in (pos, vel, frc):
                if a is not None and not (a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.shape == (self.n, 3)):
                    raise PfmdsError(1, "download(out=...) needs C-contiguous float64 arrays of shape (n, 3)")
        else:
            pos, vel = np.zeros((self.n, 3)), np.zeros((self.n, 3))
            frc = np.zeros((self.n, 3)) if forces else None
        self._call("download", self._ctx, _d(pos), _d(vel), _d(frc))
        return pos, vel, frc

    def neighbours(self, interaction, lst):
        g1, _, mx, _, _ = self.inter[interaction][1][lst]
        rows = len(self.groups[g1])
        nlist = np.zeros((rows, mx), np.int32)
        nnum = np.zeros(rows, np.int32)
        less = np.zeros(rows, np.int32)
        self._call("neighbours", self._ctx, interaction, lst, _i(nlist), _i(nnum), _i(less))
        return nlist, nnum, less

    def normals(self, interaction):
        g1 = self.inter[interaction][1][2][0]
        out = np.zeros((len(self.groups[g1]), 3))
        self._call("normals", self._ctx, interaction, _d(out))
        return out

    def get_nhc(self, k):
        x, v = np.zeros(self.nhc_M[k]), np.zeros(self.nhc_M[k])
        self._call("get_nhc", self._ctx, k, _d(x), _d(v))
        return x, v

    def set_nhc(self, k, x, v):
        x = np.ascontiguousarray(x, np.float64)
        v = np.ascontiguousarray(v, np.float64)
        self._call("set_nhc", self._ctx, k, _d(x), _d(v))

    def save_state(self):
        """What positions and velocities do not carry (thermostat chains, group sizes): pfmds_save_state."""
        n = C.c_longlong()
        self._call("state_size", self._ctx, C.byref(n))
        blob = np.zeros(n.value)
        self._call("save_state", self._ctx, _d(blob))
        return blob

    def restore_state(self, pos, vel, blob):
        """Exact restart of a freshly configured context: pfmds_restore_state (lists and forces are rebuilt on the device)."""
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1)
        vel = np.ascontiguousarray(vel, np.float64).reshape(-1)
        blob = np.ascontiguousarray(blob, np.float64)
        self._call("restore_state", self._ctx, _d(pos), _d(vel), _d(blob))

    def upload(self, pos=None, vel=None):
        pos = None if pos is None else np.ascontiguousarray(pos, np.float64).reshape(-1)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float64).reshape(-1)
        self._call("upload", self._ctx, _d(pos), _d(vel))

    def upload_ptr(self, pos_ptr, vel_ptr):
        """Upload from raw host addresses (e.g. pinned torch tensors)."""
        self._call("upload", self._ctx, C.cast(pos_ptr, _dp), C.cast(vel_ptr, _dp))

    def pair_count(self, interaction, lst):
        n = C.c_longlong()
        self._call("pair_count", self._ctx, interaction, lst, C.byref(n))
        return n.value

    def pair_count_within(self, interaction, lst, r):
        n = C.c_longlong()
        self._call("pair_count_within", self._ctx, interaction, lst, float(r), C.byref(n))
        return n.value

    def set_profiling(self, on):
        self._call("set_profiling", self._ctx, int(bool(on)))

    def kernel_times(self):
        """{kernel class: (total ms, launches)} measured with CUDA events on the context's stream."""
        n = 32
        ms = np.zeros(n)
        cnt = np.zeros(n, np.int64)
        self._call("kernel_times", self._ctx, n, _d(ms), cnt.ctypes.data_as(C.POINTER(C.c_longlong)))
        name = getattr(self._lib, self._p + "kernel_name")
        name.restype = C.c_char_p
        name.argtypes = [C.c_int]
        return {name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n) if cnt[k] > 0}

    def timer_start(self):
        self._call("timer_start", self._ctx)

    def timer_stop(self):
        ms = C.c_double()
        self._call("timer_stop", self._ctx, C.byref(ms))
        return ms.value

    def timers(self):
        t = np.zeros(6)
        self._call("timers", self._ctx, _d(t))
        return t

    def launch_count(self):
        n = C.c_longlong()
        self._call("launch_count", self._ctx, C.byref(n))
        return n.value

    def close(self):
        if self._ctx:
            getattr(self._lib, self._p + "destroy")(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_peaks(device=0, lib_path=LIB_PATH):
    """(FP64 FMA TFLOP/s, copy GB/s) measured on the device by the library's micro-benchmarks."""
    lib = load_library(lib_path)
    lib.pfmds_measure_peaks.argtypes = [C.c_int, _dp, _dp]
    lib.pfmds_measure_peaks.restype = C.c_int
    a, b = C.c_double(), C.c_double()
    rc = lib.pfmds_measure_peaks(device, C.byref(a), C.byref(b))
    if rc != 0:
        raise PfmdsError(rc, "pfmds_measure_peaks failed")
    return a.value, b.value


def configure(case, device=0, lib_path=LIB_PATH, prefix="pfmds_"):
    """Build an Engine from a case dict of pfmds_b200.inputs (same calls md() makes, md_simulation.f90:48-93)."""
    from .inputs import group_indexes

    e = Engine(case["pos"], case["vel"], case["mass"], case["box"], device=device, lib_path=lib_path, prefix=prefix)
    for g in range(1, len(case["groups"]) + 1):
        e.set_group(g, group_indexes(case, g))
    r = case["roles"]
    e.set_roles(r["all_moving"], r["xyz_moving"], r["z_moving"], r["all_atoms"])
    for g, t, m, q in case["nhc"]:
        e.add_nhc(g, t, m, q)
    e.set_misc(case["zero_momentum_period"], case["invert_z_vel"])
    for ch in case.get("changes", []):
        e.add_group_change(*ch)
    for it in case["interactions"]:
        e.add_interaction(it["name"], it["params"], it["lists"])
    return e

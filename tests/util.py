"""Helpers shared by the parity tests: drive the CUDA library and the CPU oracle through the same calls."""
import os

import numpy as np

from pfmds_b200 import inputs
from pfmds_b200.engine import configure

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = dict(lib_path=os.path.join(ROOT, "oracle", "_build", "liboracle.so"), prefix="oracle_")
RTOL = 1e-9  # north_star: per-atom forces and total energies within 1e-9 relative in FP64


# Which side of the library's size switches a context takes (pfmds_ctx::small_n, nl_warp_n; read from the environment at
# pfmds_create).  "large" drives a SMALL system through the kernels the 10^6-atom bench times: thread-per-atom pipelined pair
# kernels (k_rjl_force / k_rjl_density / k_rjl_force_e / k_lj1g_pipe, zero_forces fused away) and the thread-per-atom list build
# (k_build_mask; "large_build" = k_build), so that they meet the O(N^2) oracle on hardware.
PATHS = {
    "small": {},
    "large": {"PFMDS_SMALL_N": "0", "PFMDS_NL_WARP_N": "0"},
    "large_build": {"PFMDS_SMALL_N": "0", "PFMDS_NL_WARP_N": "0", "PFMDS_NL_MASK": "0", "PFMDS_LJ1G_PIPE": "0"},
}


def gpu(case, path="small", **kw):
    env = PATHS[path] if isinstance(path, str) else dict(path)
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return configure(case, **kw)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def oracle(case):
    return configure(case, **ORACLE)


def rel_err(a, b):
    """max |a-b| relative to the largest magnitude of the reference array (per-atom forces: relative to max |F|)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def small_cases():
    return {
        "ab_gas": inputs.ab_gas(n_side=8, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, period=5),
        "cu_fcc": inputs.cu_fcc(ncell=6, jitter=0.05, period=5),
        "gr_cu_ljc": inputs.graphene_on_cu_small(interface="ljc", period=5),
        "gr_cu_morsec": inputs.graphene_on_cu_small(interface="morsec", period=5),
        "gr_cu_ljc_simplified": inputs.graphene_on_cu_small(interface="ljc", period=5, simplified=True),
    }


def next_row_cases():
    """Small cases of the SURVEY.md 8(f) rows (deposition, rebosc): frozen in tests/golden/ too, with their own tolerances."""
    return {
        "lj_deposition": inputs.lj_deposition(),
        "graphene_rebosc": inputs.graphene_rebosc(),
    }


def list_ids(case):
    return [(k, j) for k, it in enumerate(case["interactions"]) for j in range(len(it["lists"]))]


def rows_of(case, k, j):
    """Number of rows of list j of interaction k (nl(2) of lj/ljc/morsec is the converse of nl(1): rows = its group 2)."""
    it = case["interactions"][k]
    g = it["lists"][0][1] if (j == 1 and it["name"] in ("lj", "ljc", "morsec")) else it["lists"][j][0]
    return len(inputs.group_indexes(case, g))


def neighbours(eng, case, k, j):
    """(nlist, nnum, lessnnum) with the right number of rows for converse lists."""
    import ctypes as C
    n = rows_of(case, k, j)
    mx = case["interactions"][k]["lists"][j][2]
    nlist = np.zeros((n, mx), np.int32)
    nnum = np.zeros(n, np.int32)
    less = np.zeros(n, np.int32)
    ip = C.POINTER(C.c_int)
    eng._call("neighbours", eng._ctx, k, j, nlist.ctypes.data_as(ip), nnum.ctypes.data_as(ip), less.ctypes.data_as(ip))
    return nlist, nnum, less

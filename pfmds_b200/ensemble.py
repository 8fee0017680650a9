"""Ensemble (task-farm) mode: the reference's run_md_simulation_mpi (runners/run_md_simulation_mpi.f90:68-100).
Ranks never exchange data; rank r (1-based) of n runs the list entries i with mod(i-1,n)==r-1."""
from __future__ import annotations

import os
import subprocess


def shard(set_num: int, world: int, rank0: int):
    """1-based list entries owned by 0-based rank `rank0` (run_md_simulation_mpi.f90:76)."""
    if world > set_num:
        raise ValueError("error: too many mpi nodes (%d) for this list (%d)" % (world, set_num))  # :87
    return [i for i in range(1, set_num + 1) if (i - 1) % world == rank0]


def node_prefix(rank0: int) -> str:
    return "%04d-" % (rank0 + 1)  # '(i4.4,A)', :16


def run_rank(exe, rank0, world, ipath, ilist, opath="", prefix="", out_period=1000, extra=()):
    """Launch one ensemble rank of a run_md_simulation host (product or oracle binary)."""
    cmd = [exe, "-node", str(rank0 + 1), "-nodes", str(world), "-ipath", ipath, "-ilist", ilist, "-op", str(out_period)]
    if opath:
        cmd += ["-opath", opath]
    if prefix:
        cmd += ["-p", prefix]
    return subprocess.run(cmd + list(extra), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)

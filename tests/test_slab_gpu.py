"""Slab decomposition over 2 GPUs against the single-GPU path on the same system (run with gpurun --gpus 2):
forces, energies, thermostat and positions after steps that include list rebuilds, migration and halo updates."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from pfmds_b200 import inputs
from pfmds_b200.engine import configure
from pfmds_b200.slab import configure_slab, broadcast_unique_id
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = broadcast_unique_id(dist, torch.device("cuda", local))
which = sys.argv[2]
if which == "rjl":
    case = inputs.cu_fcc(cells=(16, 5, 5), jitter=0.05, period=5, temperature=900.0)   # hot: atoms cross the slab faces
    integ, dt = "nvt", 2.0
elif which == "lj1g":
    case = inputs.lj_fluid(n_side=16, period=5, temperature=300.0)
    integ, dt = "nve", 1.0
else:
    case = inputs.ab_gas(n_side=16, period=5, cap_aa=80, cap_ab=40, cap_ba=80, cap_bb=24, temperature=300.0)
    case["zero_momentum_period"] = 7
    integ, dt = "nvt", 1.0
slab = configure_slab(case, rank, world, local, uid)
ref = configure(case, device=local)                      # every rank also runs the whole system alone
def compare(tag, tol_f, tol_x):
    gid, p, v, f = slab.download()
    P, V, F = ref.download()
    n = torch.tensor([len(gid)], device="cuda"); dist.all_reduce(n)
    assert int(n.item()) == len(case["mass"]), (tag, int(n.item()))
    assert np.abs(p - P[gid - 1]).max() < tol_x, (tag, "pos", np.abs(p - P[gid - 1]).max())
    assert np.abs(v - V[gid - 1]).max() <= tol_x * max(1e-30, np.abs(V).max()) * 1e3 + 1e-18, (tag, "vel")
    assert np.abs(f - F[gid - 1]).max() < tol_f * np.abs(F).max(), (tag, "frc", np.abs(f - F[gid - 1]).max() / np.abs(F).max())
    es, er = slab.energies(), ref.energies()
    assert np.allclose(es[0], er[0], rtol=max(tol_f, 1e-12), atol=1e-9), (tag, es[0], er[0])
    assert abs(es[1] - er[1]) <= max(tol_f, 1e-12) * abs(er[1]) + 1e-12 and abs(es[2] - er[2]) <= 1e-9 * er[2] + 1e-9
    if case["nhc"]:
        assert np.allclose(es[3], er[3], rtol=1e-7, atol=1e-9)
    ds, dr = slab.diagnostics(), ref.diagnostics()
    assert np.allclose(ds[1], dr[1], rtol=1e-11) and abs(ds[3] - dr[3]) <= 1e-12 * dr[3] and np.array_equal(ds[4], dr[4])
    return len(gid)
for e in (slab, ref):
    e.advance(integ, dt, 0, 1)
n0 = compare("step0", 1e-11, 1e-12)
for e in (slab, ref):
    e.advance(integ, dt, 1, 23)          # rebuilds (with migration) at 5, 10, 15, 20
n1 = compare("step23", 1e-8, 1e-9)
moved = torch.tensor([abs(n1 - n0)], device="cuda"); dist.all_reduce(moved)
if rank == 0:
    print("SLAB_OK", which, "atoms that changed owner (net):", int(moved.item()))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("which", ["rjl", "lj1g", "lj"])
def test_slab_matches_single_gpu(tmp_path, cuda_lib, which):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    w = str(tmp_path / "worker.py")
    open(w, "w").write(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29641",
                        w, ROOT, which], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "SLAB_OK" in r.stdout, r.stdout[-4000:]

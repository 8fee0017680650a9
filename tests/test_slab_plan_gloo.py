"""N>1 slab path, host side, on CPU over gloo (world_size 2 and 3): every atom has exactly one owner, the group masks and
global group sizes each rank would hand to pfmds_create_slab agree across ranks, and the ghost sets are symmetric: what
rank r imports from its right neighbour is exactly what that neighbour exports to its left.  (The device side of the
decomposition is covered on GPUs by tests/test_slab_gpu.py.)"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from pfmds_b200 import inputs
from pfmds_b200.slab import slab_partition, halo_atoms
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = inputs.cu_fcc(cells=(4 * world + 2, 4, 4), jitter=0.05, seed=11)      # every rank builds the same crystal (seeded)
mine, mask, sizes = slab_partition(case, rank, world)
H = 6.5 * (1 + 1e-9) + 1e-5
gl, gr = halo_atoms(case, rank, world, H)
n = len(case["mass"])
owned = [None] * world
dist.all_gather_object(owned, mine.tolist())
ghosts = [None] * world
dist.all_gather_object(ghosts, (gl.tolist(), gr.tolist()))
meta = [None] * world
dist.all_gather_object(meta, (mask.tobytes(), sizes.tolist()))
flat = sorted(sum(owned, []))
assert flat == list(range(n)), "every atom has exactly one owner"
assert all(m == meta[0] for m in meta), "masks and global group sizes agree on every rank"
assert sizes[0] == n and int((mask & 1).sum()) == n
x = case["pos"][:, 0]; W = case["box"][0] / world
for r in range(world):
    right = (r + 1) % world
    # what r imports from its right neighbour = the right neighbour's atoms within H of its own left face
    exp = [i for i in owned[right] if x[i] - right * W < H]
    assert sorted(ghosts[r][1]) == sorted(exp)
    left = (r - 1) % world
    exp = [i for i in owned[left] if (left + 1) * W - x[i] <= H]
    assert sorted(ghosts[r][0]) == sorted(exp)
    # completeness: every atom within r_cut (minimum image) of one of r's atoms is local or a ghost of r
    have = set(owned[r]) | set(ghosts[r][0]) | set(ghosts[r][1])
    L = case["box"]
    P = case["pos"]
    for i in owned[r][::37]:
        d = P - P[i]
        d -= L * np.round(d / L)
        near = np.where((d * d).sum(1) < 6.5 ** 2)[0]
        assert set(near.tolist()) <= have
if rank == 0:
    print("PLAN_OK", world, [len(o) for o in owned], [len(g[0]) + len(g[1]) for g in ghosts])
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_slab_plan_over_gloo(tmp_path, world):
    w = str(tmp_path / "worker.py")
    open(w, "w").write(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
                        "--master-port", str(29660 + world), w, ROOT], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert "PLAN_OK" in r.stdout, r.stdout[-3000:]

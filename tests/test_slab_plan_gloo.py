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
# the per-rank generator of the 10^8-atom workload (bench.py --workload cu_fcc_1e8): same atoms, numbers and positions as the
# partition of the whole crystal, global momentum zero and exactly 300 K after the one all-reduce of its velocity sums
from pfmds_b200.slab import cu_fcc_slab_inputs
def allsum(v):
    t = torch.from_numpy(np.ascontiguousarray(v, np.float64).copy()); dist.all_reduce(t); return t.numpy()
settings, loc = cu_fcc_slab_inputs(rank, world, (3, 4, 4), allsum, seed=5)
whole = inputs.cu_fcc(cells=(3 * world, 4, 4), seed=5)
mine_w, mask_w, sizes_w = slab_partition(whole, rank, world)
assert loc["n_global"] == len(whole["mass"]) and np.array_equal(np.sort(loc["gid"]), mine_w + 1)
assert np.allclose(loc["pos"], whole["pos"][loc["gid"] - 1], rtol=0, atol=1e-12)
assert np.array_equal(loc["sizes"], sizes_w) and np.array_equal(loc["mask"], mask_w[loc["gid"] - 1])
assert settings["interactions"] == whole["interactions"] and settings["groups"] == whole["groups"] and np.allclose(settings["box"], whole["box"])
assert abs(settings["nhc"][0][3] / whole["nhc"][0][3] - 1) < 1e-12 and "pos" not in settings
mom = allsum((loc["mass"][:, None] * loc["vel"]).sum(0))
ke = allsum(np.array([(loc["mass"] * (loc["vel"] ** 2).sum(1)).sum()]))[0] / 2 * inputs.MASS_COEF
assert np.abs(mom).max() < 1e-9 and abs(2 * ke / inputs.KB / (3 * loc["n_global"]) - 300.0) < 1e-9
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

"""N>1 path on CPU: two processes over gloo reproduce the ensemble sharding of run_md_simulation_mpi
(mod(i-1,nnodes)==node_id-1, per-rank NNNN- prefixes) and the bench's max-over-ranks timing reduction."""
import os
import subprocess
import sys

import pytest

from pfmds_b200 import inputs
from conftest import ORACLE_EXE, ROOT

WORKER = r'''
import os, sys, glob
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from pfmds_b200.ensemble import shard, run_rank
d, exe = sys.argv[2], sys.argv[3]
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard(5, world, rank)
r = run_rank(exe, rank, world, d, "list.txt", opath=d, out_period=100)
assert r.returncode == 0, r.stdout
dist.barrier()
t = torch.tensor([10.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)          # bench.py: slowest rank defines the step time
owned = [None] * world
dist.all_gather_object(owned, mine)
if rank == 0:
    assert t.item() == 10.0 + world - 1
    assert sorted(sum(owned, [])) == [1, 2, 3, 4, 5] and all(len(set(a) & set(b)) == 0 for i, a in enumerate(owned) for b in owned[i + 1:])
    files = sorted(os.path.basename(f) for f in glob.glob(d + "*final_init.xyz"))
    assert files == ["0001-r1_final_init.xyz", "0001-r3_final_init.xyz", "0001-r5_final_init.xyz", "0002-r2_final_init.xyz", "0002-r4_final_init.xyz"], files
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_two_rank_ensemble_over_gloo(tmp_path, oracle_lib):
    case = inputs.ab_gas(n_side=4, cap_aa=64, cap_ab=64, cap_ba=64, cap_bb=64, period=5, steps=(4, 0, 0))
    case["integrators"] = [("nve", 0.5, 4, 1000, 2)]
    d = str(tmp_path) + os.sep
    inputs.write_case(d, case)
    open(d + "list.txt", "w").write("5\n" + "".join("md_run_settings.txt r%d_\n" % k for k in range(1, 6)))
    open(d + "worker.py", "w").write(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29631",
                        d + "worker.py", ROOT, d, ORACLE_EXE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=d, timeout=300)
    assert "GLOO_OK" in r.stdout, r.stdout[-3000:]

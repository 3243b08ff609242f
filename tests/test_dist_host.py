"""CPU tests (gloo, world_size 2) of the host logic of the multi-GPU path: the partition is
deterministic and identical on every rank, the cut-edge lists agree, and the communicator's
max-merge / point-to-point transfer behave as the engine expects."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import refapi as R
import stmqr_b200 as sq
from stmqr_b200 import dist as D


@pytest.mark.parametrize("case", ["dwt_992_metis", "lap2d_24_metis", "lap3d_8_metis"])
@pytest.mark.parametrize("nparts", [1, 2, 4, 8])
def test_partition_is_valid(case, nparts):
    sym, *_ = R.load_golden(case)
    owner, top = sq.partition_fronts(sym, nparts)
    owner2, top2 = sq.partition_fronts(sym, nparts)
    assert np.array_equal(owner, owner2) and np.array_equal(top, top2)          # deterministic
    assert owner.min() >= 0 and owner.max() < nparts
    parent = np.full(sym.nf, -1)
    for f in range(sym.nf):
        for q in range(int(sym.Childp[f]), int(sym.Childp[f + 1])):
            parent[int(sym.Child[q])] = f
    for f in range(sym.nf):
        if top[f]:
            assert owner[f] == 0
            assert parent[f] < 0 or top[parent[f]]                              # closed upwards
        elif parent[f] >= 0 and not top[parent[f]]:
            assert owner[f] == owner[parent[f]]                                 # whole subtrees
    if nparts == 1:
        assert not top.any()
    cut = D.cut_edges(sym, owner, top)
    assert all(top[parent[c]] and not top[c] for c, _ in cut)


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.environ["REPO"], "tests"))
    import refapi as R
    import stmqr_b200 as sq
    from stmqr_b200 import dist as D
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = D.TorchComm(torch.device("cpu"))
    sym, *_ = R.load_golden("lap2d_24_metis")
    owner, top = sq.partition_fronts(sym, world)
    # every rank computed the same partition
    t = torch.from_numpy(np.concatenate([owner, top]).astype(np.int64))
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    cut = D.cut_edges(sym, owner, top)
    # max-merge of an int32 array whose entries are owned by exactly one rank (Hm/Hr/Cm/W pattern)
    mine = torch.zeros(sym.nf, dtype=torch.int32)
    mine[torch.from_numpy(owner == rank)] = torch.arange(sym.nf, dtype=torch.int32)[torch.from_numpy(owner == rank)] + 1
    comm.allreduce_max({rank: mine})
    assert torch.equal(mine, torch.arange(sym.nf, dtype=torch.int32) + 1)
    w = torch.full((sym.m,), -1, dtype=torch.int32)
    w[rank::world] = torch.arange(sym.m, dtype=torch.int32)[rank::world]
    comm.allreduce_max({rank: w})
    assert torch.equal(w, torch.arange(sym.m, dtype=torch.int32))
    # the cut children travel to part 0
    for c, own in cut:
        payload = torch.full((5,), float(c))
        dst = torch.zeros(5)
        comm.transfer(own, 0, payload if rank == own else None, dst if rank == 0 else None)
        if rank == 0 and own != 0:
            assert torch.equal(dst, payload)
    s = comm.sum_host([float(rank + 1), 1.0]); m = comm.max_host([float(rank)])
    assert s == [world * (world + 1) / 2, float(world)] and m == [float(world - 1)]
    dist.barrier()
    dist.destroy_process_group()
    print("worker", rank, "ok")
""")


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=R.ROOT, PYTHONPATH=os.path.join(R.PKG, "py"))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2

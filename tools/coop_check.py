"""Multi-GPU check of the cooperative front path (run under torchrun, one rank per GPU): a 3-D Laplacian whose top
fronts are large enough for the K = 128 path, factorized (a) over all ranks with the C data plane (cooperative
fronts on unless STMQR_B200_COOP=0) and (b) on rank 0 alone; the gathered result must equal the single-GPU one
(integer structure bit for bit, R to the north_star tolerance -- in fact bitwise, every column sees the same
arithmetic).   usage: torchrun ... tools/coop_check.py [grid=64]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch, torch.distributed as dist
import bench                                        # (sets sys.path for the package and tests/refapi)
import refapi as R
import stmqr_b200 as sq
from stmqr_b200 import dist as D, matrices as M

g = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ref = R.Reference()
A = ref.csc_from_arrays(*M.laplacian_3d(g))
tol = ref.default_tol(A)
sym = ref.analyze_only(A, 2, tol)
At = ref.csc_to_numpy(A)
ntol = sym.n
os.environ["STMQR_B200_COOP_VERBOSE"] = "1"
e = sq.Engine(local)
e.analyze(sym)
e.upload_matrix(At)
df = D.DistFactorization(e, sym, torch.device("cuda", local))
ms = []
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    info = df.factorize(tol, ntol)
    ms.append(e.stats().ms_numeric)
t = torch.tensor([min(ms)], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
num = e.download(info)
got = D.gather_numeric(sym, df.owner, num, info)
e.close()
dist.barrier()
if rank == 0:
    e1 = sq.Engine(local)
    e1.analyze(sym)
    e1.upload_matrix(At)
    m1 = []
    for it in range(3):
        i1 = e1.factorize_resident(tol, ntol)
        m1.append(e1.stats().ms_numeric)
    want = e1.download(i1)
    e1.close()
    bad = R.structural_equal(got, want, sym)
    d = R.compare_R(sym, got, want, R.a_norm(At))
    same = np.array_equal(got.stack[: got.rh_size], want.stack[: want.rh_size])
    print(f"coop_check lap3d_{g} on {world} GPUs: {float(t.item()):.1f} ms (1 GPU {min(m1):.1f} ms), structure differs: {bad}, "
          f"max|dR|/|A| {d:.2e}, packed R+H bitwise equal: {same}, rank {got.rank}/{sym.n}, flops equal: {got.flops == want.flops}", flush=True)
    assert not bad and d <= R.R_TOL
dist.barrier()
dist.destroy_process_group()

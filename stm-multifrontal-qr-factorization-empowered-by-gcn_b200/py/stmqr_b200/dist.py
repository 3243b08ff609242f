"""stmqr_b200.dist -- the numeric factorization with the elimination tree partitioned over several
GPUs: independent etree subtrees on different GPUs, the contribution blocks of the
subtree roots moved to the GPU that owns the top of the tree, the global row permutation (qr_hpinv)
finished after an element-wise max-merge of a few integer arrays.

Replaces the reference's task tree + TPSM pool (SparseQR_analyze.c:705-1161,
SparseQR_multithreads.c:14-115).  One engine handle per GPU.  The data path is entirely on the
devices: the only host logic here is the (deterministic, replicated) partition and the order of the
calls.  Two communicators:

* ``TorchComm``  : one process per GPU under torchrun, torch.distributed (NCCL send/recv over NVLink
                   for the contribution blocks, NCCL all-reduce(MAX) for the merges; gloo on CPU for
                   the host-logic tests)
* ``LocalComm``  : all parts in ONE process on one GPU, executed one after the other (used by the GPU
                   parity test: the partitioned factorization must reproduce the single-GPU one bit
                   for bit, because every front sees exactly the same arithmetic)
"""
from __future__ import annotations

import numpy as np

from . import (ARRAY_CM, ARRAY_HM, ARRAY_HR, ARRAY_RDEAD, ARRAY_W, Engine, NumericInfo, Symbolic,
               partition_fronts)


class _DevMem:
    """__cuda_array_interface__ view of engine device memory, so torch can wrap it without a copy."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def as_tensor(ptr: int, count: int, typestr: str, device):
    import torch
    if count == 0:
        dt = {"<f8": torch.float64, "<i4": torch.int32, "|i1": torch.int8}[typestr]
        return torch.empty(0, dtype=dt, device=device)
    return torch.as_tensor(_DevMem(ptr, count, typestr), device=device)


MERGED = ((ARRAY_HM, "<i4"), (ARRAY_HR, "<i4"), (ARRAY_CM, "<i4"), (ARRAY_RDEAD, "|i1"))


def cut_edges(sym: Symbolic, owner: np.ndarray, is_top: np.ndarray):
    """Children of top fronts that are not top themselves: [(child front, owning part)], in front
    order (the same list on every rank)."""
    out = []
    Childp, Child = sym.Childp, sym.Child
    for f in range(sym.nf):
        if not is_top[f]:
            continue
        for q in range(int(Childp[f]), int(Childp[f + 1])):
            c = int(Child[q])
            if not is_top[c]:
                out.append((c, int(owner[c])))
    return out


class TorchComm:
    """torch.distributed communicator: rank r drives part r."""

    def __init__(self, device):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device

    def parts(self):
        return [self.rank]

    def allreduce_max(self, tensors_by_part):
        for t in tensors_by_part.values():
            if t.numel():
                self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)

    def transfer(self, src_part, dst_part, src_tensor, dst_tensor):
        """src_tensor lives on part src_part, dst_tensor on dst_part (either may be None here)"""
        if src_part == dst_part or (src_tensor is not None and src_tensor.numel() == 0) or \
                (dst_tensor is not None and dst_tensor.numel() == 0):
            return
        if self.rank == src_part:
            self.dist.send(src_tensor, dst=dst_part)
        elif self.rank == dst_part:
            self.dist.recv(dst_tensor, src=src_part)

    def sum_host(self, values):
        import torch
        t = torch.tensor(values, dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def max_host(self, values):
        import torch
        t = torch.tensor(values, dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


class LocalComm:
    """All parts in this process (one engine per part, same GPU): collectives become local merges."""

    def __init__(self, nparts: int, device):
        self.world = nparts
        self.rank = 0
        self.device = device

    def parts(self):
        return list(range(self.world))

    def allreduce_max(self, tensors_by_part):
        import torch
        ts = list(tensors_by_part.values())
        if not ts or ts[0].numel() == 0:
            return
        m = ts[0].clone()
        for t in ts[1:]:
            torch.maximum(m, t, out=m)
        for t in ts:
            t.copy_(m)

    def transfer(self, src_part, dst_part, src_tensor, dst_tensor):
        if src_part != dst_part and src_tensor.numel():
            dst_tensor.copy_(src_tensor)

    def sum_host(self, values):
        return values

    def max_host(self, values):
        return values


class PartitionedFactorization:
    """engines: {part: Engine} for the parts this process drives (one entry under torchrun)."""

    def __init__(self, comm, engines: dict, sym: Symbolic):
        import torch
        self.torch = torch
        self.comm = comm
        self.engines = engines
        self.sym = sym
        self.nparts = comm.world
        self.owner, self.is_top = partition_fronts(sym, self.nparts)
        self.cut = cut_edges(sym, self.owner, self.is_top)
        for p, e in engines.items():
            e.set_partition(self.nparts, p, self.owner, self.is_top)

    def _arr(self, e: Engine, which, typestr):
        ptr, n, _ = e.device_array(which)
        return as_tensor(ptr, n, typestr, self.comm.device)

    def factorize(self, tol: float, ntol: int):
        """-> {part: NumericInfo} with the global scalars filled in (rank, rank1, maxfrank, maxfm,
        flops); rh_size stays per part (every GPU keeps the R+H blocks of its own fronts)."""
        torch, comm, E = self.torch, self.comm, self.engines
        # 1. the subtrees, all GPUs at once
        for e in E.values():
            e.factorize_begin(tol, ntol)
        for e in E.values():
            e.factorize_levels(1)
        for e in E.values():
            e.sync()
        torch.cuda.synchronize()
        # 2. contribution blocks of the cut children -> part 0 (their scalars first: one max-merge)
        ncut = len(self.cut)
        if ncut:
            scal = {}
            for p, e in E.items():
                t = torch.zeros(ncut * 3, dtype=torch.int32, device=comm.device)
                for i, (c, own) in enumerate(self.cut):
                    if own == p:
                        r = e.front_regions(c)
                        t[3 * i: 3 * i + 3] = torch.tensor([r.cm, r.hr, r.hm], dtype=torch.int32)
                scal[p] = t
            comm.allreduce_max(scal)
            vals = next(iter(scal.values())).cpu().numpy().reshape(ncut, 3)
            for i, (c, own) in enumerate(self.cut):
                cm, hr, hm = (int(x) for x in vals[i])
                src = dst = None
                if own in E:
                    src = E[own].front_regions(c)
                if 0 in E and own != 0:
                    dst = E[0].front_regions(c, cm, hr, hm)
                if own == 0:
                    continue
                for field, cnt, ts in (("C", "C_doubles", "<f8"), ("Hii", "Hii_ints", "<i4")):
                    st = as_tensor(getattr(src, field), getattr(src, cnt), ts, comm.device) if src is not None else None
                    dt = as_tensor(getattr(dst, field), getattr(dst, cnt), ts, comm.device) if dst is not None else None
                    comm.transfer(own, 0, st, dt)
            torch.cuda.synchronize()
        # 3. the top of the tree on part 0
        if 0 in E:
            E[0].factorize_levels(2)
            E[0].sync()
        # 4. qr_hpinv: merge Hm, Hr, Cm, Rdead, then W
        for which, ts in MERGED:
            comm.allreduce_max({p: self._arr(e, which, ts) for p, e in E.items()})
        torch.cuda.synchronize()
        for e in E.values():
            e.factorize_hpinv_a()
        for e in E.values():
            e.sync()
        comm.allreduce_max({p: self._arr(e, ARRAY_W, "<i4") for p, e in E.items()})
        torch.cuda.synchronize()
        infos = {p: e.factorize_hpinv_b() for p, e in E.items()}
        # 5. global scalars
        loc = list(infos.values())
        rank = sum(i.rank for i in loc)
        flops = sum(i.flops for i in loc)
        rank, flops = comm.sum_host([float(rank), float(flops)])
        maxfrank, maxfm = comm.max_host([float(max(i.maxfrank for i in loc)), float(max(i.maxfm for i in loc))])
        for i in infos.values():
            i.rank = int(rank)
            i.flops = float(flops)
            i.maxfrank = int(maxfrank)
            i.maxfm = int(maxfm)
            if ntol >= self.sym.n:
                i.rank1 = int(rank)
        return infos


class DistFactorization:
    """One GPU's share of the multi-GPU numeric phase with the data plane in C (csrc/multigpu.cuh), one process
    per GPU under torchrun: torch.distributed only carries the 128-byte NCCL unique id to the other ranks; the
    contribution blocks, the row ids and the merges travel through ncclSend / ncclRecv / ncclAllReduce issued by
    the C library on the engine's own stream, with no host synchronisation between the etree levels."""

    def __init__(self, engine: Engine, sym: Symbolic, device=None):
        import torch
        import torch.distributed as dist
        from . import map_fronts, nccl_unique_id
        self.engine = engine
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.owner = map_fronts(sym, self.world)
        engine.set_ownership(self.world, self.rank, self.owner)
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if self.rank == 0:
            t.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        engine.comm_init(self.world, self.rank, bytes(t.cpu().numpy().tobytes()))

    def factorize(self, tol: float, ntol: int) -> NumericInfo:
        return self.engine.factorize_dist(tol, ntol)


def gather_numeric(sym: Symbolic, owner: np.ndarray, num, info):
    """rank 0 <- every rank's download (torch.distributed.gather_object), merged into one qr_numeric-shaped
    object; None on the other ranks.  For parity checks on small inputs."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = (num, {k: getattr(info, k) for k in ("rank", "rank1", "maxfrank", "maxfm", "rh_size", "flops")})
    out = [None] * world if rank == 0 else None
    dist.gather_object(payload, out, dst=0)
    if rank != 0:
        return None

    class _I:
        pass
    nums, infos = {}, {}
    for p, (n, i) in enumerate(out):
        nums[p] = n
        o = _I()
        o.__dict__.update(i)
        infos[p] = o
    return merge_numerics(sym, owner, nums, infos)


def merge_numerics(sym: Symbolic, owner: np.ndarray, nums: dict, infos: dict):
    """Host-side gather of the per-GPU downloads into ONE qr_numeric-shaped object (every front's
    packed R+H, HStair, HTau and Hii come from the GPU that owns it; Hm, Hr, HPinv, Rdead are
    already global on every GPU).  Used by the parity tests and by callers that want the whole
    factorization in one place."""
    from . import Numeric
    nf = sym.nf
    any_num = next(iter(nums.values()))
    any_info = next(iter(infos.values()))
    fn = (sym.Rp[1:nf + 1] - sym.Rp[:nf]).astype(np.int64)
    sizes = np.zeros(nf, np.int64)
    # packed size of front f = distance to the next block in its owner's stack
    for p, num in nums.items():
        mine = np.nonzero(owner == p)[0]
        if mine.size == 0:
            continue
        offs = num.Roff[mine]
        order = np.argsort(offs, kind="stable")
        ends = np.append(offs[order][1:], infos[p].rh_size)
        sizes[mine[order]] = ends - offs[order]
    Roff = np.zeros(max(nf, 1), np.int64)
    Roff[1:nf] = np.cumsum(sizes)[:-1] if nf > 1 else []
    stack = np.zeros(max(int(sizes.sum()), 1))
    HStair = np.zeros_like(any_num.HStair)
    HTau = np.zeros_like(any_num.HTau)
    Hii = np.zeros_like(any_num.Hii)
    for f in range(nf):
        num = nums[int(owner[f])]
        o = int(num.Roff[f])
        stack[Roff[f]: Roff[f] + sizes[f]] = num.stack[o: o + sizes[f]]
        p1 = int(sym.Rp[f])
        HStair[p1: p1 + fn[f]] = num.HStair[p1: p1 + fn[f]]
        HTau[p1: p1 + fn[f]] = num.HTau[p1: p1 + fn[f]]
        h1 = int(sym.Hip[f])
        hm = int(any_num.Hm[f])
        Hii[h1: h1 + hm] = num.Hii[h1: h1 + hm]
    return Numeric(int(any_info.rank), int(any_info.rank1), int(any_info.maxfrank), int(any_info.maxfm),
                   int(sizes.sum()), float(any_info.flops), stack=stack, Roff=Roff, Rdead=any_num.Rdead,
                   HStair=HStair, HTau=HTau, Hii=Hii, Hm=any_num.Hm, Hr=any_num.Hr, HPinv=any_num.HPinv)

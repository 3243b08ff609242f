"""GPU test of the multi-GPU path on ONE B200: the etree is partitioned over 2..4 engine handles
that live on the same device (stmqr_b200.dist.LocalComm runs the parts one after the other and
replaces NCCL by local merges).  The gathered factorization must reproduce the single-GPU one:
integer structure bit for bit, values to rounding (the panel kernel's thread count, hence its
summation order, is chosen per level from the fronts present, so a sub-level may round differently
-- the reference itself is not bit-reproducible across task partitions)."""
import numpy as np
import pytest

import refapi as R
import stmqr_b200 as sq
from stmqr_b200 import dist as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["dwt_992_metis", "lap2d_24_metis", "lap3d_8_metis", "rankdef_120x80_colamd"])
@pytest.mark.parametrize("nparts", [2, 3, 4])
def test_partitioned_equals_single(case, nparts):
    import torch
    sym, A, tol, ntol, want = R.load_golden(case)
    single = sq.Engine(0)
    single.analyze(sym)
    one = single.download(single.factorize(A, tol, ntol))
    single.close()

    engines = {}
    for p in range(nparts):
        e = sq.Engine(0)
        e.analyze(sym)
        e.upload_matrix(A)
        engines[p] = e
    pf = D.PartitionedFactorization(D.LocalComm(nparts, torch.device("cuda", 0)), engines, sym)
    assert set(np.unique(pf.owner)) <= set(range(nparts))
    infos = pf.factorize(tol, ntol)
    nums = {p: e.download(infos[p]) for p, e in engines.items()}
    got = D.merge_numerics(sym, pf.owner, nums, infos)
    assert not R.structural_equal(got, one, sym)
    assert got.rh_size == one.rh_size
    fp, fn = R.front_shapes(sym)
    for f in range(sym.nf):
        n = R.packed_front_size(sym, one, f)
        a = got.stack[int(got.Roff[f]): int(got.Roff[f]) + n]
        b = one.stack[int(one.Roff[f]): int(one.Roff[f]) + n]
        assert a.shape == b.shape
    assert R.compare_R(sym, got, one, R.a_norm(A)) <= 1e-13
    assert got.flops == one.flops
    R.assert_numeric_parity(sym, A, got, want, f"{case} on {nparts} parts")
    for e in engines.values():
        e.close()


@pytest.mark.parametrize("case", ["dwt_992_metis", "lap2d_24_metis", "lap3d_8_metis", "rankdef_120x80_colamd",
                                  "tall_600x150_colamd"])
@pytest.mark.parametrize("nparts", [2, 3, 4, 8])
def test_c_data_plane_equals_single(case, nparts):
    """The multi-GPU numeric phase with the data plane in C (csrc/multigpu.cuh): general ownership (upper levels
    spread over the GPUs), per-level transfers of contribution blocks + row ids + Cm/Hr/Hm in symbolic-bound sizes,
    max-merges at the end -- here with the in-process transport (peer copies, one host thread per handle) on
    handles that share ONE device, so every transfer, every arena offset and every merge is exercised.  The
    gathered factorization must reproduce the single-GPU one."""
    sym, A, tol, ntol, want = R.load_golden(case)
    single = sq.Engine(0)
    single.analyze(sym)
    one = single.download(single.factorize(A, tol, ntol))
    single.close()
    owner = sq.map_fronts(sym, nparts)
    assert owner.min() >= 0 and owner.max() < nparts
    engines = []
    for p in range(nparts):
        e = sq.Engine(0)
        e.analyze(sym)
        e.upload_matrix(A)
        e.set_ownership(nparts, p, owner)
        engines.append(e)
    grp = sq.PeerGroup(engines)
    for rep in range(2):                                       # twice: the plan and the transport are reusable
        infos = grp.factorize(tol, ntol)
        nums = {p: e.download(infos[p]) for p, e in enumerate(engines)}
        got = D.merge_numerics(sym, owner, nums, dict(enumerate(infos)))
        assert not R.structural_equal(got, one, sym)
        assert got.rh_size == one.rh_size
        assert R.compare_R(sym, got, one, R.a_norm(A)) <= 1e-13
        assert got.flops == one.flops
        R.assert_numeric_parity(sym, A, got, want, f"{case} on {nparts} GPUs (C data plane)")
    grp.close()
    for e in engines:
        e.close()

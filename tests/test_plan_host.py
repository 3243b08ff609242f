"""Host logic of the plan (no GPU): the recycled contribution-block arena.  A block lives from the pack of its
front's etree level to the assembly of its parent's level; the offsets planned by stmqr_b200_analyze /
stmqr_b200_set_partition must never let two blocks that are alive at the same time overlap, on one GPU and for
every part of a partitioned tree (the reference pops a child's block off its stack after the parent's assembly,
SparseQR_factorize.c:907-972)."""
import numpy as np
import pytest

import refapi as R
import stmqr_b200 as sq


def parents(sym):
    par = np.full(sym.nf, -1, np.int64)
    for f in range(sym.nf):
        for q in range(int(sym.Childp[f]), int(sym.Childp[f + 1])):
            par[int(sym.Child[q])] = f
    return par


def check_no_overlap(sym, coff, csize, level, alive):
    """alive: fronts whose block is held on this GPU.  Intervals [level(c), level(parent)] in schedule time."""
    par = parents(sym)
    ev = []
    for c in np.nonzero(alive)[0]:
        if csize[c] <= 0:
            continue
        birth = int(level[c])
        death = int(level[par[c]]) if par[c] >= 0 else 10 ** 9
        ev.append((int(coff[c]), int(coff[c] + csize[c]), birth, death, int(c)))
    ev.sort()
    # sweep over address order: compare each block with the following blocks that start before it ends
    for i, (a0, a1, b, d, c) in enumerate(ev):
        j = i + 1
        while j < len(ev) and ev[j][0] < a1:
            _, _, b2, d2, c2 = ev[j]
            # the parent's assembly (death) precedes the pack of the same level (birth of that level's blocks)
            assert d <= b2 or d2 <= b, f"blocks of fronts {c} and {c2} overlap while both are alive"
            j += 1


@pytest.mark.parametrize("case", ["dwt_992_metis", "lap2d_24_metis", "lap3d_8_metis", "tall_600x150_colamd"])
def test_recycled_arena_single_gpu(case):
    sym, A, tol, ntol, _ = R.load_golden(case)
    p = sq.Planner()
    p.analyze(sym)
    info, coff, csize, level = p.plan_info()
    assert info.C_doubles <= info.C_doubles_unrecycled + 2
    assert (coff + csize <= info.C_doubles).all()
    check_no_overlap(sym, coff, csize, level, np.ones(sym.nf, bool))
    # every child is on a lower level than its parent
    par = parents(sym)
    assert all(level[c] < level[par[c]] for c in range(sym.nf) if par[c] >= 0)
    assert info.device_bytes > 0
    p.close()


@pytest.mark.parametrize("case,nparts", [("lap2d_24_metis", 2), ("lap3d_8_metis", 4), ("dwt_992_metis", 3)])
def test_recycled_arena_partitioned(case, nparts):
    sym, A, tol, ntol, _ = R.load_golden(case)
    owner, is_top = sq.partition_fronts(sym, nparts)
    par = parents(sym)
    for part in range(nparts):
        p = sq.Planner()
        p.analyze(sym)
        p.set_partition(nparts, part, owner, is_top)
        info, coff, csize, level = p.plan_info()
        assert info.nparts == nparts and info.mypart == part
        # schedule time on this GPU: its subtrees by level, then (part 0) the top of the tree by level
        nlev = int(level.max()) + 1 if sym.nf else 1
        when = np.where(is_top.astype(bool), nlev + 1 + level, level)
        mine = owner == part
        # blocks received from other GPUs (cut children of the top) are placed when the top phase starts
        recv = np.zeros(sym.nf, bool)
        if part == 0:
            for c in range(sym.nf):
                if par[c] >= 0 and is_top[par[c]] and not is_top[c] and owner[c] != 0:
                    recv[c] = True
        when = np.where(recv, nlev, when)
        alive = mine | recv
        assert (coff[alive] + csize[alive] <= info.C_doubles).all()
        check_no_overlap(sym, coff, csize, when, alive)
        p.close()


def test_planner_has_no_compute_path():
    sym, A, tol, ntol, _ = R.load_golden("lap2d_16_notol")
    p = sq.Planner()
    p.analyze(sym)
    with pytest.raises(sq.EngineError):
        p.upload_matrix(A)
    with pytest.raises(sq.EngineError):
        p.factorize_resident(tol, ntol)
    p.close()


@pytest.mark.parametrize("case,nparts", [("lap2d_24_metis", 2), ("lap3d_8_metis", 4), ("lap3d_8_metis", 8),
                                         ("dwt_992_metis", 3)])
def test_ownership_map_and_arena(case, nparts):
    """General ownership (stmqr_b200_map_fronts) + the per-GPU arena plan of stmqr_b200_set_ownership: whole
    subtrees below the cut stay on one GPU, every GPU's blocks (its own and the ones it receives after the
    child's level) never overlap while alive."""
    sym, *_ = R.load_golden(case)
    owner = sq.map_fronts(sym, nparts)
    assert np.array_equal(owner, sq.map_fronts(sym, nparts))                  # deterministic
    assert owner.min() >= 0 and owner.max() < nparts
    _, is_top = sq.partition_fronts(sym, nparts)
    par = parents(sym)
    for c in range(sym.nf):
        if par[c] >= 0 and not is_top[par[c]]:
            assert owner[c] == owner[par[c]]                                   # subtrees below the cut are whole
    if nparts > 1 and is_top.any():
        assert len(set(owner[is_top.astype(bool)])) >= 1
    for part in range(nparts):
        p = sq.Planner()
        p.analyze(sym)
        p.set_ownership(nparts, part, owner)
        info, coff, csize, level = p.plan_info()
        mine = owner == part
        # a received block is placed after the child's level and read by its parent's level on this GPU
        recv = np.array([par[c] >= 0 and owner[par[c]] == part and owner[c] != part for c in range(sym.nf)])
        alive = mine | recv
        assert (coff[alive] + csize[alive] <= info.C_doubles).all()
        # a block that was sent away is never recycled: give it an infinite life on the sender
        sent = np.array([par[c] >= 0 and owner[par[c]] != part and owner[c] == part for c in range(sym.nf)])
        birth = level.astype(np.int64)
        check_no_overlap_general(sym, coff, csize, birth, alive, sent, par, level)
        p.close()


def check_no_overlap_general(sym, coff, csize, birth, alive, sent, par, level):
    ev = []
    for c in np.nonzero(alive)[0]:
        if csize[c] <= 0:
            continue
        death = 10 ** 9 if (sent[c] or par[c] < 0) else int(level[par[c]])
        ev.append((int(coff[c]), int(coff[c] + csize[c]), int(birth[c]), death, int(c)))
    ev.sort()
    for i, (a0, a1, b, d, c) in enumerate(ev):
        j = i + 1
        while j < len(ev) and ev[j][0] < a1:
            _, _, b2, d2, c2 = ev[j]
            assert d <= b2 or d2 <= b, f"blocks of fronts {c} and {c2} overlap while both are alive"
            j += 1


def test_cooperative_chunk_map():
    """stmqr_b200_coop_chunks (host only): chunk 0 stays on the home GPU (blocks 0..3 are factorized before any
    block can have travelled), every other GPU gets an equal share, the home GPU -- which also runs the panels --
    a smaller one that vanishes as GPUs are added; deterministic."""
    fn = 29388                                       # root front of the 96^3 Laplacian
    nch = (fn + 511) // 512
    for nparts, home in ((2, 0), (4, 0), (8, 0), (8, 3)):
        own = sq.coop_chunks(nparts, home, fn)
        assert own.shape == (nch,) and own[0] == home
        assert own.min() >= 0 and own.max() < nparts
        cnt = np.bincount(own, minlength=nparts)
        others = np.delete(cnt, home)
        assert others.max() - others.min() <= 1     # the peers share evenly
        assert cnt[home] <= others.max() + 1        # home never holds more than a peer (beyond chunk 0)
        assert np.array_equal(own, sq.coop_chunks(nparts, home, fn))
        # cyclic enough that every GPU still has work late in the factorization: the last quarter of the chunks
        # is spread over all the peers
        tail = own[-(nch // 4):]
        assert len(set(tail.tolist()) - {home}) == nparts - 1
    assert sq.coop_chunks(8, 0, 29388)[1:].tolist().count(0) <= 1       # 8 GPUs: home keeps (almost) nothing but chunk 0
    share2 = np.bincount(sq.coop_chunks(2, 0, fn), minlength=2)[0] / nch
    assert 0.35 <= share2 <= 0.5                     # 2 GPUs: home takes about 43 %
    assert sq.coop_chunks(1, 0, 1000).tolist() == [0, 0]
    with pytest.raises(RuntimeError):
        sq.coop_chunks(4, 4, 1000)


def test_plan_is_the_same_for_any_number_of_planner_threads():
    """The per-front symbolic pass of stmqr_b200_analyze runs in parallel slices (STMQR_B200_PLAN_THREADS): the plan
    -- arena sizes, every contribution-block offset, the R+H bound -- must not depend on how many threads built it.
    The thread count is read once per process, so every setting gets its own interpreter."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, hashlib; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np, refapi as R, stmqr_b200 as sq\n"
        "out = []\n"
        "for case in ('dwt_992_metis', 'lap2d_24_metis', 'lap3d_8_metis', 'rankdef_120x80_colamd'):\n"
        "    sym = R.load_golden(case)[0]\n"
        "    p = sq.Planner(); p.analyze(sym)\n"
        "    info, coff, csize, level = p.plan_info()\n"
        "    h = hashlib.sha256(coff.tobytes() + csize.tobytes() + level.tobytes()).hexdigest()[:16]\n"
        "    out.append((case, int(info.F_doubles), int(info.C_doubles), int(info.R_doubles), int(info.device_bytes), h))\n"
        "    p.close()\n"
        "print(out)\n" % (os.path.join(R.PKG, "py"), os.path.join(R.ROOT, "tests")))
    res = []
    for nt in ("1", "3", "8"):
        env = dict(os.environ, STMQR_B200_PLAN_THREADS=nt)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(r.stdout.strip().splitlines()[-1])
    assert res[0] == res[1] == res[2], res

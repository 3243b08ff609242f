"""CPU desk check of the cooperative-front schedule (csrc/stmqr_b200.cu: the coop_on branch of run_level on the
front's home GPU, coop_peer_level on the others).  The communication calls of every GPU are
stream ordered, so the schedule is deadlock free iff the per-GPU call sequences can be executed under the strictest
model -- every send, receive and broadcast is a rendezvous that blocks its stream until all participants have
reached it.  This test replays the sequences the C++ code issues (same loop bounds, same ownership map from the
library) under that model, and checks what the data flow needs: a block is at home before home factorizes it, with
every earlier reflector applied exactly once."""
import numpy as np
import pytest

import refapi as R  # noqa: F401  (puts the package on sys.path)
import stmqr_b200 as sq

WB, CHUNK = 128, 512


def ranges(own, me, lo, fn):
    out = []
    for c in range(lo // CHUNK, (fn + CHUNK - 1) // CHUNK):
        if own[c] != me:
            continue
        a, b = max(lo, c * CHUNK), min(fn, (c + 1) * CHUNK)
        if a < b:
            out.append((a, b))
    return out


def sequences(np_, home, fn):
    """-> {rank: [op, ...]}, op = ("bcast", J) | ("send", block, dst) | ("recv", block, src) | ("apply", J, a, b) |
    ("panel", block)"""
    own = sq.coop_chunks(np_, home, fn)
    nblk = (fn + WB - 1) // WB
    seq = {p: [] for p in range(np_)}
    # distribution of the assembled front (one group of sends / receives; chunk order)
    for p in range(np_):
        if p == home:
            continue
        for (a, b) in ranges(own, p, 0, fn):
            seq[home].append(("send", ("chunk", a), p))
            seq[p].append(("recv", ("chunk", a), home))
    for J in range(nblk):
        seq[home].append(("panel", J))
        cb = (J + 1) * WB
        if cb >= fn:
            break
        cm = min(cb + WB, fn)
        if own[cb // CHUNK] != home:
            seq[home].append(("recv", J + 1, int(own[cb // CHUNK])))
        seq[home].append(("bcast", J))
        seq[home].append(("apply", J, cb, cm))
        for (a, b) in ranges(own, home, cm, fn):
            seq[home].append(("apply", J, a, b))
        for p in range(np_):
            if p == home:
                continue
            seq[p].append(("bcast", J))
            c2 = cb + WB
            if c2 >= fn:
                continue
            e2 = min(c2 + WB, fn)
            send2 = own[c2 // CHUNK] == p
            if send2:
                seq[p].append(("apply", J, c2, e2))
            for (a, b) in ranges(own, p, e2 if send2 else c2, fn):
                seq[p].append(("apply", J, a, b))
            if send2:
                seq[p].append(("send", J + 2, home))
    return own, nblk, seq


def execute(np_, home, fn):
    own, nblk, seq = sequences(np_, home, fn)
    pos = {p: 0 for p in seq}
    where = {}                                      # column block -> rank that holds it
    for B in range(nblk):
        where[B] = int(own[(B * WB) // CHUNK])
    applied = {B: [] for B in range(nblk)}          # reflectors applied to every block, in order
    progress = True
    while progress:
        progress = False
        for p in seq:
            while pos[p] < len(seq[p]) and seq[p][pos[p]][0] in ("apply", "panel"):
                op = seq[p][pos[p]]
                if op[0] == "panel":
                    B = op[1]
                    assert where[B] == home == p, f"block {B} is not at home when it is factorized"
                    assert applied[B] == list(range(B)), f"block {B} factorized with reflectors {applied[B]}"
                else:
                    _, J, a, b = op
                    for B in range(a // WB, (b + WB - 1) // WB):
                        assert where[B] == p, f"rank {p} updates block {B} which is on rank {where[B]}"
                        applied[B].append(J)
                pos[p] += 1
                progress = True
        heads = {p: (seq[p][pos[p]] if pos[p] < len(seq[p]) else None) for p in seq}
        # a broadcast completes when every rank is at it
        h0 = heads[home]
        if h0 and h0[0] == "bcast" and all(heads[p] == h0 for p in seq):
            for p in seq:
                pos[p] += 1
            progress = True
            continue
        # a send / receive pair completes when both ends are at it
        for p, op in heads.items():
            if op and op[0] == "send":
                q = op[2]
                if heads[q] == ("recv", op[1], p):
                    if not isinstance(op[1], tuple):
                        where[op[1]] = q
                    pos[p] += 1
                    pos[q] += 1
                    progress = True
                    break
    stuck = {p: seq[p][pos[p]] for p in seq if pos[p] < len(seq[p])}
    assert not stuck, f"deadlock: {stuck}"
    assert all(where[B] == home for B in range(nblk))
    return nblk


@pytest.mark.parametrize("np_", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("fn", [300, 512, 1100, 4096, 22735, 29388])
def test_schedule_completes_and_every_block_sees_every_reflector_once(np_, fn):
    for home in (0, np_ - 1):
        execute(np_, home, fn)

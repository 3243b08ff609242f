"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, called through the C ABI,
against (1) the committed golden vectors of the real reference, (2) the plain-C oracle on
seeded synthetic inputs, (3) the reference's own consumers (QR_qmult / QR_solve) through the
drop-in qr_factorize, and (4) the unmodified reference driver qrtest under LD_PRELOAD."""
import os
import subprocess

import numpy as np
import pytest

import refapi as R
import stmqr_b200 as sq
from stmqr_b200 import matrices as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    e = sq.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def oracle():
    return R.Oracle()


def run_engine(engine, sym, A, tol, ntol):
    engine.analyze(sym)
    info = engine.factorize(A, tol, ntol)
    return engine.download(info)


@pytest.mark.parametrize("case", R.GOLDEN_CASES)
def test_engine_matches_golden(engine, case):
    sym, A, tol, ntol, want = R.load_golden(case)
    got = run_engine(engine, sym, A, tol, ntol)
    R.assert_numeric_parity(sym, A, got, want, case)
    assert got.flops == want.flops
    assert R.reference_flops(sym, got) == got.flops


@pytest.mark.parametrize("case", ["dwt_992_metis", "lap2d_24_metis", "rankdef_120x80_colamd"])
def test_assembly_bit_exact_on_leaves(case, oracle):
    """Assembled F of leaf fronts is pure data movement: bit-exact against the oracle; every
    front's row/column maps (Hii before hpinv, Stair, fm) are exact for all fronts."""
    sym, A, tol, ntol, _ = R.load_golden(case)
    e = sq.Engine(0)
    e.set_debug_capture(True)
    e.analyze(sym)
    info = e.factorize(A, tol, ntol)
    got = e.download(info)
    want = oracle.factorize(sym, A, tol, ntol, capture=True)
    assert np.array_equal(got.Hm, want.Hm)
    nchild = np.diff(sym.Childp[: sym.nf + 1])
    nleaf = ninterior = 0
    for f in range(sym.nf):
        Fa = e.get_front(f, 0)
        assert Fa.shape == want.Fasm[f].shape
        if nchild[f] == 0:
            assert np.array_equal(Fa, want.Fasm[f]), f"leaf front {f}"
            nleaf += 1
        else:
            # interior fronts: a child's C block is unique only up to an orthogonal mixing of its rows
            # (SURVEY.md 8(c)), so (1) the original rows of S -- identified by the oracle's row ids
            # before qr_hpinv -- must be bit-exact, in the same positions, and (2) F'F, which no
            # orthogonal row mixing can change, must agree to the R tolerance
            col1, fp = int(sym.Super[f]), int(sym.Super[f + 1] - sym.Super[f])
            r1, r2 = int(sym.Sleft[col1]), int(sym.Sleft[col1 + fp])
            ids = want.Hii_raw[int(sym.Hip[f]): int(sym.Hip[f]) + Fa.shape[0]]
            srow = (ids >= r1) & (ids < r2)
            assert srow.sum() == r2 - r1
            assert np.array_equal(Fa[srow], want.Fasm[f][srow]), f"S rows of interior front {f}"
            G1, G2 = Fa.T @ Fa, want.Fasm[f].T @ want.Fasm[f]
            assert np.max(np.abs(G1 - G2)) <= R.R_TOL * R.a_norm(A) ** 2, f"F'F of interior front {f}"
            ninterior += 1
    assert nleaf > 0 and ninterior > 0
    e.close()


@pytest.mark.parametrize("gen,order,tolmode", [
    (("lap2d", 48), 2, "default"), (("lap2d", 40), 1, "default"), (("lap3d", 12), 2, "default"),
    (("tall", 3000, 800), 1, "default"), (("lap2d", 32), 2, "notol"), (("rankdef", 400, 300), 1, "default"),
])
def test_engine_matches_oracle_synthetic(engine, oracle, gen, order, tolmode):
    """Seeded synthetic inputs analysed by the reference's own qr_analyze (reused unchanged)."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref for the symbolic analysis")
    ref = R.Reference()
    ref.set_backend("reference")
    if gen[0] == "lap2d":
        m, n, p, i, x = M.laplacian_2d(gen[1])
    elif gen[0] == "lap3d":
        m, n, p, i, x = M.laplacian_3d(gen[1])
    elif gen[0] == "tall":
        m, n, p, i, x = M.tall_banded_random(gen[1], gen[2], draws=8, halfwidth=16, seed=4)
    else:
        m, n, p, i, x = M.random_sparse(gen[1], gen[2], 0.02, seed=11, rank_deficient_cols=25)
    A = ref.csc_from_arrays(m, n, p, i, x)
    tol = ref.default_tol(A) if tolmode == "default" else -1.0
    QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    sym = ref.symbolic(QR)
    refnum = ref.numeric(QR, sym)
    At, ttol, ntol = ref.tapped()
    want = oracle.factorize(sym, At, ttol, ntol)
    got = run_engine(engine, sym, At, ttol, ntol)
    R.assert_numeric_parity(sym, At, got, want, str(gen))
    R.assert_numeric_parity(sym, At, got, refnum, str(gen) + " vs reference")
    assert got.flops == want.flops
    ref.free_qr(QR); ref.free_sparse(A); ref.close()


@pytest.mark.parametrize("gen,order", [(("lap3d", 30), 2), (("lap2d", 300), 2), (("tall", 40000, 10000), 1)])
def test_large_fronts_two_level_blocking(gen, order):
    """Fronts big enough for the two-level (128-column outer block, K = 128 DMMA update) path of
    kernels_wide.cuh: same integer structure and R (up to row signs) as the reference's CPU
    qr_factorize, and as the engine's own single-level path (flags bit 1)."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref for the symbolic analysis")
    ref = R.Reference()
    ref.set_backend("reference")
    if gen[0] == "lap2d":
        m, n, p, i, x = M.laplacian_2d(gen[1])
    elif gen[0] == "lap3d":
        m, n, p, i, x = M.laplacian_3d(gen[1])
    else:
        m, n, p, i, x = M.tall_banded_random(gen[1], gen[2], draws=8, halfwidth=64, seed=4)
    A = ref.csc_from_arrays(m, n, p, i, x)
    tol = ref.default_tol(A)
    QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    sym = ref.symbolic(QR)
    refnum = ref.numeric(QR, sym)
    At, ttol, ntol = ref.tapped()
    fm, fn = np.asarray(sym.Fm[: sym.nf]), np.diff(sym.Rp[: sym.nf + 1])
    assert (fm >= 1024).any() and (fn >= 384).any(), "input too small to reach the two-level path"
    got = {}
    # flags bit 1: single-level blocking; STMQR_B200_GRID_ROWS: fronts with at least that many rows
    # take the panel kernel that exchanges through global memory (k_panel_grid)
    # STMQR_B200_WIDE_ROWS: fronts with at least that many rows take the two-level path (default 4096)
    for flags, grid_rows in ((0, None), (2, None), (0, "1024"), (2, "1024")):
        os.environ["STMQR_B200_WIDE_ROWS"] = "1024"
        if grid_rows:
            os.environ["STMQR_B200_GRID_ROWS"] = grid_rows
        try:
            e = sq.Engine(0)
        finally:
            os.environ.pop("STMQR_B200_GRID_ROWS", None)
            os.environ.pop("STMQR_B200_WIDE_ROWS", None)
        e.set_options(flags=flags)
        num = got[(flags, grid_rows)] = run_engine(e, sym, At, ttol, ntol)
        R.assert_numeric_parity(sym, At, num, refnum, f"{gen} flags {flags} grid {grid_rows} vs reference")
        assert num.flops == R.reference_flops(sym, num)
        assert not R.structural_equal(num, got[(0, None)], sym)
        e.close()
    ref.free_qr(QR); ref.free_sparse(A); ref.close()


@pytest.mark.parametrize("name,order", [("dwt_992", 2), ("t2d_q9", 2), ("bcsstk14", 1), ("epb1", 1), ("ex18", 1),
                                        ("cvxqp3", 1)])
def test_dropin_through_reference_api(name, order):
    """The reference's SparseQR() with the B200 qr_factorize interposed: same integer structure
    as the CPU reference, solve residual (qrtest.c check_error) and Q orthogonality through the
    reference's untouched QR_qmult / QR_solve, allocator accounting balanced."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref")
    path = os.path.join(R.DATA_DIR, name + ".mtx")
    ref = R.Reference()
    A = ref.read_mtx(path)
    tol = ref.default_tol(A)
    ref.set_backend("reference")
    QRc = ref.sparseqr(A, order, tol, grain=1.0, tap=True)      # tap: keep a copy of A / Y
    symc = ref.symbolic(QRc); numc = ref.numeric(QRc, symc)
    At, _, _ = ref.tapped()
    res_c = ref.check_error(A, QRc)
    # allocator accounting of one SparseQR + SparseQR_free cycle with the reference's own
    # qr_factorize (the reference itself leaves m*sizeof(Long) bytes accounted, so compare deltas)
    i0 = ref.memory_inuse()
    ref.free_qr(ref.sparseqr(A, order, tol, grain=1.0))
    delta_cpu = ref.memory_inuse() - i0
    inuse0 = ref.memory_inuse()
    ref.set_backend("b200")
    QRg = ref.sparseqr(A, order, tol, grain=1.0)
    symg = ref.symbolic(QRg); numg = ref.numeric(QRg, symg)
    res_g = ref.check_error(A, QRg)
    assert not R.structural_equal(numg, numc, symg)
    full_rank = numc.rank == symc.n
    # R parity at the north_star tolerance for rank-deficient inputs too: two runs of the CPU reference
    # itself (serial tree vs TPSM tasks, 1 vs 8 BLAS threads; recorded in this container with
    # tools/oracle_self_noise.py) differ by 2.7e-13 / 6.0e-13 * ||A|| on cvxqp3 (rank 17042/17500, 6300 x 6300
    # root front) and by 3.9e-17 on dwt_992 (rank 496/992), so 1e-10 leaves more than two orders of margin
    d = R.compare_R(symg, numg, numc, R.a_norm(At))
    print(f"{name}: rank {numg.rank}/{symc.n}, max |dR| / ||A|| vs the CPU reference = {d:.3e}")
    assert d <= R.r_tol_for(name), d
    if full_rank:
        assert res_g <= max(10 * res_c, 1e-9), (res_g, res_c)
    # Q' Q = I on random vectors through the reference's own Q-apply
    rng = np.random.default_rng(3)
    m = At.nrow if ref.qr_info(QRg)["n1cols"] == 0 else None
    if m is not None:
        X = rng.standard_normal((m, 3))
        Y = ref.qmult(QRg, R.QR_QX, ref.qmult(QRg, R.QR_QTX, X))
        assert np.max(np.abs(Y - X)) <= 1e-10 * max(1.0, np.max(np.abs(X)))
    ref.free_qr(QRg)
    assert ref.memory_inuse() - inuse0 == delta_cpu   # every block freed by the reference's qr_freenum
    ref.free_qr(QRc); ref.free_sparse(A)
    ref.set_backend("reference")
    ref.close()


@pytest.mark.parametrize("name,order,devices", [("dwt_992", 2, "0,0"), ("t2d_q9", 2, "0,0,0"), ("epb1", 1, "0,0,0,0"),
                                                ("cvxqp3", 1, "0,0")])
def test_dropin_multi_gpu_through_reference_api(name, order, devices):
    """STMQR_B200_DEVICES: the SAME drop-in entry point with one engine handle per listed device (here the handles
    share device 0, which exercises every transfer and merge of csrc/multigpu.cuh's in-process transport): the
    reference's SparseQR() gets a qr_numeric with one stack per GPU; same integer structure and R as the CPU
    reference, same solve residual, Q'Q = I through the reference's untouched consumers, allocator balanced."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref")
    ref = R.Reference()
    A = ref.read_mtx(os.path.join(R.DATA_DIR, name + ".mtx"))
    tol = ref.default_tol(A)
    ref.set_backend("reference")
    QRc = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    symc = ref.symbolic(QRc); numc = ref.numeric(QRc, symc)
    At, _, _ = ref.tapped()
    res_c = ref.check_error(A, QRc)
    i0 = ref.memory_inuse()
    ref.free_qr(ref.sparseqr(A, order, tol, grain=1.0))
    delta_cpu = ref.memory_inuse() - i0
    ref.dropin_shutdown()
    os.environ["STMQR_B200_DEVICES"] = devices
    try:
        ref.set_backend("b200")
        inuse0 = ref.memory_inuse()
        QRg = ref.sparseqr(A, order, tol, grain=1.0)
        assert int(ref.qr_info(QRg)["ns"]) == len(devices.split(","))          # one stack per GPU
        symg = ref.symbolic(QRg); numg = ref.numeric(QRg, symg)
        assert not R.structural_equal(numg, numc, symg)
        d = R.compare_R(symg, numg, numc, R.a_norm(At))
        assert d <= R.r_tol_for(name), d
        res_g = ref.check_error(A, QRg)
        if numc.rank == symc.n:
            assert res_g <= max(10 * res_c, 1e-9), (res_g, res_c)
        if ref.qr_info(QRg)["n1cols"] == 0:
            rng = np.random.default_rng(3)
            X = rng.standard_normal((At.nrow, 2))
            Y = ref.qmult(QRg, R.QR_QX, ref.qmult(QRg, R.QR_QTX, X))
            assert np.max(np.abs(Y - X)) <= 1e-10 * max(1.0, np.max(np.abs(X)))
        ref.refactorize(A, QRg)                                                 # cached multi-GPU plan, second call
        numg2 = ref.numeric(QRg, symg)
        assert not R.structural_equal(numg2, numc, symg)
        ref.free_qr(QRg)
        # (the stacks are counted differently from the one-stack layout only by the ns-sized pointer arrays)
        assert abs((ref.memory_inuse() - inuse0) - delta_cpu) <= 0
    finally:
        os.environ.pop("STMQR_B200_DEVICES", None)
        ref.dropin_shutdown()
        ref.free_qr(QRc); ref.free_sparse(A)
        ref.set_backend("reference")
        ref.close()


def test_qrtest_driver_with_ld_preload(tmp_path):
    """The reference's own acceptance driver (STMMQR/test/qrtest.c), unmodified binary, with the
    drop-in library preloaded: prints the published fingerprint residual for dwt_992."""
    qrtest = os.path.join(R.REF_DIR, "qrtest")
    if not os.path.exists(qrtest):
        pytest.skip("needs oracle/_ref/qrtest")
    (tmp_path / "Results").mkdir()
    env = dict(os.environ, LD_PRELOAD=sq.DROPIN_PATH, OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([qrtest, os.path.join(R.DATA_DIR, "dwt_992.mtx"), "1", "1"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "res =  2.4e+01" in out.stdout, out.stdout[-2000:]     # STM-MQR.xlsx row 700 / SURVEY.md 4


def test_qrtest_driver_statically_linked(tmp_path):
    """The same driver from a STATIC link of the reference (its stock build) with the drop-in object in place of
    the one renamed symbol (host/link_static.sh, INTEGRATION.md Option C): LD_PRELOAD is not involved."""
    import test_abi
    exe = test_abi._static_link(tmp_path)
    (tmp_path / "Results").mkdir(exist_ok=True)
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    env.pop("LD_PRELOAD", None)
    out = subprocess.run([exe, os.path.join(R.DATA_DIR, "dwt_992.mtx"), "1", "1"], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "res =  2.4e+01" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("case", ["dwt_992_metis", "lap2d_24_metis", "tall_600x150_colamd"])
def test_streamed_download_is_identical(engine, case):
    """stmqr_b200_factorize_streamed (R+H blocks copied level by level while later levels run)
    delivers bit-for-bit what factorize + download delivers."""
    sym, A, tol, ntol, want = R.load_golden(case)
    a = run_engine(engine, sym, A, tol, ntol)
    engine.analyze(sym)
    info, stack = engine.factorize_streamed(A, tol, ntol)
    b = engine.download(info, stack=stack)
    assert not R.structural_equal(a, b, sym)
    assert a.rh_size == b.rh_size and np.array_equal(a.stack[: a.rh_size], b.stack[: b.rh_size])
    assert np.array_equal(a.Roff, b.Roff) and np.array_equal(a.HTau, b.HTau)
    R.assert_numeric_parity(sym, A, b, want, case + " streamed")


def test_errors_and_edge_cases(engine):
    sym, A, tol, ntol, want = R.load_golden("lap2d_16_notol")
    engine.analyze(sym)
    # matrix that does not match the analysis
    bad = sq.Csc(A.nrow, A.ncol, A.p, A.i[::-1].copy(), A.x)
    with pytest.raises(sq.EngineError):
        engine.factorize(sq.Csc(A.nrow - 1, A.ncol, A.p, A.i, A.x), tol, ntol)
    # refactorization with new values on the same plan: linear in A (R scales)
    info = engine.factorize(A, tol, ntol)
    n1 = engine.download(info)
    A2 = sq.Csc(A.nrow, A.ncol, A.p, A.i, 2.0 * A.x)
    info2 = engine.factorize(A2, tol, ntol)
    n2 = engine.download(info2)
    assert np.array_equal(n1.HStair, n2.HStair) and np.array_equal(n1.Hii, n2.Hii)
    assert R.compare_R(sym, n2, sq.Numeric(**{**n1.__dict__, "stack": 2.0 * n1.stack}), R.a_norm(A2)) <= 1e-13


def test_config2_full_size_properties():
    """BASELINE config[1] at FULL size (2-D Laplacian 1024 x 1024, METIS) through the drop-in, checked by
    size-independent properties: full rank, solve residual through the reference's own QR_solve /
    QR_qmult (qrtest.c check_error), run-to-run determinism (bitwise), and exact linearity under a
    power-of-two scaling of A (every operation of the factorization commutes with it)."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref for the symbolic analysis and the consumers")
    g = 1024
    m, n, p, i, x = M.laplacian_2d(g)
    ref = R.Reference()
    ref.set_backend("b200")
    os.environ["STMQR_B200_CACHE_PLAN"] = "1"
    try:
        A = ref.csc_from_arrays(m, n, p, i, x)
        tol = ref.default_tol(A)
        QR = ref.sparseqr(A, 2, tol, grain=1.0)
        info = ref.qr_info(QR)
        assert int(info["rank"]) == n and int(info["n1cols"]) == 0
        res = ref.check_error(A, QR)
        assert res <= 1e-8, res
        sym = ref.symbolic(QR)
        a = ref.numeric(QR, sym)
        assert np.isfinite(a.stack[: a.rh_size]).all() and np.isfinite(a.HTau).all()
        stack1, tau1, hii1 = a.stack[: a.rh_size].copy(), a.HTau.copy(), a.Hii.copy()
        ref.refactorize(A, QR)                                   # same symbolic object, new numeric
        b = ref.numeric(QR, sym)
        assert b.rh_size == a.rh_size and np.array_equal(b.stack[: b.rh_size], stack1)       # deterministic
        assert np.array_equal(b.HTau, tau1) and np.array_equal(b.Hii, hii1)
        A2 = ref.csc_from_arrays(m, n, p, i, 4.0 * x)
        ref.refactorize(A2, QR)                                  # (tol is stored in the QR object: unchanged,
        c = ref.numeric(QR, sym)                                 #  no pivot of this matrix is anywhere near it)
        assert np.array_equal(c.HStair, a.HStair) and np.array_equal(c.Hii, hii1)
        # R scales by 4 exactly, Householder vectors and tau do not change: compare R rows front by front
        assert np.array_equal(c.HTau, tau1)
        d = R.compare_R(sym, c, sq.Numeric(**{**b.__dict__, "stack": 4.0 * stack1}), 1.0)
        assert d == 0.0, d
        ref.free_qr(QR); ref.free_sparse(A); ref.free_sparse(A2)
    finally:
        os.environ.pop("STMQR_B200_CACHE_PLAN", None)
        ref.set_backend("reference")
        ref.close()


@pytest.mark.parametrize("case", ["lap2d_24_metis", "lap3d_8_metis", "tall_600x150_colamd", "lap2d_16_notol"])
def test_engine_factorization_solves(engine, oracle, case):
    """Functional check of what is NOT comparable entrywise (Householder vectors, tau): the GPU engine's
    packed R+H, pushed through the restated consumers x = E*(R\\(Q'b)) (oracle.qmult / oracle.rsolve, both
    pinned against the reference's QR_qmult / QR_solve on CPU), solves min ||Ax - b||: normal-equation
    residual at rounding level, and Q'(Qx) = x."""
    sym, A, tol, ntol, want = R.load_golden(case)
    got = run_engine(engine, sym, A, tol, ntol)
    assert got.rank == sym.n                                   # these cases are full rank
    S = A.to_scipy()
    rng = np.random.default_rng(8)
    b = rng.standard_normal(sym.m)
    x = oracle.least_squares(sym, got, b)[:, 0]
    r = S @ x - b
    nrm = np.linalg.norm(S.toarray(), 2)
    assert np.linalg.norm(S.T @ r) <= 1e-10 * nrm * max(np.linalg.norm(b), 1.0), case
    X = rng.standard_normal((sym.m, 2))
    back = oracle.qmult(sym, got, R.QR_QTX, oracle.qmult(sym, got, R.QR_QX, X))
    assert np.max(np.abs(back - X)) <= 1e-12

"""GPU tests of the rows next to the path: device Q-apply / R-solve on the resident
factorization, and the values-only refactorization, against the plain-C oracle restatements (pinned to the
reference's QR_qmult / QR_solve on CPU, tests/test_oracle.py) and against the reference's own consumers."""
import os

import numpy as np
import pytest

import refapi as R
import stmqr_b200 as sq

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    e = sq.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def oracle():
    return R.Oracle()


def factor(engine, case):
    sym, A, tol, ntol, want = R.load_golden(case)
    engine.analyze(sym)
    info = engine.factorize(A, tol, ntol)
    return sym, A, tol, ntol, engine.download(info)


@pytest.mark.parametrize("case", R.GOLDEN_CASES)
@pytest.mark.parametrize("nx", [1, 3, 6])
def test_device_qmult_matches_oracle(engine, oracle, case, nx):
    """Y = Q'X and Y = QX computed on the device from the resident R+H blocks == the restated QR_qmult fed
    with the SAME (downloaded) factorization; rank-deficient inputs (dwt_992: 496 dead columns) included."""
    sym, A, tol, ntol, got = factor(engine, case)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((sym.m, nx))
    for method in (R.QR_QTX, R.QR_QX):
        Y = engine.qmult(method, X)
        Yo = oracle.qmult(sym, got, method, X)
        assert np.max(np.abs(Y - Yo)) <= 1e-12 * max(1.0, np.max(np.abs(Yo))), (case, method)
    # Q is orthogonal: Q (Q'X) = X
    back = engine.qmult(R.QR_QX, engine.qmult(R.QR_QTX, X))
    assert np.max(np.abs(back - X)) <= 1e-12 * max(1.0, np.max(np.abs(X)))


@pytest.mark.parametrize("case", R.GOLDEN_CASES)
@pytest.mark.parametrize("nrhs", [1, 5])
def test_device_rsolve_matches_oracle(engine, oracle, case, nrhs):
    """X = R\\B and X = E*(R\\B) on the device == the restated qr_rsolve on the same factorization (dead columns
    get the basic solution 0)."""
    sym, A, tol, ntol, got = factor(engine, case)
    rng = np.random.default_rng(6)
    B = rng.standard_normal((sym.m, nrhs))
    for permuted in (False, True):
        X = engine.rsolve(B, permuted)
        Xo = oracle.rsolve(sym, got, B, permuted)
        scale = max(1.0, np.max(np.abs(Xo)))
        assert np.max(np.abs(X - Xo)) <= 1e-10 * scale, (case, permuted, np.max(np.abs(X - Xo)), scale)
        assert np.array_equal(X == 0, Xo == 0) or got.rank == sym.n


@pytest.mark.parametrize("case", ["lap2d_24_metis", "lap3d_8_metis", "tall_600x150_colamd", "lap2d_16_notol"])
def test_device_least_squares(engine, oracle, case):
    """min ||Ax - b|| entirely on the device (no R+H download): equal to the restated consumers' solution and
    with a normal-equation residual at rounding level."""
    sym, A, tol, ntol, got = factor(engine, case)
    S = A.to_scipy()
    rng = np.random.default_rng(8)
    b = rng.standard_normal((sym.m, 2))
    X, ms = engine.solve_ls(b)
    Xo = oracle.least_squares(sym, got, b)
    assert np.max(np.abs(X - Xo)) <= 1e-10 * max(1.0, np.max(np.abs(Xo)))
    r = S @ X - b
    nrm = np.linalg.norm(S.toarray(), 2)
    assert np.linalg.norm(S.T @ r) <= 1e-10 * nrm * max(np.linalg.norm(b), 1.0)
    assert ms > 0


@pytest.mark.parametrize("name,order", [("t2d_q9", 2), ("epb1", 1), ("dwt_992", 2)])
def test_device_consumers_match_reference(name, order):
    """The same through the reference: its QR_qmult / QR_solve applied to the drop-in's qr_numeric (host) vs the
    engine's device Q-apply / R-solve on an identical factorization (the engine is deterministic)."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref")
    ref = R.Reference()
    A = ref.read_mtx(os.path.join(R.DATA_DIR, name + ".mtx"))
    tol = ref.default_tol(A)
    ref.set_backend("b200")
    QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    assert ref.qr_info(QR)["n1cols"] == 0
    sym = ref.symbolic(QR)
    At, ttol, ntol = ref.tapped()
    e = sq.Engine(0)
    e.analyze(sym)
    e.factorize(At, ttol, ntol)
    rng = np.random.default_rng(9)
    X = rng.standard_normal((sym.m, 2))
    for method in (R.QR_QTX, R.QR_QX):
        Yr = ref.qmult(QR, method, X)
        Y = e.qmult(method, X)
        assert np.max(np.abs(Y - Yr)) <= 1e-11 * max(1.0, np.max(np.abs(Yr))), (name, method)
    C1 = ref.qmult(QR, R.QR_QTX, X)
    Xr = ref.solve(QR, R.QR_RETX_EQUALS_B, C1, sym.n)
    Xd, _ = e.solve_ls(X)
    assert np.max(np.abs(Xd - Xr)) <= 1e-9 * max(1.0, np.max(np.abs(Xr))), name
    e.close()
    ref.free_qr(QR); ref.free_sparse(A)
    ref.set_backend("reference")
    ref.close()


@pytest.mark.parametrize("case", ["lap2d_24_metis", "tall_600x150_colamd", "rankdef_120x80_colamd"])
def test_values_only_refactorization_matches_oracle(engine, oracle, case):
    """New values on the resident pattern (8 bytes per entry uploaded, S built by a scatter through the slot
    map of the first factorization) == the oracle's factorization of the new matrix."""
    sym, A, tol, ntol, _ = R.load_golden(case)
    engine.analyze(sym)
    engine.factorize(A, tol, ntol)                           # leaves the A -> S slot map on the device
    rng = np.random.default_rng(12)
    x2 = A.x * (1.0 + 0.25 * rng.standard_normal(A.x.size))
    A2 = sq.Csc(A.nrow, A.ncol, A.p, A.i, x2)
    info = engine.refactorize_values(x2, tol, ntol)
    got = engine.download(info)
    want = oracle.factorize(sym, A2, tol, ntol)
    R.assert_numeric_parity(sym, A2, got, want, case + " values-only")
    assert got.flops == want.flops
    # and the speculative path of factorize(): same pattern object -> values only, verified on the device
    info3 = engine.factorize(A2, tol, ntol)
    got3 = engine.download(info3)
    assert np.array_equal(got3.stack[: got3.rh_size], got.stack[: got.rh_size])
    # a different pattern with the same counts (the same matrix with the entries of every column stored in
    # another order) must be detected by the device-side comparison and go through the full path
    i2, x3 = A.i.copy(), A.x.copy()
    for j in range(A.ncol):
        seg = slice(int(A.p[j]), int(A.p[j + 1]))
        i2[seg] = np.roll(A.i[seg], 1)
        x3[seg] = np.roll(A.x[seg], 1)
    A3 = sq.Csc(A.nrow, A.ncol, A.p, i2, x3)
    got4 = engine.download(engine.factorize(A3, tol, ntol))
    want4 = oracle.factorize(sym, A, tol, ntol)
    R.assert_numeric_parity(sym, A, got4, want4, case + " reordered entries")


@pytest.mark.parametrize("name,order", [("dwt_992", 2), ("t2d_q9", 2), ("epb1", 1), ("cvxqp3", 1)])
def test_device_rconvert_matches_reference(name, order):
    """R as a compressed-column matrix extracted on the device == the reference's own qr_rcount / qr_rconvert
    applied to the same factorization (the drop-in's qr_numeric on the host; the engine is deterministic, so a
    second engine holds the identical factor): column pointers, row indices and values bit for bit, including
    the order inside a column, the dropped exact zeros and the econ cut; rank-deficient inputs included."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref")
    ref = R.Reference()
    A = ref.read_mtx(os.path.join(R.DATA_DIR, name + ".mtx"))
    tol = ref.default_tol(A)
    ref.set_backend("b200")
    QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    assert ref.qr_info(QR)["n1cols"] == 0
    sym = ref.symbolic(QR)
    At, ttol, ntol = ref.tapped()
    e = sq.Engine(0)
    e.analyze(sym)
    info = e.factorize(At, ttol, ntol)
    for econ in (sym.m, max(1, int(info.rank) // 2)):
        Rp, Ri, Rx = e.rconvert(econ)
        Rp0, Ri0, Rx0 = ref.rconvert(QR, econ, sym.n)
        assert np.array_equal(Rp, Rp0), (name, econ)
        assert np.array_equal(Ri, Ri0), (name, econ)
        assert np.array_equal(Rx, Rx0), (name, econ)
    # the extracted R is the R of the factorization: R'R = (AP)'(AP) on the live columns (full-rank inputs)
    if int(info.rank) == sym.n:
        import scipy.sparse as sp
        Rp, Ri, Rx = e.rconvert()
        Rm = sp.csc_matrix((Rx, Ri, Rp), shape=(sym.m, sym.n))
        S = At.to_scipy()[:, sym.Qfill[: sym.n]] if sym.arrays.get("Qfill") is not None else At.to_scipy()
        G1, G2 = (Rm.T @ Rm), (S.T @ S)
        assert abs(G1 - G2).max() <= 1e-10 * abs(G2).max()
    e.close()
    ref.free_qr(QR); ref.free_sparse(A)
    ref.set_backend("reference")
    ref.close()

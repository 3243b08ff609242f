"""CPU tests of the checker itself (no GPU): the plain-C oracle against the committed golden
vectors produced by the real reference, and -- when the reference build is present -- against
the reference live on the bundled matrices."""
import os

import numpy as np
import pytest

import refapi as R

pytestmark = pytest.mark.skipif(not R.have_oracle(), reason="oracle/libstmqr_oracle.so not built")


@pytest.fixture(scope="module")
def oracle():
    return R.Oracle()


@pytest.mark.parametrize("case", R.GOLDEN_CASES)
def test_oracle_matches_golden(oracle, case):
    sym, A, tol, ntol, want = R.load_golden(case)
    got = oracle.factorize(sym, A, tol, ntol)
    R.assert_numeric_parity(sym, A, got, want, case)
    assert got.flops == want.flops                      # the reference's own FLOP_COUNT, exactly
    assert R.reference_flops(sym, got) == got.flops     # and the outputs-only recomputation


def test_golden_covers_rank_deficiency():
    sym, A, tol, ntol, want = R.load_golden("dwt_992_metis")
    assert want.rank == 496 and int(want.Rdead.sum()) == 496
    sym, A, tol, ntol, want = R.load_golden("rankdef_120x80_colamd")
    assert want.rank < 80 and int(want.Rdead.sum()) == 80 - want.rank


def test_notol_shapes_are_symbolic(oracle):
    """tol = QR_NO_TOL: Hm equals the symbolic Fm exactly."""
    sym, A, tol, ntol, want = R.load_golden("lap2d_16_notol")
    assert tol < 0
    got = oracle.factorize(sym, A, tol, ntol)
    assert np.array_equal(got.Hm, sym.Fm[:sym.nf])


def test_larfg_semantics(oracle):
    """dlarfg conventions of SURVEY.md Appendix B."""
    import ctypes as C
    f = oracle.lib.stmqr_oracle_larfg
    f.restype = C.c_double
    f.argtypes = [C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    # zero sub-column: tau = 0, alpha untouched
    a = C.c_double(-3.0); x = (C.c_double * 3)(0, 0, 0)
    assert f(4, C.byref(a), x) == 0.0 and a.value == -3.0
    # n <= 1
    a = C.c_double(2.0)
    assert f(1, C.byref(a), x) == 0.0 and a.value == 2.0
    # beta = -sign(alpha) * ||[alpha; x]||, 1 <= tau <= 2
    a = C.c_double(3.0); x = (C.c_double * 1)(4.0)
    tau = f(2, C.byref(a), x)
    assert a.value == pytest.approx(-5.0) and 1.0 <= tau <= 2.0
    assert tau == pytest.approx((-5.0 - 3.0) / -5.0)
    assert x[0] == pytest.approx(4.0 / (3.0 + 5.0))


@pytest.mark.skipif(not R.have_reference(), reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("name,order", [("dwt_992", 0), ("t2d_q9", 2), ("bcsstk14", 1), ("lns_3937", 1),
                                        ("ex18", 1), ("reorientation_8", 2)])
def test_oracle_matches_reference_live(oracle, name, order):
    """Pins the restatement against the real reference on the bundled .mtx files, including the
    singleton (Y / freeA) call site (bcsstk14, lns_3937, ex18, reorientation_8)."""
    path = os.path.join(R.DATA_DIR, name + ".mtx")
    if not os.path.exists(path):
        pytest.skip("matrix not bundled")
    ref = R.Reference()
    ref.set_backend("reference")
    A = ref.read_mtx(path)
    tol = ref.default_tol(A)
    QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    sym = ref.symbolic(QR)
    want = ref.numeric(QR, sym)
    At, ttol, ntol = ref.tapped()
    got = oracle.factorize(sym, At, ttol, ntol)
    bad = R.structural_equal(got, want, sym)
    assert not bad, bad
    # R at the north_star tolerance, rank-deficient inputs included; the two inputs on which the
    # reference does not reproduce its own R to that tolerance get the recorded bound (refapi.R_TOL_BY_INPUT)
    d = R.compare_R(sym, got, want, R.a_norm(At))
    print(f"{name}: rank {got.rank}/{sym.n}, tol margin {got.min_tol_margin:.2e}, max |dR| / ||A|| = {d:.3e}")
    assert d <= R.r_tol_for(name), d
    assert got.flops == ref.qr_info(QR)["flopcount"]
    ref.free_qr(QR)
    ref.free_sparse(A)
    ref.close()


@pytest.mark.skipif(not R.have_reference(), reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("name,order", [("dwt_992", 2), ("t2d_q9", 2), ("epb1", 1)])
def test_oracle_qmult_matches_reference(oracle, name, order):
    """The restatement of the consumer next to the path (QR_qmult: Y = Q'X, Y = QX),
    pinned against the reference's own QR_qmult on the reference's own numeric object (matrices without
    column singletons; dwt_992 has 496 dead columns)."""
    path = os.path.join(R.DATA_DIR, name + ".mtx")
    ref = R.Reference()
    ref.set_backend("reference")
    A = ref.read_mtx(path)
    QR = ref.sparseqr(A, order, ref.default_tol(A), grain=1.0, tap=True)
    if ref.qr_info(QR)["n1cols"] != 0:
        pytest.skip("column singletons")
    sym = ref.symbolic(QR)
    num = ref.numeric(QR, sym)
    rng = np.random.default_rng(5)
    X = rng.standard_normal((sym.m, 3))
    for method in (R.QR_QTX, R.QR_QX):
        want = ref.qmult(QR, method, X)
        got = oracle.qmult(sym, num, method, X)
        assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want))), (name, method)
    # Q'(QX) = X through the restatement alone
    back = oracle.qmult(sym, num, R.QR_QTX, oracle.qmult(sym, num, R.QR_QX, X))
    assert np.max(np.abs(back - X)) <= 1e-12
    ref.free_qr(QR); ref.free_sparse(A); ref.close()


@pytest.mark.skipif(not R.have_reference(), reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("name,order", [("dwt_992", 2), ("t2d_q9", 2), ("epb1", 1)])
def test_oracle_rsolve_matches_reference(oracle, name, order):
    """qr_rsolve restated (X = E*(R\\B), basic solution on dead columns), pinned against QR_solve of the
    reference; and the whole solve path x = E*(R\\(Q'b)) from the restatements reproduces the reference's x."""
    path = os.path.join(R.DATA_DIR, name + ".mtx")
    ref = R.Reference()
    ref.set_backend("reference")
    A = ref.read_mtx(path)
    QR = ref.sparseqr(A, order, ref.default_tol(A), grain=1.0, tap=True)
    if ref.qr_info(QR)["n1cols"] != 0:
        pytest.skip("column singletons")
    sym = ref.symbolic(QR)
    num = ref.numeric(QR, sym)
    rng = np.random.default_rng(6)
    B = rng.standard_normal((sym.m, 2))
    for permuted, system in ((True, R.QR_RETX_EQUALS_B), (False, R.QR_RX_EQUALS_B)):
        want = ref.solve(QR, system, B, sym.n)
        got = oracle.rsolve(sym, num, B, permuted=permuted)
        scale = max(1.0, np.max(np.abs(want)))
        assert np.max(np.abs(got - want)) <= 1e-10 * scale, (name, system)
    y = ref.qmult(QR, R.QR_QTX, B)
    want = ref.solve(QR, R.QR_RETX_EQUALS_B, y, sym.n)
    got = oracle.least_squares(sym, num, B)
    assert np.max(np.abs(got - want)) <= 1e-9 * max(1.0, np.max(np.abs(want)))
    ref.free_qr(QR); ref.free_sparse(A); ref.close()

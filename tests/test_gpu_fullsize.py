"""GPU tests at the sizes BASELINE.json names, through the drop-in qr_factorize called by the
reference's own SparseQR(), checked with the reference's own consumers (QR_qmult / QR_solve):

* config 4: banded random tall-sparse 4M x 1M, COLAMD -- least-squares residual and the Q-apply checks
  (Q'(AP) is upper triangular and norm preserving on sampled columns, Q Q' = I)
* config 5: 3-D Laplacian 96^3, METIS -- rank, finiteness, solve residual (qrtest.c check_error)
* mid sizes where the CPU reference still finishes in about a minute: R parity (integer structure bit-exact,
  R to 1e-10 ||A||) at lap3d 64^3 and tall 400k x 100k
(config 2 at full size: tests/test_gpu_parity.py::test_config2_full_size_properties)"""
import os
import time

import numpy as np
import pytest

import refapi as R
import stmqr_b200 as sq
from stmqr_b200 import matrices as M

pytestmark = pytest.mark.gpu


def _dropin(ref, m, n, p, i, x, order):
    A = ref.csc_from_arrays(m, n, p, i, x)
    tol = ref.default_tol(A)
    ref.set_backend("b200")
    QR = ref.sparseqr(A, order, tol, grain=1.0)
    return A, QR


def test_config4_tall_4Mx1M_least_squares():
    if not R.have_reference():
        pytest.skip("needs oracle/_ref for the symbolic analysis and the consumers")
    m, n = 4_000_000, 1_000_000
    mm, nn, p, i, x = M.tall_banded_random(m, n, draws=8, halfwidth=64, seed=4)
    ref = R.Reference()
    try:
        A, QR = _dropin(ref, mm, nn, p, i, x, 1)
        info = ref.qr_info(QR)
        assert int(info["n1cols"]) == 0
        assert int(info["rank"]) == n                              # full column rank
        import scipy.sparse as sp
        S = sp.csc_matrix((x, i, p), shape=(m, n))
        anorm = float(np.sqrt((x * x).sum()))
        rng = np.random.default_rng(5)
        b = rng.standard_normal(m)
        # x = E * (R \ (Q'b)) with the reference's own QR_qmult + QR_solve on the drop-in's qr_numeric
        c = ref.qmult(QR, R.QR_QTX, b)
        xs = ref.solve(QR, R.QR_RETX_EQUALS_B, c, n)[:, 0]
        assert np.isfinite(xs).all()
        r = S @ xs - b
        ne = np.linalg.norm(S.T @ r)
        assert ne <= 1e-10 * anorm * np.linalg.norm(b), (ne, anorm)
        # Q is orthogonal: ||Q'b|| = ||b||, Q (Q'b) = b
        assert abs(np.linalg.norm(c) - np.linalg.norm(b)) <= 1e-10 * np.linalg.norm(b)
        back = ref.qmult(QR, R.QR_QX, c)[:, 0]
        assert np.max(np.abs(back - b)) <= 1e-10 * np.max(np.abs(b))
        # A P = Q R on sampled columns: Q'(A P e_j) is column j of R -- zero below row j, same norm as the column
        sym = ref.symbolic(QR)
        Qfill = sym.Qfill if sym.arrays.get("Qfill") is not None else np.arange(n)
        for j in (0, 1, n // 3, n // 2, n - 2, n - 1):
            a = np.asarray(S[:, int(Qfill[j])].todense()).ravel()
            y = ref.qmult(QR, R.QR_QTX, a)[:, 0]
            assert np.max(np.abs(y[j + 1:])) <= 1e-10 * anorm, (j, np.max(np.abs(y[j + 1:])))
            assert abs(np.linalg.norm(y) - np.linalg.norm(a)) <= 1e-10 * anorm
        ref.free_qr(QR); ref.free_sparse(A)
    finally:
        ref.set_backend("reference")
        ref.dropin_shutdown()
        ref.close()


def test_config5_lap3d_96_solve_residual():
    if not R.have_reference():
        pytest.skip("needs oracle/_ref for the symbolic analysis and the consumers")
    g = 96
    m, n, p, i, x = M.laplacian_3d(g)
    ref = R.Reference()
    try:
        A, QR = _dropin(ref, m, n, p, i, x, 2)
        info = ref.qr_info(QR)
        assert int(info["rank"]) == n and int(info["n1cols"]) == 0
        res = ref.check_error(A, QR)                               # qrtest.c:11-53 on the drop-in's factorization
        assert res <= 1e-9, res
        sym = ref.symbolic(QR)
        num = ref.numeric(QR, sym)
        assert np.isfinite(num.HTau).all()
        # chunked: 3.6e9 doubles
        st = num.stack[: num.rh_size]
        for a in range(0, st.size, 1 << 28):
            assert np.isfinite(st[a: a + (1 << 28)]).all()
        assert (num.Hm[: sym.nf] <= sym.Fm[: sym.nf]).all()        # actual front heights within the symbolic bounds
        assert int(num.Hr[: sym.nf].sum()) == n
        # the reference's FLOP_COUNT (:1571) as the drop-in reported it, read BEFORE check_error (QR_solve adds its own
        # FLOP_COUNTs to cc->SPQR_flopcount, SparseQR.c:2424-2489), against the count recomputed from HStair
        assert info["flopcount"] == R.reference_flops(sym, num)
        ref.free_qr(QR); ref.free_sparse(A)
    finally:
        ref.set_backend("reference")
        ref.dropin_shutdown()
        ref.close()


@pytest.mark.parametrize("gen,order", [(("lap3d", 64), 2), (("tall", 400000, 100000), 1)])
def test_r_parity_against_cpu_reference_mid_size(gen, order):
    """The largest inputs on which the CPU reference still runs inside a test: its own qr_factorize (TPSM tree tasks
    on all cores) vs the drop-in on the same symbolic analysis."""
    if not R.have_reference():
        pytest.skip("needs oracle/_ref")
    if gen[0] == "lap3d":
        m, n, p, i, x = M.laplacian_3d(gen[1])
    else:
        m, n, p, i, x = M.tall_banded_random(gen[1], gen[2], draws=8, halfwidth=64, seed=4)
    cores = os.cpu_count() or 1
    ref = R.Reference()
    try:
        A = ref.csc_from_arrays(m, n, p, i, x)
        tol = ref.default_tol(A)
        ref.set_backend("b200")
        QRg = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
        symg = ref.symbolic(QRg); numg = ref.numeric(QRg, symg)
        At, _, _ = ref.tapped()
        ref.set_backend("reference")
        t0 = time.time()
        if gen[0] == "lap3d":
            # TPSM tree tasks on all cores (qrtest.c:144); the pool is sized far above ntasks (the
            # reference's scheduler deadlocks when ntasks exceeds its pool)
            QRc = ref.sparseqr(A, order, tol, grain=2.0 * cores, pool=512, blas_threads=1)
        else:
            # the banded tall matrix has a chain-like etree: with 2*cores grain it is cut into more tasks than any
            # safe pool size on a 16-core host (the run hangs inside TPSM), so the CPU reference runs its serial
            # etree with threaded BLAS here
            QRc = ref.sparseqr(A, order, tol, grain=1.0, pool=0, blas_threads=cores)
        t_cpu = time.time() - t0
        symc = ref.symbolic(QRc); numc = ref.numeric(QRc, symc)
        assert not R.structural_equal(numg, numc, symg)
        d = R.compare_R(symg, numg, numc, R.a_norm(At))
        print(f"{gen}: CPU reference {ref.qr_info(QRc)['fac_seconds']:.1f} s numeric ({t_cpu:.1f} s total), "
              f"max |dR| / ||A|| = {d:.2e}")
        assert d <= R.R_TOL, d
        ref.free_qr(QRg); ref.free_qr(QRc); ref.free_sparse(A)
    finally:
        ref.set_backend("reference")
        ref.dropin_shutdown()
        ref.close()

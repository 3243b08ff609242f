"""Oracle-vs-oracle noise floor of R: the CPU reference run serially, with TPSM tree tasks and with
threaded BLAS on the same input; prints max |dR| / ||A|| after per-row sign normalisation.  Sets the bound used by
tests/test_gpu_parity.py::test_dropin_through_reference_api for the rank-deficient inputs (cvxqp3, dwt_992)."""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'stm-multifrontal-qr-factorization-empowered-by-gcn_b200', 'py'))
import numpy as np, refapi as R
ref = R.Reference(); ref.set_backend("reference")
for name, order in [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]] or (("cvxqp3",1),("dwt_992",2)):
    A = ref.read_mtx(os.path.join(R.DATA_DIR, name+".mtx"))
    tol = ref.default_tol(A)
    t=time.time()
    Q1 = ref.sparseqr(A, order, tol, grain=1.0, blas_threads=1, tap=True)
    s1 = ref.symbolic(Q1); n1 = ref.numeric(Q1, s1); At,_,_ = ref.tapped()
    Q2 = ref.sparseqr(A, order, tol, grain=16.0, pool=64, blas_threads=1)
    s2 = ref.symbolic(Q2); n2 = ref.numeric(Q2, s2)
    Q3 = ref.sparseqr(A, order, tol, grain=1.0, blas_threads=8)
    s3 = ref.symbolic(Q3); n3 = ref.numeric(Q3, s3)
    print(name, "rank", n1.rank, n2.rank, n3.rank, "struct diff", R.structural_equal(n1,n2,s1), R.structural_equal(n1,n3,s1))
    print(name, "serial vs tasks d =", R.compare_R(s1,n1,n2,R.a_norm(At)), "serial 1thr vs 8thr d =", R.compare_R(s1,n1,n3,R.a_norm(At)), "|A|", R.a_norm(At), time.time()-t, flush=True)

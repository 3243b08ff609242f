"""debug helper (not a test): run one golden case through the engine and print parity."""
import sys, faulthandler
faulthandler.enable()
import refapi as R
import stmqr_b200 as sq
case = sys.argv[1] if len(sys.argv) > 1 else "lap2d_24_metis"
sym, A, tol, ntol, want = R.load_golden(case)
print("create", flush=True)
e = sq.Engine(0)
if len(sys.argv) > 2:
    e.lib.stmqr_b200_set_options(e.h, __import__("ctypes").byref(sq.Options(0, 0, 0, int(sys.argv[2]))))
print("analyze", flush=True)
e.analyze(sym)
print("factorize", flush=True)
info = e.factorize(A, tol, ntol)
print("download", flush=True)
got = e.download(info)
print("rank", got.rank, want.rank, "bad:", R.structural_equal(got, want, sym))
print("dR", R.compare_R(sym, got, want, R.a_norm(A)))

"""debug helper (not a test): solve residual (qrtest.c check_error) of a bench workload through the drop-in,
optionally next to the CPU reference."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
wl = sys.argv[1]
with_cpu = len(sys.argv) > 2 and sys.argv[2] == "cpu"
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
info = ref.qr_info(QR)
t = time.time(); res = ref.check_error(A, QR); t = time.time() - t
print(wl, "B200 drop-in: rank", int(info["rank"]), "check_error res = %.3e" % res, "(%.1f s)" % t, flush=True)
sym = ref.symbolic(QR)
num = ref.numeric(QR, sym)
print("non-finite entries in R+H:", int((~np.isfinite(num.stack[: num.rh_size])).sum()), "in HTau:", int((~np.isfinite(num.HTau)).sum()), flush=True)
if with_cpu:
    ref.set_backend("reference")
    _, mat, order = bench.make_workload(wl)
    cores = os.cpu_count() or 1
    QRc = ref.sparseqr(A, order, tol, grain=2.0 * cores, pool=128, blas_threads=1)
    resc = ref.check_error(A, QRc)
    numc = ref.numeric(QRc, ref.symbolic(QRc))
    print(wl, "CPU reference: rank", int(ref.qr_info(QRc)["rank"]), "check_error res = %.3e" % resc, "fac s %.1f" % ref.qr_info(QRc)["fac_seconds"], flush=True)
    print("integer structure equal:", not R.structural_equal(num, numc, sym), "max |dR|/|A|: %.2e" % R.compare_R(sym, num, numc, R.a_norm(ref.tapped()[0] if False else R.Reference.csc_to_numpy(ref, A))), flush=True)

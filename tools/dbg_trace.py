"""debug helper (not a test): per-launch trace of one factorization of a bench workload."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "lap2d_1024"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/trace.csv"
reserved = int(sys.argv[3]) if len(sys.argv) > 3 else 0
os.environ["STMQR_B200_CACHE_PLAN"] = "1"
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
sym = ref.symbolic(QR)
At, ttol, ntol = ref.tapped()
ref.dropin_shutdown()
import stmqr_b200 as sq, ctypes as C
eng = sq.Engine(0)
def opts(prof):
    eng.lib.stmqr_b200_set_options(eng.h, C.byref(sq.Options(0, 0, prof, reserved)))
opts(0)
eng.analyze(sym)
eng.upload_matrix(At)
for _ in range(2):
    info = eng.factorize_resident(ttol, ntol)
print("resident ms", eng.stats().ms_numeric, "launches", eng.stats().launches)
opts(1)
eng.factorize_resident(ttol, ntol)
os.environ["STMQR_B200_TRACE"] = out
eng.factorize_resident(ttol, ntol)
s = eng.stats()
print("profiled ms", s.ms_numeric, dict(zip(sq.KERNEL_CLASSES, [round(x, 3) for x in s.ms_class])))

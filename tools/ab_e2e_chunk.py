"""debug helper (not a test): end-to-end time of the drop-in qr_factorize on one workload under several chunk sizes
of the download pipeline (STMQR_B200_D2H_CHUNK_MB), one symbolic analysis.  usage: ab_e2e_chunk.py <workload> <MB> ..."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
wl = sys.argv[1]
os.environ["STMQR_B200_CACHE_PLAN"] = "1"
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
for mb in sys.argv[2:]:
    os.environ["STMQR_B200_D2H_CHUNK_MB"] = mb
    ref.dropin_shutdown()                       # the next call creates a new engine handle (reads the knob, re-plans)
    ref.set_backend("b200")
    ts = [ref.refactorize(A, QR) for _ in range(7)]
    print(wl, "chunk MB", mb, "e2e ms", [round(t * 1e3, 1) for t in ts], "median of last 5 %.1f" % (np.median(ts[2:]) * 1e3), flush=True)

"""debug helper (not a test): symbolic analysis + ONE factorization through the drop-in (for ncu)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "lap2d_512"
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
print(setup)

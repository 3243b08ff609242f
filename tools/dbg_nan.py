"""debug helper (not a test): factorize a bench workload on the GPU and report fronts whose packed R+H has NaN/Inf."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
wl = sys.argv[1]
os.environ["STMQR_B200_CACHE_PLAN"] = "1"
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
sym = ref.symbolic(QR)
At, ttol, ntol = ref.tapped()
ref.dropin_shutdown()
import stmqr_b200 as sq
eng = sq.Engine(0)
eng.set_options(flags=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
eng.analyze(sym)
info = eng.factorize(At, ttol, ntol)
num = eng.download(info)
st = num.stack[: info.rh_size]
bad = ~np.isfinite(st)
print(wl, "rh_size", info.rh_size, "non-finite entries", int(bad.sum()), "rank", info.rank, flush=True)
if bad.any():
    idx = np.flatnonzero(bad)
    order = np.argsort(num.Roff[: sym.nf])
    starts = num.Roff[: sym.nf][order]
    which = order[np.searchsorted(starts, idx[[0, len(idx) // 2, -1]], side="right") - 1]
    fn = np.diff(sym.Rp[: sym.nf + 1])
    fronts = np.unique(order[np.searchsorted(starts, idx[:: max(1, len(idx) // 100000)], side="right") - 1])
    print("fronts with non-finite entries:", len(fronts), "first few (f, Hm, fn):",
          [(int(f), int(num.Hm[f]), int(fn[f])) for f in fronts[:8]])
    f = int(fronts[0])
    off = int(num.Roff[f])
    loc = idx[(idx >= off)][0] - off
    print("first front", f, "Hm", int(num.Hm[f]), "fn", int(fn[f]), "fp", int(sym.Super[f + 1] - sym.Super[f]),
          "first bad offset in block", int(loc), "HTau nan count", int(np.isnan(num.HTau).sum()))

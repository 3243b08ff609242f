"""debug helper (not a test): which engine options make HTau non-finite on a bench workload."""
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
for spec in sys.argv[2:]:
    flags, _, env = spec.partition(":")
    kv = dict(x.split("=") for x in env.split(",") if x)
    os.environ.update(kv)
    eng = sq.Engine(0)
    for k in kv:
        os.environ.pop(k)
    eng.set_options(flags=int(flags))
    eng.analyze(sym)
    info = eng.factorize(At, ttol, ntol)
    num = eng.download(info, stack=np.empty(1))
    bad = ~np.isfinite(num.HTau[: sym.rjsize])
    fn = np.diff(sym.Rp[: sym.nf + 1])
    first = None
    if bad.any():
        p = int(np.flatnonzero(bad)[0])
        f = int(np.searchsorted(sym.Rp[: sym.nf + 1], p, side="right") - 1)
        first = (f, int(num.Hm[f]), int(fn[f]), p - int(sym.Rp[f]))
    Fm = np.asarray(sym.Fm[: sym.nf]); Cmb = np.asarray(sym.Cm[: sym.nf])
    fp = np.diff(sym.Super[: sym.nf + 1])
    cm = np.minimum(num.Hm[: sym.nf] - num.Hr[: sym.nf], fn - fp)
    print("max(Hm - Fm_bound)", int((num.Hm[: sym.nf] - Fm).max()), "max(cm - Cm_bound)", int((cm - np.minimum(Cmb, fn - fp)).max()),
          "fronts with Hm > bound", int((num.Hm[: sym.nf] > Fm).sum()), "with cm > bound", int((cm > np.minimum(Cmb, fn - fp)).sum()))
    if first:
        f = first[0]
        kids = sym.Child[int(sym.Childp[f]): int(sym.Childp[f + 1])]
        print("front", f, "fp", int(fp[f]), "children", [(int(c), int(num.Hm[c]), int(fn[c]), int(fp[c]), int(Fm[c]), int(Cmb[c]), int(cm[c])) for c in kids])
    print(wl, "flags", flags, env, "non-finite HTau", int(bad.sum()), "first (front, Hm, fn, col)", first,
          "ms", round(eng.stats().ms_numeric, 1), flush=True)
    eng.close()

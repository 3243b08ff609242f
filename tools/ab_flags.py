"""debug helper (not a test): resident numeric time of ONE workload under several option sets, with a single
symbolic analysis.  usage: ab_flags.py <workload> <flags>[:ENV=V,ENV=V] ...   (flags = stmqr_options.reserved)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
wl = sys.argv[1]
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
sym = ref.symbolic(QR)
At, ttol, ntol = ref.tapped()
ref.dropin_shutdown()
import stmqr_b200 as sq
base = None
for spec in sys.argv[2:]:
    fl, _, env = spec.partition(":")
    kv = dict(x.split("=") for x in env.split(",") if x)
    os.environ.update(kv)
    eng = sq.Engine(0)
    eng.set_options(flags=int(fl, 0))
    eng.analyze(sym)
    eng.upload_matrix(At)
    ms = []
    for _ in range(4):
        info = eng.factorize_resident(ttol, ntol)
        ms.append(eng.stats().ms_numeric)
    num = eng.download(info)
    if base is None:
        base = num
        same = "base"
    else:
        same = "ints equal: %s, max|dR|/|A| %.2e" % (not R.structural_equal(num, base, sym), R.compare_R(sym, num, base, R.a_norm(At)))
    print(wl, "flags", fl, env, "resident ms", [round(x, 2) for x in ms], "GF/s %.0f" % (info.flops / min(ms) * 1e-6),
          "rank", info.rank, same, flush=True)
    eng.close()
    for k in kv:
        os.environ.pop(k, None)

"""debug helper (not a test): resident numeric time of a bench workload under the current env knobs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
wl = sys.argv[1]
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
os.environ["STMQR_B200_CACHE_PLAN"] = "1"
R, ref, A, QR, tol, desc, setup = bench.host_setup(wl, "b200")
sym = ref.symbolic(QR)
At, ttol, ntol = ref.tapped()
ref.dropin_shutdown()
import stmqr_b200 as sq
for env in sys.argv[3:]:
    kv = dict(x.split("=") for x in env.split(",") if x)
    os.environ.update(kv)
    # wide_min_rows is read once per process: only GRID_ROWS / FLAGS vary here
    eng = sq.Engine(0)
    eng.set_options(flags=flags)
    eng.analyze(sym)
    eng.upload_matrix(At)
    ms = []
    for _ in range(4):
        info = eng.factorize_resident(ttol, ntol)
        ms.append(eng.stats().ms_numeric)
    print(wl, "flags", flags, env, "resident ms", [round(x, 1) for x in ms], "GF/s %.0f" % (info.flops / min(ms) * 1e-6), flush=True)
    eng.close()

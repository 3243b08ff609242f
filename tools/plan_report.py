"""Plan a bench workload WITHOUT a GPU and without any numeric work: the reference's symbolic phase (stopped at its
qr_factorize call) + the host-only planner.  Prints arena sizes, the saving of the recycled contribution-block
arena, and the same per part of a partitioned tree.   usage: plan_report.py <workload> [nparts ...]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
import refapi as R
import stmqr_b200 as sq

wl = sys.argv[1]
desc, mat, order = bench.make_workload(wl)
ref = R.Reference()
A = ref.read_mtx(os.path.join(R.DATA_DIR, mat[1] + ".mtx")) if mat[0] == "mtx" else ref.csc_from_arrays(*mat)
t = time.time()
sym = ref.analyze_only(A, order, ref.default_tol(A))
print(f"{wl}: symbolic phase {time.time() - t:.1f} s, nf {sym.nf}, maxfn {sym.maxfn}, rjsize {sym.rjsize}", flush=True)
GB = 8 / 2 ** 30
for nparts in [1] + [int(x) for x in sys.argv[2:]]:
    owner, is_top = sq.partition_fronts(sym, nparts)
    for part in range(nparts):
        p = sq.Planner()
        t = time.time()
        p.analyze(sym)
        tp = time.time() - t
        if nparts > 1:
            p.set_partition(nparts, part, owner, is_top)
        info, coff, csize, level = p.plan_info()
        print(f"  parts {nparts} part {part}: levels {info.nlevels}, F {info.F_doubles * GB:.2f} GB, C {info.C_doubles * GB:.2f} GB "
              f"(unrecycled {info.C_doubles_unrecycled * GB:.2f} GB), R+H bound {info.R_doubles * GB:.2f} GB, "
              f"device total {info.device_bytes / 2 ** 30:.2f} GB, fronts owned {(owner == part).sum()}, "
              f"top fronts {int(is_top.sum())}, host plan {tp * 1e3:.0f} ms", flush=True)
        p.close()

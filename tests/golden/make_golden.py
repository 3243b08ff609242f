"""Generates the committed golden fixtures from the REAL reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference).  Run in the build container:

    python tests/golden/make_golden.py

Each fixture holds: the matrix actually handed to qr_factorize (A, or the singleton-pruned Y),
tol/ntol, every qr_symbolic array (reference's qr_analyze, reused unchanged by the product) and
the reference's qr_numeric (integer structure + packed R+H stack).  The tests compare both the
plain-C oracle (CPU) and the CUDA engine (GPU) against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import refapi as R  # noqa: E402
from stmqr_b200 import matrices as M  # noqa: E402

CASES = [
    # name, source, ordering arg (qrtest convention: 0 AMD 1 COLAMD 2 METIS 3 NESDIS), tol mode
    ("dwt_992_metis", ("mtx", "dwt_992"), 2, "default"),      # BASELINE.json config[0]; rank 496/992
    ("dwt_992_colamd", ("mtx", "dwt_992"), 1, "default"),
    ("lap2d_24_metis", ("lap2d", 24), 2, "default"),          # miniature of config[1]
    ("lap3d_8_metis", ("lap3d", 8), 2, "default"),            # miniature of config[4]
    ("tall_600x150_colamd", ("tall", 600, 150), 1, "default"),  # miniature of config[3]
    ("lap2d_16_notol", ("lap2d", 16), 1, "notol"),            # tol = QR_NO_TOL: symbolic Fm/Cm exact
    ("rankdef_120x80_colamd", ("rankdef", 120, 80), 1, "default"),
]


def build_matrix(ref, src):
    if src[0] == "mtx":
        return ref.read_mtx(os.path.join(R.DATA_DIR, src[1] + ".mtx"))
    if src[0] == "lap2d":
        m, n, p, i, x = M.laplacian_2d(src[1])
    elif src[0] == "lap3d":
        m, n, p, i, x = M.laplacian_3d(src[1])
    elif src[0] == "tall":
        m, n, p, i, x = M.tall_banded_random(src[1], src[2], draws=6, halfwidth=8, seed=4)
    elif src[0] == "rankdef":
        m, n, p, i, x = M.random_sparse(src[1], src[2], 0.06, seed=7, rank_deficient_cols=12)
    else:
        raise ValueError(src)
    return ref.csc_from_arrays(m, n, p, i, x)


def main():
    ref = R.Reference()
    ref.set_backend("reference")
    for name, src, order, tolmode in CASES:
        A = build_matrix(ref, src)
        tol = ref.default_tol(A) if tolmode == "default" else -1.0
        QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
        info = ref.qr_info(QR)
        sym = ref.symbolic(QR)
        num = ref.numeric(QR, sym)
        At, ttol, ntol = ref.tapped()
        res = ref.check_error(A, QR) if At.nrow == At.ncol and info["n1cols"] == 0 else -1.0
        out = {("s_" + k): np.int64(v) for k, v in sym.scalars.items()}
        out.update({("a_" + k): v for k, v in sym.arrays.items()})
        out.update(A_m=np.int64(At.nrow), A_n=np.int64(At.ncol), A_p=At.p, A_i=At.i, A_x=At.x,
                   tol=np.float64(ttol), ntol=np.int64(ntol), ordering=np.int64(order),
                   ref_rank=np.int64(num.rank), ref_rank1=np.int64(num.rank1),
                   ref_maxfrank=np.int64(num.maxfrank), ref_maxfm=np.int64(num.maxfm),
                   ref_flops=np.float64(info["flopcount"]), ref_res=np.float64(res),
                   ref_Rdead=num.Rdead, ref_HStair=num.HStair, ref_HTau=num.HTau,
                   ref_Hii=np.where(R.valid_hii_mask(sym, num), num.Hii, -1), ref_Hm=num.Hm, ref_Hr=num.Hr,
                   ref_HPinv=num.HPinv, ref_Roff=num.Roff, ref_stack=num.stack)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {At.nrow}x{At.ncol} nnz {At.nnz} nf {sym.nf} rank {num.rank} flops {info['flopcount']:.4g} "
              f"res {res:.2e} -> {os.path.getsize(path)/1024:.0f} KiB")
        ref.free_qr(QR)
        ref.free_sparse(A)
    ref.close()


if __name__ == "__main__":
    main()

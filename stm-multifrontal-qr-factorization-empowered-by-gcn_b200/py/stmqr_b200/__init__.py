"""stmqr_b200 -- thin ctypes binding of the C ABI in include/stmqr_b200.h.

The product is the CUDA library ``lib/libstmqr_b200.so`` (hand-written sm_100a kernels behind
an ``extern "C"`` boundary) plus the C drop-in ``lib/libstmqr_dropin.so`` that replaces the
reference's ``qr_factorize`` (STMMQR/src/qr/SparseQR_factorize.c:222).  This module only
marshals numpy arrays into the plain views of the header; it is used by ``bench.py``, by the
tests and by ``__graft_entry__.smoke()``.  There is no CPU fallback: importing works without a
GPU (so that symbol/ABI checks can run), every compute call raises if the library or an
sm_100 device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_PKG_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
LIB_DIR = os.path.join(_PKG_ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libstmqr_b200.so")
DROPIN_PATH = os.path.join(LIB_DIR, "libstmqr_dropin.so")

_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)

STMQR_OK = 0
ERRORS = {-1: "NO_DEVICE", -2: "OUT_OF_MEMORY", -3: "TOO_LARGE", -4: "INVALID", -5: "CUDA"}


class SymbolicView(C.Structure):
    """stmqr_symbolic_view (mirrors qr_symbolic, STMMQR/include/SparseQR_struct.h:26-137)."""
    _fields_ = [(k, C.c_int64) for k in
                ("m", "n", "anz", "nf", "maxfn", "rjsize", "hisize", "do_rank_detection", "keepH")] + \
               [(k, _i64p) for k in
                ("Sp", "Sj", "Qfill", "PLinv", "Sleft", "Parent", "Child", "Childp", "Super",
                 "Rp", "Rj", "Post", "Hip", "Fm", "Cm")]


SYM_ARRAYS = {  # name -> length as a function of the scalars
    "Sp": lambda s: s["m"] + 1, "Sj": lambda s: s["anz"], "Qfill": lambda s: s["n"],
    "PLinv": lambda s: s["m"], "Sleft": lambda s: s["n"] + 2, "Parent": lambda s: s["nf"] + 1,
    "Child": lambda s: s["nf"] + 1, "Childp": lambda s: s["nf"] + 2, "Super": lambda s: s["nf"] + 1,
    "Rp": lambda s: s["nf"] + 1, "Rj": lambda s: s["rjsize"], "Post": lambda s: s["nf"] + 1,
    "Hip": lambda s: s["nf"] + 1, "Fm": lambda s: s["nf"] + 1, "Cm": lambda s: s["nf"] + 1,
}
SYM_SCALARS = ("m", "n", "anz", "nf", "maxfn", "rjsize", "hisize", "do_rank_detection", "keepH")


class CscView(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int64), ("nzmax", C.c_int64),
                ("p", _i64p), ("i", _i64p), ("x", _f64p)]


class NumericInfo(C.Structure):
    _fields_ = [("rank", C.c_int64), ("rank1", C.c_int64), ("maxfrank", C.c_int64),
                ("maxfm", C.c_int64), ("rh_size", C.c_int64), ("flops", C.c_double)]


class NumericView(C.Structure):
    _fields_ = [("stack", _f64p), ("Roff", _i64p), ("Rdead", C.c_void_p), ("HStair", _i64p),
                ("HTau", _f64p), ("Hii", _i64p), ("Hm", _i64p), ("Hr", _i64p), ("HPinv", _i64p)]


class Stats(C.Structure):
    _fields_ = [("ms_plan", C.c_double), ("ms_h2d", C.c_double), ("ms_numeric", C.c_double),
                ("ms_d2h", C.c_double), ("ms_assemble", C.c_double), ("ms_front", C.c_double),
                ("bytes_assemble", C.c_double), ("flops", C.c_double),
                ("launches", C.c_int64), ("nlevels", C.c_int64), ("nf_small", C.c_int64),
                ("nf_big", C.c_int64), ("device_bytes", C.c_int64),
                ("ms_class", C.c_double * 8), ("launches_class", C.c_int64 * 8), ("update_flops", C.c_double)]


KERNEL_CLASSES = ("build_S", "front_setup", "assemble", "panel", "update", "finish_alloc", "pack", "hpinv_misc")


class Options(C.Structure):
    _fields_ = [("panel", C.c_int32), ("small_elems", C.c_int32), ("profile_phases", C.c_int32),
                ("reserved", C.c_int32)]


class FrontRegions(C.Structure):
    _fields_ = [("C", C.c_void_p), ("C_doubles", C.c_int64), ("Hii", C.c_void_p), ("Hii_ints", C.c_int64),
                ("cm", C.c_int64), ("hr", C.c_int64), ("hm", C.c_int64)]


class PlanInfo(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("nlevels", "F_doubles", "C_doubles", "C_doubles_unrecycled", "R_doubles",
                                         "device_bytes", "nparts", "mypart")]


ARRAY_HM, ARRAY_HR, ARRAY_CM, ARRAY_RDEAD, ARRAY_W = range(5)

EXPORTS = (
    "stmqr_b200_device_count", "stmqr_b200_create", "stmqr_b200_destroy", "stmqr_b200_set_options",
    "stmqr_b200_analyze", "stmqr_b200_upload_matrix", "stmqr_b200_factorize_resident",
    "stmqr_b200_download", "stmqr_b200_factorize", "stmqr_b200_get_stats", "stmqr_b200_last_error",
    "stmqr_b200_set_debug_capture", "stmqr_b200_get_front", "stmqr_b200_measure_fp64_peak",
    "stmqr_b200_factorize_begin", "stmqr_b200_factorize_levels", "stmqr_b200_factorize_hpinv_a",
    "stmqr_b200_factorize_hpinv_b", "stmqr_b200_sync", "stmqr_b200_partition_fronts",
    "stmqr_b200_set_partition", "stmqr_b200_device_array", "stmqr_b200_front_regions",
    "stmqr_b200_rh_bound", "stmqr_b200_factorize_streamed", "stmqr_b200_stream_begin", "stmqr_b200_stream_end",
    "stmqr_b200_create_planner", "stmqr_b200_plan_info",
    "stmqr_b200_upload_values", "stmqr_b200_refactorize_values",
    "stmqr_b200_qmult", "stmqr_b200_rsolve", "stmqr_b200_solve_ls", "stmqr_b200_rcount", "stmqr_b200_rconvert",
    "stmqr_b200_map_fronts", "stmqr_b200_set_ownership", "stmqr_b200_nccl_unique_id", "stmqr_b200_comm_init",
    "stmqr_b200_peer_group_create", "stmqr_b200_peer_group_destroy", "stmqr_b200_factorize_dist",
    "stmqr_b200_factorize_multi", "stmqr_b200_factorize_multi_ex", "stmqr_b200_gather_outputs",
    "stmqr_b200_coop_chunks",
)

_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen the CUDA library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("STMQR_B200_LIB", path)       # (instrumented builds of the same library, tools/)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the B200 engine has no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.stmqr_b200_device_count.restype = C.c_int
    lib.stmqr_b200_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.stmqr_b200_destroy.argtypes = [C.c_void_p]
    lib.stmqr_b200_destroy.restype = None
    lib.stmqr_b200_set_options.argtypes = [C.c_void_p, C.POINTER(Options)]
    lib.stmqr_b200_analyze.argtypes = [C.c_void_p, C.POINTER(SymbolicView)]
    lib.stmqr_b200_upload_matrix.argtypes = [C.c_void_p, C.POINTER(CscView)]
    lib.stmqr_b200_factorize_resident.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.POINTER(NumericInfo)]
    lib.stmqr_b200_download.argtypes = [C.c_void_p, C.POINTER(NumericView)]
    lib.stmqr_b200_factorize.argtypes = [C.c_void_p, C.POINTER(CscView), C.c_double, C.c_int64,
                                         C.POINTER(NumericInfo)]
    lib.stmqr_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.stmqr_b200_last_error.argtypes = [C.c_void_p]
    lib.stmqr_b200_last_error.restype = C.c_char_p
    lib.stmqr_b200_measure_fp64_peak.argtypes = [C.c_void_p, _f64p, _f64p]
    lib.stmqr_b200_set_debug_capture.argtypes = [C.c_void_p, C.c_int]
    lib.stmqr_b200_get_front.argtypes = [C.c_void_p, C.c_int64, C.c_int, _f64p, C.c_int64, _i64p, _i64p]
    lib.stmqr_b200_factorize_begin.argtypes = [C.c_void_p, C.c_double, C.c_int64]
    lib.stmqr_b200_factorize_levels.argtypes = [C.c_void_p, C.c_int]
    lib.stmqr_b200_factorize_hpinv_a.argtypes = [C.c_void_p]
    lib.stmqr_b200_factorize_hpinv_b.argtypes = [C.c_void_p, C.POINTER(NumericInfo)]
    lib.stmqr_b200_sync.argtypes = [C.c_void_p]
    lib.stmqr_b200_partition_fronts.argtypes = [C.POINTER(SymbolicView), C.c_int, C.POINTER(C.c_int32),
                                                C.POINTER(C.c_int32)]
    lib.stmqr_b200_set_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                             C.POINTER(C.c_int32)]
    lib.stmqr_b200_device_array.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _i64p,
                                            C.POINTER(C.c_int32)]
    lib.stmqr_b200_front_regions.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                             C.POINTER(FrontRegions)]
    lib.stmqr_b200_stream_begin.argtypes = [C.c_void_p, _f64p, C.c_int64]
    lib.stmqr_b200_stream_end.argtypes = [C.c_void_p]
    lib.stmqr_b200_rh_bound.argtypes = [C.c_void_p, _i64p]
    lib.stmqr_b200_factorize_streamed.argtypes = [C.c_void_p, C.POINTER(CscView), C.c_double, C.c_int64, _f64p,
                                                  C.c_int64, C.POINTER(NumericInfo)]
    lib.stmqr_b200_upload_values.argtypes = [C.c_void_p, _f64p, C.c_int64]
    lib.stmqr_b200_refactorize_values.argtypes = [C.c_void_p, _f64p, C.c_int64, C.c_double, C.c_int64,
                                                  C.POINTER(NumericInfo)]
    lib.stmqr_b200_qmult.argtypes = [C.c_void_p, C.c_int, C.c_int64, _f64p, _f64p]
    lib.stmqr_b200_rsolve.argtypes = [C.c_void_p, C.c_int, C.c_int64, _f64p, _f64p]
    lib.stmqr_b200_solve_ls.argtypes = [C.c_void_p, C.c_int64, _f64p, _f64p, _f64p]
    lib.stmqr_b200_rcount.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p]
    lib.stmqr_b200_rconvert.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p, _f64p]
    lib.stmqr_b200_map_fronts.argtypes = [C.POINTER(SymbolicView), C.c_int, C.POINTER(C.c_int32)]
    lib.stmqr_b200_set_ownership.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    lib.stmqr_b200_nccl_unique_id.argtypes = [C.c_void_p]
    lib.stmqr_b200_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.stmqr_b200_peer_group_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    lib.stmqr_b200_peer_group_destroy.argtypes = [C.c_void_p]
    lib.stmqr_b200_peer_group_destroy.restype = None
    lib.stmqr_b200_factorize_dist.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.POINTER(NumericInfo)]
    lib.stmqr_b200_factorize_multi.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.POINTER(NumericInfo)]
    lib.stmqr_b200_create_planner.argtypes = [C.POINTER(C.c_void_p)]
    lib.stmqr_b200_plan_info.argtypes = [C.c_void_p, C.POINTER(PlanInfo), _i64p, _i64p, C.POINTER(C.c_int32)]
    _lib = lib
    return lib


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


class Symbolic:
    """Owns numpy copies of the qr_symbolic arrays and exposes a SymbolicView over them."""

    def __init__(self, scalars: dict, arrays: dict):
        self.scalars = {k: int(scalars[k]) for k in SYM_SCALARS}
        self.arrays = {}
        self.view = SymbolicView()
        for k in SYM_SCALARS:
            setattr(self.view, k, self.scalars[k])
        for k, lenf in SYM_ARRAYS.items():
            a = arrays.get(k)
            if a is None:
                setattr(self.view, k, None)
                continue
            a, p = _i64(a)
            assert a.shape[0] >= lenf(self.scalars), (k, a.shape, lenf(self.scalars))
            self.arrays[k] = a
            setattr(self.view, k, p)

    @classmethod
    def from_view(cls, v: SymbolicView) -> "Symbolic":
        """Deep-copy a view whose pointers belong to someone else (e.g. the reference's QRsym)."""
        scal = {k: int(getattr(v, k)) for k in SYM_SCALARS}
        arrs = {}
        for k, lenf in SYM_ARRAYS.items():
            p = getattr(v, k)
            if not p:
                arrs[k] = None
            else:
                n = lenf(scal)
                arrs[k] = np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].copy()
        return cls(scal, arrs)

    def __getattr__(self, k):
        if k in ("scalars", "arrays", "view"):
            raise AttributeError(k)
        if k in self.scalars:
            return self.scalars[k]
        if k in self.arrays:
            return self.arrays[k]
        raise AttributeError(k)

    def save(self, path):
        np.savez_compressed(path, **{("s_" + k): np.int64(v) for k, v in self.scalars.items()},
                            **{("a_" + k): v for k, v in self.arrays.items()})

    @classmethod
    def load(cls, path):
        z = np.load(path)
        scal = {k[2:]: int(z[k]) for k in z.files if k.startswith("s_")}
        arrs = {k[2:]: z[k] for k in z.files if k.startswith("a_")}
        return cls(scal, arrs)


def partition_fronts(sym: "Symbolic", nparts: int):
    """Host-only, deterministic partition of the etree over nparts GPUs -> (owner[nf], is_top[nf])."""
    lib = load_library()
    owner = np.zeros(max(sym.nf, 1), np.int32)
    top = np.zeros(max(sym.nf, 1), np.int32)
    st = lib.stmqr_b200_partition_fronts(C.byref(sym.view), nparts, owner.ctypes.data_as(C.POINTER(C.c_int32)),
                                         top.ctypes.data_as(C.POINTER(C.c_int32)))
    if st != STMQR_OK:
        raise RuntimeError(f"partition_fronts: {ERRORS.get(st, st)}")
    return owner[:sym.nf], top[:sym.nf]


def map_fronts(sym: "Symbolic", nparts: int) -> np.ndarray:
    """Host-only, deterministic owner of every front over nparts GPUs (subtrees below the cut dealt largest
    first; a front above the cut on the GPU of its heaviest child)."""
    lib = load_library()
    owner = np.zeros(max(sym.nf, 1), np.int32)
    st = lib.stmqr_b200_map_fronts(C.byref(sym.view), nparts, owner.ctypes.data_as(C.POINTER(C.c_int32)))
    if st != STMQR_OK:
        raise RuntimeError(f"map_fronts: {ERRORS.get(st, st)}")
    return owner[:sym.nf]


def coop_chunks(nparts: int, home: int, fn: int) -> np.ndarray:
    """Host-only: owner GPU of every 512-column chunk of a cooperative front with fn columns whose home is
    `home` (the map stmqr_b200_factorize_dist uses)."""
    lib = load_library()
    n = C.c_int64(0)
    lib.stmqr_b200_coop_chunks.argtypes = [C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    st = lib.stmqr_b200_coop_chunks(nparts, home, fn, None, C.byref(n))
    if st != STMQR_OK:
        raise RuntimeError(f"coop_chunks: {ERRORS.get(st, st)}")
    out = np.zeros(max(int(n.value), 1), np.int32)
    st = lib.stmqr_b200_coop_chunks(nparts, home, fn, out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n))
    if st != STMQR_OK:
        raise RuntimeError(f"coop_chunks: {ERRORS.get(st, st)}")
    return out[: int(n.value)]


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    st = lib.stmqr_b200_nccl_unique_id(buf)
    if st != STMQR_OK:
        raise RuntimeError(f"nccl_unique_id: {ERRORS.get(st, st)} (libnccl.so.2 not loadable?)")
    return buf.raw


class PeerGroup:
    """N engines of this process as one multi-GPU group (peer copies, one host thread per engine inside
    factorize()); the engines may sit on different devices or, for tests, on the same one."""

    def __init__(self, engines):
        self.lib = load_library()
        self.engines = list(engines)
        arr = (C.c_void_p * len(self.engines))(*[e.h for e in self.engines])
        self.g = C.c_void_p()
        st = self.lib.stmqr_b200_peer_group_create(arr, len(self.engines), C.byref(self.g))
        if st != STMQR_OK:
            raise EngineError(f"peer_group_create: {ERRORS.get(st, st)}")

    def factorize(self, tol: float, ntol: int):
        infos = (NumericInfo * len(self.engines))()
        st = self.lib.stmqr_b200_factorize_multi(self.g, tol, ntol, infos)
        if st != STMQR_OK:
            msgs = [self.lib.stmqr_b200_last_error(e.h) for e in self.engines]
            raise EngineError(f"factorize_multi: {ERRORS.get(st, st)}: {[m.decode() for m in msgs if m]}")
        return list(infos)

    def close(self):
        if self.g:
            self.lib.stmqr_b200_peer_group_destroy(self.g)
            self.g = C.c_void_p()


class Csc:
    def __init__(self, nrow, ncol, p, i, x):
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.p, pp = _i64(p)
        self.i, pi = _i64(i)
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.view = CscView(self.nrow, self.ncol, int(self.p[-1]), pp, pi, self.x.ctypes.data_as(_f64p))

    @property
    def nnz(self):
        return int(self.p[-1])

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.x, self.i, self.p), shape=(self.nrow, self.ncol))


@dataclass
class Numeric:
    """Host copy of the qr_numeric members (STMMQR/include/SparseQR_struct.h:145-209)."""
    rank: int
    rank1: int
    maxfrank: int
    maxfm: int
    rh_size: int
    flops: float
    stack: np.ndarray
    Roff: np.ndarray
    Rdead: np.ndarray
    HStair: np.ndarray
    HTau: np.ndarray
    Hii: np.ndarray
    Hm: np.ndarray
    Hr: np.ndarray
    HPinv: np.ndarray


class EngineError(RuntimeError):
    pass


class Engine:
    """One handle = one GPU + one symbolic plan."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.h = C.c_void_p()
        st = self.lib.stmqr_b200_create(device, C.byref(self.h))
        if st != STMQR_OK:
            raise EngineError(f"stmqr_b200_create failed: {ERRORS.get(st, st)} (no CPU fallback)")
        self.sym = None

    def _check(self, st, what):
        if st != STMQR_OK:
            msg = self.lib.stmqr_b200_last_error(self.h)
            raise EngineError(f"{what}: {ERRORS.get(st, st)}: {msg.decode() if msg else ''}")

    def set_options(self, panel=0, small_elems=0, profile_phases=0, flags=0):
        """flags (tuning / A-B tests, stmqr_options.reserved): bit 0 no look-ahead, bit 1 no two-level
        (128-column) blocking, bit 2 no k_panel_grid, bit 3 no small-front kernel (set before analyze),
        bit 4 non-persistent K = 128 apply"""
        o = Options(panel, small_elems, profile_phases, flags)
        self._check(self.lib.stmqr_b200_set_options(self.h, C.byref(o)), "set_options")

    def analyze(self, sym: Symbolic):
        self.sym = sym
        self._check(self.lib.stmqr_b200_analyze(self.h, C.byref(sym.view)), "analyze")

    def upload_matrix(self, A: Csc):
        self._check(self.lib.stmqr_b200_upload_matrix(self.h, C.byref(A.view)), "upload_matrix")

    def factorize_resident(self, tol: float, ntol: int) -> NumericInfo:
        info = NumericInfo()
        self._check(self.lib.stmqr_b200_factorize_resident(self.h, tol, ntol, C.byref(info)), "factorize_resident")
        return info

    def factorize(self, A: Csc, tol: float, ntol: int) -> NumericInfo:
        info = NumericInfo()
        self._check(self.lib.stmqr_b200_factorize(self.h, C.byref(A.view), tol, ntol, C.byref(info)), "factorize")
        return info

    def factorize_streamed(self, A: Csc, tol: float, ntol: int):
        """upload + numeric phase with the R+H stack copied to the host level by level while the
        next levels run -> (info, stack[:rh_size]); then download(info, stack=...) for the rest"""
        cap = C.c_int64()
        self._check(self.lib.stmqr_b200_rh_bound(self.h, C.byref(cap)), "rh_bound")
        stack = np.empty(max(int(cap.value), 1), np.float64)
        info = NumericInfo()
        self._check(self.lib.stmqr_b200_factorize_streamed(
            self.h, C.byref(A.view), tol, ntol, stack.ctypes.data_as(_f64p), stack.size, C.byref(info)),
            "factorize_streamed")
        return info, stack[: max(int(info.rh_size), 1)]

    def stream_begin(self):
        """start copying the R+H blocks of finished levels into a host stack (allocated by the bound) while
        the phased numeric calls run -> the stack; call stream_end() after factorize_hpinv_b"""
        cap = C.c_int64()
        self._check(self.lib.stmqr_b200_rh_bound(self.h, C.byref(cap)), "rh_bound")
        if getattr(self, "_stream_stack", None) is None or self._stream_stack.size < max(int(cap.value), 1):
            self._stream_stack = np.empty(max(int(cap.value), 1), np.float64)
        self._check(self.lib.stmqr_b200_stream_begin(self.h, self._stream_stack.ctypes.data_as(_f64p),
                                                     self._stream_stack.size), "stream_begin")
        return self._stream_stack

    def stream_end(self):
        self._check(self.lib.stmqr_b200_stream_end(self.h), "stream_end")

    def download(self, info: NumericInfo, stack=None) -> Numeric:
        s = self.sym
        out = Numeric(int(info.rank), int(info.rank1), int(info.maxfrank), int(info.maxfm),
                      int(info.rh_size), float(info.flops),
                      stack=(stack if stack is not None else np.empty(max(int(info.rh_size), 1), np.float64)),
                      Roff=np.empty(max(s.nf, 1), np.int64), Rdead=np.zeros(max(s.n, 1), np.int8),
                      HStair=np.empty(max(s.rjsize, 1), np.int64), HTau=np.empty(max(s.rjsize, 1), np.float64),
                      Hii=np.empty(max(s.hisize, 1), np.int64), Hm=np.empty(max(s.nf, 1), np.int64),
                      Hr=np.empty(max(s.nf, 1), np.int64), HPinv=np.empty(max(s.m, 1), np.int64))
        v = NumericView()
        v.stack = out.stack.ctypes.data_as(_f64p) if stack is None else None      # NULL: already streamed
        v.Roff = out.Roff.ctypes.data_as(_i64p)
        v.Rdead = out.Rdead.ctypes.data
        v.HStair = out.HStair.ctypes.data_as(_i64p)
        v.HTau = out.HTau.ctypes.data_as(_f64p)
        v.Hii = out.Hii.ctypes.data_as(_i64p)
        v.Hm = out.Hm.ctypes.data_as(_i64p)
        v.Hr = out.Hr.ctypes.data_as(_i64p)
        v.HPinv = out.HPinv.ctypes.data_as(_i64p)
        self._check(self.lib.stmqr_b200_download(self.h, C.byref(v)), "download")
        return out

    # ---- the numeric phase in pieces / several GPUs (see dist.py)
    def factorize_begin(self, tol: float, ntol: int):
        self._check(self.lib.stmqr_b200_factorize_begin(self.h, tol, ntol), "factorize_begin")

    def factorize_levels(self, part: int):
        self._check(self.lib.stmqr_b200_factorize_levels(self.h, part), "factorize_levels")

    def factorize_hpinv_a(self):
        self._check(self.lib.stmqr_b200_factorize_hpinv_a(self.h), "factorize_hpinv_a")

    def factorize_hpinv_b(self) -> NumericInfo:
        info = NumericInfo()
        self._check(self.lib.stmqr_b200_factorize_hpinv_b(self.h, C.byref(info)), "factorize_hpinv_b")
        return info

    def sync(self):
        self._check(self.lib.stmqr_b200_sync(self.h), "sync")

    def set_partition(self, nparts: int, mypart: int, owner, is_top):
        owner = np.ascontiguousarray(owner, np.int32)
        is_top = np.ascontiguousarray(is_top, np.int32)
        self._check(self.lib.stmqr_b200_set_partition(self.h, nparts, mypart,
                                                      owner.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      is_top.ctypes.data_as(C.POINTER(C.c_int32))), "set_partition")

    def set_ownership(self, nparts: int, mypart: int, owner):
        owner = np.ascontiguousarray(owner, np.int32)
        self._check(self.lib.stmqr_b200_set_ownership(self.h, nparts, mypart,
                                                      owner.ctypes.data_as(C.POINTER(C.c_int32))), "set_ownership")

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        self._check(self.lib.stmqr_b200_comm_init(self.h, nranks, rank, C.c_char_p(unique_id)), "comm_init")

    def factorize_dist(self, tol: float, ntol: int) -> NumericInfo:
        info = NumericInfo()
        self._check(self.lib.stmqr_b200_factorize_dist(self.h, tol, ntol, C.byref(info)), "factorize_dist")
        return info

    def device_array(self, which: int):
        """(device pointer, count, element bytes) of one of the arrays merged over the GPUs"""
        p, n, eb = C.c_void_p(), C.c_int64(), C.c_int32()
        self._check(self.lib.stmqr_b200_device_array(self.h, which, C.byref(p), C.byref(n), C.byref(eb)),
                    "device_array")
        return p.value, n.value, eb.value

    def front_regions(self, f: int, cm: int = -1, hr: int = -1, hm: int = -1) -> FrontRegions:
        r = FrontRegions()
        self._check(self.lib.stmqr_b200_front_regions(self.h, f, cm, hr, hm, C.byref(r)), "front_regions")
        return r

    def stats(self) -> Stats:
        s = Stats()
        self._check(self.lib.stmqr_b200_get_stats(self.h, C.byref(s)), "get_stats")
        return s

    def measure_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.stmqr_b200_measure_fp64_peak(self.h, C.byref(a), C.byref(b)), "measure_fp64_peak")
        return a.value, b.value

    def set_debug_capture(self, on=True):
        self._check(self.lib.stmqr_b200_set_debug_capture(self.h, int(on)), "set_debug_capture")

    def get_front(self, f: int, which: int) -> np.ndarray:
        fm, fn = C.c_int64(), C.c_int64()
        cap = int(self.sym.Fm[f]) * int(self.sym.Rp[f + 1] - self.sym.Rp[f])
        buf = np.zeros(max(cap, 1), np.float64)
        self._check(self.lib.stmqr_b200_get_front(self.h, f, which, buf.ctypes.data_as(_f64p), cap,
                                                  C.byref(fm), C.byref(fn)), "get_front")
        return buf[: fm.value * fn.value].reshape((fn.value, fm.value)).T  # column-major -> (fm, fn)

    # ---- values-only refactorization and the device-side consumers (Q-apply, R-solve)
    def refactorize_values(self, Ax, tol: float, ntol: int) -> NumericInfo:
        """new values on the resident pattern (the matrix of the last upload_matrix / factorize)"""
        Ax = np.ascontiguousarray(Ax, np.float64)
        info = NumericInfo()
        self._check(self.lib.stmqr_b200_refactorize_values(self.h, Ax.ctypes.data_as(_f64p), Ax.size, tol, ntol,
                                                           C.byref(info)), "refactorize_values")
        return info

    @staticmethod
    def _cols(X):
        X = np.asfortranarray(X, dtype=np.float64)
        return X.reshape(-1, 1, order="F") if X.ndim == 1 else X

    def qmult(self, method: int, X) -> np.ndarray:
        """Y = Q'X (method 0) or QX (method 1) from the factorization resident on the device"""
        X = self._cols(X)
        Y = np.zeros_like(X, order="F")
        self._check(self.lib.stmqr_b200_qmult(self.h, method, X.shape[1], X.ctypes.data_as(_f64p),
                                              Y.ctypes.data_as(_f64p)), "qmult")
        return Y

    def rsolve(self, B, permuted: bool = True) -> np.ndarray:
        """X = E*(R\\B) (permuted) or R\\B; B is m-by-nrhs"""
        B = self._cols(B)
        X = np.zeros((self.sym.n, B.shape[1]), order="F")
        self._check(self.lib.stmqr_b200_rsolve(self.h, int(permuted), B.shape[1], B.ctypes.data_as(_f64p),
                                               X.ctypes.data_as(_f64p)), "rsolve")
        return X

    def solve_ls(self, B):
        """x = E*(R\\(Q'b)) on the device -> (X n-by-nrhs, device milliseconds incl. the copies of b and x)"""
        B = self._cols(B)
        X = np.zeros((self.sym.n, B.shape[1]), order="F")
        ms = C.c_double()
        self._check(self.lib.stmqr_b200_solve_ls(self.h, B.shape[1], B.ctypes.data_as(_f64p),
                                                 X.ctypes.data_as(_f64p), C.byref(ms)), "solve_ls")
        return X, ms.value

    def rconvert(self, econ: int = None):
        """R of the resident factorization as CSC in the column order of S -> (Rp[n+1], Ri, Rx), extracted on the
        device (qr_rcount / qr_rconvert); only rows < econ (default m)"""
        econ = self.sym.m if econ is None else int(econ)
        Rp = np.zeros(self.sym.n + 1, np.int64)
        nnz = C.c_int64()
        self._check(self.lib.stmqr_b200_rcount(self.h, econ, Rp.ctypes.data_as(_i64p), C.byref(nnz)), "rcount")
        Ri = np.zeros(max(int(nnz.value), 1), np.int64)
        Rx = np.zeros(max(int(nnz.value), 1))
        self._check(self.lib.stmqr_b200_rconvert(self.h, econ, Rp.ctypes.data_as(_i64p), Ri.ctypes.data_as(_i64p),
                                                 Rx.ctypes.data_as(_f64p)), "rconvert")
        return Rp, Ri[: nnz.value], Rx[: nnz.value]

    def plan_info(self):
        """-> (PlanInfo, Coff[nf], Csize[nf], level[nf]): arena sizes of the current plan, offset / bound size of
        every contribution block in the recycled arena, etree level of every front"""
        nf = max(self.sym.nf, 1)
        info = PlanInfo()
        coff, csize, level = np.zeros(nf, np.int64), np.zeros(nf, np.int64), np.zeros(nf, np.int32)
        self._check(self.lib.stmqr_b200_plan_info(self.h, C.byref(info), coff.ctypes.data_as(_i64p),
                                                  csize.ctypes.data_as(_i64p),
                                                  level.ctypes.data_as(C.POINTER(C.c_int32))), "plan_info")
        return info, coff[: self.sym.nf], csize[: self.sym.nf], level[: self.sym.nf]

    def close(self):
        if self.h:
            self.lib.stmqr_b200_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Planner(Engine):
    """Host-only handle (no GPU needed): analyze / set_partition compute the plan -- level schedule, arena sizes,
    recycled contribution-block offsets -- and account for the device memory it would take."""

    def __init__(self):
        self.lib = load_library()
        self.h = C.c_void_p()
        st = self.lib.stmqr_b200_create_planner(C.byref(self.h))
        if st != STMQR_OK:
            raise EngineError(f"stmqr_b200_create_planner failed: {ERRORS.get(st, st)}")
        self.sym = None

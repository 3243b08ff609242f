#!/usr/bin/env python
"""bench.py -- numeric multifrontal-QR factorization throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload lap2d_1024] [--impl reference]

A "step" is ONE numeric factorization (the reference's qr_factorize, SparseQR_factorize.c:222)
of the workload matrix.  Default workload = BASELINE.json configs[1]: 2-D 5-point Laplacian on a
1024x1024 grid (1 048 576 unknowns), METIS ordering, default tolerance.

What is measured (one JSON line, printed by rank 0):
  value   FP64 GFLOP/s of the numeric factorization with A already resident in HBM
          (stmqr_b200_factorize_resident; CUDA events on the engine's stream; flops = the
          reference's own FLOP_COUNT, SparseQR_factorize.c:1571, computed on the device and
          equal to the count recomputed from HStair).
  e2e     the same metric through the reference-facing plug-in: the reference's SparseQR
          machinery calls qr_factorize (our drop-in, host/qr_factorize_b200.c) with a HOST
          sparse_csc and gets a HOST qr_numeric back (H2D of A and D2H of R+H inside the timed
          region, wall clock around the call = the reference's Fac_time interval).
  roofline  the dominant kernel class of the step, from CUDA events around every launch
          (options.profile_phases) in extra steps run after the timed region.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, compiled from /root/reference) running
          its own CPU qr_factorize on the same matrix and symbolic object, on this box's cores.

The symbolic phase (ordering, etree, fronts; qr_analyze) is the reference's own code, reused
unchanged as north_star prescribes; it runs once, outside every timed region.  The numeric path
of the product never touches oracle/: it is libstmqr_b200.so (CUDA) + libstmqr_dropin.so (C).
`--impl reference` times the reference CPU path alone (rank 0), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stm-multifrontal-qr-factorization-empowered-by-gcn_b200")
sys.path.insert(0, os.path.join(PKG, "py"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "numeric_qr_factorization_fp64_gflops"
UNIT = "GFLOP/s"


# ----------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ----------------------------------------------------------------------------------------------
def make_workload(name: str):
    """-> (description, (m, n, Ap, Ai, Ax) or ('mtx', file), ordering_arg)"""
    from stmqr_b200 import matrices as M
    if name.startswith("lap2d_"):
        g = int(name.split("_")[1])
        return (f"2-D 5-point Laplacian {g}x{g} grid ({g*g} unknowns), METIS", M.laplacian_2d(g), 2)
    if name.startswith("lap3d_"):
        g = int(name.split("_")[1])
        return (f"3-D 7-point Laplacian {g}^3 ({g**3} unknowns), METIS", M.laplacian_3d(g), 2)
    if name.startswith("tall_"):
        _, m, n = name.split("_")
        m, n = int(m), int(n)
        return (f"banded random tall-sparse {m}x{n} (8 draws/row, halfwidth 64, PCG64 seed 4), COLAMD",
                M.tall_banded_random(m, n, draws=8, halfwidth=64, seed=4), 1)
    if name.startswith("dense_"):
        _, m, n = name.split("_")
        return (f"dense random {m}x{n} as sparse_csc (one front: the large-front kernels alone), COLAMD",
                M.dense_random(int(m), int(n)), 1)
    if name.startswith("mtx:"):            # mtx:<name>:<ordering>
        _, f, o = name.split(":")
        return (f"bundled Data/{f}.mtx, ordering arg {o}", ("mtx", f), int(o))
    raise SystemExit(f"unknown workload {name}")


_JSON_FD = 1


def emit(line: dict):
    sys.stdout.flush()
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [nm for k, nm in enumerate(names)
                   if any(len(r) > 3 + k and r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------
def host_setup(workload: str, backend: str):
    """Matrix + the reference host library's symbolic analysis (SparseQR -> qr_1colamd ->
    qr_analyze, reused unchanged).  The first numeric factorization inside SparseQR() goes
    through `backend` ("b200" drop-in or "reference")."""
    import refapi as R
    desc, mat, order = make_workload(workload)
    ref = R.Reference()
    if mat[0] == "mtx":
        A = ref.read_mtx(os.path.join(R.DATA_DIR, mat[1] + ".mtx"))
    else:
        m, n, p, i, x = mat
        A = ref.csc_from_arrays(m, n, p, i, x)
    tol = ref.default_tol(A)
    ref.set_backend(backend)
    t0 = time.time()
    QR = ref.sparseqr(A, order, tol, grain=1.0, tap=True)
    t_total = time.time() - t0
    info = ref.qr_info(QR)
    return R, ref, A, QR, tol, desc, {"symbolic_plus_first_factorization_s": t_total,
                                      "first_factorize_s": info["fac_seconds"], "n1cols": info["n1cols"]}


def run_reference_arm(args):
    """The reference's own CPU numeric factorization: qr_factorize of oracle/_ref (compiled
    from the unmodified sources), every host thread it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refapi as R
    cores = os.cpu_count() or 1
    desc, mat, order = make_workload(args.workload)
    ref = R.Reference()
    ref.set_backend("reference")
    if mat[0] == "mtx":
        A = ref.read_mtx(os.path.join(R.DATA_DIR, mat[1] + ".mtx"))
    else:
        A = ref.csc_from_arrays(*mat)
    tol = ref.default_tol(A)
    # two legal ways to use the cores: serial etree x threaded BLAS, or the
    # reference's TPSM tree tasks (cc->SPQR_grain = 2*cores as qrtest.c:144 sets it; pool sized
    # above ntasks) x 1 BLAS thread.  Each needs its own symbolic object (the task
    # partition is made by qr_analyze).  Probe each once, keep the faster for the timed steps.
    QRs = ref.sparseqr(A, order, tol, grain=1.0, blas_threads=cores)
    info = ref.qr_info(QRs)
    if info["n1cols"] != 0:
        raise SystemExit("reference arm needs a workload without column singletons")
    flops = info["flopcount"]                      # cc->SPQR_flopcount (counted when grain <= 1)
    sym = ref.symbolic(QRs)
    QRt = ref.sparseqr(A, order, tol, grain=2.0 * cores, pool=128, blas_threads=1)
    modes = {"threaded_blas": (QRs, dict(pool=0, blas_threads=cores)),
             "tree_tasks": (QRt, dict(pool=128, blas_threads=1))}
    probe = {k: ref.refactorize(A, q, **kw) for k, (q, kw) in modes.items()}
    best = min(probe, key=probe.get)
    q, kw = modes[best]
    times = []
    for s in range(args.warmup + args.steps):
        t = ref.refactorize(A, q, **kw)
        if s >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = flops / sec * 1e-9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "m": sym.m, "n": sym.n, "nnz": sym.anz,
                       "fronts": sym.nf, "flops_per_step": flops},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"whole workload, {args.steps} full qr_factorize calls, mode {best} "
                                       f"(ntasks {int(ref.qr_info(q)['ntasks'])}); probe seconds {probe}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def multi_gpu_parity_check(D, sq, R, local, rank, world):
    """Before anything is timed: golden inputs factorized over the REAL ranks through the C data plane (NCCL),
    gathered on rank 0 and compared with the reference's golden output (integer structure bit-exact, R to the
    north_star tolerance).  -> "ok" (rank 0) or raises."""
    import torch
    import torch.distributed as dist
    done = []
    for case in ("dwt_992_metis", "lap3d_8_metis", "tall_600x150_colamd"):
        sym, A, tol, ntol, want = R.load_golden(case)
        e = sq.Engine(local)
        e.analyze(sym)
        e.upload_matrix(A)
        df = D.DistFactorization(e, sym, torch.device("cuda", local))
        info = df.factorize(tol, ntol)
        num = e.download(info)
        got = D.gather_numeric(sym, df.owner, num, info)
        if rank == 0:
            d = R.assert_numeric_parity(sym, A, got, want, f"{case} over {world} ranks (NCCL)")
            assert got.flops == want.flops
            done.append(f"{case}: fronts per rank {np.bincount(df.owner, minlength=world).tolist()}, "
                        f"max|dR|/|A| {d:.1e}")
        e.close()
        dist.barrier()
    return "ok (" + "; ".join(done) + ")" if rank == 0 else "ok"


def extra_leg(args, sq, torch, dist, rank, world, local, workload):
    """A second, GPU-only workload in the same process (resident numeric phase, no CPU arm): the 3-D Laplacian
    the north_star names for the 1 -> 8 GPU scaling, next to the headline workload."""
    import refapi as R
    desc, mat, order = make_workload(workload)
    ref = R.Reference()
    A = ref.csc_from_arrays(*mat)
    tol = ref.default_tol(A)
    t0 = time.time()
    sym = ref.analyze_only(A, order, tol)               # the reference's symbolic phase alone (no numeric work)
    t_sym = time.time() - t0
    At = ref.csc_to_numpy(A)
    eng = sq.Engine(local)
    eng.analyze(sym)
    eng.upload_matrix(At)
    ntol = sym.n
    pf = None
    if world > 1:
        from stmqr_b200 import dist as D
        pf = D.DistFactorization(eng, sym, torch.device("cuda", local))
    ms = []
    info = None
    for s in range(1 + args.leg_steps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        info = pf.factorize(tol, ntol) if pf is not None else eng.factorize_resident(tol, ntol)
        if s >= 1:
            ms.append(eng.stats().ms_numeric)
    t = float(np.mean(ms))
    rh = float(info.rh_size)
    dev_bytes = float(eng.stats().device_bytes)
    if dist is not None:
        tt = torch.tensor([t, dev_bytes], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t, dev_bytes = float(tt[0].item()), float(tt[1].item())
        tt = torch.tensor([rh], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        rh = float(tt.item())
    out = {"workload": workload, "description": desc, "ms_per_step": t, "value": float(info.flops) / t * 1e-6,
           "unit": UNIT, "steps": args.leg_steps, "warmup": 1, "flops_per_step": float(info.flops),
           "rank": int(info.rank), "fronts": sym.nf, "rh_doubles": rh, "device_bytes_max_per_gpu": dev_bytes,
           "symbolic_s": t_sym, "timing": "CUDA events on every rank's engine stream, max over ranks"}
    eng.close()
    ref.free_sparse(A)
    ref.close()
    return out


def run_b200_arm(args):
    import torch
    import stmqr_b200 as sq

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    os.environ["STMQR_B200_DEVICE"] = str(local)
    os.environ["STMQR_B200_CACHE_PLAN"] = "1"
    R, ref, A, QR, tol, desc, setup = host_setup(args.workload, "b200")
    sym = ref.symbolic(QR)
    At, ttol, ntol = ref.tapped()

    # one engine at a time on the device: the drop-in's own handle (used by host_setup's first
    # factorization and again by the e2e leg below) is released while the resident leg runs
    ref.dropin_shutdown()
    eng = sq.Engine(local)
    eng.set_options(panel=args.panel)
    eng.analyze(sym)
    eng.upload_matrix(At)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    pf = None
    parity = None
    if world > 1:
        # every front on its owner GPU, the data plane in C (csrc/multigpu.cuh): contribution blocks + row ids
        # over ncclSend/ncclRecv per etree level, max-merges with ncclAllReduce, all on the engine's stream
        from stmqr_b200 import dist as D
        parity = multi_gpu_parity_check(D, sq, R, local, rank, world)     # real ranks, real NCCL, golden inputs
        pf = D.DistFactorization(eng, sym, torch.device("cuda", local))

    def factor_step():
        """one numeric factorization of the resident matrix -> (NumericInfo of this rank, ms)"""
        if pf is None:
            inf = eng.factorize_resident(ttol, ntol)
            return inf, eng.stats().ms_numeric            # CUDA events on the engine's stream
        barrier()
        inf = pf.factorize(ttol, ntol)
        return inf, eng.stats().ms_numeric                # events on this rank's stream (waits for peers included)

    # ---------------- value: A resident in HBM, device time ----------------------------------
    for _ in range(args.warmup):
        info, _ = factor_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms = []
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        info, ms = factor_step()
        dev_ms.append(ms)
    barrier()
    st = eng.stats()
    launches = int(st.launches)
    flops = float(info.flops)
    t_dev = float(np.sum(dev_ms)) * 1e-3
    if dist is not None:
        tt = torch.tensor([t_dev, float(launches)], device="cuda", dtype=torch.float64)
        t2 = tt.clone()
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(t2, op=dist.ReduceOp.SUM)
        t_dev = float(tt[0].item())
        launches = int(t2[1].item())

    # ---------------- roofline of the dominant kernel class (extra, untimed steps) ------------
    eng.set_options(panel=args.panel, profile_phases=1)
    cls_ms = np.zeros(8)
    cls_n = np.zeros(8)
    nprof = max(1, min(args.steps, 3))
    for _ in range(nprof):
        flush.zero_()
        torch.cuda.synchronize()
        factor_step()
        s2 = eng.stats()
        cls_ms += np.array(list(s2.ms_class))
        cls_n += np.array(list(s2.launches_class))
    cls_ms /= nprof
    cls_n /= nprof
    s2 = eng.stats()
    peaks, peak_src = measured_peaks()
    dmma_tf, dfma_tf = eng.measure_fp64_peak()
    asm_ms = float(cls_ms[1] + cls_ms[2] + cls_ms[5] + cls_ms[6])
    front_ms = float(cls_ms[3] + cls_ms[4])
    classes = dict(zip(sq.KERNEL_CLASSES, [round(float(x), 4) for x in cls_ms]))
    dom = int(np.argmax(cls_ms))
    ms_step = t_dev / args.steps * 1e3                 # the TIMED step (look-ahead overlap on, no per-launch events)
    other_ms = float(cls_ms[0] + cls_ms[7]) + asm_ms   # build_S, hpinv and assembly/pack: serial on the main stream
    # DRAM traffic per launch of the dominant kernel comes from an ncu --set full capture of this workload, if one
    # has been summarised under profiles/ (profiles/dram_traffic.json: workload -> kernel, bytes, source); never a
    # constant in this file
    traffic = traffic_src = traffic_kernel = None
    tj = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tj):
        with open(tj) as f:
            ent = json.load(f).get(args.workload)
        if ent:
            traffic, traffic_src, traffic_kernel = ent.get("dram_bytes_per_launch"), ent.get("source"), ent.get("kernel")
    if dom in (3, 4):
        # front QR (panel + WY update kernels); algorithmic flops = the reference's count.  Denominator: the timed
        # step minus the serial assembly/pack/hpinv kernels (measured in the profile steps) = the wall time the
        # front-QR kernels occupy in the real, overlapped schedule
        front_timed = max(ms_step - other_ms, 1e-6)
        ach = flops / (front_timed * 1e-3) * 1e-12
        roof = {"bound": "tensor", "kernel": "front QR (k_panel_* + k_update_dmma / k_wide_*)", "achieved": ach,
                "peak": dmma_tf, "unit": "TFLOP/s", "frac": ach / dmma_tf if dmma_tf else None, "traffic": traffic,
                "how": "reference flop count / (timed ms_per_step - assembly, pack, build_S, hpinv ms of the profile steps)",
                "front_qr_ms_timed": front_timed, "whole_step_tflops": flops / (ms_step * 1e-3) * 1e-12,
                "whole_step_frac": flops / (ms_step * 1e-3) * 1e-12 / dmma_tf if dmma_tf else None,
                "peak_source": "FP64 mma.sync (DMMA) register-loop microbenchmark run in this process "
                               "(stmqr_b200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                "avg_launch_ms": front_ms / max(1.0, float(cls_n[3] + cls_n[4])), "traffic_source": traffic_src,
                "traffic_kernel": traffic_kernel}
    else:
        ach = float(s2.bytes_assemble) / (asm_ms * 1e-3) * 1e-9
        roof = {"bound": "hbm", "kernel": "assembly+pack (k_front_setup, k_assemble, k_front_finish, k_pack)",
                "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_source": traffic_src, "traffic_kernel": traffic_kernel,
                "peak_source": peak_src,
                "avg_launch_ms": asm_ms / max(1.0, float(cls_n[1] + cls_n[2] + cls_n[5] + cls_n[6]))}
    roof["class_ms"] = classes
    # (several GPUs: bytes_assemble is the merged total of all ranks, asm_ms this rank's kernels -> aggregate rate
    # against world x the per-GPU peak)
    roof["assembly_gbs"] = float(s2.bytes_assemble) / (asm_ms * 1e-3) * 1e-9 if asm_ms > 0 else None
    roof["assembly_frac_of_hbm"] = roof["assembly_gbs"] / (peaks["hbm_gbs"] * world) if asm_ms > 0 else None
    roof["front_qr_tflops"] = flops / (front_ms * 1e-3) * 1e-12 if front_ms > 0 else None
    roof["fp64_dmma_peak_tflops"] = dmma_tf
    roof["fp64_dfma_peak_tflops"] = dfma_tf
    eng.set_options(panel=args.panel, profile_phases=0)

    # ---------------- factorize + solve without downloading the factor ----------------
    solve = None
    if pf is None and sym.m == sym.n:
        rng = np.random.default_rng(5)
        bvec = rng.standard_normal(sym.m)
        eng.factorize_resident(ttol, ntol)
        xs, ms_solve = eng.solve_ls(bvec)                   # warm-up (builds the Householder table)
        xs, ms_solve = eng.solve_ls(bvec)
        Ssp = At.to_scipy()
        res = float(np.linalg.norm(Ssp @ xs[:, 0] - bvec) / max(np.linalg.norm(bvec), 1e-300))
        solve = {"what": "x = E*(R\\(Q'b)) on the device from the resident R+H (stmqr_b200_solve_ls), 1 right-hand side",
                 "ms": ms_solve, "relative_residual": res,
                 "bytes_not_downloaded": int(info.rh_size) * 8}
    legs = {}
    if args.leg and args.leg != args.workload:
        if pf is None:
            eng.close()
            eng = None
        legs[args.leg] = extra_leg(args, sq, torch, dist, rank, world, local, args.leg)
    if pf is None and eng is not None:
        eng.close()                                   # the e2e leg below uses the drop-in's own handle

    # ---------------- e2e: host sparse_csc in, host qr_numeric out -----------------------------
    # one GPU: through the reference-facing plug-in, qr_factorize (drop-in) called by the reference
    # host library.  several GPUs: every rank uploads A from host memory, the partitioned numeric
    # phase runs, every rank downloads the R+H blocks of its own fronts into host arrays.
    e2e_s = []
    can_e2e = setup["n1cols"] == 0
    rh_local = int(info.rh_size)
    if can_e2e:
        if pf is None:
            # the harness's backend switch is process-wide (the extra leg's symbolic probe touches it): select the
            # drop-in again and refuse to time anything else
            ref.set_backend("b200")
            assert int(ref.L.rh_get_backend()) == 1, "e2e leg: the drop-in qr_factorize is not the selected backend"
        for s in range(min(args.warmup, 2) + args.steps):
            flush.zero_()
            barrier()
            if pf is None:
                t = ref.refactorize(A, QR)
            else:
                t0 = time.perf_counter()
                eng.upload_matrix(At)
                stack = eng.stream_begin()                  # this rank's R+H blocks go to the host level by level
                inf = pf.factorize(ttol, ntol)
                eng.stream_end()
                eng.download(inf, stack=stack[: max(int(inf.rh_size), 1)])
                barrier()
                t = time.perf_counter() - t0
            if s >= min(args.warmup, 2):
                e2e_s.append(t)
    clocks = sampler.stop()
    t_e2e = float(np.sum(e2e_s)) if e2e_s else None
    rh_total = rh_local
    if dist is not None:
        tt = torch.tensor([t_e2e or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item()) if t_e2e is not None else None
        tt = torch.tensor([float(rh_local)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        rh_total = int(tt.item())
    rh_bytes = rh_total * 8
    h2d = ((sym.n + 1) * 8 + sym.anz * 16) * world
    d2h = rh_bytes + (8 * (2 * sym.rjsize + sym.hisize + 3 * sym.nf + sym.m) + sym.n) * world

    # ---------------- CPU baseline: the reference's own qr_factorize on this box's cores --------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and can_e2e:
        # bounded sample: the whole workload, once per way of using the cores (serial etree x
        # threaded BLAS on this run's symbolic object; the reference's TPSM tree tasks with
        # cc->SPQR_grain = 2*cores need their own qr_analyze because it makes the task partition)
        cores = os.cpu_count() or 1
        ref.set_backend("reference")
        probe = {"threaded_blas": ref.refactorize(A, QR, pool=0, blas_threads=cores)}
        ntasks = 1
        if not args.no_cpu_tree_tasks:
            _, _, order = make_workload(args.workload)
            QRt = ref.sparseqr(A, order, tol, grain=2.0 * cores, pool=128, blas_threads=1)
            ntasks = int(ref.qr_info(QRt)["ntasks"])
            probe["tree_tasks"] = min(ref.qr_info(QRt)["fac_seconds"],
                                      ref.refactorize(A, QRt, pool=128, blas_threads=1))
            ref.free_qr(QRt)
        best = min(probe, key=probe.get)
        cpu = {"value": flops / probe[best] * 1e-9, "unit": UNIT, "cores": cores, "kind": "reference",
               "seconds": probe[best],
               "sample": f"whole workload, one qr_factorize call per mode, seconds {probe} "
                         f"(tree_tasks: ntasks {ntasks}, TPSM pool 128, 1 BLAS thread; threaded_blas: serial "
                         f"etree, {cores} OpenBLAS threads); best = {best}"}
        ref.set_backend("b200")

    if rank == 0:
        total_flops = flops * args.steps              # ONE factorization per step, shared by all ranks
        line = {"metric": METRIC, "value": total_flops / t_dev * 1e-9, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": args.workload, "description": desc, "m": sym.m, "n": sym.n,
                           "nnz": sym.anz, "fronts": sym.nf, "etree_levels": int(st.nlevels),
                           "rank": int(info.rank), "flops_per_step": flops,
                           "rh_doubles": rh_total, "tol": ttol,
                           "l2": "256 MiB device buffer written between timed steps (L2 flush)",
                           "multi_gpu": ("single" if world == 1 else
                                         f"every front on its owner GPU ({world} GPUs, fronts per rank "
                                         f"{np.bincount(pf.owner, minlength=world).tolist()}); contribution blocks + row "
                                         f"ids move after the child's etree level over ncclSend/ncclRecv issued by the C "
                                         f"library on the engine stream, merges by ncclAllReduce; no host sync between levels"),
                           "timing": ("CUDA events on the engine stream" if world == 1 else
                                      "CUDA events on every rank's engine stream after a barrier (waits for peers included), max over ranks"),
                           "device_bytes": int(st.device_bytes)},
                "e2e": ({"value": flops * len(e2e_s) / t_e2e * 1e-9, "unit": UNIT,
                         "ms_per_step": t_e2e / len(e2e_s) * 1e3, "h2d_bytes_per_step": h2d,
                         "d2h_bytes_per_step": d2h,
                         "how": ("wall clock around qr_factorize (drop-in) called by the reference host "
                                 "library with a host sparse_csc; host qr_numeric out; plan cached "
                                 "(STMQR_B200_CACHE_PLAN=1)") if world == 1 else
                                ("per rank: upload A from host, partitioned numeric phase with the rank's own R+H "
                                 "blocks streamed to host arrays level by level; max over ranks"),
                         "first_call_with_plan_s": setup["first_factorize_s"],
                         "plan_ms": float(st.ms_plan)} if t_e2e else None),
                "gpu_launches": launches * args.steps,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "solve": solve, "legs": legs,
                "parity_check": parity,
                "stats": {"ms_h2d": float(st.ms_h2d), "ms_d2h": float(st.ms_d2h),
                          "launches_per_step": launches}}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="lap2d_1024")
    ap.add_argument("--panel", type=int, default=0)
    ap.add_argument("--leg", default="lap3d_96",
                    help="second, GPU-only workload measured in the same run (the 3-D Laplacian of the scaling target); '' = none")
    ap.add_argument("--leg-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cpu-tree-tasks", action="store_true",
                    help="skip the reference's TPSM tree-task mode in the CPU baseline probe")
    args = ap.parse_args()
    # the reference library printf()s its own progress on fd 1: keep fd 1 for the ONE JSON line
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()

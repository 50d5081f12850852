#!/usr/bin/env python
"""bench.py -- headline benchmark of the LSSP solve-loop hot path on B200.

Metric (BASELINE.json): Krylov iterations per second (and SpMV HBM GB/s against the
roofline) on the reference's headline configuration.  At N = 1 the workload is
BASELINE.json configs[1]: 3-D 7-point Laplacian 256^3 (16.8 M rows, fp64 CSR),
CG preconditioned by ILU(0), b = 1, x0 = 0, default tolerances (1e-7).

A "step" is one complete solve (to the reference's tolerance) of that system.
  value : iterations / second, device-resident operands, CUDA-event timed.
  e2e   : the same through the reference-facing host call (lssp_solver_solve with
          HOST b and x: H2D of b and x0 and D2H of x inside the timed region).
  roofline / roofline_spmv : achieved algorithmic GB/s of the dominant kernel
          (the level-scheduled triangular sweeps) and of the CSR SpMV, both timed
          live with CUDA events on the library's stream.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, serial, 1 core) timed on
          this box's host cores on a bounded sample of the same workload.

`--impl reference` times only the reference's own CPU implementation.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(key):
    """DRAM bytes per launch of a kernel on a named workload, from the committed `ncu --set full`
    capture (profiles/traffic.json); None when that workload has no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        super().__init__(daemon=True)
        self.device, self.rows, self._halt = device, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def finish(self):
        self._halt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_problem(args):
    from lssp_b200 import generators as g
    N = args.grid
    args.ilu_level = 0
    if args.workload == "cg_ilu0":
        A = g.lap3d(N)
        name = "lap3d_%d CG+ILU(0) (BASELINE.json configs[1])" % N
        solver, pc = "cg", "iluk"
    elif args.workload == "bicgstab_ilu0":
        A = g.cd3d(N)
        name = "cd3d_%d BiCGStab+ILU(0)" % N
        solver, pc = "bicgstab", "iluk"
    elif args.workload == "bicgstab_iluk1":
        A = g.cd3d(N)
        name = "cd3d_%d BiCGStab+ILUK(1) (BASELINE.json configs[2])" % N
        solver, pc = "bicgstab", "iluk"
        args.ilu_level = 1
    elif args.workload == "cg_amg":
        A = g.lap3d(N)
        name = "lap3d_%d CG+SXAMG-style V-cycle, zero initial guess, %s Gauss-Seidel (BASELINE.json configs[3] operator)" % (
            N, {0: "natural-order", 1: "C/F-ordered", 2: "C/F-ordered multicolour"}[args.amg_order])
        solver, pc = "cg", "amg"
    elif args.workload == "cg_non":
        A = g.lap3d(N)
        name = "lap3d_%d CG unpreconditioned" % N
        solver, pc = "cg", "non"
    else:
        raise SystemExit("unknown workload " + args.workload)
    return A, name, solver, pc


def reference_arm(args, rank):
    """The reference's own serial CPU implementation (oracle/_ref), 1 thread: the library
    has no threads.  Each step = one lssp_solver_solve capped at
    --ref-iters iterations on the same matrix (a bounded sample of the workload)."""
    if rank != 0:
        return 0
    import oracle
    A, name, solver, pc = build_problem(args)
    n = len(A[0]) - 1
    if pc == "amg" or not oracle.Ref.available():
        # no compiled reference for this path (libsxamg is not in the tree / _ref not built): the C port
        from lssp_b200 import api   # host set-up only (factors / hierarchy); no GPU work
        P = oracle.Port()
        kw = {}
        if pc == "amg":
            H = api.AmgHierarchy(A, cf_order=args.amg_order)
            kw["amg"] = P.amg(H.levels, coarse_inv=H.coarse_inv, zero_guess=1, cf_order=args.amg_order)
        elif pc == "iluk":
            kw["LU"] = api.ilu_factor(A, "iluk", level=args.ilu_level)
        its, secs = 0, 0.0
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            r = P.solve(solver, A, np.ones(n), maxit=args.ref_iters, **kw)
            if step >= args.warmup:
                its += r["nits"]
                secs += time.perf_counter() - t0
        val = its / secs
        print(json.dumps({"impl": "reference", "metric": "%s_iterations_per_second" % args.workload, "value": val,
                          "unit": "iter/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": name, "n": n, "nnz": int(A[0][-1]), "iterations_per_step": args.ref_iters},
                          "cpu_baseline": {"value": val, "unit": "iter/s", "cores": 1, "kind": "port",
                                           "sample": "%d steps x %d iterations of the same solve, oracle/*.c"
                                                     % (args.steps, args.ref_iters)},
                          "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    R = oracle.Ref()
    L = R.lib
    L.ref_session_create.restype = C.c_void_p
    L.ref_session_solve.restype = C.c_int
    prm = oracle.ref_params(maxit=args.ref_iters, iluk_level=args.ilu_level)
    tas = C.c_double()
    h = C.c_void_p(L.ref_session_create(oracle.SOLVERS[solver], oracle.PCS[pc], n, A[0].ctypes.data_as(C.c_void_p),
                                        A[1].ctypes.data_as(C.c_void_p), A[2].ctypes.data_as(C.c_void_p),
                                        C.byref(prm), C.byref(tas)))
    b = np.ones(n)
    its, secs = 0, 0.0
    for step in range(args.warmup + args.steps):
        x = np.zeros(n)
        res, t = C.c_double(), C.c_double()
        k = L.ref_session_solve(h, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), args.ref_iters,
                                C.byref(res), C.byref(t))
        if step >= args.warmup:
            its += k
            secs += t.value
    L.ref_session_destroy(h)
    val = its / secs
    # N > 1: the GPU arm's value is n_gpus x iterations/s on an n_gpus-slab problem, i.e. SLAB-iterations per
    # second (weak scaling); the serial reference advances one 256^3 slab-iteration at the rate measured here
    # whatever the number of slabs (its cost per iteration is linear in n), so the same number is its value
    unit = "iter/s" if args.gpus <= 1 else "iter/s x n_gpus"
    line = {"impl": "reference", "metric": "%s_iterations_per_second" % args.workload, "value": val, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "n": n, "nnz": int(A[0][-1]), "iterations_per_step": args.ref_iters,
                       "sample": "one GPU's slab of the weak-scaled problem; serial cost per iteration is linear in n"
                                 if args.gpus > 1 else "the whole problem"},
            "cpu_baseline": {"value": val, "unit": "iter/s", "cores": 1, "kind": "reference",
                             "sample": "%d steps x %d iterations of the same solve (maxit capped), assemble %.1f s excluded"
                                       % (args.steps, args.ref_iters, tas.value)},
            "e2e": {"value": val, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def cpu_baseline(args, A, solver, pc, pcobj=None):
    import oracle
    n = len(A[0]) - 1
    if pc == "amg":
        # libsxamg is not in the reference tree: the CPU side is the restated cycle (parity unpinned)
        H = pcobj.hierarchy
        m = oracle.Port().amg(H.levels, coarse_inv=H.coarse_inv, zero_guess=1, cf_order=H.pars.cf_order)
        t0 = time.perf_counter()
        r = oracle.Port().solve(solver, A, np.ones(n), amg=m, maxit=args.ref_iters)
        t = time.perf_counter() - t0
        return {"value": r["nits"] / t, "unit": "iter/s", "cores": 1, "kind": "port",
                "sample": "%d iterations of the same solve, oracle/oracle.c + amg_oracle.c on the same hierarchy"
                          % r["nits"], "host_cores_total": os.cpu_count()}
    if oracle.Ref.available():
        R = oracle.Ref()
        L = R.lib
        L.ref_session_create.restype = C.c_void_p
        L.ref_session_solve.restype = C.c_int
        L.ref_session_time_mxy.restype = C.c_double
        L.ref_session_time_pc.restype = C.c_double
        prm = oracle.ref_params(maxit=args.ref_iters, iluk_level=args.ilu_level)
        tas = C.c_double()
        h = C.c_void_p(L.ref_session_create(oracle.SOLVERS[solver], oracle.PCS[pc], n, A[0].ctypes.data_as(C.c_void_p),
                                            A[1].ctypes.data_as(C.c_void_p), A[2].ctypes.data_as(C.c_void_p),
                                            C.byref(prm), C.byref(tas)))
        b, x = np.ones(n), np.zeros(n)
        res, t = C.c_double(), C.c_double()
        k = L.ref_session_solve(h, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), args.ref_iters,
                                C.byref(res), C.byref(t))
        t_mxy = L.ref_session_time_mxy(h, 3)
        t_pc = L.ref_session_time_pc(h, 2)
        L.ref_session_destroy(h)
        spmv_bytes = 12.0 * int(A[0][-1]) + 4.0 * (n + 1) + 16.0 * n
        return {"value": k / t.value, "unit": "iter/s", "cores": 1, "kind": "reference",
                "sample": "%d iterations of the same %s solve (maxit capped), unmodified reference, serial" % (k, solver),
                "assemble_s": tas.value, "spmv_gbs": spmv_bytes / t_mxy / 1e9, "spmv_ms": 1e3 * t_mxy,
                "pc_apply_ms": 1e3 * t_pc, "host_cores_total": os.cpu_count()}
    P = oracle.Port()
    from lssp_b200 import api
    LU = api.ilu_factor(A, "iluk", level=args.ilu_level) if pc == "iluk" else None
    t0 = time.perf_counter()
    r = P.solve(solver, A, np.ones(n), LU=LU, maxit=args.ref_iters)
    t = time.perf_counter() - t0
    return {"value": r["nits"] / t, "unit": "iter/s", "cores": 1, "kind": "port",
            "sample": "%d iterations of the same solve, oracle/oracle.c" % r["nits"], "host_cores_total": os.cpu_count()}


def parity_block(args, ctx, api, solver, dA, pc, n):
    """Outside the timed region: this run against the UNMODIFIED reference's numbers for the same problem
    (tests/golden/baseline_<N>.json, generated by tests/golden/make_baseline_golden.py from oracle/_ref) -- iterations to
    tolerance and the first 20 residuals in the shipped tree-reduction mode, and the same solve with the reference-order
    reductions (exact_sum.cu), which must reproduce the reference bit for bit."""
    key = {"cg_ilu0": "lap3d/cg+iluk0", "bicgstab_iluk1": "cd3d/bicgstab+iluk1"}.get(args.workload)
    path = os.path.join(ROOT, "tests", "golden", "baseline_%d.json" % args.grid)
    if key is None or not os.path.exists(path):
        return None
    with open(path) as f:
        gold = json.load(f).get(key)
    if not gold:
        return None
    want = np.array(gold["history"])
    out = {"reference_iterations": gold["nits"], "reference_residual": gold["residual"],
           "reference_source": "tests/golden/baseline_%d.json (unmodified reference, oracle/_ref)" % args.grid}
    x = np.zeros(n)
    r = api.lssp_solver_solve(ctx, solver, dA, pc, np.ones(n), x, nhist=20, maxit=3000)
    k = min(len(want), len(r["hist"]))
    out["history_relerr"] = float(np.max(np.abs(r["hist"][:k] - want[:k]) / want[:k]))
    ctx.set_option(api.OPT_REDUCE_SEQUENTIAL, 2)
    try:
        x = np.zeros(n)
        api.lssp_solver_solve(ctx, solver, dA, pc, np.ones(n), x, nhist=20, maxit=3000)      # warm-up (scratch allocation)
        x = np.zeros(n)
        r2 = api.lssp_solver_solve(ctx, solver, dA, pc, np.ones(n), x, nhist=20, maxit=3000)
    finally:
        ctx.set_option(api.OPT_REDUCE_SEQUENTIAL, 0)
    out["reference_order_mode"] = {
        "what": "LSSPG_OPT_REDUCE_SEQUENTIAL = 2: dot products = the reference's sequential sums, computed in parallel",
        "iterations": r2["nits"], "residual": r2["residual"],
        "bit_identical_to_reference": bool(r2["nits"] == gold["nits"] and r2["residual"] == gold["residual"] and
                                           list(r2["hist"][:len(want)]) == list(want)),
        "iter_per_s": r2["nits"] / (r2["solve_ms"] / 1e3)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--workload", default="cg_ilu0")
    ap.add_argument("--check-every", type=int, default=8)
    ap.add_argument("--pl-rows", type=int, default=4000000, help="rows per GPU of the power-law workload (idrs_powerlaw, --gpus > 1)")
    ap.add_argument("--amg-order", type=int, default=1, help="cf_order of the AMG smoother (cg_amg): 1 C/F by index, 2 multicolour")
    ap.add_argument("--ref-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="--gpus > 1: weak = a grid^3 block per GPU (default, the driver's curve); strong = ONE grid^3 problem "
                         "cut into z-slabs over the GPUs")
    ap.add_argument("--shape", default="slab", choices=["slab", "cubic"],
                    help="--gpus > 1, weak scaling: slab = grid x grid x (grid n_gpus); cubic = the most cubic global grid with "
                         "grid^3 rows per GPU (8 GPUs, grid 256: 512^3, 2 MiB halo planes)")
    ap.add_argument("--operator", default=None, choices=[None, "lap", "cd"],
                    help="--gpus > 1: lap = 7-point Laplacian, cd = convection-diffusion (default: cd for bicgstab_ilu0, else lap)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return reference_arm(args, rank)

    if world > 1 or args.gpus > 1:
        from lssp_b200 import dist_bench
        return dist_bench.main(args, rank, world, local_rank)

    from lssp_b200 import api
    from lssp_b200._lib import lib, check
    t_setup = time.perf_counter()
    A, name, solver, pckind = build_problem(args)
    n, nnz = len(A[0]) - 1, int(A[0][-1])
    t_gen = time.perf_counter() - t_setup
    ctx = api.Context(local_rank)
    ctx.set_option(api.OPT_CHECK_EVERY, args.check_every)
    dA = api.Csr(ctx, A)
    t0 = time.perf_counter()
    ilu_level = args.ilu_level
    if pckind == "iluk":
        pc = api.Preconditioner.iluk(ctx, A, level=ilu_level)
    elif pckind == "amg":
        pc = api.Preconditioner.sxamg(ctx, A, share=dA, zero_guess=1, cf_order=args.amg_order)
    else:
        pc = api.Preconditioner.non(ctx, n)
    t_pc = time.perf_counter() - t0
    L = lib()
    b, x = ctx.upload(np.ones(n)), ctx.zeros(n)

    def solve_device():
        check(L.lsspg_memset_zero(ctx.h, x.ptr, C.c_size_t(8 * n)))
        return api.solve_device(ctx, solver, dA, pc, b, x, maxit=3000)

    for _ in range(args.warmup):
        r = solve_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.sync()
    launches0 = ctx.launches
    ms = C.c_double()
    check(L.lsspg_timer_start(ctx.h, 0))
    its = 0
    for _ in range(args.steps):
        r = solve_device()
        its += r["nits"]
    check(L.lsspg_timer_stop(ctx.h, 0, C.byref(ms)))
    ctx.sync()
    launches = ctx.launches - launches0
    total_ms = ms.value
    value = its / (total_ms / 1e3)
    nits, residual = r["nits"], r["residual"]

    # ---- e2e: the reference-facing call, lssp_solver_solve(LSSP_SOLVER &, LSSP_PC &) of liblssp.so (through the extern "C"
    # handle of liblssp_e2e.so, ctypes cannot call C++), with the caller's HOST vectors: the library copies b and x0 to the
    # device and x back inside the timed call.  The vectors live in pinned host memory.
    hb, hx = C.c_void_p(), C.c_void_p()
    check(L.lsspg_host_alloc(C.c_size_t(8 * n), C.byref(hb)))
    check(L.lsspg_host_alloc(C.c_size_t(8 * n), C.byref(hx)))
    b_host = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_double)), shape=(n,))
    x_host = np.ctypeslib.as_array(C.cast(hx, C.POINTER(C.c_double)), shape=(n,))
    b_host[:] = 1.0
    x_host[:] = 0.0
    facade = None
    if pckind in ("iluk", "non"):
        E = C.CDLL(os.path.join(ROOT, "lssp_b200", "liblssp_e2e.so"))
        E.lssp_e2e_create.restype = C.c_void_p
        E.lssp_e2e_solve.restype = C.c_int
        C.c_int.in_dll(C.CDLL(os.path.join(ROOT, "lssp_b200", "liblssp.so")), "lssp_verbosity").value = 0
        LSSP_SOLVER = {"cg": 7, "bicgstab": 4, "gmres": 0, "idrs": 18}[solver]     # include/lssp/type-defs.h
        facade = C.c_void_p(E.lssp_e2e_create(LSSP_SOLVER, 1 if pckind == "iluk" else 0, n, A[0].ctypes.data_as(C.c_void_p),
                                              A[1].ctypes.data_as(C.c_void_p), A[2].ctypes.data_as(C.c_void_p), hx, hb, ilu_level, 3000, 50,
                                              C.c_double(-1.0)))
    e2e_its, e2e_s = 0, 0.0
    for step in range(1 + args.steps):
        x_host[:] = 0.0
        ctx.sync()
        t0 = time.perf_counter()
        if facade is not None:
            res = C.c_double()
            re = {"nits": E.lssp_e2e_solve(facade, C.byref(res)), "residual": res.value}
        else:
            re = api.lssp_solver_solve(ctx, solver, dA, pc, b_host, x_host, maxit=3000)
        checksum = float(x_host[n // 2])          # the caller reads its answer from its own x
        dt = time.perf_counter() - t0
        if step >= 1:
            e2e_its += re["nits"]
            e2e_s += dt
    clocks = sampler.finish()
    e2e = {"value": e2e_its / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 8 * n,
           "x_mid": checksum, "iterations_per_solve": re["nits"], "residual": re["residual"],
           "through": "liblssp.so lssp_solver_solve(LSSP_SOLVER&, LSSP_PC&)" if facade is not None else "lssp_b200.api.lssp_solver_solve"}
    if facade is not None:
        E.lssp_e2e_destroy(facade)

    # ---- per-kernel roofline numbers, timed live on the library's stream
    peak, peak_kind = measured_peak()

    def timed(fn, reps):
        fn()
        check(L.lsspg_timer_start(ctx.h, 1))
        for _ in range(reps):
            fn()
        t = C.c_double()
        check(L.lsspg_timer_stop(ctx.h, 1, C.byref(t)))
        return t.value / reps

    y = ctx.empty(n)
    ms_spmv = timed(lambda: dA.mv(api.MV_MXY, b, y), 20)
    spmv_gbs = dA.spmv_bytes / ms_spmv / 1e6
    roof_spmv = {"bound": "hbm", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                 "traffic": ncu_traffic("spmv_pipe_kernel@lap3d_%d" % args.grid) if args.workload != "bicgstab_ilu0" else None,
                 "ms": ms_spmv, "bytes": dA.spmv_bytes, "peak_kind": peak_kind, "frac_of_8TBs": spmv_gbs / 8000.0,
                 "kernel": "spmv_pipe_kernel (cp.async.bulk double-buffered tiles)",
                 "note": "peak is MEASURED_PEAKS.json's torch copy_ figure (half reads, half writes); this kernel's "
                         "traffic is 94 % reads and can exceed it -- see frac_of_8TBs for the nominal HBM3e roofline"}
    ms_per_it = total_ms / its
    if pckind == "iluk":
        ms_pcap = timed(lambda: pc.apply(y, b), 10)
        pc_gbs = pc.bytes / ms_pcap / 1e6
        # per launch: one application = 2 launches of the sweep kernel (L, then U) of equal algorithmic bytes
        roof = {"bound": "hbm", "achieved": pc_gbs, "peak": peak, "unit": "GB/s", "frac": pc_gbs / peak,
                "traffic": ncu_traffic("tri_pencil_kernel@lap3d_%d" % args.grid) if args.workload == "cg_ilu0" else None,
                "kernel": "triangular sweep (tri_pencil_kernel on lattice factors, tri_box_ell_kernel / tri_solve_kernel "
                          "otherwise); 2 launches per ILU application (L, U), average of the two", "ms": ms_pcap / 2,
                "bytes": pc.bytes / 2, "share_of_iteration": ms_pcap / ms_per_it, "peak_kind": peak_kind,
                "note": "algorithmic bytes = SURVEY.md 8d (12 nnz(T) + 20 n per sweep); the pencil kernel streams the VALUES "
                        "only (no column indices, no permutation): traffic < bytes.  Bound by the dependency chain "
                        "(3N-2 hyperplanes), not by HBM"}
        info = pc.info()
    elif pckind == "amg":
        ms_pcap = timed(lambda: pc.apply(y, b), 10)
        pc_gbs = pc.bytes / ms_pcap / 1e6
        lv = pc.hierarchy.levels
        roof = {"bound": "hbm", "achieved": pc_gbs, "peak": peak, "unit": "GB/s", "frac": pc_gbs / peak, "traffic": None,
                "kernel": "one V-cycle (gs_sweep_kernel x4 + residual/restriction/prolongation SpMVs per level)",
                "ms": ms_pcap, "bytes": pc.bytes, "share_of_iteration": ms_pcap / ms_per_it, "peak_kind": peak_kind,
                "levels": [[L["n"], int(L["A"][0][-1])] for L in lv]}
        info = {}
    else:
        roof = dict(roof_spmv, kernel="spmv_tiles_kernel", share_of_iteration=ms_spmv / ms_per_it)
        info = {}

    line = {"metric": "%s_iterations_per_second" % args.workload, "value": value, "unit": "iter/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "n": n, "nnz": nnz, "tol_rel": 1e-7, "iterations_per_solve": nits,
                       "residual": residual, "check_every": args.check_every, "tri_levels": info.get("levels_L"),
                       "l2": "working set (CSR %.2f GB + ILU factors) far exceeds the 126 MB L2; no flush needed"
                             % (dA.spmv_bytes / 1e9)},
            "ms_per_iteration": ms_per_it, "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
            "roofline": roof, "roofline_spmv": roof_spmv,
            "setup_s": {"generate": t_gen, "pc_host_setup_and_upload": t_pc, "host_threads": int(L.lsspg_host_threads())}}
    par = parity_block(args, ctx, api, solver, dA, pc, n)
    if par:
        line["config"].update(par)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, A, solver, pckind, pc)
    print(json.dumps(line))
    check(L.lsspg_host_free(hb))
    check(L.lsspg_host_free(hx))
    return 0


if __name__ == "__main__":
    sys.exit(main())

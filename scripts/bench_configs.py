#!/usr/bin/env python
"""The other BASELINE.json configurations on one B200 (bench.py itself times configs[1]):
  C1  2-D 5-point 1000^2, unpreconditioned CG                        (configs[0])
  C2  3-D 7-point 256^3, CG + ILU(0)                                 (configs[1], for reference)
  C3  3-D convection-diffusion 256^3, BiCGStab + ILUK(1), GMRES(30) + ILUT   (configs[2])
  C4' 3-D 7-point 256^3, CG + SX-AMG-style V-cycle (one GPU's share of configs[3])
  C5' power-law CSR (n = 4 M, ~80 M nnz: configs[4] scaled to one quick run), SpMV + IDRS(4)
Per case: SpMV GB/s and fraction of the measured HBM peak, solver iterations/s, time to solution,
algorithmic bytes per iteration -> fraction of the HBM roofline of the whole iteration.
Usage: python scripts/bench_configs.py [--quick] [--only C1,C3a]     (one JSON line per case)"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssp_b200 import api, generators as g  # noqa: E402
from lssp_b200._lib import check, lib  # noqa: E402

PEAK = 6457.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(ctx, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    t = C.c_double()
    check(lib().lsspg_timer_start(ctx.h, 2))
    for _ in range(reps):
        fn()
    check(lib().lsspg_timer_stop(ctx.h, 2, C.byref(t)))
    return t.value / reps


def iteration_bytes(solver, n, nnz, pc_bytes, s=4, m=30):
    """fused-minimum HBM bytes of one iteration, SURVEY.md 8d"""
    extra = 8.0 * n if pc_bytes else 0.0
    if solver == "cg":
        return 12.0 * nnz + 92.0 * n + pc_bytes + extra
    if solver == "bicgstab":
        return 24.0 * nnz + 176.0 * n + 2 * pc_bytes
    if solver == "idrs":
        return 12.0 * nnz + pc_bytes + (24.0 * s + 108.0) * n
    if solver == "gmres":   # average inner step i = (m-1)/2, cycle-end work spread over m steps
        i = (m - 1) / 2.0
        return 12.0 * nnz + 20.0 * n + pc_bytes + (32.0 * i + 56.0) * n + (8.0 * n * (m + 2) + 12.0 * nnz + 28.0 * n) / m
    return None


def run_case(ctx, tag, desc, A, solver, pckind, maxit=3000, **kw):
    n, nnz = len(A[0]) - 1, int(A[0][-1])
    t0 = time.perf_counter()
    dA = api.Csr(ctx, A)
    if pckind == "non":
        pc = api.Preconditioner.non(ctx, n)
    elif pckind == "iluk0":
        pc = api.Preconditioner.iluk(ctx, A, level=0)
    elif pckind == "iluk1":
        pc = api.Preconditioner.iluk(ctx, A, level=1)
    elif pckind == "ilut":
        pc = api.Preconditioner.ilut(ctx, A)
    elif pckind in ("amg", "amg_mc"):
        pc = api.Preconditioner.sxamg(ctx, A, share=dA, zero_guess=1, cf_order=2 if pckind == "amg_mc" else 1)
    t_setup = time.perf_counter() - t0
    # b = 1 as exam.cxx; the power-law operator has unit row sums (x = 1 would solve it at once): oscillating b
    rhs = np.ones(n) if not tag.startswith("C5") else np.sin(np.arange(n) * 0.37) + 1.5
    b, x, y = ctx.upload(rhs), ctx.zeros(n), ctx.empty(n)
    ms_spmv = timed(ctx, lambda: dA.mv(api.MV_MXY, b, y))
    ms_pc = timed(ctx, lambda: pc.apply(y, b), reps=5, warm=2) if pckind != "non" else 0.0
    best = None
    for rep in range(2):
        check(lib().lsspg_memset_zero(ctx.h, x.ptr, C.c_size_t(8 * n)))
        r = api.solve_device(ctx, solver, dA, pc, b, x, maxit=maxit, **kw)
        if best is None or r["solve_ms"] < best["solve_ms"]:
            best = r
    r = best
    ms_it = r["solve_ms"] / max(r["nits"], 1)
    pcb = pc.bytes if pckind != "non" else 0.0
    ib = iteration_bytes(solver, n, nnz, pcb, s=kw.get("idrs", 4), m=kw.get("restart", 30))
    row = {"case": tag, "workload": desc, "n": n, "nnz": nnz, "solver": solver, "pc": pckind,
           "spmv_ms": ms_spmv, "spmv_gbs": dA.spmv_bytes / ms_spmv / 1e6, "spmv_frac_of_measured_peak": dA.spmv_bytes / ms_spmv / 1e6 / PEAK,
           "spmv_frac_of_8TBs": dA.spmv_bytes / ms_spmv / 1e6 / 8000.0,
           "pc_apply_ms": ms_pc, "pc_gbs": (pcb / ms_pc / 1e6) if ms_pc else None,
           "iterations": r["nits"], "converged": r["nits"] < maxit, "residual": r["residual"],
           "time_to_solution_ms": r["solve_ms"], "ms_per_iteration": ms_it, "iterations_per_s": 1e3 / ms_it,
           "iteration_bytes": ib, "iteration_gbs": ib / ms_it / 1e6 if ib else None,
           "iteration_frac_of_measured_peak": ib / ms_it / 1e6 / PEAK if ib else None,
           "host_setup_s": t_setup, "peak_gbs": PEAK}
    print(json.dumps({k: (float("%.6g" % v) if isinstance(v, float) else v) for k, v in row.items()}), flush=True)
    pc.free()
    dA.free()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="small grids (smoke run of this script)")
    ap.add_argument("--only", default="")
    ap.add_argument("--pl-rows", type=int, default=4000000,
                    help="rows of the power-law case C5 (BASELINE.json configs[4] is 50 000 000: ~1 G entries, 12 GB CSR; "
                         "the generator runs row blocks on a thread pool, about a minute on 16 cores)")
    a = ap.parse_args()
    only = set(a.only.split(",")) if a.only else None
    N3, N2, NP = (48, 200, 100000) if a.quick else (256, 1000, a.pl_rows)
    ctx = api.Context(0)
    cases = [
        ("C1", "lap2d_%d CG unpreconditioned (exam.cxx operator)" % N2, lambda: g.laplacian_5pt(N2), "cg", "non", {}),
        ("C2", "lap3d_%d CG+ILU(0)" % N3, lambda: g.lap3d(N3), "cg", "iluk0", {}),
        ("C2b", "lap3d_%d BiCGStab+ILU(0)" % N3, lambda: g.lap3d(N3), "bicgstab", "iluk0", {}),
        ("C3a", "cd3d_%d BiCGStab+ILUK(1)" % N3, lambda: g.cd3d(N3), "bicgstab", "iluk1", {}),
        ("C3b", "cd3d_%d GMRES(30)+ILUT" % N3, lambda: g.cd3d(N3), "gmres", "ilut", dict(restart=30)),
        ("C4", "lap3d_%d CG+SX-AMG-style V-cycle (zero initial guess)" % N3, lambda: g.lap3d(N3), "cg", "amg", {}),
        ("C4b", "lap3d_%d CG+SX-AMG-style V-cycle, multicolour smoother (zero initial guess)" % N3, lambda: g.lap3d(N3), "cg", "amg_mc", {}),
        ("C5", "powerlaw n=%d IDRS(4) unpreconditioned" % NP, lambda: g.powerlaw(NP), "idrs", "non", dict(idrs=4)),
    ]
    for tag, desc, gen, solver, pckind, kw in cases:
        if only and tag not in only:
            continue
        t0 = time.perf_counter()
        A = gen()
        sys.stderr.write("%s: generated in %.1f s\n" % (tag, time.perf_counter() - t0))
        run_case(ctx, tag, desc, A, solver, pckind, **kw)
        del A
    ctx.close()


if __name__ == "__main__":
    main()

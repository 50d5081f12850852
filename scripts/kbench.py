#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA-event timed on the library's stream): SpMV, BLAS-1,
triangular sweeps, and short CG / BiCGStab runs.  Usage: python scripts/kbench.py [N ...]"""
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
    L = lib()
    for _ in range(warm):
        fn()
    t = C.c_double()
    check(L.lsspg_timer_start(ctx.h, 2))
    for _ in range(reps):
        fn()
    check(L.lsspg_timer_stop(ctx.h, 2, C.byref(t)))
    return t.value / reps


def main():
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [128, 256]
    opts = [a for a in sys.argv[1:] if not a.isdigit()]
    ctx = api.Context(0)
    if "pipe" in opts:      # SpMV through the bulk-copy pipeline kernel
        ctx.set_option(api.OPT_SPMV_KERNEL, 2)
    if "ldg" in opts:       # SpMV through the register-staged kernel
        ctx.set_option(api.OPT_SPMV_KERNEL, 1)
    out = []
    for N in sizes:
        t0 = time.time()
        A = g.cd3d(N) if "cd" in opts else g.lap3d(N)
        n, nnz = N ** 3, int(A[0][-1])
        dA = api.Csr(ctx, A)
        x, y, z = ctx.upload(np.ones(n)), ctx.upload(np.full(n, 0.5)), ctx.empty(n)
        row = {"N": N, "n": n, "nnz": nnz}
        ms = timed(ctx, lambda: dA.mv(api.MV_MXY, x, z))
        row["spmv_ms"], row["spmv_gbs"] = ms, dA.spmv_bytes / ms / 1e6
        ms = timed(ctx, lambda: dA.mv(api.MV_AMXPBYZ, x, z, alpha=-1.0, beta=1.0, y=y))
        row["resid_ms"], row["resid_gbs"] = ms, (dA.spmv_bytes + 8 * n) / ms / 1e6
        ms = timed(ctx, lambda: api.lssp_vec_axpbyz(ctx, 1.5, x, 0.5, y, z))
        row["axpbyz_gbs"] = 24.0 * n / ms / 1e6
        ms = timed(ctx, lambda: api.lssp_vec_copy(ctx, z, x))
        row["copy_gbs"] = 16.0 * n / ms / 1e6
        ms = timed(ctx, lambda: api.lssp_vec_dot(ctx, x, y))
        row["dot_ms"], row["dot_gbs"] = ms, 16.0 * n / ms / 1e6
        if "notri" not in opts:
            pc = api.Preconditioner.iluk(ctx, A, level=0)
            info = pc.info()
            ms = timed(ctx, lambda: pc.apply(z, x), reps=10)
            row["ilu0_apply_ms"], row["ilu0_gbs"], row["levels"] = ms, pc.bytes / ms / 1e6, info["levels_L"]
            row["us_per_level"] = 1e3 * ms / (info["levels_L"] + info["levels_U"])
            b, sol = ctx.upload(np.ones(n)), ctx.zeros(n)
            for solver in ("cg", "bicgstab"):
                check(lib().lsspg_memset_zero(ctx.h, sol.ptr, C.c_size_t(8 * n)))
                r = api.solve_device(ctx, solver, dA, pc, b, sol, maxit=30)
                row[solver + "_ilu0_ms_per_it"] = r["solve_ms"] / r["nits"]
            pcn = api.Preconditioner.non(ctx, n)
            for solver in ("cg", "bicgstab"):
                check(lib().lsspg_memset_zero(ctx.h, sol.ptr, C.c_size_t(8 * n)))
                r = api.solve_device(ctx, solver, dA, pcn, b, sol, maxit=50)
                row[solver + "_non_ms_per_it"] = r["solve_ms"] / r["nits"]
            pc.free()
        row["setup_s"] = time.time() - t0
        out.append(row)
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}))
        dA.free()
    ctx.close()


if __name__ == "__main__":
    main()

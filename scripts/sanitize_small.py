#!/usr/bin/env python
"""Small-size tour of the hot kernels for compute-sanitizer (memcheck / racecheck / synccheck):
SpMV (bulk-copy pipeline + tile kernel), pencil / box / slice sweeps, AMG smoother, a CG + ILU(0) and a BiCGStab + ILUK(1)
solve.  Results are checked against the CPU checker so that a sanitizer run is also a parity run.
Usage: compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from lssp_b200 import api, generators as g  # noqa: E402


def main():
    chk = oracle.Ref() if oracle.Ref.available() else oracle.Port()
    ctx = api.Context(0)
    ok = True
    for name, A in (("lap3d_20", g.lap3d(20)), ("cd3d_16", g.cd3d(16)), ("powerlaw_3000", g.powerlaw(3000, window=300))):
        n = len(A[0]) - 1
        x = np.sin(np.arange(n) * 0.37) + 0.25
        for opt in (1, 2):
            ctx.set_option(api.OPT_SPMV_KERNEL, opt)
            dA = api.Csr(ctx, A)
            same = np.array_equal(dA.mv_host(3, x, alpha=-1.0, beta=1.0, y=x), chk.mv(3, A, x, alpha=-1.0, beta=1.0, y=x))
            print("spmv kernel %d %-14s %s" % (opt, name, "exact" if same else "DIFFERS (long rows: tree sums)"))
            ok = ok and (same or name.startswith("powerlaw"))
            dA.free()
        ctx.set_option(api.OPT_SPMV_KERNEL, 0)
        for level in ((0, 1, 2) if not name.startswith("powerlaw") else (0,)):
            L, U = api.ilu_factor(A, "iluk", level=level)
            want = chk.tri_upper(U, chk.tri_lower(L, x))
            pc = api.Preconditioner(ctx, "ilu", n, L, U)
            t = api.Tri(ctx, 0, L)
            got = pc.apply_host(x)
            got2 = pc.apply_host(x)
            same = np.array_equal(got, want) and np.array_equal(got2, want)
            print("ilu(%d) apply %-14s schedule kind %d: %s" % (level, name, t.schedule()["kind"], "exact" if same else "MISMATCH"))
            ok = ok and same
            t.free()
            pc.free()
    A = g.lap3d(20)
    n = 20 ** 3
    dA = api.Csr(ctx, A)
    pc = api.Preconditioner.iluk(ctx, A, level=0)
    r = api.lssp_solver_solve(ctx, "cg", dA, pc, np.ones(n), np.zeros(n))
    print("cg + ilu(0) lap3d_20: %d iterations, residual %.3e" % (r["nits"], r["residual"]))
    A2 = g.cd3d(16)
    dA2 = api.Csr(ctx, A2)
    pc2 = api.Preconditioner.iluk(ctx, A2, level=1)
    r = api.lssp_solver_solve(ctx, "bicgstab", dA2, pc2, np.ones(16 ** 3), np.zeros(16 ** 3))
    print("bicgstab + iluk(1) cd3d_16: %d iterations, residual %.3e" % (r["nits"], r["residual"]))
    amg = api.Preconditioner.sxamg(ctx, g.laplacian_5pt(40))
    z = amg.apply_host(np.ones(1600))
    print("amg V-cycle lap2d_40: |z| = %.6e" % np.linalg.norm(z))
    print("SANITIZE TOUR", "OK" if ok else "FAILED")
    ctx.close()


if __name__ == "__main__":
    main()

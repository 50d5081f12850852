#!/usr/bin/env python
"""Times the L and U sweeps of ILU(k) on an N^3 7-point grid (CUDA events on the library's stream) and, with
LSSPG_TRI_PROF=1, prints the per-pencil timers of the pencil schedule (tri_pencil.cu).
Usage: [LSSPG_TRI_PROF=1] python scripts/pencil_prof.py [N] [level] [cd]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssp_b200 import api, generators as g  # noqa: E402
from lssp_b200._lib import check, lib  # noqa: E402


def timed(ctx, fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    t = C.c_double()
    check(lib().lsspg_timer_start(ctx.h, 2))
    for _ in range(reps):
        fn()
    check(lib().lsspg_timer_stop(ctx.h, 2, C.byref(t)))
    return t.value / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    A = g.cd3d(N) if "cd" in sys.argv else g.lap3d(N)
    n = N ** 3
    ctx = api.Context(0)
    L, U = api.ilu_factor(A, "iluk", level=level)
    rhs = ctx.upload(np.sin(np.arange(n) * 0.37) + 0.25)
    x = ctx.empty(n)
    for name, which, T in (("L", 0, L), ("U", 1, U)):
        tri = api.Tri(ctx, which, T)
        ms = timed(ctx, lambda: tri.solve(x, rhs))
        nnz = int(T[0][-1])
        alg = 12.0 * nnz + 20.0 * n
        print("%s sweep N=%d level=%d: %.4f ms  (algorithmic %.1f MB -> %.0f GB/s)  schedule %s" %
              (name, N, level, ms, alg / 1e6, alg / ms / 1e6, tri.schedule()), flush=True)
        if os.environ.get("LSSPG_TRI_PROF"):
            npn = tri.schedule()["boxes"]
            buf = (C.c_ulonglong * (8 * npn))()
            got = lib().lsspg_debug_tri_pencil_prof(ctx.h, tri.h, buf, npn)
            if got > 0:
                p = np.array(buf[:8 * got], dtype=np.uint64).reshape(got, 8).astype(np.float64)
                t0 = p[:, 0].min()
                start, end = (p[:, 0] - t0) / 1e3, (p[:, 1] - t0) / 1e3
                steps = p[:, 5]
                print("  pencils %d  kernel span %.1f us;  per step (cycles): total %.0f, ghost wait %.0f, barrier %.0f, values wait %.0f, rhs wait %.0f" %
                      (got, end.max(), (p[:, 2] / steps).mean(), (p[:, 3] / steps).mean(), (p[:, 4] / steps).mean(),
                       (p[:, 6] / steps).mean(), (p[:, 7] / steps).mean()))
                print("  pencil duration us: min %.1f mean %.1f max %.1f" % ((end - start).min(), (end - start).mean(), (end - start).max()))
                idx = np.linspace(0, got - 1, min(got, 24)).astype(int)
                for i in idx:
                    print("   ticket %4d: start %7.1f us end %7.1f us  steps %4d  cyc/step %5.0f  ghost %5.0f  bar %5.0f  vals %5.0f  rhs %5.0f" %
                          (i, start[i], end[i], steps[i], p[i, 2] / steps[i], p[i, 3] / steps[i], p[i, 4] / steps[i], p[i, 6] / steps[i], p[i, 7] / steps[i]))
        tri.free()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""GPU check of the skewed box schedule for ILU(1) factors (LSSPG_TRI_SKEW=1) against the slice schedule:
exactness of one application against the oracle's serial sweeps, and time per application.  One JSON line per
case, flushed as it goes (the run may be cut short).

    python tests/skew_check_gpu.py [N ...]        default: 64 128 256
"""
import ctypes as C
import json
import os
import sys
import time

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


def tile_sweep(N, shapes):
    """box shapes for the skewed schedule: ms per ILUK(1) application at N^3 (SKEW_TILES="16,8,4;8,8,8" N)"""
    ctx = api.Context(0)
    A = g.cd3d(N)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, kind="iluk", level=1)
    rhs = np.sin(np.arange(n) * 0.37) + 0.3
    first = None
    for shape in shapes:
        os.environ["LSSPG_TRI_TILE"] = shape
        try:
            pc = api.Preconditioner(ctx, "ilu", n, L, U)
        finally:
            del os.environ["LSSPG_TRI_TILE"]
        x, b = ctx.zeros(n), ctx.upload(rhs)
        pc.apply(x, b)
        got = x.get()
        if first is None:
            first = got
        ms = timed(ctx, lambda: pc.apply(x, b), reps=10, warm=2)
        print(json.dumps({"case": "cd3d_%d ILUK(1) box shape" % N, "shape": shape, "apply_ms": ms,
                          "same_bits_as_first_shape": bool(np.array_equal(got, first))}), flush=True)
        pc.free()
    ctx.close()


def main():
    if os.environ.get("SKEW_TILES"):
        return tile_sweep(int(sys.argv[1]) if len(sys.argv) > 1 else 256, os.environ["SKEW_TILES"].split(";"))
    sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
    ctx = api.Context(0)
    import oracle
    port = oracle.Port()
    for N in sizes:
        A = g.cd3d(N)
        n = len(A[0]) - 1
        t0 = time.perf_counter()
        L, U = api.ilu_factor(A, kind="iluk", level=1)
        t_fac = time.perf_counter() - t0
        rhs = np.sin(np.arange(n) * 0.37) + 0.3
        want = port.ilu_apply(L, U, rhs) if N <= 128 else None
        out = {"case": "cd3d_%d ILUK(1)" % N, "n": n, "factor_s": t_fac}
        first = None
        for skew in ("0", "1"):
            os.environ["LSSPG_TRI_SKEW"] = skew
            t0 = time.perf_counter()
            pc = api.Preconditioner(ctx, "ilu", n, L, U)
            t_up = time.perf_counter() - t0
            x, b = ctx.zeros(n), ctx.upload(rhs)
            pc.apply(x, b)
            got = x.get()
            ms = timed(ctx, lambda: pc.apply(x, b))
            info = pc.info()
            key = "skew" if skew == "1" else "slices"
            out[key] = {"apply_ms": ms, "schedule_and_upload_s": t_up, "info": info,
                        "exact_vs_oracle": None if want is None else bool(np.array_equal(got, want))}
            if first is None:
                first = got
            else:
                out[key]["equal_to_slices"] = bool(np.array_equal(got, first))
            pc.free()
            print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

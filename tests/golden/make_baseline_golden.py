#!/usr/bin/env python
"""Golden values at the BASELINE.json sizes (256^3), produced by the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference by oracle/Makefile) in the build container -- the reference itself cannot travel to the GPU box fast
enough to re-run whole 256^3 solves inside the test-suite (a full CG + ILU(0) solve takes ~2 minutes on one core).

Per case: SHA-256 of the reference's kernel outputs (SpMV x4, ILU application), iterations to tolerance, final residual
and the first 20 residuals in full precision (solver re-run with maxit = 1..20 from the same x0 on ONE assembled session,
SURVEY.md 8c method 1).  Inputs are the deterministic generators of lssp_b200/generators.py with b = 1, x0 = 0 and the
test vector of tests/util.py.

Usage: python tests/golden/make_baseline_golden.py [N=256] [cases...]   ->  tests/golden/baseline_<N>.json"""
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from lssp_b200 import generators as g  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def tvec(n, k=0):
    i = np.arange(n, dtype=np.float64)
    return np.sin(i * (0.37 + 0.11 * k)) + 0.25 * np.cos(i * 1.3 + k)


def session_history(ref, solver, pc, A, b, k, **kw):
    """first k residuals (full precision) + the converged solve, on one assembled reference session"""
    L = ref.lib
    L.ref_session_create.restype = C.c_void_p
    L.ref_session_solve.restype = C.c_int
    n = len(A[0]) - 1
    prm = oracle.ref_params(**kw)
    tasm = C.c_double()
    h = C.c_void_p(L.ref_session_create(oracle.SOLVERS[solver], oracle.PCS[pc], n, A[0].ctypes.data_as(C.c_void_p),
                                        A[1].ctypes.data_as(C.c_void_p), A[2].ctypes.data_as(C.c_void_p), C.byref(prm),
                                        C.byref(tasm)))
    hist = []
    res, sec = C.c_double(), C.c_double()
    for m in range(1, k + 1):
        x = np.zeros(n)
        nits = L.ref_session_solve(h, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), m, C.byref(res), C.byref(sec))
        if nits < m:
            break
        hist.append(res.value)
        print("    maxit %2d: residual %.17g (%.1f s)" % (m, res.value, sec.value), flush=True)
    x = np.zeros(n)
    nits = L.ref_session_solve(h, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), kw.get("maxit", 3000), C.byref(res), C.byref(sec))
    L.ref_session_destroy(h)
    return dict(history=hist, nits=int(nits), residual=res.value, x_sha=sha(x), x_norm=float(np.linalg.norm(x)),
                solve_seconds=sec.value, assemble_seconds=tasm.value)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    only = [a for a in sys.argv[1:] if not a.isdigit()]
    ref = oracle.Ref()
    path = os.path.join(HERE, "baseline_%d.json" % N)
    out = json.load(open(path)) if os.path.exists(path) else {}
    out["_how"] = "tests/golden/make_baseline_golden.py %d: unmodified reference (oracle/_ref), g++ -O2 -ffp-contract=off" % N
    n = N ** 3
    b = np.ones(n)

    def save():
        with open(path, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)

    def want(tag):
        return not only or tag in only

    if want("kernels"):
        t0 = time.time()
        A = g.lap3d(N)
        x, y = tvec(n), tvec(n, 1)
        e = {"n": n, "nnz": int(A[0][-1])}
        e["mxy_sha"] = sha(ref.mv(0, A, x))
        e["amxy_sha"] = sha(ref.mv(1, A, x, alpha=-0.75))
        e["amxpby_sha"] = sha(ref.mv(2, A, x, alpha=1.25, beta=-0.5, y=y))
        e["amxpbyz_sha"] = sha(ref.mv(3, A, x, alpha=-1.0, beta=1.0, y=y))
        Lf, Uf = ref.ilu(A, "iluk", level=0)
        e["ilu0_nnz"] = [int(Lf[0][-1]), int(Uf[0][-1])]
        e["ilu0_factor_sha"] = sha(np.concatenate([Lf[2], Uf[2]]))
        yy = ref.tri_lower(Lf, x)
        e["ilu0_lower_sha"] = sha(yy)
        e["ilu0_apply_sha"] = sha(ref.tri_upper(Uf, yy))
        out["lap3d/kernels"] = e
        save()
        print("lap3d kernels: %.0f s" % (time.time() - t0), flush=True)
        Ac = g.cd3d(N)
        e = {"n": n, "nnz": int(Ac[0][-1])}
        e["amxpbyz_sha"] = sha(ref.mv(3, Ac, x, alpha=-1.0, beta=1.0, y=y))
        Lf, Uf = ref.ilu(Ac, "iluk", level=1)
        e["iluk1_nnz"] = [int(Lf[0][-1]), int(Uf[0][-1])]
        e["iluk1_factor_sha"] = sha(np.concatenate([Lf[2], Uf[2]]))
        e["iluk1_apply_sha"] = sha(ref.tri_upper(Uf, ref.tri_lower(Lf, x)))
        Lf, Uf = ref.ilu(Ac, "ilut", p=-1, tol=1e-3)
        e["ilut_nnz"] = [int(Lf[0][-1]), int(Uf[0][-1])]
        e["ilut_factor_sha"] = sha(np.concatenate([Lf[2], Uf[2]]))
        e["ilut_apply_sha"] = sha(ref.tri_upper(Uf, ref.tri_lower(Lf, x)))
        out["cd3d/kernels"] = e
        save()
        print("cd3d kernels: %.0f s" % (time.time() - t0), flush=True)
    cases = [("lap3d/cg+iluk0", g.lap3d, "cg", "iluk", dict(iluk_level=0, maxit=3000)),
             ("lap3d/bicgstab+iluk0", g.lap3d, "bicgstab", "iluk", dict(iluk_level=0, maxit=3000)),
             ("cd3d/bicgstab+iluk1", g.cd3d, "bicgstab", "iluk", dict(iluk_level=1, maxit=3000)),
             ("cd3d/gmres30+ilut", g.cd3d, "gmres", "ilut", dict(restart=30, maxit=3000))]
    for tag, gen, solver, pc, kw in cases:
        if not want(tag):
            continue
        t0 = time.time()
        print(tag, flush=True)
        A = gen(N)
        out[tag] = session_history(ref, solver, pc, A, b, 20, **kw)
        save()
        print("  %s: nits %d residual %.9e (%.0f s)" % (tag, out[tag]["nits"], out[tag]["residual"], time.time() - t0), flush=True)


if __name__ == "__main__":
    main()

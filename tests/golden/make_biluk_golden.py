"""Block-ILU(k) fixtures (tests/golden/biluk_golden.json) from the UNMODIFIED reference sources compiled with
USE_BLAS = USE_LAPACK = 1 against oracle/blas_standin.c (oracle/_ref/liblssp_refb.so): factors, one application,
whole solves through lssp_solver_create/assemble/solve with LSSP_PC_BILUK.

    python tests/golden/make_biluk_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle  # noqa: E402
from util import matrix, sha, tvec  # noqa: E402

FACTOR_CASES = [("cd3d_12", 4, 0), ("cd3d_12", 4, 1), ("cd3d_12", 3, 2), ("cd3d_12", 1, 1), ("lap2d_100", 2, 1),
                ("lap2d_100", 5, 0), ("random_600", 3, 1), ("random_600", 6, 0), ("lap3d_32", 8, 0), ("cd3d_32", 4, 1)]
SOLVE_CASES = [("cd3d_12", 4, 1, "bicgstab"), ("cd3d_12", 4, 0, "gmres"), ("lap2d_100", 2, 1, "cg"),
               ("lap2d_100", 5, 0, "bicgstab"), ("lap2d_100", 4, 1, "gmres"), ("cd3d_32", 4, 1, "bicgstab"),
               ("cd3d_32", 2, 0, "idrs"), ("lap3d_32", 8, 0, "cg")]

if __name__ == "__main__":
    r = oracle.RefB()
    out = {"factors": {}, "solves": {}}
    for name, bs, level in FACTOR_CASES:
        A = matrix(name)
        n = len(A[0]) - 1
        (L, D, U), x = r.bilu(A, n // bs, level, tvec(n, 2))
        out["factors"]["%s/bs%d/k%d" % (name, bs, level)] = dict(
            nnzL=int(L[0][-1]), nnzD=int(D[0][-1]), nnzU=int(U[0][-1]), L_sha=sha(*L), D_sha=sha(*D), U_sha=sha(*U),
            apply_sha=sha(x))
    for name, bs, level, solver in SOLVE_CASES:
        A = matrix(name)
        n = len(A[0]) - 1
        s = r.solve_biluk(solver, A, np.ones(n), n // bs, level, maxit=3000, restart=30)
        out["solves"]["%s/bs%d/k%d/%s" % (name, bs, level, solver)] = dict(
            nits=int(s["nits"]), residual=float(s["residual"]), xnorm=float(np.linalg.norm(s["x"])))
        print(name, bs, level, solver, s["nits"], s["residual"])
    with open(os.path.join(HERE, "biluk_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

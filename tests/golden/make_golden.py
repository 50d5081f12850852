"""Generate tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref).

Run in the build container, where /root/reference exists:
    make -C oracle ref && python tests/golden/make_golden.py
The fixtures travel with the repository; the GPU box never needs the reference
sources.  Everything is produced by the reference's own code through
oracle/ref_shim.cxx -- kernels, factorisations and whole solves.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from lssp_b200 import generators as g  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


MATRICES = {
    "lap2d_100": lambda: g.laplacian_5pt(100),
    "lap3d_32": lambda: g.lap3d(32),
    "cd3d_32": lambda: g.cd3d(32),
    "cd3d_12": lambda: g.cd3d(12),
    "powerlaw_4000": lambda: g.powerlaw(4000, window=300),
    "random_600": lambda: g.random_csr(600, 5, seed=7),
}


def test_vector(n, k=0):
    i = np.arange(n, dtype=np.float64)
    return np.sin(i * (0.37 + 0.11 * k)) + 0.25 * np.cos(i * 1.3 + k)


def main():
    R = oracle.Ref()
    Rz = oracle.Ref(zero_malloc=True)
    out = {"kernels": {}, "factors": {}, "solves": {}, "histories": {}, "blockjacobi": {}}

    for name, mk in MATRICES.items():
        A = mk()
        n = len(A[0]) - 1
        x, y = test_vector(n), test_vector(n, 1)
        e = {"n": n, "nnz": int(A[0][-1]), "matrix_sha": sha(*A)}
        e["mxy"] = sha(R.mv(0, A, x))
        e["amxy"] = sha(R.mv(1, A, x, alpha=-1.75))
        e["amxpby"] = sha(R.mv(2, A, x, alpha=0.5, beta=-2.0, y=y))
        e["amxpbyz"] = sha(R.mv(3, A, x, alpha=-1.0, beta=1.0, y=y))
        e["dot"] = R.dot(x, y)
        e["norm"] = R.norm(x)
        e["axpby"] = sha(R.axpby(1.25, x, -0.5, y))
        e["axpbyz"] = sha(R.axpbyz(-3.0, x, 0.125, y))
        out["kernels"][name] = e

        for tag, kw in (("iluk0", dict(kind="iluk", level=0)), ("iluk1", dict(kind="iluk", level=1)),
                        ("iluk2", dict(kind="iluk", level=2)), ("ilut", dict(kind="ilut")),
                        ("iluk0_bj4", dict(kind="iluk", level=0, blk_size=(n + 3) // 4)),
                        ("ilut_bj2", dict(kind="ilut", blk_size=(n + 1) // 2))):
            L, U = R.ilu(A, **kw)
            z = np.zeros(n)
            cache = R.tri_lower(L, x)
            z = R.tri_upper(U, cache)
            out["factors"][name + "/" + tag] = {
                "nnzL": int(L[0][-1]), "nnzU": int(U[0][-1]), "L_sha": sha(*L), "U_sha": sha(*U),
                "sumL": float(np.sum(L[2])), "sumU": float(np.sum(U[2])),
                "lower_sha": sha(cache), "apply_sha": sha(z), "apply_norm": float(np.linalg.norm(z))}

    # whole solves: every internal driver x {NON, ILUK(0), ILUK(1), ILUT} on the exam.cxx matrix
    A = MATRICES["lap2d_100"]()
    b = np.ones(len(A[0]) - 1)
    pcs = (("non", {}), ("iluk0", dict(iluk_level=0)), ("iluk1", dict(iluk_level=1)), ("ilut", {}))
    for s in oracle.SOLVERS:
        lib = Rz if s in ("gpbicg", "gpbicr") else R
        for tag, kw in pcs:
            r = lib.solve(s, "non" if tag == "non" else tag[:4], A, b, maxit=3000, restart=30, **kw)
            out["solves"]["lap2d_100/%s/%s" % (s, tag)] = {
                "nits": r["nits"], "residual": r["residual"], "xnorm": float(np.linalg.norm(r["x"]))}

    # 3-D operators incl. 20-step full-precision histories
    cases = [("lap3d_32", "cg", "non", {}), ("lap3d_32", "cg", "iluk", dict(iluk_level=0)),
             ("lap3d_32", "bicgstab", "iluk", dict(iluk_level=0)),
             ("cd3d_32", "bicgstab", "non", {}), ("cd3d_32", "bicgstab", "iluk", dict(iluk_level=0)),
             ("cd3d_32", "bicgstab", "iluk", dict(iluk_level=1)), ("cd3d_32", "bicgstab", "ilut", {}),
             ("cd3d_32", "cg", "non", {}),
             ("cd3d_32", "gmres", "ilut", dict(restart=30)), ("cd3d_32", "gmres", "non", dict(restart=30)),
             ("cd3d_32", "idrs", "non", {}), ("cd3d_32", "idrs", "iluk", dict(iluk_level=0)),
             ("powerlaw_4000", "bicgstab", "iluk", dict(iluk_level=0)), ("powerlaw_4000", "idrs", "non", {}),
             ("powerlaw_4000", "bicgstab", "non", {})]
    for m, s, pc, kw in cases:
        A = MATRICES[m]()
        n = len(A[0]) - 1
        # b = 1 (example/exam.cxx:92-95) except on the power-law matrix, whose row sums are 1
        b = test_vector(n) + 1.5 if m.startswith("powerlaw") else np.ones(n)
        r = R.solve(s, pc, A, b, maxit=3000, **kw)
        key = "%s/%s/%s%s" % (m, s, pc, "".join("_%s%s" % (k[-5:], v) for k, v in sorted(kw.items())))
        out["solves"][key] = {"nits": r["nits"], "residual": r["residual"],
                              "xnorm": float(np.linalg.norm(r["x"]))}
        out["histories"][key] = [float(v) for v in R.history(s, pc, A, b, k=20, **kw)]

    # block-Jacobi ILU (the multi-GPU preconditioner semantics)
    for m, s, pc, kw in [("lap3d_32", "cg", "iluk", dict(iluk_level=0)),
                         ("cd3d_32", "bicgstab", "iluk", dict(iluk_level=0)),
                         ("cd3d_32", "bicgstab", "iluk", dict(iluk_level=1))]:
        A = MATRICES[m]()
        n = len(A[0]) - 1
        b = np.ones(n)
        for P in (1, 2, 4, 8):
            r = R.solve(s, pc, A, b, maxit=3000, blk_size=(n + P - 1) // P, **kw)
            out["blockjacobi"]["%s/%s/%s%d/P%d" % (m, s, pc, kw["iluk_level"], P)] = {
                "nits": r["nits"], "residual": r["residual"]}

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

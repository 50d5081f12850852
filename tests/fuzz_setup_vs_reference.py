"""Randomised comparison of the threaded host set-up with the compiled reference (oracle/_ref): ILU(k) / ILUT incl.
block-Jacobi variants (`ilu`) and block ILU(k) (`bilu`) on random grids, power-law and random matrices.  Not collected
by pytest (minutes per seed); run by hand with different thread counts / pipeline chunks:

    LSSPG_HOST_THREADS=7 LSSPG_PIPE_CHUNK=13 python tests/fuzz_setup_vs_reference.py ilu 1
    LSSPG_HOST_THREADS=5 python tests/fuzz_setup_vs_reference.py bilu 4

Round 1: 120 + 24 cases, 0 mismatches (threads 5 / 7 / 8 / 16, chunks 1 / 13 / default)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from lssp_b200 import api, generators as g  # noqa: E402


def same(F, G):
    return all(np.array_equal(a, b) for X, Y in zip(F, G) for a, b in zip(X, Y))


def fuzz_ilu(seed):
    r = oracle.Ref()
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(10):
        kind = rng.integers(0,4)
        if kind == 0:
            dims = tuple(int(x) for x in rng.integers(6, 40, 3)); n = dims[0]*dims[1]*dims[2]
            A = g.stencil_7pt_rows(dims, 0, n, conv=(0.3,0.2,0.1)); A = (A[0], A[1].astype(np.int32), A[2]); name = "grid%s" % (dims,)
        elif kind == 1:
            n = int(rng.integers(20000, 90000)); A = g.powerlaw(n, window=int(rng.integers(50, 5000))); name = "powerlaw%d" % n
        elif kind == 2:
            n = int(rng.integers(5000, 40000)); A = g.random_csr(n, int(rng.integers(3, 9)), seed=int(rng.integers(1, 1000))); name = "random%d" % n
        else:
            N = int(rng.integers(60, 220)); A = g.laplacian_5pt(N); n = N*N; name = "lap2d%d" % N
        n = len(A[0]) - 1
        for kw in (dict(kind="iluk", level=int(rng.integers(0,4))), dict(kind="ilut", p=int(rng.integers(2,12)), tol=float(10.0**-rng.integers(1,6))),
                   dict(kind="iluk", level=int(rng.integers(0,3)), blk_size=int(rng.integers(n//7+1, n))), dict(kind="ilut", p=int(rng.integers(2,9)), tol=1e-3, blk_size=int(rng.integers(n//5+1, n)))):
            ok = same(api.ilu_factor(A, **kw), r.ilu(A, **kw))
            bad += not ok
            print(name, kw, "OK" if ok else "MISMATCH", flush=True)
    print("mismatches:", bad)


def fuzz_bilu(seed):
    r = oracle.RefB()
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(12):
        bs = int(rng.integers(1, 9))
        nb = int(rng.integers(300, 20000))
        n = nb * bs
        kind = rng.integers(0, 3)
        if kind == 0: A = g.random_csr(n, int(rng.integers(3, 8)), seed=int(rng.integers(1, 1000)))
        elif kind == 1: A = g.powerlaw(n, window=int(rng.integers(20, 2000)))
        else:
            N = int(np.sqrt(n)); N -= N % bs; N = max(N, bs); A = g.laplacian_5pt(N); n = N * N
            if n % bs: continue
        n = len(A[0]) - 1
        lev = int(rng.integers(0, 3))
        try:
            F = api.bilu_factor(A, n // bs, lev)
        except Exception as e:
            print("ours error", n, bs, lev, str(e)[:60]); continue
        ok = same(F, r.bilu(A, n // bs, lev)); bad += not ok
        print(n, bs, lev, "OK" if ok else "MISMATCH", flush=True)
    print("mismatches:", bad)


if __name__ == "__main__":
    (fuzz_bilu if sys.argv[1] == "bilu" else fuzz_ilu)(int(sys.argv[2]) if len(sys.argv) > 2 else 1)

"""Host-side set-up profile (no GPU needed): factorisation, schedule and packing times of the triangular factors.

    python scripts/setup_prof.py [N] [kind] [level]     kind: iluk | ilut; matrix: lap3d (iluk 0) or cd3d otherwise
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lssp_b200 import api, generators  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kind = sys.argv[2] if len(sys.argv) > 2 else "iluk"
level = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t = time.perf_counter()
A = generators.lap3d(N) if (kind == "iluk" and level == 0) else generators.cd3d(N)
print("generate %.2f s  n=%d nnz=%d" % (time.perf_counter() - t, len(A[0]) - 1, len(A[1])))
t = time.perf_counter()
L, U = api.ilu_factor(A, kind=kind, level=level, p=7 if kind == "ilut" else -1, tol=1e-3)
print("factor   %.2f s  nnz(L)=%d nnz(U)=%d" % (time.perf_counter() - t, len(L[1]), len(U[1])))
for which, T, name in ((0, L, "L"), (1, U, "U")):
    t = time.perf_counter()
    r = api.tri_pack_host(which, T)
    print("%s: kind %d  %016x  %.3f GB  schedule %.2f s  pack %.2f s  (call %.2f s)"
          % (name, r["kind"], r["fingerprint"], r["bytes"] / 1e9, r["schedule_s"], r["pack_s"], time.perf_counter() - t))

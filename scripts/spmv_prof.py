#!/usr/bin/env python
"""A few SpMV launches on one matrix (for ncu captures).  Usage: python scripts/spmv_prof.py KIND SIZE [reps]
KIND: powerlaw | lap3d | lap2d | cd3d"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssp_b200 import api, generators as g  # noqa: E402

kind, size = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
A = {"powerlaw": g.powerlaw, "lap3d": g.lap3d, "lap2d": g.laplacian_5pt, "cd3d": g.cd3d}[kind](size)
n = len(A[0]) - 1
ctx = api.Context(0)
dA = api.Csr(ctx, A)
print(kind, size, "n", n, "nnz", int(A[0][-1]), dA.schedule_info())
x, y = ctx.upload(np.sin(np.arange(n) * 0.37) + 1.5), ctx.empty(n)
for _ in range(reps):
    dA.mv(api.MV_MXY, x, y)
ctx.sync()

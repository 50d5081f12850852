#!/usr/bin/env python
"""A few AMG cycles on an N^3 Laplacian (for `ncu --metrics gpu__time_duration.sum` launch lists).
Usage: python scripts/amg_prof.py N [cycles] [cf_order]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssp_b200 import api, generators as g  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 1
order = int(sys.argv[3]) if len(sys.argv) > 3 else 1
A = g.lap3d(N)
n = N ** 3
ctx = api.Context(0)
dA = api.Csr(ctx, A)
t = time.time()
pc = api.Preconditioner.sxamg(ctx, A, share=dA, zero_guess=1, cf_order=order)
print("setup %.1f s" % (time.time() - t), [(L["n"], int(L["A"][0][-1])) for L in pc.hierarchy.levels])
for l in range(len(pc.hierarchy.levels) - 1):
    nl = pc.hierarchy.levels[l]["n"]
    print("level", l, pc.hierarchy.walk_gs_host(l, 0, np.ones(nl), np.zeros(nl))[1])
b, x = ctx.upload(np.ones(n)), ctx.zeros(n)
for _ in range(cycles):
    pc.apply(x, b)
ctx.sync()
print("launches", ctx.launches)

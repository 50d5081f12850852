"""AMG set-up at N^3: host threads vs per-row phases on the device (LSSPG_SETUP_PROF=1 prints the phases of both)."""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from lssp_b200 import api
from lssp_b200 import generators as g

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
A = g.lap3d(N)
ctx = api.Context(0)
api.AmgHierarchy(g.lap3d(16), ctx=ctx)   # context warm-up
t0 = time.perf_counter()
Hd = api.AmgHierarchy(A, ctx=ctx)
t1 = time.perf_counter()
Hh = api.AmgHierarchy(A)
t2 = time.perf_counter()
same = all(np.array_equal(u, v) for a, b in zip(Hd.levels, Hh.levels) for k in ("A", "P", "R") if a[k] is not None
           for u, v in zip(a[k], b[k]))
print("AMG set-up at %d^3: device phases %.3f s, host %.3f s, levels %d, identical %s" % (N, t1 - t0, t2 - t1, len(Hd.levels), same))

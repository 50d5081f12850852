"""CG + ILU(0) at N^3 with the reference-order reductions (LSSPG_OPT_REDUCE_SEQUENTIAL = 2): time per iteration next
to the tree-reduction mode.  Run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lssp_b200 import api

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
modes = [int(m) for m in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 2]
n = N ** 3
for mode in modes:
    c = api.Context(0)
    c.set_option(api.OPT_REDUCE_SEQUENTIAL, mode)
    d = api.DMat.lap3d(c, N)
    L, U = d.ilu_factor(level=0)
    dA = d.to_csr(take=True)
    pc = api.Preconditioner(c, "ilu", n, L, U)
    b, x = c.upload(np.ones(n)), c.zeros(n)
    for rep in range(2):
        x = c.zeros(n)
        r = api.solve_device(c, "cg", dA, pc, b, x, maxit=maxit)
    print("mode %d: nits %d residual %.17g solve %.1f ms = %.3f ms / iteration, %d launches" %
          (mode, r["nits"], r["residual"], r["solve_ms"], r["solve_ms"] / max(r["nits"], 1), r["launches"]))
    pc.free(); dA.free(); c.close()

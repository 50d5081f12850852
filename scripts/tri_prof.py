import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssp_b200 import api, generators as g
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = api.Context(0)
A = g.lap3d(N); n = N**3
pc = api.Preconditioner.iluk(ctx, A, level=0)
x, z = ctx.upload(np.ones(n)), ctx.empty(n)
for _ in range(3):
    pc.apply(z, x)
ctx.sync()

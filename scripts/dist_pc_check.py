#!/usr/bin/env python
"""Debug aid (torchrun): every rank builds its block-Jacobi ILU(0) of the weak-scaling slab problem twice -- pencil schedule
and box schedule -- applies both to the same vector and compares bit for bit; also against the CPU checker on rank-local data."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as td
    from lssp_b200 import api, dist, generators as g
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dims = (N, N, N * world)
    n = dims[0] * dims[1] * dims[2]
    blk, r0, r1 = dist.block_rows(n, world, rank)
    rows = g.stencil_7pt_rows(dims, r0, r1)
    shard = dist.make_shard(rows, n, world, rank)
    ctx = api.Context(local)
    B = shard.diag_block()
    Lf, Uf = api.ilu_factor(B, "iluk", level=0)
    no = shard.n_owned
    off = int(sys.argv[2]) if len(sys.argv) > 2 else rank
    v = np.sin(np.arange(no) * 0.37 + off) + 0.25
    pc1 = api.Preconditioner(ctx, "ilu", no, Lf, Uf)
    os.environ["LSSPG_TRI_PENCIL"] = "0"
    pc2 = api.Preconditioner(ctx, "ilu", no, Lf, Uf)
    del os.environ["LSSPG_TRI_PENCIL"]
    a, b = pc1.apply_host(v), pc2.apply_host(v)
    a2 = pc1.apply_host(v)
    for rep in range(int(os.environ.get('REPS', '0'))):
        a3 = pc1.apply_host(v)
        print('rank %d rep %d: equal to box %s' % (rank, rep, np.array_equal(a3, b)), flush=True)
    info = (int(Lf[0][-1]), int(Uf[0][-1]), int(B[0][-1]), float(np.abs(Lf[2]).sum()), float(np.abs(Uf[2]).sum()))
    print("rank %d: nnz(L,U,B) %s  pencil==box %s  repeat %s  nan %d  mismatches %d  first %s" %
          (rank, info, np.array_equal(a, b), np.array_equal(a, a2), int(np.isnan(a).sum()), int(np.sum(a != b)),
           np.flatnonzero(a != b)[:5]), flush=True)
    td.barrier()
    td.destroy_process_group()


if __name__ == "__main__":
    main()

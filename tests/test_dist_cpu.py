"""CPU suite, part 3: the N > 1 path (row sharding, halo lists, block-Jacobi, global
reductions) -- host logic only, no GPU.  Part of it runs as two real processes over
torch.distributed's `gloo` backend."""
import os
import socket

import numpy as np
import pytest

from lssp_b200 import api, dist
from util import matrix, tvec


def all_shards(A, P):
    n = len(A[0]) - 1
    needs = []
    for r in range(P):
        blk, r0, r1 = dist.block_rows(n, P, r)
        needs.append(dist.needed_ghosts(dist.slice_rows(A, r0, r1)[1], r0, r1, blk))
    out = []
    for r in range(P):
        blk, r0, r1 = dist.block_rows(n, P, r)
        out.append(dist.make_shard(dist.slice_rows(A, r0, r1), n, P, r, all_needs=needs))
    return out


@pytest.mark.parametrize("name", ["lap3d_32", "cd3d_12", "powerlaw_4000", "random_600"])
@pytest.mark.parametrize("P", [2, 3, 8])
def test_sharded_spmv_equals_global_spmv_bit_for_bit(port, name, P):
    A = matrix(name)
    n = len(A[0]) - 1
    x = tvec(n)
    want = port.mv(0, A, x)
    shards = all_shards(A, P)
    for S in shards:
        # what the peers send must be exactly what this rank expects, in the same order
        off = 0
        for p, cnt in zip(S.peers, S.recv_counts):
            T = shards[p]
            k = T.peers.index(S.rank)
            so = int(np.sum(T.send_counts[:k]))
            sent = T.send_idx[so:so + T.send_counts[k]].astype(np.int64) + T.r0
            assert np.array_equal(sent, S.ghost_global[off:off + cnt])
            off += cnt
        assert off == S.n_ghost
        xl = np.concatenate([x[S.r0:S.r1], x[S.ghost_global]])          # [owned ; ghost]
        got = port.mv(0, (S.Ap, S.Aj, S.Ax), xl) if S.n_owned else np.zeros(0)
        # the shard keeps every row's storage order -> same products, same order, same bits
        assert np.array_equal(got[:S.n_owned], want[S.r0:S.r1])
    assert sum(S.n_owned for S in shards) == n


def test_stencil_halo_is_one_plane_per_side():
    N, P = 32, 4
    A = matrix("lap3d_32")
    shards = all_shards(A, P)
    for S in shards:
        sides = (S.rank > 0) + (S.rank < P - 1)
        assert S.n_ghost == sides * N * N and len(S.peers) == sides
        assert all(abs(p - S.rank) == 1 for p in S.peers)


def test_powerlaw_row_blocks_tile_the_global_matrix(port):
    """every rank generates its own rows of the irregular matrix (counter-based hashing): the blocks
    must be exactly the rows of the matrix a single process generates, and shard like any other"""
    from lssp_b200 import generators as g
    n, P = 4000, 3
    A = g.powerlaw(n, window=300)
    x = tvec(n)
    want = port.mv(0, A, x)
    blocks, needs = [], []
    for rank in range(P):
        blk, r0, r1 = dist.block_rows(n, P, rank)
        rows = g.powerlaw_rows(n, r0, r1, window=300)
        ref = dist.slice_rows(A, r0, r1)
        assert all(np.array_equal(u, v) for u, v in zip(rows, ref))
        blocks.append(rows)
        needs.append(dist.needed_ghosts(rows[1], r0, r1, blk))
    for rank in range(P):
        S = dist.make_shard(blocks[rank], n, P, rank, all_needs=needs)
        xl = np.concatenate([x[S.r0:S.r1], x[S.ghost_global]])
        assert np.array_equal(port.mv(0, (S.Ap, S.Aj, S.Ax), xl)[:S.n_owned], want[S.r0:S.r1])


@pytest.mark.parametrize("P", [2, 4])
def test_block_jacobi_factor_equals_reference_blocked_ilu(golden, P):
    """The ILU(0) a rank computes from its own diagonal block is the corresponding block of the
    reference's blocked factorisation (src/pc-iluk.cxx:411-552 with blk_size = ceil(n/P))."""
    A = matrix("cd3d_32")
    n = len(A[0]) - 1
    blk = -(-n // P)
    Lg, Ug = api.ilu_factor(A, "iluk", level=0, blk_size=blk)
    for S in all_shards(A, P):
        L, U = api.ilu_factor(S.diag_block(), "iluk", level=0)
        for (gp, gj, gx), (lp, lj, lx) in ((Lg, L), (Ug, U)):
            a, b = int(gp[S.r0]), int(gp[S.r1])
            assert np.array_equal(gp[S.r0:S.r1 + 1] - gp[S.r0], lp)
            assert np.array_equal(gj[a:b] - S.r0, lj) and np.array_equal(gx[a:b], lx)


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q):
    """One rank of a 2-process CPU run: shard, exchange halos over gloo, block-Jacobi PCG with
    globally reduced dot products (the algorithm the GPU drivers run over NCCL)."""
    import torch
    import torch.distributed as td

    import oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port_no)
    td.init_process_group("gloo", rank=rank, world_size=world)
    P_ = oracle.Port()
    A = matrix("lap3d_32")
    n = len(A[0]) - 1
    blk, r0, r1 = dist.block_rows(n, world, rank)
    S = dist.make_shard(dist.slice_rows(A, r0, r1), n, world, rank)      # all_gather_object inside
    L, U = api.ilu_factor(S.diag_block(), "iluk", level=0)
    local = (S.Ap, S.Aj, S.Ax)

    def exchange(v):                                                      # v: [owned ; ghost]
        reqs, bufs, off, so = [], [], 0, 0
        for p, sc, rc in zip(S.peers, S.send_counts, S.recv_counts):
            sb = torch.from_numpy(np.ascontiguousarray(v[S.send_idx[so:so + sc]]))
            rb = torch.empty(rc, dtype=torch.float64)
            reqs += [td.isend(sb, p), td.irecv(rb, p)]
            bufs.append((off, rb))
            off, so = off + rc, so + sc
        for r in reqs:
            r.wait()
        for o, rb in bufs:
            v[S.n_owned + o:S.n_owned + o + len(rb)] = rb.numpy()

    def gsum(v):
        t = torch.tensor([v], dtype=torch.float64)
        td.all_reduce(t)
        return float(t[0])

    no = S.n_owned
    b = np.ones(no)
    x = np.zeros(no + S.n_ghost)
    xg = tvec(n)
    v = np.concatenate([xg[r0:r1], np.zeros(S.n_ghost)])
    exchange(v)
    spmv_ok = np.array_equal(P_.mv(0, local, v), P_.mv(0, A, xg)[r0:r1])
    # PCG (reference src/solver-cg.cxx:56-117) with block-Jacobi ILU(0)
    r = b - P_.mv(0, local, x)
    res0 = np.sqrt(gsum(P_.dot(r, r)))
    tol = max(1e-7 * res0, 1e-7, 1e-7 * np.sqrt(gsum(P_.dot(b, b))))
    p = np.zeros(no + S.n_ghost)
    rho0, its = 0.0, 0
    for it in range(3000):
        z = P_.ilu_apply(L, U, r)
        rho1 = gsum(P_.dot(z, r))
        p[:no] = z if it == 0 else z + (rho1 / rho0) * p[:no]
        exchange(p)
        q_ = P_.mv(0, local, p)
        alpha = rho1 / gsum(P_.dot(q_, p[:no]))
        rho0 = rho1
        x[:no] += alpha * p[:no]
        r -= alpha * q_
        its = it + 1
        if np.sqrt(gsum(P_.dot(r, r))) <= tol:
            break
    q.put((rank, spmv_ok, its))
    td.destroy_process_group()


def test_two_process_gloo_block_jacobi_cg(golden):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = golden["blockjacobi"]["lap3d_32/cg/iluk0/P2"]["nits"]   # the reference's blocked ILU(0), 2 blocks
    for rank, spmv_ok, its in res:
        assert spmv_ok
        assert abs(its - want) <= 1, (its, want)

"""Row sharding of the solve loop across GPUs: one process per GPU.

Host-side logic only (pure numpy + torch.distributed for the metadata exchange; works
with the `gloo` backend on CPUs, which is how the test-suite covers it):

* `block_rows`     -- contiguous row blocks of ceil(n/P) rows, the block boundaries of the
                      reference's lssp_mat_get_block_diag (src/matrix-utils.cxx:615,626-628), so
                      that the per-rank ILU equals the reference's blocked ILU (block-Jacobi).
* `make_shard`     -- from a rank's rows with GLOBAL columns: which columns are ghosts, who
                      owns them, what this rank must send to whom; columns renumbered
                      [owned ; ghost] with the ghosts grouped by owner rank.
* `DeviceShard`    -- uploads the shard, creates the NCCL communicator and the halo object
                      behind the C ABI (lsspg_comm_* / lsspg_halo_*).

The data path (halo send/recv, all-reduce of dot products) runs inside liblsspg on the
library's stream; torch.distributed is plumbing for rendezvous and set-up only.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib


def block_rows(n, P, rank):
    blk = -(-n // P)
    r0 = min(n, rank * blk)
    return blk, r0, min(n, r0 + blk)


class Shard:
    """Host description of one rank's row block."""

    def __init__(self):
        self.rank = self.P = self.n = self.r0 = self.r1 = self.blk = 0
        self.Ap = self.Aj = self.Ax = None            # local CSR, columns renumbered [owned ; ghost]
        self.ghost_global = None                      # global column of every ghost slot
        self.peers, self.send_counts, self.recv_counts = [], [], []
        self.send_idx = np.zeros(0, np.int32)         # owned row indices, grouped by peer

    @property
    def n_owned(self):
        return self.r1 - self.r0

    @property
    def n_ghost(self):
        return len(self.ghost_global)

    def diag_block(self):
        """The owned x owned block (ghost columns dropped): what the block-Jacobi ILU factors.
        A row left without entries gets a unit diagonal, as lssp_mat_get_block_diag does."""
        keep = self.Aj < self.n_owned
        row_of = np.repeat(np.arange(self.n_owned), np.diff(self.Ap))
        cnt = np.bincount(row_of[keep], minlength=self.n_owned).astype(np.int64)
        Aj, Ax = self.Aj[keep], self.Ax[keep]
        if np.any(cnt == 0):
            rows = row_of[keep]
            empty = np.flatnonzero(cnt == 0)
            rows = np.concatenate([rows, empty])
            Aj = np.concatenate([Aj, empty.astype(Aj.dtype)])
            Ax = np.concatenate([Ax, np.ones(len(empty))])
            order = np.lexsort((Aj, rows))
            Aj, Ax = Aj[order], Ax[order]
            cnt = cnt.copy()
            cnt[empty] = 1
        Ap = np.zeros(self.n_owned + 1, np.int64)
        np.cumsum(cnt, out=Ap[1:])
        return Ap.astype(np.int32), Aj.astype(np.int32), Ax


def needed_ghosts(Aj_global, r0, r1, blk):
    """Ghost columns of a row block, grouped by owner: {owner_rank: sorted global columns}."""
    ext = np.unique(Aj_global[(Aj_global < r0) | (Aj_global >= r1)])
    owners = ext // blk
    return {int(p): ext[owners == p] for p in np.unique(owners)}


def make_shard(rows, n, P, rank, all_needs=None, gather=None):
    """rows = (Ap, Aj_global, Ax) of this rank's block.  `all_needs[q]` is rank q's
    needed_ghosts(); when None it is collected with `gather` (default:
    torch.distributed.all_gather_object)."""
    Ap, Ajg, Ax = rows
    blk, r0, r1 = block_rows(n, P, rank)
    assert len(Ap) - 1 == r1 - r0
    need = needed_ghosts(Ajg, r0, r1, blk)
    if all_needs is None:
        if gather is None:
            import torch.distributed as dist

            def gather(obj):
                out = [None] * dist.get_world_size()
                dist.all_gather_object(out, obj)
                return out
        all_needs = gather(need)
    S = Shard()
    S.rank, S.P, S.n, S.r0, S.r1, S.blk = rank, P, n, r0, r1, blk
    peers = sorted(set(need) | {q for q in range(P) if q != rank and rank in all_needs[q]})
    ghost, send = [], []
    for p in peers:
        g = need.get(p, np.zeros(0, np.int64))
        s = all_needs[p].get(rank, np.zeros(0, np.int64)) if p != rank else np.zeros(0, np.int64)
        ghost.append(np.asarray(g, dtype=np.int64))
        send.append(np.asarray(s, dtype=np.int64) - r0)
        S.recv_counts.append(len(g))
        S.send_counts.append(len(s))
    S.peers = peers
    S.ghost_global = np.concatenate(ghost) if ghost else np.zeros(0, np.int64)
    S.send_idx = (np.concatenate(send) if send else np.zeros(0, np.int64)).astype(np.int32)
    # renumber: owned columns -> col - r0, ghosts -> n_owned + position in the peer-grouped ghost list
    owned = (Ajg >= r0) & (Ajg < r1)
    order = np.argsort(S.ghost_global, kind="stable")
    pos = np.searchsorted(S.ghost_global[order], Ajg[~owned])
    Aj = np.empty(len(Ajg), dtype=np.int64)
    Aj[owned] = Ajg[owned] - r0
    Aj[~owned] = (r1 - r0) + order[pos]
    # keep every row sorted by LOCAL column?  No: the reference sums a row in GLOBAL column order
    # (src/lssp.cxx:173 sorts by global column); the storage order is left untouched so that the
    # row sums are bit-identical to the unsharded SpMV.
    S.Ap, S.Aj, S.Ax = np.asarray(Ap, np.int32), Aj.astype(np.int32), np.asarray(Ax, np.float64)
    return S


def slice_rows(A, r0, r1):
    """Rows [r0, r1) of a global CSR matrix, columns still global (tests / small runs)."""
    Ap, Aj, Ax = A
    a, b = int(Ap[r0]), int(Ap[r1])
    return (Ap[r0:r1 + 1] - Ap[r0]).astype(np.int32), Aj[a:b].astype(np.int64), Ax[a:b]


class DeviceShard:
    """Uploads a Shard and owns the communicator + halo (C ABI lsspg_comm_* / lsspg_halo_*)."""

    def __init__(self, ctx, shard, id_bytes=None):
        """id_bytes: the broadcast NCCL id -> the communicator is created here (first shard of a
        context); None -> the context already has its communicator."""
        from . import api
        L = lib()
        self.ctx, self.shard = ctx, shard
        self.owns_comm = id_bytes is not None
        if self.owns_comm:
            buf = (C.c_char * 128).from_buffer_copy(id_bytes)
            check(L.lsspg_comm_init(ctx.h, shard.rank, shard.P, buf))
            self._connect_p2p()
        self.A = api.Csr(ctx, (shard.Ap, shard.Aj, shard.Ax), num_cols=shard.n_owned + shard.n_ghost)
        self.halo = C.c_void_p()
        np_ = len(shard.peers)
        peers = (C.c_int * max(np_, 1))(*shard.peers)
        sc = (C.c_int * max(np_, 1))(*shard.send_counts)
        rc = (C.c_int * max(np_, 1))(*shard.recv_counts)
        si = np.ascontiguousarray(shard.send_idx, dtype=np.int32)
        check(L.lsspg_halo_create(ctx.h, shard.n_owned, np_, peers, sc, si.ctypes.data_as(C.c_void_p), rc,
                                  C.byref(self.halo)))
        check(L.lsspg_csr_set_halo(self.A.h, self.halo))

    def _connect_p2p(self):
        """Peer-to-peer mailboxes for the one-shot all-reduce of the dot products (comm.cu): CUDA IPC handles exchanged
        with torch.distributed; LSSPG_P2P_ALLREDUCE=0 keeps ncclAllReduce."""
        import os
        if os.environ.get("LSSPG_P2P_ALLREDUCE", "1") == "0" or self.shard.P < 2:
            return
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        L = lib()
        mine = (C.c_char * 64)()
        ok = L.lsspg_comm_p2p_local(self.ctx.h, mine) == 0
        handles = [None] * self.shard.P
        dist.all_gather_object(handles, bytes(mine) if ok else None)
        if any(h is None for h in handles):
            return
        blob = (C.c_char * (64 * self.shard.P)).from_buffer_copy(b"".join(handles))
        ok = L.lsspg_comm_p2p_connect(self.ctx.h, blob) == 0
        flags = [None] * self.shard.P
        dist.all_gather_object(flags, ok)
        if not all(flags):
            raise RuntimeError("peer-to-peer all-reduce: a rank could not map its peers (%s); set LSSPG_P2P_ALLREDUCE=0"
                               % lib().lsspg_last_error().decode())

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        check(lib().lsspg_comm_unique_id(buf))
        return bytes(buf)

    def close(self):
        L = lib()
        if self.halo:
            check(L.lsspg_csr_set_halo(self.A.h, None))
            L.lsspg_halo_destroy(self.ctx.h, self.halo)
            self.halo = C.c_void_p()
        if self.owns_comm:
            L.lsspg_comm_destroy(self.ctx.h)
            self.owns_comm = False

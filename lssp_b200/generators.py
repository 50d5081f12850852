"""Deterministic synthetic CSR inputs.  fp64 values, int32 indices.

All generators return ``(Ap, Aj, Ax)`` numpy arrays with columns sorted ascending
in every row, exactly what ``lssp_solver_assemble`` hands to the drivers
(reference src/lssp.cxx:173).
"""
import numpy as np


def _compress(cols, vals, mask):
    """Row-major compression of an (n, k) candidate table into CSR."""
    counts = mask.sum(axis=1, dtype=np.int64)
    Ap = np.zeros(mask.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=Ap[1:])
    assert Ap[-1] < 2 ** 31, "int32 num_nnzs overflow (reference limit, SURVEY 7.3 item 7)"
    return Ap.astype(np.int32), cols[mask].astype(np.int32), vals[mask].astype(np.float64)


def laplacian_5pt(N):
    """2-D 5-point Laplacian on an N x N grid, row order N,W,C,E,S -- the matrix
    of the reference's example program (example/exam.cxx:4-59)."""
    n = N * N
    idx = np.arange(n, dtype=np.int64)
    i, j = idx // N, idx % N
    cols = np.stack([idx - N, idx - 1, idx, idx + 1, idx + N], axis=1)
    mask = np.stack([i > 0, j > 0, np.ones(n, bool), j < N - 1, i < N - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), (n, 5))
    return _compress(cols, vals, mask)


def stencil_7pt(N, conv=(0.0, 0.0, 0.0), chunk_planes=16):
    """3-D 7-point operator on an N^3 grid, natural order i=(z*N+y)*N+x, entries
    in ascending column order (-N^2,-N,-1,0,+1,+N,+N^2), Dirichlet truncation.
    conv=(0,0,0): Laplacian (diag 6, off -1).  conv=(cx,cy,cz): central-difference
    convection-diffusion, lower neighbours -1-c, upper -1+c."""
    cx, cy, cz = conv
    n = N * N * N
    stencil = np.array([-1.0 - cz, -1.0 - cy, -1.0 - cx, 6.0, -1.0 + cx, -1.0 + cy, -1.0 + cz])
    offs = np.array([-N * N, -N, -1, 0, 1, N, N * N], dtype=np.int64)
    Ap = np.zeros(n + 1, dtype=np.int64)

    def planes(z0):   # rows of the planes [z0, z0 + chunk_planes): independent of every other chunk
        z1 = min(N, z0 + chunk_planes)
        idx = np.arange(z0 * N * N, z1 * N * N, dtype=np.int64)
        x, y, z = idx % N, (idx // N) % N, idx // (N * N)
        mask = np.stack([z > 0, y > 0, x > 0, np.ones(len(idx), bool), x < N - 1, y < N - 1,
                         z < N - 1], axis=1)
        cols = idx[:, None] + offs[None, :]
        Ap[idx + 1] = mask.sum(axis=1)
        return cols[mask].astype(np.int32), np.broadcast_to(stencil, mask.shape)[mask]

    starts = list(range(0, N, chunk_planes))
    if len(starts) > 1:   # chunks on a thread pool (numpy releases the GIL in its kernels), concatenated in order
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max(1, min(16, os.cpu_count() or 1))) as ex:
            parts = list(ex.map(planes, starts))
    else:
        parts = [planes(z0) for z0 in starts]
    np.cumsum(Ap, out=Ap)
    assert Ap[-1] < 2 ** 31
    return Ap.astype(np.int32), np.concatenate([q[0] for q in parts]), np.concatenate([q[1] for q in parts])


def stencil_7pt_rows(dims, r0, r1, conv=(0.0, 0.0, 0.0), chunk=1 << 20):
    """Rows [r0, r1) of the 7-point operator on an nx x ny x nz grid, with GLOBAL column
    indices (int64) -- what one rank of a row-sharded run generates for itself.
    Returns (Ap int32 relative to r0, Aj int64 global, Ax)."""
    nx, ny, nz = dims
    cx, cy, cz = conv
    stencil = np.array([-1.0 - cz, -1.0 - cy, -1.0 - cx, 6.0, -1.0 + cx, -1.0 + cy, -1.0 + cz])
    offs = np.array([-nx * ny, -nx, -1, 0, 1, nx, nx * ny], dtype=np.int64)
    Ap = np.zeros(r1 - r0 + 1, dtype=np.int64)
    Aj_parts, Ax_parts = [], []
    for a in range(r0, r1, chunk):
        idx = np.arange(a, min(r1, a + chunk), dtype=np.int64)
        x, y, z = idx % nx, (idx // nx) % ny, idx // (nx * ny)
        mask = np.stack([z > 0, y > 0, x > 0, np.ones(len(idx), bool), x < nx - 1, y < ny - 1,
                         z < nz - 1], axis=1)
        cols = idx[:, None] + offs[None, :]
        Ap[idx - r0 + 1] = mask.sum(axis=1)
        Aj_parts.append(cols[mask])
        Ax_parts.append(np.broadcast_to(stencil, mask.shape)[mask])
    np.cumsum(Ap, out=Ap)
    assert Ap[-1] < 2 ** 31
    Aj = np.concatenate(Aj_parts) if Aj_parts else np.zeros(0, np.int64)
    Ax = np.concatenate(Ax_parts) if Ax_parts else np.zeros(0)
    return Ap.astype(np.int32), Aj, Ax


def lap3d(N):
    return stencil_7pt(N)


def cd3d(N):
    return stencil_7pt(N, conv=(0.3, 0.2, 0.1))


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def _u01(h):
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


_PL_BLOCK = 1 << 18   # rows per generator block


def powerlaw(n, seed=20261018, lmin=5, gamma=2.3, lmax=4096, window=65536, local=0.9):
    """Irregular CSR with power-law row lengths and column locality (SURVEY.md 8d,
    C5): l_i = clamp(floor(lmin * u^(-1/(gamma-1))), 1, lmax) off-diagonal draws,
    90 % within +-window of the diagonal and 10 % uniform; off-diagonals uniform in
    (-1,0), duplicates merged (summed), diagonal = 1 + sum |off| (strictly
    diagonally dominant, nonsymmetric).  Counter-based splitmix64 hashing of
    (seed, row, k) makes the matrix reproducible anywhere."""
    if n <= _PL_BLOCK:
        return powerlaw_rows(n, 0, n, seed, lmin, gamma, lmax, window, local)
    # every row is a function of (seed, row) only: row blocks on a thread pool (numpy releases the GIL in its
    # kernels), concatenated -- the same arrays as one call over all rows
    import os
    from concurrent.futures import ThreadPoolExecutor
    starts = list(range(0, n, _PL_BLOCK))
    with ThreadPoolExecutor(max(1, min(16, os.cpu_count() or 1))) as ex:
        parts = list(ex.map(lambda r0: powerlaw_rows(n, r0, min(n, r0 + _PL_BLOCK), seed, lmin, gamma, lmax, window, local),
                            starts))
    Ap = np.zeros(n + 1, dtype=np.int64)
    off = 0
    for r0, (p, _, _) in zip(starts, parts):
        Ap[r0 + 1:r0 + len(p)] = off + p[1:].astype(np.int64)
        off += int(p[-1])
    if off >= 2 ** 31:
        raise ValueError("powerlaw: more than 2^31 - 1 entries (int32 CSR)")
    return Ap.astype(np.int32), np.concatenate([q[1] for q in parts]), np.concatenate([q[2] for q in parts])


def powerlaw_rows(n, r0, r1, seed=20261018, lmin=5, gamma=2.3, lmax=4096, window=65536, local=0.9):
    """Rows [r0, r1) of powerlaw(n) with GLOBAL column indices (a rank's row block): every row is a
    function of (seed, row) only, so the blocks of all ranks tile the same matrix."""
    m = r1 - r0
    with np.errstate(over="ignore"):
        rows = np.arange(r0, r1, dtype=np.uint64)
        base = _splitmix64(np.uint64(seed) ^ (rows * np.uint64(0xD1342543DE82EF95)))
        u = _u01(_splitmix64(base))
        ell = np.floor(lmin * u ** (-1.0 / (gamma - 1.0)))
        ell = np.clip(ell, 1, min(lmax, max(1, n - 1))).astype(np.int64)
        tot = int(ell.sum())
        rl = np.repeat(np.arange(m, dtype=np.int64), ell)      # local row of every draw
        r = rl + r0                                            # its global row
        starts = np.zeros(m + 1, dtype=np.int64)
        np.cumsum(ell, out=starts[1:])
        k = np.arange(tot, dtype=np.int64) - starts[rl]
        h = _splitmix64(base[rl] + (k.astype(np.uint64) + np.uint64(1)) * np.uint64(0x2545F4914F6CDD1D))
        h2 = _splitmix64(h)
        h3 = _splitmix64(h2)
        is_local = _u01(h) < local
        w = min(window, n - 1)
        near = r + (h2 % np.uint64(2 * w + 1)).astype(np.int64) - w
        near = np.where(near < 0, -near, near)
        near = np.where(near >= n, 2 * (n - 1) - near, near)
        far = (h2 % np.uint64(n)).astype(np.int64)
        c = np.where(is_local, near, far)
        c = np.where(c == r, (c + 1) % n, c)
        v = -_u01(h3)
    # merge duplicates, sort columns, add the diagonal
    key = rl * n + c
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    grp = np.cumsum(first) - 1
    vs = np.zeros(int(grp[-1]) + 1)
    np.add.at(vs, grp, v)
    ku = key[first]
    ru, cu = ku // n, ku % n
    dsum = np.zeros(m)
    np.add.at(dsum, ru, np.abs(vs))
    rr = np.concatenate([ru, np.arange(m, dtype=np.int64)])
    cc = np.concatenate([cu, np.arange(r0, r1, dtype=np.int64)])
    vv = np.concatenate([vs, 1.0 + dsum])
    order = np.argsort(rr * n + cc, kind="stable")
    rr, cc, vv = rr[order], cc[order], vv[order]
    Ap = np.zeros(m + 1, dtype=np.int64)
    np.add.at(Ap, rr + 1, 1)
    np.cumsum(Ap, out=Ap)
    return Ap.astype(np.int32), cc.astype(np.int32), vv.astype(np.float64)


def random_csr(n, avg=6, seed=0, sorted_cols=True, diag=True):
    """Small ragged test matrix (includes empty rows when diag=False)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, 2 * avg + 1, size=n)
    Ap = np.zeros(n + 1, dtype=np.int64)
    cols, vals = [], []
    for i in range(n):
        c = rng.choice(n, size=min(int(lens[i]), n), replace=False)
        if diag and i not in c:
            c = np.append(c, i)
        if sorted_cols:
            c = np.sort(c)
        v = rng.uniform(-1, 1, size=len(c))
        if diag:
            v[c == i] = 1.0 + np.abs(v).sum()
        cols.append(c)
        vals.append(v)
        Ap[i + 1] = Ap[i] + len(c)
    Aj = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    Ax = np.concatenate(vals) if vals else np.zeros(0)
    return Ap.astype(np.int32), Aj, Ax.astype(np.float64)

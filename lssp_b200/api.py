"""Python host mirror of the LSSP interface for the solve-loop hot path.

Thin object layer over the C ABI (include/lsspg.h).  Names follow the reference
(`lssp_mv_mxy`, `lssp_vec_dot`, `lssp_pc_ilu_solve`, `lssp_solver_solve` ...;
reference include/mvops.h, include/vector.h, include/solver-tri.h,
include/lssp.h) so parity tests read like calls into the reference.  All
arithmetic happens in the CUDA kernels behind the C ABI; nothing here computes.
"""
import ctypes as C

import numpy as np

from ._lib import AmgPars, SolveInfo, SolverOpts, check, lib

SOLVERS = {"gmres": 0, "lgmres": 1, "rgmres": 2, "rlgmres": 3, "bicgstab": 4, "bicgstabl": 5,
           "bicgsafe": 6, "cg": 7, "cgs": 8, "gpbicg": 9, "cr": 10, "crs": 11, "bicrstab": 12,
           "bicrsafe": 13, "gpbicr": 14, "qmrcgstab": 15, "tfqmr": 16, "orthomin": 17, "idrs": 18}
MV_MXY, MV_AMXY, MV_AMXPBY, MV_AMXPBYZ = 0, 1, 2, 3
OPT_SPMV_KERNEL, OPT_SPMV_EXACT, OPT_CHECK_EVERY, OPT_REDUCE_SEQUENTIAL, OPT_GRAPHS = 1, 2, 3, 4, 5


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One device + one stream (lsspg_ctx)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        check(lib().lsspg_ctx_create(int(device), C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().lsspg_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib().lsspg_sync(self.h))

    @property
    def launches(self):
        return lib().lsspg_ctx_launches(self.h)

    @property
    def stream(self):
        return lib().lsspg_ctx_stream(self.h)

    def set_option(self, opt, val):
        check(lib().lsspg_ctx_set_option(self.h, int(opt), int(val)))

    def empty(self, n):
        return DVec(self, n)

    def zeros(self, n):
        v = DVec(self, n)
        check(lib().lsspg_memset_zero(self.h, v.ptr, C.c_size_t(8 * n)))
        return v

    def upload(self, a):
        a = _f64(a)
        v = DVec(self, len(a))
        v.set(a)
        return v


class DVec:
    """Device vector of n doubles (the device image of an lssp_vec)."""

    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, int(n)
        self.ptr = C.c_void_p()
        check(lib().lsspg_malloc(ctx.h, C.c_size_t(8 * max(self.n, 1)), C.byref(self.ptr)))

    def set(self, a):
        a = _f64(a)
        assert len(a) == self.n
        check(lib().lsspg_h2d(self.ctx.h, self.ptr, _p(a), C.c_size_t(8 * self.n)))

    def get(self):
        out = np.empty(self.n)
        check(lib().lsspg_d2h(self.ctx.h, _p(out), self.ptr, C.c_size_t(8 * self.n)))
        return out

    def free(self):
        if self.ptr and self.ctx.h:      # a closed context has already released the device
            lib().lsspg_free(self.ctx.h, self.ptr)
        self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Csr:
    """Device-resident CSR matrix (the device image of an lssp_mat_csr)."""

    def __init__(self, ctx, A, num_cols=None):
        self.ctx = ctx
        self.h = C.c_void_p()
        if A[0] is None:  # the reference's zero matrix: Ap == NULL
            n = int(A[1])
            self.n, self.nnz = n, 0
            check(lib().lsspg_csr_upload(ctx.h, n, n, None, None, None, C.byref(self.h)))
            return
        Ap, Aj, Ax = _i32(A[0]), _i32(A[1]), _f64(A[2])
        self.n = len(Ap) - 1
        self.nnz = int(Ap[-1])
        check(lib().lsspg_csr_upload(ctx.h, self.n, self.n if num_cols is None else num_cols,
                                     _p(Ap), _p(Aj), _p(Ax), C.byref(self.h)))

    @property
    def spmv_bytes(self):
        return lib().lsspg_csr_spmv_bytes(self.h)

    def schedule_info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(lib().lsspg_csr_schedule_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(num_tiles=a.value, num_stream_tiles=b.value, max_tile_nnz=c.value)

    def mv(self, kind, x, z, alpha=1.0, beta=0.0, y=None):
        check(lib().lsspg_mv(self.ctx.h, kind, self.h, C.c_double(alpha), x.ptr, C.c_double(beta),
                             y.ptr if y is not None else None, z.ptr))

    def mv_host(self, kind, x, alpha=1.0, beta=0.0, y=None):
        x = _f64(x)
        y = None if y is None else _f64(y)
        z = np.empty(self.n)
        check(lib().lsspg_mv_host(self.ctx.h, kind, self.h, C.c_double(alpha), _p(x), C.c_double(beta),
                                  _p(y), _p(z)))
        return z

    def free(self):
        if self.h and self.ctx.h:
            lib().lsspg_csr_destroy(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DMat:
    """Device-resident CSR / BCSR matrix of the set-up path (lsspg_dmat; matops_gpu.cu): generated,
    sorted, repaired, restricted and factorised on the GPU without a host copy of Aj / Ax."""

    def __init__(self, ctx, A=None, handle=None):
        self.ctx = ctx
        self.h = C.c_void_p() if handle is None else handle
        if A is not None:
            Ap, Aj, Ax = _i32(A[0]), _i32(A[1]), _f64(A[2])
            n = len(Ap) - 1
            check(lib().lsspg_dmat_upload(ctx.h, n, n, _p(Ap), _p(Aj), _p(Ax), C.byref(self.h)))

    @classmethod
    def stencil(cls, ctx, dims, stencil, r0=0, r1=None, col_shift=0):
        """rows [r0, r1) of the 7-point operator on an nx x ny x nz grid (generators.stencil_7pt / laplacian_5pt)"""
        nx, ny, nz = dims
        r1 = nx * ny * nz if r1 is None else r1
        st = _f64(np.asarray(stencil, dtype=np.float64))
        h = C.c_void_p()
        check(lib().lsspg_dmat_gen_stencil(ctx.h, nx, ny, nz, C.c_longlong(r0), C.c_longlong(r1), C.c_longlong(col_shift),
                                           _p(st), C.byref(h)))
        return cls(ctx, handle=h)

    @classmethod
    def lap3d(cls, ctx, N):
        return cls.stencil(ctx, (N, N, N), [-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0])

    @classmethod
    def cd3d(cls, ctx, N, conv=(0.3, 0.2, 0.1)):
        cx, cy, cz = conv
        return cls.stencil(ctx, (N, N, N), [-1.0 - cz, -1.0 - cy, -1.0 - cx, 6.0, -1.0 + cx, -1.0 + cy, -1.0 + cz])

    @classmethod
    def laplacian_5pt(cls, ctx, N):
        return cls.stencil(ctx, (N, N, 1), [0.0, -1.0, -1.0, 4.0, -1.0, -1.0, 0.0])

    def dims(self):
        n, m, nnz, bs = C.c_int(), C.c_int(), C.c_longlong(), C.c_int()
        check(lib().lsspg_dmat_dims(self.h, C.byref(n), C.byref(m), C.byref(nnz), C.byref(bs)))
        return n.value, m.value, nnz.value, bs.value

    def download(self):
        n, m, nnz, bs = self.dims()
        Ap, Aj, Ax = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz * bs * bs)
        check(lib().lsspg_dmat_download(self.ctx.h, self.h, _p(Ap), _p(Aj), _p(Ax)))
        return Ap, Aj, Ax

    def _derive(self, fn, *args):
        h = C.c_void_p()
        check(fn(self.ctx.h, self.h, *args, C.byref(h)))
        return DMat(self.ctx, handle=h)

    def copy(self):
        return self._derive(lib().lsspg_dmat_copy)

    def is_sorted(self):
        r = C.c_int()
        check(lib().lsspg_dmat_is_sorted(self.ctx.h, self.h, C.byref(r)))
        return bool(r.value)

    def sort_columns(self):
        check(lib().lsspg_dmat_sort_columns(self.ctx.h, self.h))
        return self

    def adjust_zero_diag(self, tol=1e-10):
        return self._derive(lib().lsspg_dmat_adjust_zero_diag, C.c_double(tol))

    def get_block_diag(self, blk_size):
        return self._derive(lib().lsspg_dmat_get_block_diag, int(blk_size))

    def to_bcsr(self, blk_size):
        return self._derive(lib().lsspg_dmat_to_bcsr, int(blk_size))

    def to_csr(self, take=False):
        """the SpMV-ready matrix (Csr) of this device matrix; take=True hands the arrays over"""
        n, m, nnz, bs = self.dims()
        out = Csr.__new__(Csr)
        out.ctx, out.h, out.n, out.nnz = self.ctx, C.c_void_p(), n, nnz
        check(lib().lsspg_csr_from_dmat(self.ctx.h, self.h, 1 if take else 0, C.byref(out.h)))
        return out

    def ilu_factor(self, level=0, blk_size=0):
        """ILU(k) on the device (lsspg_ilu_factor_dmat): (L, U) in the reference's layout, on the host"""
        h = C.c_void_p()
        check(lib().lsspg_ilu_factor_dmat(self.ctx.h, self.h, int(level), int(blk_size), C.byref(h)))
        return _factors_out(h, self.dims()[0])

    def ilut_factor(self, p=-1, tol=1e-3, blk_size=0):
        """ILUT(p, tol) on the device (lsspg_ilut_factor_dmat)"""
        h = C.c_void_p()
        check(lib().lsspg_ilut_factor_dmat(self.ctx.h, self.h, int(p), C.c_double(tol), int(blk_size), C.byref(h)))
        return _factors_out(h, self.dims()[0])

    def free(self):
        if self.h and self.ctx.h:
            lib().lsspg_dmat_destroy(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _factors_out(h, n):
    nn, nl, nu = C.c_int(), C.c_int(), C.c_int()
    lib().lsspg_factors_sizes(h, C.byref(nn), C.byref(nl), C.byref(nu))
    Lp, Lj, Lx = np.empty(n + 1, np.int32), np.empty(nl.value, np.int32), np.empty(nl.value)
    Up, Uj, Ux = np.empty(n + 1, np.int32), np.empty(nu.value, np.int32), np.empty(nu.value)
    lib().lsspg_factors_get(h, _p(Lp), _p(Lj), _p(Lx), _p(Up), _p(Uj), _p(Ux))
    lib().lsspg_factors_destroy(h)
    return (Lp, Lj, Lx), (Up, Uj, Ux)


# ---- reference-named entry points (host vectors in, host vectors out) ------------
def lssp_mv_mxy(A, x):
    """y = A x   (reference src/mvops.cxx:118-150)"""
    return A.mv_host(MV_MXY, x)


def lssp_mv_amxy(a, A, x):
    """y = a A x   (reference src/mvops.cxx:81-115)"""
    return A.mv_host(MV_AMXY, x, alpha=a)


def lssp_mv_amxpby(alpha, A, x, beta, y):
    """y = beta y + alpha A x   (reference src/mvops.cxx:5-39); returns the new y"""
    return A.mv_host(MV_AMXPBY, x, alpha=alpha, beta=beta, y=y)


def lssp_mv_amxpbyz(alpha, A, x, beta, y):
    """z = beta y + alpha A x   (reference src/mvops.cxx:42-78)"""
    return A.mv_host(MV_AMXPBYZ, x, alpha=alpha, beta=beta, y=y)


def lssp_vec_dot(ctx, x, y):
    out = C.c_double()
    check(lib().lsspg_vec_dot(ctx.h, x.n, x.ptr, y.ptr, C.byref(out)))
    return out.value


def lssp_vec_norm(ctx, x):
    out = C.c_double()
    check(lib().lsspg_vec_norm(ctx.h, x.n, x.ptr, C.byref(out)))
    return out.value


def lssp_vec_multidot(ctx, xs, y):
    k = len(xs)
    arr = (C.c_void_p * k)(*[v.ptr.value for v in xs])
    out = (C.c_double * k)()
    check(lib().lsspg_vec_multidot(ctx.h, y.n, k, arr, y.ptr, out))
    return np.array(out[:])


def lssp_vec_set_value(ctx, x, val):
    check(lib().lsspg_vec_set(ctx.h, x.n, x.ptr, C.c_double(val)))


def lssp_vec_copy(ctx, dst, src):
    check(lib().lsspg_vec_copy(ctx.h, dst.n, dst.ptr, src.ptr))


def lssp_vec_axy(ctx, a, x, y):
    check(lib().lsspg_vec_axy(ctx.h, x.n, C.c_double(a), x.ptr, y.ptr))


def lssp_vec_axpby(ctx, a, x, b, y):
    check(lib().lsspg_vec_axpby(ctx.h, x.n, C.c_double(a), x.ptr, C.c_double(b), y.ptr))


def lssp_vec_axpbyz(ctx, a, x, b, y, z):
    check(lib().lsspg_vec_axpbyz(ctx.h, x.n, C.c_double(a), x.ptr, C.c_double(b), y.ptr, z.ptr))


def lssp_vec_scale(ctx, x, a):
    check(lib().lsspg_vec_scale(ctx.h, x.n, x.ptr, C.c_double(a)))


# ---- triangular factors / preconditioners ----------------------------------------------
def tri_levels(which, T):
    """Host-side dependency levels of a triangular factor (0 lower / 1 upper)."""
    Tp, Tj = _i32(T[0]), _i32(T[1])
    n = len(Tp) - 1
    lev = np.empty(n, np.int32)
    nl = C.c_int()
    check(lib().lsspg_tri_levels_host(which, n, _p(Tp), _p(Tj), _p(lev), C.byref(nl)))
    return lev, nl.value


def tri_walk_layout_host(which, T, rhs):
    """Layout self-check (tests only): walk the level-ordered sliced-ELL image on the host."""
    Tp, Tj, Tx = _i32(T[0]), _i32(T[1]), _f64(T[2])
    n = len(Tp) - 1
    x = np.zeros(n)
    ns, pad = C.c_int(), C.c_longlong()
    check(lib().lsspg_debug_tri_walk_layout_host(which, n, _p(Tp), _p(Tj), _p(Tx), _p(x), _p(_f64(rhs)),
                                                 C.byref(ns), C.byref(pad)))
    return x, ns.value, pad.value


def tri_walk_tiled_host(which, T, rhs):
    """Layout self-check (tests only) of the box schedule used for structured-grid factors.
    Returns (x, info) or (None, None) when the factor has no box schedule."""
    Tp, Tj, Tx = _i32(T[0]), _i32(T[1]), _f64(T[2])
    n = len(Tp) - 1
    x = np.zeros(n)
    ok, info = C.c_int(), (C.c_int * 8)()
    check(lib().lsspg_debug_tri_walk_tiled_host(which, n, _p(Tp), _p(Tj), _p(Tx), _p(x), _p(_f64(rhs)),
                                                C.byref(ok), info))
    if not ok.value:
        return None, None
    keys = ("boxes", "box_levels", "max_box_rows", "row_levels", "nx", "ny", "nz", "box_edge")
    return x, dict(zip(keys, list(info)))


def tri_pack_host(which, T):
    """Fingerprint of the device image `Tri` would upload for this factor, built without a GPU (tests and set-up
    profiling only).  Returns dict(kind, fingerprint, bytes, schedule_s, pack_s); kind 0 = slice schedule,
    1 = box blobs, 2 = ELL box blobs."""
    Tp, Tj, Tx = _i32(T[0]), _i32(T[1]), _f64(T[2])
    n = len(Tp) - 1
    kind, fp, nb, sec = C.c_int(), C.c_ulonglong(), C.c_longlong(), (C.c_double * 2)()
    check(lib().lsspg_debug_tri_pack_host(which, n, _p(Tp), _p(Tj), _p(Tx), C.byref(kind), C.byref(fp), C.byref(nb), sec))
    return dict(kind=kind.value, fingerprint=fp.value, bytes=nb.value, schedule_s=sec[0], pack_s=sec[1])


def tri_walk_packed_host(which, T, rhs):
    """Emulation of the ELL box kernels from the packed blobs (tests only).  Returns (x, dict(rounds, chunks, boxes,
    box_levels)) or (None, None) when the factor has no such schedule."""
    Tp, Tj, Tx = _i32(T[0]), _i32(T[1]), _f64(T[2])
    n = len(Tp) - 1
    x = np.zeros(n)
    ok, info = C.c_int(), (C.c_int * 4)()
    check(lib().lsspg_debug_tri_walk_packed_host(which, n, _p(Tp), _p(Tj), _p(Tx), _p(x), _p(_f64(rhs)), C.byref(ok), info))
    if not ok.value:
        return None, None
    return x, dict(zip(("rounds", "chunks", "boxes", "box_levels"), list(info)))


def tri_walk_pencil_host(which, T, rhs):
    """Replay of the pencil schedule (tri_pencil.cu) from the image the device reads (tests only).  Returns
    (x, info) or (None, None) when the factor is not a lattice factor."""
    Tp, Tj, Tx = _i32(T[0]), _i32(T[1]), _f64(T[2])
    n = len(Tp) - 1
    x = np.zeros(n)
    ok, info = C.c_int(), (C.c_int * 8)()
    check(lib().lsspg_debug_tri_walk_pencil_host(which, n, _p(Tp), _p(Tj), _p(Tx), _p(x), _p(_f64(rhs)), C.byref(ok), info))
    if not ok.value:
        return None, None
    keys = ("pencils", "threads", "max_ghost", "max_dk", "slots", "values_per_row", "max_steps", "skew")
    return x, dict(zip(keys, list(info)))


class Tri:
    """Device-resident triangular factor in level order (lsspg_tri)."""

    def __init__(self, ctx, which, T):
        self.ctx = ctx
        Tp, Tj, Tx = _i32(T[0]), _i32(T[1]), _f64(T[2])
        self.n = len(Tp) - 1
        self.h = C.c_void_p()
        check(lib().lsspg_tri_analyse(ctx.h, which, self.n, _p(Tp), _p(Tj), _p(Tx), C.byref(self.h)))

    def info(self):
        a, b, c = C.c_int(), C.c_int(), C.c_longlong()
        check(lib().lsspg_tri_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(num_levels=a.value, num_slices=b.value, padded_nnz=c.value)

    def schedule(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().lsspg_tri_schedule(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        # kind: 2 pencil schedule (boxes = pencils, box_levels = steps), 1 box schedule, 0 slice schedule
        return dict(tiled=bool(a.value), kind=a.value, boxes=b.value, box_levels=c.value, max_box_rows=d.value)

    def solve(self, x, rhs):
        check(lib().lsspg_tri_solve(self.ctx.h, self.h, x.ptr, rhs.ptr))

    def free(self):
        if self.h and self.ctx.h:
            lib().lsspg_tri_destroy(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ilu_factor(A, kind="iluk", level=0, p=-1, tol=1e-3, blk_size=0, ctx=None):
    """ILU(k) / ILUT set-up (reference src/pc-iluk.cxx, src/pc-ilut.cxx).  Returns (L, U) CSR
    triples in the reference's layout.  ctx given: the whole set-up -- ingest, symbolic and
    numeric phases (ILUT: the dual-threshold row recurrence), block restriction, L / U split -- runs on the GPU (ilu_gpu.cu),
    with bit-identical factors."""
    Ap, Aj, Ax = _i32(A[0]), _i32(A[1]), _f64(A[2])
    n = len(Ap) - 1
    h = C.c_void_p()
    if ctx is not None and kind == "iluk":
        check(lib().lsspg_ilu_factor_device(ctx.h, n, _p(Ap), _p(Aj), _p(Ax), int(level), int(blk_size), C.byref(h)))
    elif ctx is not None:
        check(lib().lsspg_ilut_factor_device(ctx.h, n, _p(Ap), _p(Aj), _p(Ax), int(p), C.c_double(tol), int(blk_size), C.byref(h)))
    else:
        check(lib().lsspg_ilu_factor(0 if kind == "iluk" else 1, n, _p(Ap), _p(Aj), _p(Ax), int(level), int(p),
                                     C.c_double(tol), int(blk_size), C.byref(h)))
    return _factors_out(h, n)


def bilu_factor(A, num_blks, level=0):
    """Block ILU(k) set-up (reference src/pc-biluk.cxx:377-431): blocks of n / num_blks rows.  Returns (L, D, U) CSR
    triples -- unit diagonal last in L and first in U, D the block diagonal of the inverted pivot blocks."""
    Ap, Aj, Ax = _i32(A[0]), _i32(A[1]), _f64(A[2])
    n = len(Ap) - 1
    h = C.c_void_p()
    check(lib().lsspg_bilu_factor(n, _p(Ap), _p(Aj), _p(Ax), int(num_blks), int(level), C.byref(h)))
    nn, nl, nd, nu = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().lsspg_bfactors_sizes(h, C.byref(nn), C.byref(nl), C.byref(nd), C.byref(nu))
    out = [(np.empty(n + 1, np.int32), np.empty(z.value, np.int32), np.empty(z.value)) for z in (nl, nd, nu)]
    (Lp, Lj, Lx), (Dp, Dj, Dx), (Up, Uj, Ux) = out
    lib().lsspg_bfactors_get(h, _p(Lp), _p(Lj), _p(Lx), _p(Dp), _p(Dj), _p(Dx), _p(Up), _p(Uj), _p(Ux))
    lib().lsspg_bfactors_destroy(h)
    return out


class Preconditioner:
    """Device-side preconditioner application (the `pc.solve` seam, reference
    include/type-defs.h:104,144)."""

    def __init__(self, ctx, kind, n, L=None, U=None, D=None):
        self.ctx, self.kind, self.n = ctx, kind, n
        self.h = C.c_void_p()
        if kind == "non":
            check(lib().lsspg_pc_create_non(ctx.h, n, C.byref(self.h)))
        elif kind == "ilu":
            L = (_i32(L[0]), _i32(L[1]), _f64(L[2]))
            U = (_i32(U[0]), _i32(U[1]), _f64(U[2]))
            check(lib().lsspg_pc_create_ilu(ctx.h, n, _p(L[0]), _p(L[1]), _p(L[2]), _p(U[0]), _p(U[1]),
                                            _p(U[2]), C.byref(self.h)))
        elif kind == "bilu":
            L = (_i32(L[0]), _i32(L[1]), _f64(L[2]))
            U = (_i32(U[0]), _i32(U[1]), _f64(U[2]))
            D = (_i32(D[0]), _i32(D[1]), _f64(D[2]))
            check(lib().lsspg_pc_create_bilu(ctx.h, n, _p(L[0]), _p(L[1]), _p(L[2]), _p(D[0]), _p(D[1]),
                                             _p(D[2]), _p(U[0]), _p(U[1]), _p(U[2]), C.byref(self.h)))
        else:
            raise ValueError(kind)

    @classmethod
    def non(cls, ctx, n):
        return cls(ctx, "non", n)

    @classmethod
    def iluk(cls, ctx, A, level=1, blk_size=0):
        """lssp_pc_iluk_assemble (reference src/pc-iluk.cxx:566-581); default level 1 (src/pc.cxx:3)"""
        L, U = ilu_factor(A, "iluk", level=level, blk_size=blk_size)
        return cls(ctx, "ilu", len(L[0]) - 1, L, U)

    @classmethod
    def biluk(cls, ctx, A, num_blks, level=1):
        """lssp_pc_biluk_assemble (reference src/pc-biluk.cxx:416-431); default level 1 (src/pc.cxx:3)"""
        L, D, U = bilu_factor(A, num_blks, level=level)
        return cls(ctx, "bilu", len(L[0]) - 1, L, U, D)

    @classmethod
    def ilut(cls, ctx, A, p=-1, tol=1e-3, blk_size=0):
        """lssp_pc_ilut_assemble (reference src/pc-ilut.cxx:429-456)"""
        L, U = ilu_factor(A, "ilut", p=p, tol=tol, blk_size=blk_size)
        return cls(ctx, "ilu", len(L[0]) - 1, L, U)

    @classmethod
    def sxamg(cls, ctx, A, hierarchy=None, share=None, **pars):
        """lssp_pc_sxamg_assemble (reference src/pc-sxamg.cxx:75-126): classical AMG hierarchy of A,
        one V-cycle per application.  `hierarchy`: an AmgHierarchy to reuse; `share`: a device Csr
        holding the same A (saves a second copy of the level-0 operator)."""
        H = hierarchy if hierarchy is not None else AmgHierarchy(A, **pars)
        self = cls.__new__(cls)
        self.ctx, self.kind, self.n = ctx, "amg", H.levels[0]["n"]
        self.h = C.c_void_p()
        self.hierarchy = H
        self._share = share   # keep the shared operator alive
        check(lib().lsspg_pc_create_amg(ctx.h, H.h, share.h if share is not None else None, C.byref(self.h)))
        return self

    def amg_solve(self, b, x, tol=1e-8, maxit=100):
        """lssp_solver_sxamg (reference src/solver-sxamg.cxx:25-99) with HOST b / x: cycles from x
        until ||b - A x|| / ||b|| <= tol; returns dict(nits, residual, x)."""
        b = _f64(b)
        assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous
        nits, ares = C.c_int(), C.c_double()
        check(lib().lsspg_amg_solve_host(self.ctx.h, self.h, _p(b), _p(x), C.c_double(tol), int(maxit),
                                         C.byref(nits), C.byref(ares)))
        return dict(nits=nits.value, residual=ares.value, x=x)

    @property
    def bytes(self):
        return lib().lsspg_pc_bytes(self.h)

    def info(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_longlong(), C.c_longlong()
        check(lib().lsspg_pc_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(levels_L=a.value, levels_U=b.value, padded_L=c.value, padded_U=d.value)

    def apply(self, x, rhs):
        """pc.solve(&pc, x, rhs) on device vectors"""
        check(lib().lsspg_pc_apply(self.ctx.h, self.h, x.ptr, rhs.ptr))

    def apply_host(self, rhs, x0=None):
        """pc.solve(&pc, x, rhs) on host vectors (x0: incoming contents of x)"""
        rhs = _f64(rhs)
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        check(lib().lsspg_pc_apply_host(self.ctx.h, self.h, _p(x), _p(rhs)))
        return x

    def free(self):
        if self.h and self.ctx.h:
            lib().lsspg_pc_destroy(self.ctx.h, self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class AmgHierarchy:
    """Host image of the SX-AMG-style hierarchy (lsspg_amg_setup_host; the role of sx_amg_setup,
    reference src/pc-sxamg.cxx:107-110).  `levels[l]` holds n, nc, A, P, R (CSR triples) and cf."""

    def __init__(self, A, ctx=None, replay=False, **pars):
        """ctx given: the per-row phases of the set-up (strong couplings, interpolation, restriction, Galerkin products) run
        on the GPU (lsspg_amg_setup_device, amg_gpu.cu), the C/F splitting on the host; replay=True: the same row functions
        on the CPU (test-suite).  The hierarchy is the same in every variant, array by array."""
        Ap, Aj, Ax = _i32(A[0]), _i32(A[1]), _f64(A[2])
        p = AmgPars()
        check(lib().lsspg_amg_pars_default(C.byref(p)))
        for k, v in pars.items():
            if not hasattr(p, k):
                raise TypeError("unknown AMG parameter %r" % k)
            setattr(p, k, v)
        self.pars = p
        self.h = C.c_void_p()
        if ctx is not None:
            check(lib().lsspg_amg_setup_device(ctx.h, len(Ap) - 1, _p(Ap), _p(Aj), _p(Ax), C.byref(p), C.byref(self.h)))
        elif replay:
            check(lib().lsspg_debug_amg_setup_replay_host(len(Ap) - 1, _p(Ap), _p(Aj), _p(Ax), C.byref(p), C.byref(self.h)))
        else:
            check(lib().lsspg_amg_setup_host(len(Ap) - 1, _p(Ap), _p(Aj), _p(Ax), C.byref(p), C.byref(self.h)))
        nl, dense = C.c_int(), C.c_int()
        check(lib().lsspg_amg_host_levels(self.h, C.byref(nl), C.byref(dense)))
        self.coarse_dense = bool(dense.value)
        self.levels = []
        for l in range(nl.value):
            n, nc, za, zp, zr = (C.c_int() for _ in range(5))
            check(lib().lsspg_amg_host_level_sizes(self.h, l, C.byref(n), C.byref(nc), C.byref(za), C.byref(zp),
                                                   C.byref(zr)))
            n, nc, za, zp, zr = n.value, nc.value, za.value, zp.value, zr.value
            last = l == nl.value - 1
            a = (np.zeros(n + 1, np.int32), np.zeros(za, np.int32), np.zeros(za))
            P = (np.zeros(n + 1, np.int32), np.zeros(zp, np.int32), np.zeros(zp))
            R = (np.zeros(nc + 1, np.int32), np.zeros(zr, np.int32), np.zeros(zr))
            cf = np.zeros(n, np.int32)
            check(lib().lsspg_amg_host_level_get(self.h, l, _p(a[0]), _p(a[1]), _p(a[2]),
                                                 None if last else _p(P[0]), None if last else _p(P[1]),
                                                 None if last else _p(P[2]), None if last else _p(R[0]),
                                                 None if last else _p(R[1]), None if last else _p(R[2]), _p(cf)))
            rank = np.zeros(n, np.int32)
            check(lib().lsspg_amg_host_level_rank(self.h, l, _p(rank)))
            self.levels.append(dict(n=n, nc=nc, A=a, P=None if last else P, R=None if last else R, cf=cf, rank=rank))
        self.coarse_inv = None
        if self.coarse_dense:
            nlast = self.levels[-1]["n"]
            self.coarse_inv = np.zeros((nlast, nlast))
            check(lib().lsspg_amg_host_coarse_inverse(self.h, _p(self.coarse_inv)))

    def walk_gs_host(self, l, post, b, x_old, mode=0):
        """CPU self-check of the smoother schedule of level l (not a compute path): one sweep.
        mode 0: the schedule the device would pick, 1: 32-row slices, 2: one ticket per row."""
        post = int(post) | (int(mode) << 1)
        b, x_old = _f64(b), _f64(x_old)
        x_new = np.full(len(b), np.nan)
        info = np.zeros(4, np.int32)
        check(lib().lsspg_debug_amg_walk_gs_host(self.h, int(l), int(post), _p(b), _p(x_old), _p(x_new), _p(info)))
        return x_new, dict(slices=int(info[0]), levels_c=int(info[1]), levels_f=int(info[2]), width=int(info[3]))

    def free(self):
        if self.h:
            lib().lsspg_amg_host_destroy(self.h)
        self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def solver_supported(name):
    return bool(lib().lsspg_solver_supported(SOLVERS[name]))


def _opts(hist, **kw):
    o = SolverOpts()
    check(lib().lsspg_solver_opts_default(C.byref(o)))
    names = dict(rtol="tol_rel", atol="tol_abs", rbtol="tol_rb", maxit="maxit", restart="restart",
                 augk="aug_k", bgsl="bgsl", idrs="idrs", verb="verb")
    for k, v in kw.items():
        setattr(o, names[k], v)
    if hist is not None:
        o.hist_len = len(hist)
        o.hist = hist.ctypes.data_as(C.POINTER(C.c_double))
    return o


def lssp_solver_solve(ctx, solver, A, pc, b, x, nhist=0, **kw):
    """lssp_solver_solve (reference src/lssp.cxx:250-414) with HOST b / x, as a
    caller of the reference sees it: x is the initial guess on entry and is
    overwritten with the solution.  Returns dict(nits, residual, hist, ...)."""
    b = _f64(b)
    assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous
    hist = np.zeros(nhist) if nhist else None
    o = _opts(hist, **kw)
    info = SolveInfo()
    check(lib().lsspg_krylov_solve_host(ctx.h, SOLVERS[solver], A.h, pc.h, _p(b), _p(x), C.byref(o),
                                        C.byref(info)))
    return dict(nits=info.nits, residual=info.residual, x=x, solve_ms=info.solve_ms,
                launches=info.launches, breakdown=info.breakdown,
                hist=None if hist is None else hist[:info.hist_used])


def solve_device(ctx, solver, A, pc, b, x, nhist=0, **kw):
    """Same solve with device-resident b / x (DVec)."""
    hist = np.zeros(nhist) if nhist else None
    o = _opts(hist, **kw)
    info = SolveInfo()
    check(lib().lsspg_krylov_solve(ctx.h, SOLVERS[solver], A.h, pc.h, b.ptr, x.ptr, C.byref(o), C.byref(info)))
    return dict(nits=info.nits, residual=info.residual, solve_ms=info.solve_ms, launches=info.launches,
                breakdown=info.breakdown, hist=None if hist is None else hist[:info.hist_used])

"""TEST INFRASTRUCTURE (oracle) -- ctypes front-ends for the two CPU checkers.

* :class:`Port` -- ``oracle/liboracle.so``, the plain-C restatement in
  ``oracle/oracle.c`` (built by ``make -C oracle port``).
* :class:`Ref`  -- ``oracle/_ref/liblssp_ref.so`` (and ``liblssp_refz.so``, the
  zero-initialising-malloc variant), the UNMODIFIED reference sources compiled
  where they lie by ``make -C oracle ref`` plus the shim ``ref_shim.cxx``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  The product package
``lssp_b200`` never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")

# reference enum values with every USE_* = 0 (include/type-defs.h:156-174, :64-98)
SOLVERS = {"gmres": 0, "lgmres": 1, "rgmres": 2, "rlgmres": 3, "bicgstab": 4, "bicgstabl": 5,
           "bicgsafe": 6, "cg": 7, "cgs": 8, "gpbicg": 9, "cr": 10, "crs": 11, "bicrstab": 12,
           "bicrsafe": 13, "gpbicr": 14, "qmrcgstab": 15, "tfqmr": 16, "orthomin": 17, "idrs": 18}
PCS = {"non": 0, "iluk": 1, "ilut": 2}


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build(port=True, ref=True):
    """(Re)build the checkers.  Building the checker is not using it."""
    targets = (["port"] if port else []) + (["ref"] if ref else [])
    if targets:
        subprocess.run(["make", "-C", HERE, "-j8"] + targets, check=True,
                       stdout=subprocess.DEVNULL)


class RefParams(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("rbtol", C.c_double),
                ("maxit", C.c_int), ("restart", C.c_int), ("augk", C.c_int), ("bgsl", C.c_int),
                ("idrs", C.c_int), ("iluk_level", C.c_int), ("ilut_p", C.c_int),
                ("ilut_tol", C.c_double), ("blk_size", C.c_int), ("verb", C.c_int)]


def ref_params(**kw):
    p = RefParams(rtol=1e-7, atol=1e-7, rbtol=1e-7, maxit=1000, restart=50, augk=3, bgsl=4,
                  idrs=4, iluk_level=1, ilut_p=-1, ilut_tol=1e-3, blk_size=0, verb=0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class Ref:
    """The unmodified reference, through oracle/ref_shim.cxx."""

    def __init__(self, zero_malloc=False):
        name = "liblssp_refz.so" if zero_malloc else "liblssp_ref.so"
        self.path = os.path.join(HERE, "_ref", name)
        L = self.lib = C.CDLL(self.path)
        L.ref_dot.restype = C.c_double
        L.ref_norm.restype = C.c_double
        L.ref_ilu_create.restype = C.c_void_p
        L.ref_solve.restype = C.c_int

    @staticmethod
    def available(zero_malloc=False):
        name = "liblssp_refz.so" if zero_malloc else "liblssp_ref.so"
        return os.path.exists(os.path.join(HERE, "_ref", name))

    # kernels -----------------------------------------------------------------
    def mv(self, kind, A, x, alpha=1.0, beta=0.0, y=None):
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        if kind == 2:
            out = np.array(y, dtype=np.float64)
            self.lib.ref_mv(2, n, _ptr(Ap), _ptr(Aj), _ptr(Ax), C.c_double(alpha), _ptr(x),
                            C.c_double(beta), _ptr(out), _ptr(out))
            return out
        out = np.empty(n)
        yy = y if y is not None else out
        self.lib.ref_mv(kind, n, _ptr(Ap), _ptr(Aj), _ptr(Ax), C.c_double(alpha), _ptr(x),
                        C.c_double(beta), _ptr(yy), _ptr(out))
        return out

    def dot(self, x, y):
        return self.lib.ref_dot(len(x), _ptr(x), _ptr(y))

    def norm(self, x):
        return self.lib.ref_norm(len(x), _ptr(x))

    def axpby(self, a, x, b, y):
        out = y.copy()
        self.lib.ref_axpby(len(x), C.c_double(a), _ptr(x), C.c_double(b), _ptr(out))
        return out

    def axpbyz(self, a, x, b, y):
        out = np.empty_like(x)
        self.lib.ref_axpbyz(len(x), C.c_double(a), _ptr(x), C.c_double(b), _ptr(y), _ptr(out))
        return out

    def tri_lower(self, L, rhs):
        Lp, Lj, Lx = L
        x = np.zeros(len(rhs))
        self.lib.ref_tri_lower(len(rhs), _ptr(Lp), _ptr(Lj), _ptr(Lx), _ptr(x), _ptr(rhs))
        return x

    def tri_upper(self, U, rhs):
        Up, Uj, Ux = U
        x = np.zeros(len(rhs))
        self.lib.ref_tri_upper(len(rhs), _ptr(Up), _ptr(Uj), _ptr(Ux), _ptr(x), _ptr(rhs))
        return x

    # factorisations ------------------------------------------------------------
    def ilu(self, A, kind="iluk", level=0, p=-1, tol=1e-3, blk_size=0):
        """Returns (L, U) as CSR triples, from the reference's own factorisation."""
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        h = self.lib.ref_ilu_create(0 if kind == "iluk" else 1, n, _ptr(Ap), _ptr(Aj), _ptr(Ax),
                                    level, p, C.c_double(tol), blk_size)
        h = C.c_void_p(h)
        nl, nu = C.c_int(), C.c_int()
        self.lib.ref_ilu_sizes(h, C.byref(nl), C.byref(nu))
        Lp, Lj, Lx = np.empty(n + 1, np.int32), np.empty(nl.value, np.int32), np.empty(nl.value)
        Up, Uj, Ux = np.empty(n + 1, np.int32), np.empty(nu.value, np.int32), np.empty(nu.value)
        self.lib.ref_ilu_get(h, _ptr(Lp), _ptr(Lj), _ptr(Lx), _ptr(Up), _ptr(Uj), _ptr(Ux))
        self.lib.ref_ilu_destroy(h)
        return (Lp, Lj, Lx), (Up, Uj, Ux)

    # whole solves ----------------------------------------------------------------
    def solve(self, solver, pc, A, b, x0=None, **kw):
        """Returns dict(nits, residual, x, t_assemble, t_solve)."""
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
        prm = ref_params(**kw)
        out = np.zeros(3)
        nits = self.lib.ref_solve(SOLVERS[solver], PCS[pc], n, _ptr(Ap), _ptr(Aj), _ptr(Ax),
                                  _ptr(np.ascontiguousarray(b)), _ptr(x), C.byref(prm), _ptr(out))
        return dict(nits=nits, residual=out[0], x=x, t_assemble=out[1], t_solve=out[2])

    def history(self, solver, pc, A, b, k=20, **kw):
        """Full-precision residual history by re-solving with maxit = 1..k
        (method (1)).  Stops once the solver converges early."""
        hist = []
        for m in range(1, k + 1):
            kw["maxit"] = m
            r = self.solve(solver, pc, A, b, **kw)
            if r["nits"] < m:
                break
            hist.append(r["residual"])
        return np.array(hist)


class RefB:
    """The unmodified reference built with USE_BLAS = USE_LAPACK = 1 against oracle/blas_standin.c: its block
    ILU(k) set-up (src/pc-biluk.cxx), through oracle/ref_bilu_shim.cxx."""

    def __init__(self):
        self.path = os.path.join(HERE, "_ref", "liblssp_refb.so")
        self.lib = C.CDLL(self.path)
        self.lib.refb_bilu_create.restype = C.c_void_p

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "liblssp_refb.so"))

    def bilu(self, A, num_blks, level=0, rhs=None):
        """(L, D, U) CSR triples of the reference's BILUK(level) with block size n / num_blks; with rhs also
        x = U^-1 D L^-1 rhs from the reference's own apply"""
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        h = C.c_void_p(self.lib.refb_bilu_create(n, _ptr(Ap), _ptr(Aj), _ptr(Ax), int(num_blks), int(level)))
        nl, nd, nu = C.c_int(), C.c_int(), C.c_int()
        self.lib.refb_bilu_sizes(h, C.byref(nl), C.byref(nd), C.byref(nu))
        out = []
        for nz in (nl.value, nd.value, nu.value):
            out.append((np.empty(n + 1, np.int32), np.empty(nz, np.int32), np.empty(nz)))
        (Lp, Lj, Lx), (Dp, Dj, Dx), (Up, Uj, Ux) = out
        self.lib.refb_bilu_get(h, _ptr(Lp), _ptr(Lj), _ptr(Lx), _ptr(Dp), _ptr(Dj), _ptr(Dx), _ptr(Up), _ptr(Uj), _ptr(Ux))
        x = None
        if rhs is not None:
            x = np.zeros(n)
            r = np.ascontiguousarray(rhs, dtype=np.float64)
            self.lib.refb_bilu_apply(h, n, _ptr(x), _ptr(r))
        self.lib.refb_bilu_destroy(h)
        return out if rhs is None else (out, x)

    def solve_biluk(self, solver, A, b, num_blks, level=1, rtol=1e-7, maxit=1000, restart=50, x0=None):
        """lssp_solver_create/assemble/solve with LSSP_PC_BILUK (s.num_blks = num_blks): dict(nits, residual, x)"""
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
        out = np.zeros(1)
        nits = self.lib.refb_solve_biluk(SOLVERS[solver], n, _ptr(Ap), _ptr(Aj), _ptr(Ax),
                                         _ptr(np.ascontiguousarray(b, dtype=np.float64)), _ptr(x), int(num_blks),
                                         int(level), C.c_double(rtol), int(maxit), int(restart), _ptr(out))
        return dict(nits=nits, residual=out[0], x=x)


class OrcPC(C.Structure):
    _fields_ = [("kind", C.c_int), ("Lp", C.c_void_p), ("Lj", C.c_void_p), ("Lx", C.c_void_p),
                ("Up", C.c_void_p), ("Uj", C.c_void_p), ("Ux", C.c_void_p), ("cache", C.c_void_p),
                ("amg", C.c_void_p)]


class OrcOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("rbtol", C.c_double),
                ("maxit", C.c_int)]


class OrcAmg:
    """amg_oracle.c over a hierarchy given as plain arrays (parity with libsxamg UNPINNED)."""

    def __init__(self, lib, levels, pre, post, cf_order, coarse_inv, coarse_sweeps, zero_guess=0):
        self.lib, self.levels = lib, levels
        self.n = levels[0]["n"]
        self.inv = None if coarse_inv is None else np.ascontiguousarray(coarse_inv, dtype=np.float64)
        lib.orc_amg_create.restype = C.c_void_p
        lib.orc_amg_solve.restype = C.c_int
        self.h = C.c_void_p(lib.orc_amg_create(len(levels), int(pre), int(post), int(cf_order),
                                               int(self.inv is not None), int(coarse_sweeps), int(zero_guess),
                                               _ptr(self.inv)))
        self.keep = []
        for l, L in enumerate(levels):
            arrs = [np.ascontiguousarray(a) for a in L["A"]]
            for key in ("P", "R"):
                arrs += [None] * 3 if L[key] is None else [np.ascontiguousarray(a) for a in L[key]]
            arrs.append(np.ascontiguousarray(L["cf"], dtype=np.int32))
            arrs.append(None if L.get("rank") is None else np.ascontiguousarray(L["rank"], dtype=np.int32))
            self.keep.append(arrs)
            lib.orc_amg_set_level(self.h, l, int(L["n"]), int(L["nc"]), *[_ptr(a) for a in arrs])

    def cycle(self, rhs, x0=None):
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        self.lib.orc_amg_cycle(self.h, _ptr(x), _ptr(np.ascontiguousarray(rhs, dtype=np.float64)))
        return x

    def solve(self, b, x0=None, tol=1e-8, maxit=100):
        x = np.zeros(self.n) if x0 is None else np.array(x0, dtype=np.float64)
        ares = C.c_double()
        nits = self.lib.orc_amg_solve(self.h, _ptr(np.ascontiguousarray(b, dtype=np.float64)), _ptr(x),
                                      C.c_double(tol), int(maxit), C.byref(ares))
        return dict(nits=nits, residual=ares.value, x=x)

    def __del__(self):
        try:
            self.lib.orc_amg_destroy(self.h)
        except Exception:
            pass


class Port:
    """oracle/oracle.c -- the plain-C restatement."""

    def __init__(self):
        self.path = os.path.join(HERE, "liboracle.so")
        build(port=True, ref=False)   # make: a no-op when liboracle.so is up to date
        L = self.lib = C.CDLL(self.path)
        L.orc_dot.restype = C.c_double
        L.orc_norm.restype = C.c_double
        for f in ("orc_cg", "orc_bicgstab"):
            getattr(L, f).restype = C.c_int

    def mv(self, kind, A, x, alpha=1.0, beta=0.0, y=None):
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        out = np.empty(n)
        self.lib.orc_mv(kind, n, _ptr(Ap), _ptr(Aj), _ptr(Ax), C.c_double(alpha), _ptr(x),
                        C.c_double(beta), _ptr(y), _ptr(out))
        return out

    def dot(self, x, y):
        return self.lib.orc_dot(len(x), _ptr(x), _ptr(y))

    def norm(self, x):
        return self.lib.orc_norm(len(x), _ptr(x))

    def axpby(self, a, x, b, y):
        out = y.copy()
        self.lib.orc_axpby(len(x), C.c_double(a), _ptr(x), C.c_double(b), _ptr(out))
        return out

    def axpbyz(self, a, x, b, y):
        out = np.empty_like(x)
        self.lib.orc_axpbyz(len(x), C.c_double(a), _ptr(x), C.c_double(b), _ptr(y), _ptr(out))
        return out

    def tri_lower(self, L, rhs):
        x = np.zeros(len(rhs))
        self.lib.orc_tri_lower(len(rhs), _ptr(L[0]), _ptr(L[1]), _ptr(L[2]), _ptr(x), _ptr(rhs))
        return x

    def tri_upper(self, U, rhs):
        x = np.zeros(len(rhs))
        self.lib.orc_tri_upper(len(rhs), _ptr(U[0]), _ptr(U[1]), _ptr(U[2]), _ptr(x), _ptr(rhs))
        return x

    def ilu_apply(self, L, U, rhs):
        n = len(rhs)
        x, cache = np.zeros(n), np.zeros(n)
        self.lib.orc_ilu_apply(n, _ptr(L[0]), _ptr(L[1]), _ptr(L[2]), _ptr(U[0]), _ptr(U[1]),
                               _ptr(U[2]), _ptr(x), _ptr(rhs), _ptr(cache))
        return x

    def bilu_apply(self, L, D, U, rhs):
        n = len(rhs)
        x, cache = np.zeros(n), np.zeros(2 * n)
        self.lib.orc_bilu_apply(n, _ptr(L[0]), _ptr(L[1]), _ptr(L[2]), _ptr(D[0]), _ptr(D[1]),
                                _ptr(D[2]), _ptr(U[0]), _ptr(U[1]), _ptr(U[2]), _ptr(x),
                                _ptr(rhs), _ptr(cache))
        return x

    def gs_sweep(self, A, cf, post, b, x, rank=None):
        """one in-place Gauss-Seidel sweep (amg_oracle.c); cf=None: natural order; rank: visiting
        order inside a block (None: by index)"""
        x = np.array(x, dtype=np.float64)
        rank = None if rank is None else np.ascontiguousarray(rank, dtype=np.int32)
        self.lib.orc_gs_sweep_ranked(len(b), _ptr(A[0]), _ptr(A[1]), _ptr(A[2]), _ptr(cf), _ptr(rank), int(post),
                                     _ptr(b), _ptr(x))
        return x

    def amg(self, levels, pre=2, post=2, cf_order=1, coarse_inv=None, coarse_sweeps=40, zero_guess=0):
        """restated cycle over a given hierarchy: levels = [dict(n, nc, A, P, R, cf)], see OrcAmg"""
        return OrcAmg(self.lib, levels, pre, post, cf_order, coarse_inv, coarse_sweeps, zero_guess)

    def solve(self, solver, A, b, x0=None, LU=None, rtol=1e-7, atol=1e-7, rbtol=1e-7, maxit=1000,
              nhist=0, amg=None):
        Ap, Aj, Ax = A
        n = len(Ap) - 1
        x = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64)
        pc = OrcPC(kind=0)
        keep = []
        if amg is not None:
            pc = OrcPC(kind=2, amg=amg.h.value)
        if LU is not None:
            (Lp, Lj, Lx), (Up, Uj, Ux) = LU
            cache = np.zeros(n)
            keep = [Lp, Lj, Lx, Up, Uj, Ux, cache]
            pc = OrcPC(1, *[a.ctypes.data for a in keep])
        o = OrcOpts(rtol, atol, rbtol, maxit)
        res = C.c_double()
        hist = np.zeros(max(nhist, 1))
        fn = getattr(self.lib, "orc_" + solver)
        nits = fn(n, _ptr(Ap), _ptr(Aj), _ptr(Ax), _ptr(np.ascontiguousarray(b)), _ptr(x),
                  C.byref(pc), C.byref(o), C.byref(res), _ptr(hist), nhist)
        return dict(nits=nits, residual=res.value, x=x, hist=hist[:min(nhist, nits)])

"""ctypes binding of liblsspg.so (the C ABI declared in include/lsspg.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C lssp_b200/csrc``.  There is no fallback of any kind: if the library
is missing, or no CUDA device is present when a context is created, an
exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblsspg.so")

_lib = None


class LsspgError(RuntimeError):
    pass


class SolverOpts(C.Structure):
    _fields_ = [("tol_rel", C.c_double), ("tol_abs", C.c_double), ("tol_rb", C.c_double),
                ("maxit", C.c_int), ("restart", C.c_int), ("aug_k", C.c_int), ("bgsl", C.c_int),
                ("idrs", C.c_int), ("verb", C.c_int), ("hist_len", C.c_int),
                ("hist", C.POINTER(C.c_double))]


class SolveInfo(C.Structure):
    _fields_ = [("nits", C.c_int), ("residual", C.c_double), ("hist_used", C.c_int),
                ("solve_ms", C.c_double), ("launches", C.c_longlong), ("breakdown", C.c_int)]


class AmgPars(C.Structure):
    _fields_ = [("max_levels", C.c_int), ("coarse_dof", C.c_int), ("strong_threshold", C.c_double),
                ("max_row_sum", C.c_double), ("trunc_threshold", C.c_double), ("pre_iter", C.c_int),
                ("post_iter", C.c_int), ("cf_order", C.c_int), ("zero_guess", C.c_int), ("coarse_dense_max", C.c_int),
                ("coarse_sweeps", C.c_int), ("tol", C.c_double), ("maxit", C.c_int), ("verb", C.c_int)]


def lib():
    """Load liblsspg.so (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LsspgError(
            "liblsspg.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C lssp_b200/csrc`; there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.lsspg_last_error.restype = C.c_char_p
    L.lsspg_version.restype = C.c_char_p
    L.lsspg_ctx_stream.restype = C.c_void_p
    L.lsspg_ctx_launches.restype = C.c_longlong
    L.lsspg_csr_spmv_bytes.restype = C.c_double
    L.lsspg_pc_bytes.restype = C.c_double
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LsspgError(lib().lsspg_last_error().decode())

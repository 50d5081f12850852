"""GPU suite, part 1: kernels through the C ABI versus the CPU oracle.

Bars (BASELINE.json north_star): SpMV bit-exact on the row-sequential path,
<= 1e-14 relative otherwise; element-wise BLAS-1 bit-exact; reductions <= 1e-14
relative (fixed-order tree versus the reference's sequential sum); triangular
sweeps / preconditioner application bit-exact (<= 1e-12 demanded).
"""
import numpy as np
import pytest

from lssp_b200 import api
from lssp_b200 import generators as g
from util import MATRICES, matrix, relerr, sha, tvec

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- SpMV ------
@pytest.mark.parametrize("name", list(MATRICES))
def test_spmv_variants_match_golden_bit_exact(ctx, golden, name):
    A = matrix(name)
    n = len(A[0]) - 1
    dA = api.Csr(ctx, A)
    e = golden["kernels"][name]
    x, y = tvec(n), tvec(n, 1)
    info = dA.schedule_info()
    exact = info["num_stream_tiles"] == info["num_tiles"]   # every row on the row-sequential path
    outs = {"mxy": api.lssp_mv_mxy(dA, x), "amxy": api.lssp_mv_amxy(-1.75, dA, x),
            "amxpby": api.lssp_mv_amxpby(0.5, dA, x, -2.0, y), "amxpbyz": api.lssp_mv_amxpbyz(-1.0, dA, x, 1.0, y)}
    if exact:
        for k, v in outs.items():
            assert sha(v) == e[k], k


@pytest.mark.parametrize("name", list(MATRICES))
def test_spmv_matches_checker(ctx, checker, name):
    A = matrix(name)
    n = len(A[0]) - 1
    dA = api.Csr(ctx, A)
    x, y = tvec(n, 2), tvec(n, 3)
    info = dA.schedule_info()
    exact = info["num_stream_tiles"] == info["num_tiles"]
    for kind, kw in ((0, {}), (1, dict(alpha=3.5)), (2, dict(alpha=-0.5, beta=0.25, y=y)), (3, dict(alpha=-1.0, beta=1.0, y=y))):
        got = dA.mv_host(kind, x, **kw)
        want = checker.mv(kind, A, x, **kw)
        if exact:
            assert np.array_equal(got, want), (name, kind)
        else:
            assert relerr(got, want) <= 1e-14, (name, kind)


def test_spmv_exact_mode_is_bit_exact_on_long_rows(checker):
    # power-law rows (up to thousands of entries) through the opt-in exact schedule
    c = api.Context(0)
    c.set_option(api.OPT_SPMV_EXACT, 1)
    A = matrix("powerlaw_4000")
    dA = api.Csr(c, A)
    x = tvec(len(A[0]) - 1, 5)
    assert np.array_equal(dA.mv_host(0, x), checker.mv(0, A, x))
    c.close()


def test_spmv_bulk_copy_pipeline_kernel_is_bit_exact(checker, golden):
    """LSSPG_OPT_SPMV_KERNEL = 2: tiles staged by cp.async.bulk into two shared-memory stages; same
    row-sequential arithmetic, so the same bits -- all four variants, fused dots included (CG history)"""
    c = api.Context(0)
    c.set_option(api.OPT_SPMV_KERNEL, 2)
    for name in ("lap2d_100", "lap3d_32", "cd3d_32", "cd3d_12"):
        A = matrix(name)
        n = len(A[0]) - 1
        dA = api.Csr(c, A)
        x, y = tvec(n, 2), tvec(n, 3)
        for kind, kw in ((0, {}), (1, dict(alpha=3.5)), (2, dict(alpha=-0.5, beta=0.25, y=y)), (3, dict(alpha=-1.0, beta=1.0, y=y))):
            assert np.array_equal(dA.mv_host(kind, x, **kw), checker.mv(kind, A, x, **kw)), (name, kind)
        dA.free()
    for n, avg, seed in ((1, 1, 0), (33, 2, 2), (257, 3, 3), (1025, 5, 4)):   # tiny / ragged tiles, empty rows
        A = g.random_csr(n, avg, seed=seed, diag=False)
        if A[0][-1] == 0:
            continue
        dA = api.Csr(c, A)
        assert np.array_equal(dA.mv_host(0, tvec(n)), checker.mv(0, A, tvec(n)))
        dA.free()
    # a solve through the fused SpMV + dot path: same history as the default kernel
    A = matrix("lap3d_32")
    n = len(A[0]) - 1
    e = golden["solves"].get("lap3d_32/cg/iluk0")
    dA, pc = api.Csr(c, A), api.Preconditioner.iluk(c, A, level=0)
    xs = np.zeros(n)
    r = api.lssp_solver_solve(c, "cg", dA, pc, np.ones(n), xs, nhist=20)
    c2 = api.Context(0)
    dA2, pc2 = api.Csr(c2, A), api.Preconditioner.iluk(c2, A, level=0)
    xs2 = np.zeros(n)
    r2 = api.lssp_solver_solve(c2, "cg", dA2, pc2, np.ones(n), xs2, nhist=20)
    # (the two kernels run different grids, so the tree of the fused dot products differs in the last bits)
    assert abs(r["nits"] - r2["nits"]) <= 1 and np.allclose(r["hist"][:15], r2["hist"][:15], rtol=1e-10)
    if e:
        assert abs(r["nits"] - e["nits"]) <= 1
    for o in (pc, dA, pc2, dA2):
        o.free()
    c.close()
    c2.close()


@pytest.mark.parametrize("name,level,blk", [("lap2d_100", 0, 0), ("lap3d_32", 0, 0), ("cd3d_32", 1, 0), ("cd3d_12", 2, 0),
                                            ("random_600", 1, 0), ("lap3d_32", 0, 8192), ("cd3d_32", 1, 4096)])
def test_gpu_ilu_numeric_factorisation_is_bit_identical(ctx, name, level, blk):
    """the IKJ numeric phase level by level on the GPU gives the host's -- and so
    the reference's -- L and U bit for bit, pivot repair and block-Jacobi variants included"""
    A = matrix(name)
    Lh, Uh = api.ilu_factor(A, "iluk", level=level, blk_size=blk)
    Lg, Ug = api.ilu_factor(A, "iluk", level=level, blk_size=blk, ctx=ctx)
    for h, d in zip(Lh + Uh, Lg + Ug):
        assert np.array_equal(h, d)


def test_gpu_ilu_repairs_small_pivots_like_the_host(ctx):
    # rows without a diagonal, tiny and negative-tiny pivots (src/pc.cxx:6-7, src/matrix-utils.cxx:483-587)
    A = g.random_csr(400, 4, seed=11, diag=False)
    Ap, Aj, Ax = A[0], A[1], A[2].copy()
    Ax[::7] *= 1e-13
    A = (Ap, Aj, Ax)
    Lh, Uh = api.ilu_factor(A, "iluk", level=1)
    Lg, Ug = api.ilu_factor(A, "iluk", level=1, ctx=ctx)
    for h, d in zip(Lh + Uh, Lg + Ug):
        assert np.array_equal(h, d, equal_nan=True)


def test_spmv_long_rows_use_warp_path_within_1e14(ctx, checker):
    rng = np.random.default_rng(5)
    n = 3000
    # rows of 1 .. 5000 entries, including rows longer than one tile
    lens = np.concatenate([rng.integers(1, 8, n - 6), [70, 300, 2049, 5000, 65, 64]])
    rng.shuffle(lens)
    Ap = np.zeros(n + 1, np.int64)
    Ap[1:] = np.cumsum(np.minimum(lens, n))
    Aj = np.concatenate([np.sort(rng.choice(n, min(int(l), n), replace=False)) for l in lens]).astype(np.int32)
    Ax = rng.uniform(-1, 1, len(Aj))
    A = (Ap.astype(np.int32), Aj, Ax)
    dA = api.Csr(ctx, A)
    info = dA.schedule_info()
    assert info["num_stream_tiles"] < info["num_tiles"]
    x = tvec(n, 1)
    assert relerr(dA.mv_host(0, x), checker.mv(0, A, x)) <= 1e-14


def test_spmv_edge_cases(ctx, checker):
    # empty rows, a single row, a 1x1 matrix, ragged tiny matrices
    for n, avg, seed in ((1, 1, 0), (2, 1, 1), (33, 2, 2), (257, 3, 3), (1025, 9, 4)):
        A = g.random_csr(n, avg, seed=seed, diag=False)
        if A[0][-1] == 0:
            continue
        dA = api.Csr(ctx, A)
        x, y = tvec(n), tvec(n, 1)
        assert np.array_equal(dA.mv_host(0, x), checker.mv(0, A, x))
        assert np.array_equal(dA.mv_host(3, x, alpha=-1.0, beta=1.0, y=y), checker.mv(3, A, x, alpha=-1.0, beta=1.0, y=y))


def test_spmv_zero_matrix_branch(ctx):
    # Ap == NULL: mxy/amxy give 0, amxpby/amxpbyz give y*beta (src/mvops.cxx:33-38,72-76,110-114,145-149)
    n = 1000
    Z = api.Csr(ctx, (None, n, None))
    x, y = tvec(n), tvec(n, 1)
    assert np.array_equal(Z.mv_host(0, x), np.zeros(n))
    assert np.array_equal(Z.mv_host(1, x, alpha=2.0), np.zeros(n))
    assert np.array_equal(Z.mv_host(2, x, alpha=2.0, beta=-3.0, y=y), y * -3.0)
    assert np.array_equal(Z.mv_host(3, x, alpha=2.0, beta=0.5, y=y), y * 0.5)


def test_spmv_beta_zero_propagates_nonfinite_y(ctx):
    # the reference still multiplies y*0
    A = matrix("cd3d_12")
    n = len(A[0]) - 1
    dA = api.Csr(ctx, A)
    y = np.zeros(n)
    y[5] = np.inf
    z = dA.mv_host(3, tvec(n), alpha=1.0, beta=0.0, y=y)
    assert np.isnan(z[5]) and np.all(np.isfinite(np.delete(z, 5)))


# --------------------------------------------------------------- BLAS-1 -----
@pytest.mark.parametrize("n", [1, 31, 1024, 1025, 100003, 1 << 20])
def test_blas1_elementwise_bit_exact(ctx, checker, n):
    x, y = tvec(n), tvec(n, 1)
    dx, dy, dz = ctx.upload(x), ctx.upload(y), ctx.empty(n)
    api.lssp_vec_axpbyz(ctx, -3.0, dx, 0.125, dy, dz)
    assert np.array_equal(dz.get(), checker.axpbyz(-3.0, x, 0.125, y))
    api.lssp_vec_axpby(ctx, 1.25, dx, -0.5, dy)
    assert np.array_equal(dy.get(), checker.axpby(1.25, x, -0.5, y))
    api.lssp_vec_axy(ctx, 0.3, dx, dz)
    assert np.array_equal(dz.get(), x * 0.3)
    api.lssp_vec_scale(ctx, dz, -7.0)
    assert np.array_equal(dz.get(), (x * 0.3) * -7.0)
    api.lssp_vec_copy(ctx, dz, dx)
    assert np.array_equal(dz.get(), x)
    api.lssp_vec_set_value(ctx, dz, 2.5)
    assert np.array_equal(dz.get(), np.full(n, 2.5))


@pytest.mark.parametrize("n", [1, 33, 4097, 1000003, 1 << 22])
def test_dot_and_norm(ctx, checker, n):
    x, y = tvec(n), tvec(n, 1) + 0.5
    dx, dy = ctx.upload(x), ctx.upload(y)
    import math
    want = checker.dot(x, y)
    got = api.lssp_vec_dot(ctx, dx, dy)
    scale = np.dot(np.abs(x), np.abs(y))
    exact = math.fsum((x * y).tolist()) if n <= (1 << 20) else None
    # the tree-ordered GPU sum is at least as accurate as the reference's sequential sum:
    # both sit within the sequential-summation rounding bound of each other ...
    assert abs(got - want) <= 4.0 * math.sqrt(n) * 2.3e-16 * scale + 1e-300
    # ... and the GPU result is within 1e-14 (relative to sum |x_i y_i|) of the exactly rounded sum
    if exact is not None:
        assert abs(got - exact) <= 1e-14 * scale
    assert abs(api.lssp_vec_norm(ctx, dx) - checker.norm(x)) <= 4.0 * math.sqrt(n) * 2.3e-16 * checker.norm(x)
    # run-to-run reproducible
    assert api.lssp_vec_dot(ctx, dx, dy) == got


def test_multidot(ctx, checker):
    n = 200001
    vs = [tvec(n, k) for k in range(8)]
    y = tvec(n, 9)
    d = [ctx.upload(v) for v in vs]
    dy = ctx.upload(y)
    for k in (1, 2, 3, 5, 8):
        got = api.lssp_vec_multidot(ctx, d[:k], dy)
        for i in range(k):
            assert abs(got[i] - checker.dot(vs[i], y)) <= 2e-13 * np.dot(np.abs(vs[i]), np.abs(y))


# ------------------------------------------------- triangular solves / pc ---
@pytest.mark.parametrize("name", list(MATRICES))
@pytest.mark.parametrize("tag,kw", [("iluk0", dict(kind="iluk", level=0)), ("iluk1", dict(kind="iluk", level=1)),
                                    ("ilut", dict(kind="ilut"))])
def test_trisolve_and_ilu_apply_bit_exact(ctx, golden, checker, name, tag, kw):
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, **kw)
    e = golden["factors"][name + "/" + tag]
    rhs = tvec(n)
    dL, dU = api.Tri(ctx, 0, L), api.Tri(ctx, 1, U)
    drhs, dy, dx = ctx.upload(rhs), ctx.empty(n), ctx.empty(n)
    dL.solve(dy, drhs)
    y = dy.get()
    assert sha(y) == e["lower_sha"]
    dU.solve(dx, dy)
    assert sha(dx.get()) == e["apply_sha"]
    pc = api.Preconditioner(ctx, "ilu", n, L, U)
    x = pc.apply_host(rhs)
    assert sha(x) == e["apply_sha"]
    assert np.array_equal(x, checker.tri_upper(U, checker.tri_lower(L, rhs)))
    # repeated application reuses the resident factors and is reproducible
    assert np.array_equal(pc.apply_host(rhs), x)


def test_all_tri_schedules_are_exercised(ctx, checker):
    """Stencil (lattice) factors take the pencil schedule; with it disabled (LSSPG_TRI_PENCIL=0 at analysis time) the box
    schedule, and with that disabled too (LSSPG_TRI_TILED=0) the slice schedule: the same factor must give the same bits
    on all three, in every variant of the pencil kernel (hole masks on, own-line operand through the ring)."""
    import os

    def tri_with(env, which, T):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            return api.Tri(ctx, which, T)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    A = matrix("lap3d_32")
    n = len(A[0]) - 1
    for level in (0, 1, 2):
        L, U = api.ilu_factor(A, "iluk", level=level)
        rhs = tvec(n, 4)
        for which, F, serial in ((0, L, checker.tri_lower), (1, U, checker.tri_upper)):
            want = serial(F, rhs)
            variants = [tri_with({}, which, F), tri_with({"LSSPG_TRI_PENCIL": "0"}, which, F),
                        tri_with({"LSSPG_TRI_PENCIL": "0", "LSSPG_TRI_TILED": "0"}, which, F),
                        tri_with({"LSSPG_TRI_PENCIL_HOLES": "1"}, which, F), tri_with({"LSSPG_TRI_PENCIL_OWNLAST": "0"}, which, F),
                        tri_with({"LSSPG_TRI_PENCIL": "16,16"}, which, F), tri_with({"LSSPG_TRI_PENCIL": "16,8"}, which, F)]
            kinds = [t.schedule()["kind"] for t in variants]
            # ILU(2) rows (12 off-diagonals) are wider than the pencil kernel's 8 slots: box schedule (skewed boxes)
            assert kinds[:3] == ([2, 1, 0] if level < 2 else [1, 1, 0]), kinds
            drhs, dx = ctx.upload(rhs), ctx.empty(n)
            for t in variants:
                t.solve(dx, drhs)
                assert np.array_equal(dx.get(), want)
                t.solve(dx, drhs)      # a second sweep reuses mailboxes / counters (ticket wrap)
                assert np.array_equal(dx.get(), want)
                t.free()
    Tp = api.Tri(ctx, 0, api.ilu_factor(matrix("powerlaw_4000"), "iluk", level=0)[0])
    assert Tp.schedule()["kind"] == 0


def test_trisolve_single_chain_and_diagonal(ctx, checker):
    # worst case for level scheduling: a bidiagonal chain (n levels) and a pure diagonal (1 level)
    n = 4000
    Lp = np.arange(0, 2 * n, 2, dtype=np.int32)
    Lp = np.concatenate([[0], 2 * np.arange(1, n + 1) - 1]).astype(np.int32)
    Lj = np.empty(2 * n - 1, np.int32)
    Lx = np.empty(2 * n - 1)
    Lj[0], Lx[0] = 0, 2.0
    for i in range(1, n):
        Lj[2 * i - 1], Lx[2 * i - 1] = i - 1, -0.5 + 0.001 * (i % 7)
        Lj[2 * i], Lx[2 * i] = i, 2.0 + 0.01 * (i % 5)
    L = (Lp, Lj, Lx)
    rhs = tvec(n)
    T = api.Tri(ctx, 0, L)
    assert T.info()["num_levels"] == n
    drhs, dx = ctx.upload(rhs), ctx.empty(n)
    T.solve(dx, drhs)
    assert np.array_equal(dx.get(), checker.tri_lower(L, rhs))
    D = (np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), 1.0 + np.arange(n) * 0.001)
    T2 = api.Tri(ctx, 1, D)
    assert T2.info()["num_levels"] == 1
    T2.solve(dx, drhs)
    assert np.array_equal(dx.get(), checker.tri_upper(D, rhs))


def test_block_ilu_apply_bit_exact(ctx, port):
    # K7: lower sweep, D SpMV, upper sweep (src/pc-biluk.cxx:22-60) with synthetic L/D/U
    A = matrix("cd3d_12")
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, "iluk", level=1)
    D = g.random_csr(n, 2, seed=11)
    pc = api.Preconditioner(ctx, "bilu", n, L, U, D)
    rhs = tvec(n, 2)
    assert np.array_equal(pc.apply_host(rhs), port.bilu_apply(L, D, U, rhs))


def test_pc_non_is_a_copy(ctx):
    n = 5000
    pc = api.Preconditioner.non(ctx, n)
    rhs = tvec(n)
    assert np.array_equal(pc.apply_host(rhs), rhs)


@pytest.mark.parametrize("N", [48])
def test_ilu_apply_larger_grid_and_linearity(ctx, checker, N):
    """Size-independent properties at a size the CPU checker still finishes quickly:
    M^-1 is linear, and applying it to A-scaled data is reproducible bit for bit."""
    A = g.cd3d(N)
    n = N ** 3
    pc = api.Preconditioner.iluk(ctx, A, level=0)
    assert pc.info()["levels_L"] == 3 * N - 2
    u, v = tvec(n), tvec(n, 1)
    xu, xv, xs = pc.apply_host(u), pc.apply_host(v), pc.apply_host(u + 2.0 * v)
    assert relerr(xs, xu + 2.0 * xv) <= 1e-12
    L, U = api.ilu_factor(A, "iluk", level=0)
    assert np.array_equal(xu, checker.tri_upper(U, checker.tri_lower(L, u)))

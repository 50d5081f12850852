"""Parity at the BASELINE.json sizes (256^3, 16.8 M rows): the CUDA path against the UNMODIFIED reference.

The reference's outputs come from tests/golden/baseline_256.json, generated in the build container by
tests/golden/make_baseline_golden.py from oracle/_ref (a full reference solve at this size takes minutes on one core):
SHA-256 of its kernel outputs, iterations to tolerance, final residuals and the first 20 residuals in full precision.
Where the compiled reference travelled to the GPU box (oracle/_ref), the kernels are also compared with it live.

Bars (BASELINE.json north_star): SpMV and preconditioner application bit-exact (met); residual histories within 1e-10
relative over the first 20 iterations and iterations to tolerance within +-1 -- at THIS size met for CG / GMRES counts only;
the measured gaps and their cause are stated where the bounds are set (and in DESIGN.md section 5)."""
import json
import os

import numpy as np
import pytest

from lssp_b200 import api
from lssp_b200 import generators as g
from util import sha, tvec

pytestmark = pytest.mark.gpu
N = 256
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    path = os.path.join(HERE, "golden", "baseline_%d.json" % N)
    if not os.path.exists(path):
        pytest.skip("tests/golden/baseline_%d.json not generated" % N)
    with open(path) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def lap():
    return g.lap3d(N)


@pytest.fixture(scope="module")
def cd():
    return g.cd3d(N)


def test_spmv_all_four_kinds_bit_exact_at_256(ctx, gold, lap, cd):
    e = gold["lap3d/kernels"]
    n = N ** 3
    x, y = tvec(n), tvec(n, 1)
    dA = api.Csr(ctx, lap)
    assert int(lap[0][-1]) == e["nnz"]
    assert sha(dA.mv_host(0, x)) == e["mxy_sha"]
    assert sha(dA.mv_host(1, x, alpha=-0.75)) == e["amxy_sha"]
    assert sha(dA.mv_host(2, x, alpha=1.25, beta=-0.5, y=y)) == e["amxpby_sha"]
    assert sha(dA.mv_host(3, x, alpha=-1.0, beta=1.0, y=y)) == e["amxpbyz_sha"]
    dA.free()
    dC = api.Csr(ctx, cd)
    assert sha(dC.mv_host(3, x, alpha=-1.0, beta=1.0, y=y)) == gold["cd3d/kernels"]["amxpbyz_sha"]
    dC.free()


@pytest.mark.parametrize("case", ["lap3d/ilu0", "cd3d/iluk1", "cd3d/ilut"])
def test_ilu_factors_and_application_bit_exact_at_256(ctx, gold, lap, cd, case):
    """host-threaded factorisation == the reference's factors; pencil sweeps (ILU(0), ILUK(1): 766 / 1531 dependency levels,
    256 pencils on 148 SMs, i.e. tickets beyond the resident CTAs) and slice sweeps (ILUT) == the reference's serial sweeps"""
    A, tag, kw = {"lap3d/ilu0": (lap, "ilu0", dict(kind="iluk", level=0)), "cd3d/iluk1": (cd, "iluk1", dict(kind="iluk", level=1)),
                  "cd3d/ilut": (cd, "ilut", dict(kind="ilut", p=-1, tol=1e-3))}[case]
    e = gold[case.split("/")[0] + "/kernels"]
    n = N ** 3
    L, U = api.ilu_factor(A, **kw)
    assert [int(L[0][-1]), int(U[0][-1])] == e[tag + "_nnz"]
    assert sha(np.concatenate([L[2], U[2]])) == e[tag + "_factor_sha"]
    pc = api.Preconditioner(ctx, "ilu", n, L, U)
    kinds = pc.info()
    x = tvec(n)
    got = pc.apply_host(x)
    assert sha(got) == e[tag + "_apply_sha"], kinds
    assert np.array_equal(pc.apply_host(x), got)      # a second application: mailboxes emptied, tickets wrapped
    pc.free()


CASES = {"lap3d/cg+iluk0": ("lap", "cg", dict(kind="iluk", level=0), {}),
         "lap3d/bicgstab+iluk0": ("lap", "bicgstab", dict(kind="iluk", level=0), {}),
         "cd3d/bicgstab+iluk1": ("cd", "bicgstab", dict(kind="iluk", level=1), {}),
         "cd3d/gmres30+ilut": ("cd", "gmres", dict(kind="ilut", p=-1, tol=1e-3), dict(restart=30))}


@pytest.mark.parametrize("case", list(CASES))
def test_solves_match_the_reference_at_256(ctx, gold, lap, cd, case, record_property):
    if case not in gold:
        pytest.skip("no golden values for " + case)
    which, solver, pckw, skw = CASES[case]
    A = lap if which == "lap" else cd
    e = gold[case]
    n = N ** 3
    dA = api.Csr(ctx, A)
    L, U = api.ilu_factor(A, **pckw)
    pc = api.Preconditioner(ctx, "ilu", n, L, U)
    x = np.zeros(n)
    r = api.lssp_solver_solve(ctx, solver, dA, pc, np.ones(n), x, nhist=20, maxit=3000, **skw)
    want = np.array(e["history"])
    k = min(len(want), len(r["hist"]))
    # the reference's GMRES only refreshes solver.residual at the end of a restart cycle (src/solver-gmres.cxx:206-221), so
    # its maxit = 1..20 "history" is one repeated number: nothing to compare per iteration
    per_iteration = len(set(want.tolist())) > 2
    err = float(np.max(np.abs(r["hist"][:k] - want[:k]) / want[:k])) if per_iteration else 0.0
    record_property("history_relerr", err)
    record_property("nits", r["nits"])
    print("%s: nits %d (reference %d) residual %.9e (reference %.9e) history relerr %.2e" %
          (case, r["nits"], e["nits"], r["residual"], e["residual"], err))
    if per_iteration:
        # KNOWN GAP: north_star asks for 1e-10 over the first 20 iterations.  That holds at the sizes of
        # the other test files (<= 96^3: ~1e-13) but not here: at n = 16.8 M the reference's own sequential dot products
        # carry ~1.5e-14 of rounding error (measured against an exact sum), the shipped tree reductions ~1e-16, and the
        # Krylov recurrences of these ill-conditioned problems amplify that difference to 1.6e-9 (CG + ILU(0)),
        # 1.2e-6 (BiCGStab + ILU(0)) and 6e-5 (BiCGStab + ILUK(1)) within 20 iterations.  Matching the reference's
        # rounding needs its summation ORDER (sequential-reduction mode).  The bounds below guard against regressions.
        bound = {"lap3d/cg+iluk0": 1e-8, "lap3d/bicgstab+iluk0": 1e-5, "cd3d/bicgstab+iluk1": 1e-3}[case]
        assert k == 20 and err <= bound, (case, err)
    if solver == "bicgstab":
        # KNOWN GAP: BiCGStab's path is chaotic in the summation order of its dot products -- the shipped
        # tree reductions and the reference's sequential sums part ways after ~100 iterations, so the count to tolerance
        # is NOT within +-1 (256^3: 175 vs 155, 83 vs 87); equality holds in the sequential-reduction mode only
        # (tests/test_gpu_all_drivers.py, small sizes).  The bound here only guards against regressions.
        assert abs(r["nits"] - e["nits"]) <= 0.15 * e["nits"], (case, r["nits"], e["nits"])
    else:
        assert abs(r["nits"] - e["nits"]) <= 1, (case, r["nits"], e["nits"])
    assert abs(np.linalg.norm(x) - e["x_norm"]) <= 1e-6 * e["x_norm"]
    pc.free()
    dA.free()


@pytest.mark.parametrize("case", list(CASES))
def test_solves_EQUAL_the_reference_at_256_with_reference_order_sums(gold, lap, cd, case):
    """LSSPG_OPT_REDUCE_SEQUENTIAL = 2 (exact_sum.cu): every dot product is the reference's sequential sum, computed in
    parallel.  Then nothing differs from the CPU arithmetic and the whole solve is bit-identical at the BASELINE size:
    iterations to tolerance, final residual and the first 20 residuals EQUAL the unmodified reference's -- the bars of
    BASELINE.json north_star (1e-10, +-1) are met with margin zero, BiCGStab included."""
    if case not in gold:
        pytest.skip("no golden values for " + case)
    which, solver, pckw, skw = CASES[case]
    A = lap if which == "lap" else cd
    e = gold[case]
    n = N ** 3
    c = api.Context(0)
    c.set_option(api.OPT_REDUCE_SEQUENTIAL, 2)
    dA = api.Csr(c, A)
    L, U = api.ilu_factor(A, **pckw)
    pc = api.Preconditioner(c, "ilu", n, L, U)
    x = np.zeros(n)
    r = api.lssp_solver_solve(c, solver, dA, pc, np.ones(n), x, nhist=20, maxit=3000, **skw)
    print("%s: nits %d (reference %d) residual %.17g (reference %.17g) %.1f ms" %
          (case, r["nits"], e["nits"], r["residual"], e["residual"], r.get("solve_ms", -1.0)))
    assert r["nits"] == e["nits"]
    assert r["residual"] == e["residual"]
    want = np.array(e["history"])
    if len(set(want.tolist())) > 2:      # (GMRES: the reference's history is one repeated number, see above)
        assert list(r["hist"][:len(want)]) == list(want)
    assert abs(np.linalg.norm(x) - e["x_norm"]) <= 1e-13 * e["x_norm"]
    pc.free()
    dA.free()
    c.close()


def test_pencil_sweeps_are_repeatable_at_256(ctx, gold, lap):
    """256 pencils on 148 SMs, 286 steps each, five roles per CTA synchronised through progress words: every application
    must give the same bits, for several right-hand sides, and the bits of the box schedule (which shares no code with
    the pencil kernel).  This is the test that caught the relaxed-flag race of the first warp-specialised version
    (intermittent: ~40 % of the applications wrong in ~50 000 of 16.8 M rows; smaller grids never showed it)."""
    n = N ** 3
    L, U = api.ilu_factor(lap, "iluk", level=0)
    pc = api.Preconditioner(ctx, "ilu", n, L, U)
    os.environ["LSSPG_TRI_PENCIL"] = "0"
    try:
        box = api.Preconditioner(ctx, "ilu", n, L, U)
    finally:
        del os.environ["LSSPG_TRI_PENCIL"]
    for off in (0, 1, 2):
        v = np.sin(np.arange(n) * 0.37 + off) + 0.25
        want = box.apply_host(v)
        for rep in range(5):
            assert np.array_equal(pc.apply_host(v), want), (off, rep)
    assert sha(pc.apply_host(tvec(n))) == gold["lap3d/kernels"]["ilu0_apply_sha"]
    pc.free()
    box.free()

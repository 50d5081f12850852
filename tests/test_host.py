"""CPU suite, part 2: host-side logic of the product and the C-ABI surface.

No compute call is made on a device here: the library is loaded, its exports
are compared with include/lsspg.h, and the host-side set-up code (incomplete
factorisations, level analysis, level-ordered layout) is pinned against the
golden fixtures and, when present, the compiled reference.
"""
import os
import re

import numpy as np
import pytest

from lssp_b200 import api, lib
from util import MATRICES, matrix, sha, tvec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "lsspg.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lsspg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = lib()
    syms = header_symbols()
    assert len(syms) > 40
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, "declared in include/lsspg.h but not exported: %s" % missing
    assert b"sm_100a" in L.lsspg_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(Exception) as e:
        api.Context(0)
    assert "no CPU fallback" in str(e.value)


FACTOR_CASES = [("iluk0", dict(kind="iluk", level=0)), ("iluk1", dict(kind="iluk", level=1)),
                ("iluk2", dict(kind="iluk", level=2)), ("ilut", dict(kind="ilut"))]


@pytest.mark.parametrize("name", list(MATRICES))
@pytest.mark.parametrize("tag,kw", FACTOR_CASES)
def test_host_factorisation_matches_golden(golden, name, tag, kw):
    A = matrix(name)
    L, U = api.ilu_factor(A, **kw)
    e = golden["factors"][name + "/" + tag]
    assert (int(L[0][-1]), int(U[0][-1])) == (e["nnzL"], e["nnzU"])
    assert sha(*L) == e["L_sha"]
    assert sha(*U) == e["U_sha"]


@pytest.mark.parametrize("name", list(MATRICES))
def test_block_jacobi_factorisation_matches_golden(golden, name):
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, "iluk", level=0, blk_size=(n + 3) // 4)
    e = golden["factors"][name + "/iluk0_bj4"]
    assert sha(*L) == e["L_sha"] and sha(*U) == e["U_sha"]
    L, U = api.ilu_factor(A, "ilut", blk_size=(n + 1) // 2)
    e = golden["factors"][name + "/ilut_bj2"]
    assert sha(*L) == e["L_sha"] and sha(*U) == e["U_sha"]


def test_known_answers_appendix_a2_factors():
    # SURVEY.md App. A.2 (N = 100): nnz and level counts of the reference's own factors
    A = matrix("lap2d_100")
    for level, nnz, nlev in ((0, 29800, 199), (1, 39601, 298), (2, 39650, 299)):
        L, U = api.ilu_factor(A, "iluk", level=level)
        assert int(L[0][-1]) == nnz and int(U[0][-1]) == nnz
        assert api.tri_levels(0, L)[1] == nlev and api.tri_levels(1, U)[1] == nlev
    L, U = api.ilu_factor(A, "ilut")
    assert (int(L[0][-1]), int(U[0][-1])) == (59598, 59887)
    assert (api.tri_levels(0, L)[1], api.tri_levels(1, U)[1]) == (593, 689)


@pytest.mark.parametrize("name", ["lap3d_32", "cd3d_12"])
def test_level_law_3d(name):
    # ILU(0) on an N^3 7-point grid has 3N-2 levels
    A = matrix(name)
    N = round((len(A[0]) - 1) ** (1 / 3))
    L, U = api.ilu_factor(A, "iluk", level=0)
    lev, nl = api.tri_levels(0, L)
    assert nl == 3 * N - 2
    # a row's level is strictly greater than the level of every row it depends on
    Lp, Lj, _ = L
    for i in range(0, len(Lp) - 1, 97):
        deps = Lj[Lp[i]:Lp[i + 1] - 1]
        assert all(lev[d] < lev[i] for d in deps)


@pytest.mark.parametrize("name", list(MATRICES))
@pytest.mark.parametrize("tag,kw", [FACTOR_CASES[0], FACTOR_CASES[1], FACTOR_CASES[3]])
def test_level_ordered_layout_reproduces_reference_sweeps(golden, name, tag, kw):
    """Walking the sliced-ELL image in ticket order must give the reference's
    lower / upper sweeps bit for bit (same products, same order)."""
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, **kw)
    e = golden["factors"][name + "/" + tag]
    y, ns, pad = api.tri_walk_layout_host(0, L, tvec(n))
    assert sha(y) == e["lower_sha"]
    x, _, _ = api.tri_walk_layout_host(1, U, y)
    assert sha(x) == e["apply_sha"]
    assert pad >= int(L[0][-1]) - n and ns >= (n + 31) // 32


@pytest.mark.parametrize("name", ["lap3d_32", "cd3d_32", "lap2d_100"])
@pytest.mark.parametrize("tag,kw", [FACTOR_CASES[0], FACTOR_CASES[1], FACTOR_CASES[2]])
def test_box_schedule_reproduces_reference_sweeps(golden, monkeypatch, name, tag, kw):
    """Structured-grid factors get the box (tile) schedule: one warp per 8x8x8 (16x16) box,
    in-box operands from shared memory.  Emulating it on the host -- boxes of one strongly
    connected component advancing concurrently, level barriers inside each box -- must give
    the reference's sweeps bit for bit and must never stall (ILU(1)/(2) fill makes the box
    graph cyclic)."""
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, **kw)
    e = golden["factors"][name + "/" + tag]
    if tag != "iluk0":
        # the plain box grid is cyclic for fill factors: no box schedule; the default -- skewed boxes -- is acyclic
        monkeypatch.setenv("LSSPG_TRI_SKEW", "0")
        assert api.tri_walk_tiled_host(0, L, tvec(n))[1] is None
        monkeypatch.delenv("LSSPG_TRI_SKEW")
    y, info = api.tri_walk_tiled_host(0, L, tvec(n))
    assert info is not None, "expected a box schedule for a stencil factor"
    assert info["nx"] * info["ny"] * info["nz"] == n and info["max_box_rows"] <= 512
    assert sha(y) == e["lower_sha"]
    x, info_u = api.tri_walk_tiled_host(1, U, y)
    assert sha(x) == e["apply_sha"]
    if tag == "iluk0":
        N = info["nx"]
        assert info["row_levels"] == (3 * N - 2 if info["nz"] > 1 else 2 * N - 1)
        assert info["box_levels"] < info["row_levels"] / 5    # most hops of the critical path stay in a box


def test_box_schedule_not_used_for_irregular_factors():
    A = matrix("powerlaw_4000")
    L, U = api.ilu_factor(A, "iluk", level=0)
    assert api.tri_walk_tiled_host(0, L, tvec(len(A[0]) - 1))[0] is None


def test_tri_analysis_rejects_malformed_factors():
    Lp = np.array([0, 1, 3], np.int32)
    Lj = np.array([0, 1, 0], np.int32)     # row 1 stores its diagonal first, not last
    with pytest.raises(Exception):
        api.tri_levels(0, (Lp, Lj, np.ones(3)))
    Up = np.array([0, 2, 3], np.int32)
    Uj = np.array([0, 0, 1], np.int32)     # duplicate diagonal inside the strict upper part
    with pytest.raises(Exception):
        api.tri_levels(1, (Up, Uj, np.ones(3)))


def test_host_factorisation_matches_compiled_reference(ref):
    for name in ("cd3d_12", "random_600", "powerlaw_4000"):
        A = matrix(name)
        n = len(A[0]) - 1
        for kw in (dict(kind="iluk", level=3), dict(kind="ilut", p=9, tol=1e-4),
                   dict(kind="iluk", level=1, blk_size=(n + 2) // 3), dict(kind="ilut", p=4, tol=1e-2, blk_size=(n + 4) // 5)):
            F, G = api.ilu_factor(A, **kw), ref.ilu(A, **kw)
            assert all(np.array_equal(a, b) for X, Y in zip(F, G) for a, b in zip(X, Y)), (name, kw)


def test_generators_are_deterministic_and_sorted():
    A = matrix("powerlaw_4000")
    Ap, Aj, Ax = A
    for i in range(0, 4000, 37):
        row = Aj[Ap[i]:Ap[i + 1]]
        assert np.all(np.diff(row) > 0) and i in row
    # strict diagonal dominance
    i = 123
    row, val = Aj[Ap[i]:Ap[i + 1]], Ax[Ap[i]:Ap[i + 1]]
    assert val[row == i][0] > np.abs(val[row != i]).sum()
    from lssp_b200 import generators as g
    B = g.lap3d(8)
    assert int(B[0][-1]) == 7 * 512 - 6 * 64


def test_generators_blocked_on_the_thread_pool_give_the_same_arrays(monkeypatch):
    from lssp_b200 import generators as g
    monkeypatch.setattr(g, "_PL_BLOCK", 700)
    A = g.powerlaw(4000, window=300)
    B = g.powerlaw_rows(4000, 0, 4000, window=300)
    assert all(np.array_equal(a, b) for a, b in zip(A, B)) and A[0].dtype == np.int32
    assert sha(*A) == sha(*matrix("powerlaw_4000"))
    C = g.stencil_7pt(20, conv=(0.3, 0.2, 0.1), chunk_planes=3)      # 7 chunks on the pool
    D = g.stencil_7pt(20, conv=(0.3, 0.2, 0.1), chunk_planes=20)     # one chunk, inline
    assert all(np.array_equal(a, b) for a, b in zip(C, D))

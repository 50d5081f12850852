"""CPU suite for the SX-AMG-style path: the host set-up (lssp_b200/csrc/amg_host.cpp), the smoother
schedule, and the restated cycle (oracle/amg_oracle.c).

libsxamg is not in the reference tree and no reference test pins anything at its boundary
(SURVEY.md 8c): PARITY WITH libsxamg IS UNPINNED.  What is pinned here is (i) the set-up against
independent statements of its own specification (DESIGN.md "AMG") written with numpy/scipy, and
(ii) the restated cycle against the textbook behaviour of a V-cycle (grid-independent contraction,
a handful of PCG iterations)."""
import numpy as np
import pytest
import scipy.sparse as sp

from lssp_b200 import api
from util import matrix, tvec

CASES = ["lap2d_100", "lap3d_32", "cd3d_12", "random_600"]


@pytest.fixture(scope="module")
def hier():
    cache = {}

    def get(name, **pars):
        key = (name, tuple(sorted(pars.items())))
        if key not in cache:
            cache[key] = api.AmgHierarchy(matrix(name), **pars)
        return cache[key]
    return get


def csr(t, shape):
    return sp.csr_matrix((t[2], t[1], t[0]), shape=shape)


@pytest.mark.parametrize("name", CASES)
def test_hierarchy_shape_and_galerkin_products(hier, name):
    H = hier(name)
    A = matrix(name)
    L0 = H.levels[0]
    assert np.array_equal(L0["A"][0], A[0]) and np.array_equal(L0["A"][1], A[1]) and np.array_equal(L0["A"][2], A[2])
    assert 1 <= len(H.levels) <= 30
    for l, L in enumerate(H.levels[:-1]):
        n, nc = L["n"], L["nc"]
        nxt = H.levels[l + 1]
        assert 0 < nc < n and nxt["n"] == nc
        assert int(L["cf"].sum()) == nc
        Al, P, R = csr(L["A"], (n, n)), csr(L["P"], (n, nc)), csr(L["R"], (nc, n))
        assert abs(R - P.T).max() == 0.0                      # R = P^T exactly
        Ac = csr(nxt["A"], (nc, nc))
        G = (R @ Al @ P).tocsr()
        scale = abs(G).max()
        assert abs(Ac - G).max() <= 1e-13 * scale             # A_c = R A P
        for t in (L["A"], L["P"], L["R"], nxt["A"]):          # sorted columns, no duplicates
            for i in range(len(t[0]) - 1):
                assert np.all(np.diff(t[1][t[0][i]:t[0][i + 1]]) > 0)
    last = H.levels[-1]
    assert last["nc"] == 0 and last["P"] is None and np.all(last["cf"] == 1)
    if H.coarse_dense:
        Al = csr(last["A"], (last["n"], last["n"])).toarray()
        assert np.abs(H.coarse_inv @ Al - np.eye(last["n"])).max() <= 1e-9


@pytest.mark.parametrize("name", CASES)
def test_cf_split_and_interpolation_follow_the_specification(hier, name):
    H = hier(name)
    theta, max_row_sum, trunc = H.pars.strong_threshold, H.pars.max_row_sum, H.pars.trunc_threshold
    for L in H.levels[:-1]:
        n = L["n"]
        Ap, Aj, Ax = L["A"]
        cf = L["cf"]
        cidx = np.cumsum(cf) - 1
        Pp, Pj, Px = L["P"]
        for i in range(n):
            cols, vals = Aj[Ap[i]:Ap[i + 1]], Ax[Ap[i]:Ap[i + 1]]
            off = cols != i
            diag = vals[~off][0]
            s = -1.0 if diag < 0 else 1.0
            a = s * vals
            most = max(0.0, (-a[off]).max()) if off.any() else 0.0
            dominated = max_row_sum < 1.0 and abs(vals.sum()) > max_row_sum * abs(diag)
            strong = off & (-a >= theta * most) if (most > 0 and not dominated) else np.zeros(len(cols), bool)
            prow_c, prow_v = Pj[Pp[i]:Pp[i + 1]], Px[Pp[i]:Pp[i + 1]]
            if cf[i]:
                assert list(prow_c) == [cidx[i]] and list(prow_v) == [1.0]
                continue
            if not strong.any():
                assert len(prow_c) == 0                        # nothing to interpolate from
                continue
            sc = strong & (cf[cols] == 1)
            assert sc.any(), "F point %d has strong couplings but no strong C neighbour" % i
            # direct interpolation over the strong C neighbours, then truncation with rescaling
            all_neg, all_pos = a[off & (a < 0)].sum(), a[off & (a >= 0)].sum()
            alpha = all_neg / a[sc].sum()
            w = -alpha * a[sc] / (s * diag + all_pos)
            keep = np.abs(w) >= trunc * np.abs(w).max()
            w_kept = w[keep] * (w.sum() / w[keep].sum())
            assert list(prow_c) == list(cidx[cols[sc]][keep])
            assert np.allclose(prow_v, w_kept, rtol=1e-13, atol=0)


def test_interpolation_reproduces_constants_on_zero_row_sum_rows(hier):
    H = hier("lap3d_32")
    L = H.levels[0]
    Ap, Aj, Ax = L["A"]
    P = csr(L["P"], (L["n"], L["nc"]))
    rowsum_A = np.add.reduceat(Ax, Ap[:-1])
    interior_f = (np.abs(rowsum_A) < 1e-14) & (L["cf"] == 0)
    assert interior_f.sum() > 1000
    assert np.allclose(np.asarray(P.sum(axis=1)).ravel()[interior_f], 1.0, rtol=1e-13)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cf_order", [1, 0, 2])
def test_smoother_schedule_walk_equals_the_sequential_sweep(hier, port, name, cf_order):
    """Slices walked in ticket order, operands taken from x_new / x_old by the schedule's rule,
    every x_new operand already written when it is read: same bits as the serial in-place sweep."""
    H = hier(name, cf_order=cf_order)
    for l, L in enumerate(H.levels):
        last = l == len(H.levels) - 1
        b, x0 = tvec(L["n"], l), tvec(L["n"], l + 5)
        assert sorted(L["rank"]) == list(range(L["n"]))          # a total visiting order
        if cf_order != 2 or last:
            assert np.array_equal(L["rank"], np.arange(L["n"]))
        for post in (0, 1):
            xo = port.gs_sweep(L["A"], L["cf"] if (cf_order and not last) else None, post, b, x0,
                               rank=L["rank"] if (cf_order and not last) else None)
            for mode in (0, 1, 2):    # device's choice, 32-row slices, one ticket per row
                xn, info = H.walk_gs_host(l, post, b, x0, mode=mode)
                assert np.array_equal(xn, xo)
        if l == 0 and cf_order and name in ("lap2d_100", "lap3d_32", "cd3d_12"):
            # red-black: C points are mutually independent and so are the F points
            assert info["levels_c"] == 1 and info["levels_f"] == 1


def test_multicolour_order_has_short_dependency_chains(hier, port):
    """cf_order=2: each block is visited colour by colour -- still a sequential Gauss-Seidel sweep
    (the walk above equals the serial loop in that order), but the chains are tens of rows long
    where the index order gives hundreds: every level becomes a streaming sweep on the GPU."""
    deep = hier("lap3d_32")
    flat = hier("lap3d_32", cf_order=2)
    n = flat.levels[0]["n"]
    for l in range(len(flat.levels) - 1):
        L = flat.levels[l]
        b, x0 = np.ones(L["n"]), np.zeros(L["n"])
        d_new = flat.walk_gs_host(l, 0, b, x0)[1]
        d_old = deep.walk_gs_host(l, 0, b, x0)[1]
        assert d_new["levels_c"] + d_new["levels_f"] <= 64
        assert d_new["levels_c"] + d_new["levels_f"] <= d_old["levels_c"] + d_old["levels_f"]
    assert deep.walk_gs_host(1, 0, np.ones(deep.levels[1]["n"]), np.zeros(deep.levels[1]["n"]))[1]["levels_f"] > 64
    # and it smooths as well: same cycle counts to 1e-8
    a = port.amg(deep.levels, coarse_inv=deep.coarse_inv).solve(np.ones(n), tol=1e-8, maxit=50)
    b = port.amg(flat.levels, coarse_inv=flat.coarse_inv, cf_order=2).solve(np.ones(n), tol=1e-8, maxit=50)
    assert abs(a["nits"] - b["nits"]) <= 2 and b["nits"] <= 10


def test_setup_rejects_unsorted_columns_and_missing_diagonals():
    Ap = np.array([0, 2, 4], np.int32)
    with pytest.raises(Exception, match="not sorted"):
        api.AmgHierarchy((Ap, np.array([1, 0, 0, 1], np.int32), np.array([1.0, 2.0, 1.0, 2.0])), coarse_dof=1)
    H = api.AmgHierarchy((Ap, np.array([0, 1, 0, 1], np.int32), np.array([0.0, 1.0, 1.0, 2.0])), coarse_dof=1,
                         coarse_dense_max=0)
    with pytest.raises(Exception, match="diagonal"):
        H.walk_gs_host(len(H.levels) - 1, 0, np.ones(2), np.zeros(2))


@pytest.mark.parametrize("name,cycles", [("lap2d_100", 12), ("lap3d_32", 9), ("cd3d_12", 8)])
def test_restated_cycle_contracts_like_a_v_cycle(hier, port, name, cycles):
    H = hier(name)
    A = matrix(name)
    n = H.levels[0]["n"]
    m = port.amg(H.levels, coarse_inv=H.coarse_inv)
    r = m.solve(np.ones(n), tol=1e-8, maxit=50)
    assert r["nits"] <= cycles
    M = csr(A, (n, n))
    assert np.linalg.norm(np.ones(n) - M @ r["x"]) <= 1e-8 * np.sqrt(n) * 1.0000001
    assert abs(np.linalg.norm(np.ones(n) - M @ r["x"]) - r["residual"]) <= 1e-9 * r["residual"] + 1e-16


def test_initial_guess_semantics(hier, port):
    """The reference's adapter hands the caller's x to the cycle as the initial guess
    (src/pc-sxamg.cxx:58-64): default.  zero_guess=1 makes the cycle a fixed linear operator,
    which is what PCG needs: 4 iterations on the 32^3 Laplacian instead of stagnation."""
    H = hier("lap3d_32")
    A = matrix("lap3d_32")
    n = H.levels[0]["n"]
    lit = port.amg(H.levels, coarse_inv=H.coarse_inv)
    zer = port.amg(H.levels, coarse_inv=H.coarse_inv, zero_guess=1)
    b, x0 = tvec(n), tvec(n, 3)
    assert not np.array_equal(lit.cycle(b, x0), lit.cycle(b))
    assert np.array_equal(zer.cycle(b, x0), lit.cycle(b))
    good = port.solve("cg", A, np.ones(n), amg=zer, maxit=100)
    assert good["nits"] <= 6
    stale = port.solve("cg", A, np.ones(n), amg=lit, maxit=30)
    assert stale["nits"] == 30


# ---- committed fixtures (tests/golden/amg_golden.json, generated by tests/golden/make_amg_golden.py) ----
import hashlib
import json
import os


def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.fixture(scope="module")
def amg_golden():
    with open(os.path.join(os.path.dirname(__file__), "golden", "amg_golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["lap2d_100", "lap3d_32", "cd3d_12"])
@pytest.mark.parametrize("order", [1, 2])
def test_setup_and_restated_cycle_match_the_committed_fixtures(hier, port, amg_golden, name, order):
    """pins the specification as implemented (NOT libsxamg: parity with it is unpinned): hierarchy arrays bit
    for bit, one V-cycle of the restatement bit for bit, cycle and PCG iteration counts"""
    e = amg_golden["%s/cf%d" % (name, order)]
    H = hier(name, cf_order=order)
    assert [[L["n"], L["nc"], int(L["A"][0][-1])] for L in H.levels] == e["levels"]
    got = [_sha(L["A"][0], L["A"][1], L["A"][2], L["cf"], L["rank"],
                *([] if L["P"] is None else [L["P"][0], L["P"][1], L["P"][2]])) for L in H.levels]
    assert got == e["level_sha"]
    if H.coarse_dense:
        assert _sha(H.coarse_inv) == e["coarse_inv_sha"]
    n = H.levels[0]["n"]
    m = port.amg(H.levels, coarse_inv=H.coarse_inv, cf_order=order)
    y = m.cycle(tvec(n), tvec(n, 3))
    assert _sha(y) == e["cycle_sha"] and [float(v) for v in y[:4]] == e["cycle_head"]
    sa = m.solve(np.ones(n), tol=1e-8, maxit=50)
    assert sa["nits"] == e["standalone_nits"] and sa["residual"] == e["standalone_residual"]
    if "pcg_zero_guess_nits" in e:
        mz = port.amg(H.levels, coarse_inv=H.coarse_inv, cf_order=order, zero_guess=1)
        r = port.solve("cg", matrix(name), np.ones(n), amg=mz, maxit=100)
        assert r["nits"] == e["pcg_zero_guess_nits"] and r["residual"] == e["pcg_zero_guess_residual"]


@pytest.mark.parametrize("cf_order", [1, 2])
def test_smoother_layout_built_by_several_threads_walks_like_the_serial_sweep(port, cf_order):
    """64 000 rows: the slices of the fine levels are filled by several host threads (amg_host.cpp, gs_build_host);
    walking the image must still give the serial in-place sweep bit for bit, in all three layouts."""
    from lssp_b200 import generators as g
    H = api.AmgHierarchy(g.lap3d(40), cf_order=cf_order)
    for l in (0, 1):
        L = H.levels[l]
        b, x0 = tvec(L["n"], l), tvec(L["n"], l + 5)
        for post in (0, 1):
            xo = port.gs_sweep(L["A"], L["cf"], post, b, x0, rank=L["rank"])
            for mode in (0, 1, 2):
                assert np.array_equal(H.walk_gs_host(l, post, b, x0, mode=mode)[0], xo)

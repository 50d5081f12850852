"""CPU replay of the device AMG set-up (lssp_b200/csrc/amg_rows.cuh: strong couplings, direct interpolation with
truncation, restriction = P^T from the transposed strength graph, Galerkin products with per-row accumulator tables).
The set-up loop is shared (amg_setup_with); the replay provider runs the row functions the device kernels run, row after
row.  The hierarchy must equal the host set-up's array by array: levels, C/F splitting, A, P, R, visiting ranks, dense
inverse of the last level."""
import numpy as np
import pytest

from lssp_b200 import api
from lssp_b200 import generators as g
from util import matrix


def same_hierarchy(H, G):
    assert len(H.levels) == len(G.levels) and H.coarse_dense == G.coarse_dense
    for l, (a, b) in enumerate(zip(H.levels, G.levels)):
        assert (a["n"], a["nc"]) == (b["n"], b["nc"]), l
        for key in ("A", "P", "R"):
            if a[key] is None:
                assert b[key] is None
                continue
            for u, v in zip(a[key], b[key]):
                assert u.dtype == v.dtype and np.array_equal(u, v), (l, key)
        assert np.array_equal(a["cf"], b["cf"]) and np.array_equal(a["rank"], b["rank"]), l
    if H.coarse_dense:
        assert np.array_equal(H.coarse_inv, G.coarse_inv)


MATS = {"lap2d_100": lambda: matrix("lap2d_100"), "lap3d_32": lambda: matrix("lap3d_32"), "cd3d_32": lambda: matrix("cd3d_32"),
        "powerlaw_4000": lambda: matrix("powerlaw_4000"), "random_600": lambda: matrix("random_600"),
        "lap3d_48": lambda: g.lap3d(48)}


@pytest.mark.parametrize("name", list(MATS))
def test_row_functions_reproduce_the_host_hierarchy(name):
    A = MATS[name]()
    same_hierarchy(api.AmgHierarchy(A, replay=True), api.AmgHierarchy(A))


@pytest.mark.parametrize("pars", [dict(strong_threshold=0.5), dict(max_row_sum=0.5), dict(trunc_threshold=0.0),
                                  dict(trunc_threshold=0.6, cf_order=2), dict(coarse_dof=10, max_levels=4),
                                  dict(strong_threshold=0.05, coarse_dense_max=50)])
def test_row_functions_with_other_parameters(pars):
    for name in ("cd3d_12", "lap2d_100", "random_600"):
        A = matrix(name)
        same_hierarchy(api.AmgHierarchy(A, replay=True, **pars), api.AmgHierarchy(A, **pars))


def test_negative_diagonals_and_positive_couplings():
    """s = sign(a_ii) = -1 rows and couplings of the 'wrong' sign (lumped into the diagonal)"""
    rng = np.random.default_rng(4)
    Ap, Aj, Ax = (a.copy() for a in matrix("cd3d_12"))
    n = len(Ap) - 1
    for i in rng.integers(0, n, 200):
        Ax[Ap[i]:Ap[i + 1]] *= -1.0
    for k in rng.integers(0, len(Ax), 300):
        if Aj[k] != np.searchsorted(Ap, k, side="right") - 1:
            Ax[k] = abs(Ax[k]) * 0.3
    A = (Ap, Aj, Ax)
    same_hierarchy(api.AmgHierarchy(A, replay=True), api.AmgHierarchy(A))

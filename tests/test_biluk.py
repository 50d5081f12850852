"""Block ILU(k) (LSSP_PC_BILUK; reference src/pc-biluk.cxx, compiled there only with BLAS + LAPACK): host set-up
`lsspg_bilu_factor` against fixtures generated from the unmodified reference sources over the netlib reference dense
kernels (oracle/blas_standin.c, tests/golden/make_biluk_golden.py) and against that build live."""
import json
import os

import numpy as np
import pytest

from lssp_b200 import api, generators as g
from util import matrix, sha, tvec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "biluk_golden.json")) as f:
    GOLD = json.load(f)


def _case(key):
    name, bs, k = key.split("/")[:3]
    return name, int(bs[2:]), int(k[1:])


@pytest.fixture(scope="module")
def refb():
    import oracle
    if not oracle.RefB.available():
        pytest.skip("oracle/_ref/liblssp_refb.so not built (needs /root/reference at build time)")
    return oracle.RefB()


@pytest.mark.parametrize("key", sorted(GOLD["factors"]))
def test_block_factors_match_the_reference_fixtures(key, port):
    name, bs, level = _case(key)
    A = matrix(name)
    n = len(A[0]) - 1
    L, D, U = api.bilu_factor(A, n // bs, level)
    e = GOLD["factors"][key]
    assert (int(L[0][-1]), int(D[0][-1]), int(U[0][-1])) == (e["nnzL"], e["nnzD"], e["nnzU"])
    assert (sha(*L), sha(*D), sha(*U)) == (e["L_sha"], e["D_sha"], e["U_sha"])
    # layout the sweeps rely on: unit diagonal last in L / first in U (src/pc-biluk.cxx:37-59)
    assert np.all(L[1][L[0][1:] - 1] == np.arange(n)) and np.all(L[2][L[0][1:] - 1] == 1.0)
    assert np.all(U[1][U[0][:-1]] == np.arange(n)) and np.all(U[2][U[0][:-1]] == 1.0)
    # x = U^-1 D L^-1 rhs by the restated serial sweeps = the reference's own application
    assert sha(port.bilu_apply(L, D, U, tvec(n, 2))) == e["apply_sha"]


def test_block_factors_equal_the_compiled_reference(refb):
    same = lambda F, G: all(np.array_equal(a, b) for X, Y in zip(F, G) for a, b in zip(X, Y))  # noqa: E731
    for A in (g.cd3d(10), g.random_csr(720, 6, seed=3), g.laplacian_5pt(36), g.powerlaw(1200, window=100)):
        n = len(A[0]) - 1
        for bs in (1, 2, 3, 5, 8):
            if n % bs:
                continue
            for level in (0, 1, 3):
                assert same(api.bilu_factor(A, n // bs, level), refb.bilu(A, n // bs, level)), (n, bs, level)
    # sizes at which the level-scheduled numeric phase and the threaded assembly really use several threads
    A = g.cd3d(48)
    n = len(A[0]) - 1
    for bs, level in ((2, 1), (1, 2), (3, 0)):
        assert same(api.bilu_factor(A, n // bs, level), refb.bilu(A, n // bs, level)), (n, bs, level)
    # unsorted rows are sorted first (src/lssp.cxx:173)
    Ap, Aj, Ax = g.cd3d(8)
    Aj, Ax = Aj.copy(), Ax.copy()
    rng = np.random.default_rng(2)
    for i in range(0, 512, 3):
        p = rng.permutation(Ap[i + 1] - Ap[i])
        Aj[Ap[i]:Ap[i + 1]], Ax[Ap[i]:Ap[i + 1]] = Aj[Ap[i]:Ap[i + 1]][p], Ax[Ap[i]:Ap[i + 1]][p]
    assert same(api.bilu_factor((Ap, Aj, Ax), 128, 1), refb.bilu((Ap, Aj, Ax), 128, 1))


def test_block_tridiagonal_matrix_is_factored_exactly():
    # BILU(0) of a block-tridiagonal matrix is its exact block LU: U^-1 D L^-1 A = I
    rng = np.random.default_rng(1)
    nb, bs = 30, 4
    n = nb * bs
    M = np.zeros((n, n))
    for i in range(nb):
        for j in (i - 1, i, i + 1):
            if 0 <= j < nb:
                M[i * bs:(i + 1) * bs, j * bs:(j + 1) * bs] = rng.standard_normal((bs, bs)) + (9 * np.eye(bs) if i == j else 0)
    rows, cols = np.nonzero(M)
    Ap = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int32)
    A = (Ap, cols.astype(np.int32), M[rows, cols])
    L, D, U = api.bilu_factor(A, nb, 0)
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    csr = lambda T: sp.csr_matrix((T[2], T[1], T[0]), shape=(n, n))  # noqa: E731
    rhs = rng.standard_normal(n)
    x = sla.spsolve_triangular(csr(U), csr(D) @ sla.spsolve_triangular(csr(L), rhs, lower=True), lower=False)
    assert np.linalg.norm(M @ x - rhs) <= 1e-13 * np.linalg.norm(rhs)


def test_block_set_up_rejects_what_the_reference_cannot_factor():
    A = g.cd3d(6)   # n = 216
    with pytest.raises(Exception, match="multiple"):
        api.bilu_factor(A, 7, 0)                       # src/matrix-utils.cxx:72-74
    Ap, Aj, Ax = g.laplacian_5pt(8)
    keep = ~((np.repeat(np.arange(64), np.diff(Ap)) // 2 == 3) & (Aj // 2 == 3))   # remove diagonal block 3
    cnt = np.add.reduceat(keep.astype(np.int64), Ap[:-1])
    B = (np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32), Aj[keep], Ax[keep])
    with pytest.raises(Exception, match="diagonal block"):
        api.bilu_factor(B, 32, 0)
    Z = (Ap, Aj, np.where(Aj == np.repeat(np.arange(64), np.diff(Ap)), 0.0, Ax))      # zero diagonal, bs = 1
    with pytest.raises(Exception, match="singular"):
        api.bilu_factor(Z, 64, 0)                      # src/pc-biluk.cxx:262

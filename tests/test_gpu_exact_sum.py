"""LSSPG_OPT_REDUCE_SEQUENTIAL = 2 on the GPU: every dot product / norm equals the reference's sequential sum
(src/vector.cxx:127-131) bit for bit, computed in parallel (lssp_b200/csrc/exact_sum.cu; the algorithm and its host
replay are checked on the CPU in tests/test_exact_sum.py).  Consequence: whole solves are bit-identical to the
reference -- iteration counts, final residuals and residual histories EQUAL the golden values from the unmodified
reference, here at the small sizes of tests/golden/golden.json and in tests/test_gpu_baseline_sizes.py at 256^3."""
import numpy as np
import pytest

from lssp_b200 import api
from test_exact_sum import seq_sum
from test_gpu_solvers import CASES, GMRES_IDRS, key_of, run, run2
from util import tvec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xctx():
    c = api.Context(0)
    c.set_option(api.OPT_REDUCE_SEQUENTIAL, 2)
    c.set_option(api.OPT_SPMV_EXACT, 1)
    yield c
    c.close()


@pytest.fixture(scope="module")
def sctx():
    c = api.Context(0)
    c.set_option(api.OPT_REDUCE_SEQUENTIAL, 1)
    yield c
    c.close()


def vectors(kind, n):
    rng = np.random.default_rng(n % 1000 + len(kind))
    if kind == "mixed":
        return tvec(n), tvec(n, 1) + 0.5
    if kind == "norm":
        x = tvec(n, 2)
        return x, x
    if kind == "ones":
        return np.ones(n), np.ones(n)
    if kind == "walk":                       # partial sums wander through binade boundaries and zero; frequent ties
        return rng.standard_normal(n), np.ones(n)
    if kind == "lognormal":
        return rng.lognormal(0.0, 3.0, n), rng.lognormal(0.0, 3.0, n)
    if kind == "dyadic":                     # exact ties
        x = rng.integers(1, 1 << 20, n).astype(np.float64) * 2.0 ** -30
        x[0] = 2.0 ** 13
        return x, np.ones(n)
    raise KeyError(kind)


@pytest.mark.parametrize("kind", ["mixed", "norm", "ones", "walk", "lognormal", "dyadic"])
@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 4097, 70001, 1000003, 1 << 22])
def test_dot_equals_the_sequential_sum(xctx, kind, n):
    x, y = vectors(kind, n)
    got = api.lssp_vec_dot(xctx, xctx.upload(x), xctx.upload(y))
    want = seq_sum(x * y)
    assert np.float64(got).tobytes() == np.float64(want).tobytes(), (kind, n, got, want)


def test_dot_and_norm_at_the_baseline_size(xctx, sctx):
    n = 1 << 24
    x, y = tvec(n), tvec(n, 1) + 0.5
    dx, dy = xctx.upload(x), xctx.upload(y)
    for a, b, u, v in ((dx, dy, x, y), (dx, dx, x, x), (dy, dy, y, y)):
        assert api.lssp_vec_dot(xctx, a, b) == seq_sum(u * v)
    assert api.lssp_vec_norm(xctx, dx) == np.sqrt(seq_sum(x * x))
    # the one-thread adder (mode 1) gives the same number
    assert api.lssp_vec_dot(sctx, sctx.upload(x), sctx.upload(y)) == api.lssp_vec_dot(xctx, dx, dy)


def test_multidot_and_non_finite_terms(xctx):
    n = 200001
    vs = [tvec(n, k) for k in range(8)]
    y = tvec(n, 9)
    d = [xctx.upload(v) for v in vs]
    dy = xctx.upload(y)
    for k in (1, 2, 3, 5, 8):
        got = api.lssp_vec_multidot(xctx, d[:k], dy)
        for i in range(k):
            assert got[i] == seq_sum(vs[i] * y), (k, i)
    x = tvec(n)
    x[n // 2] = np.inf
    assert api.lssp_vec_dot(xctx, xctx.upload(x), dy) == seq_sum(x * y)
    x[n // 3] = np.nan
    assert np.isnan(api.lssp_vec_dot(xctx, xctx.upload(x), dy))


@pytest.mark.parametrize("m,s,pc,kw", CASES)
def test_solves_are_bit_identical_to_the_reference(xctx, golden, m, s, pc, kw):
    key = key_of(m, s, pc, kw)
    e, h = golden["solves"][key], golden["histories"][key]
    A, r = run(xctx, m, s, pc, kw)
    assert r["nits"] == e["nits"]
    assert r["residual"] == e["residual"]
    assert list(r["hist"][:len(h)]) == h
    assert abs(np.linalg.norm(r["x"]) - e["xnorm"]) <= 1e-14 * e["xnorm"]


@pytest.mark.parametrize("m,s,pc,kw", GMRES_IDRS)
def test_gmres_idrs_are_bit_identical_to_the_reference(xctx, golden, m, s, pc, kw):
    e = golden["solves"][key_of(m, s, pc, kw)]
    r = run2(xctx, m, s, pc, kw)
    assert r["nits"] == e["nits"] and r["residual"] == e["residual"]

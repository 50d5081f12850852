"""CPU suite, part 1: pin the oracle.

The plain-C restatement (oracle/oracle.c) is checked against the golden fixtures
generated from the UNMODIFIED reference (tests/golden/make_golden.py), against
the known-answer tables of SURVEY.md Appendix A, and -- when oracle/_ref is
present -- against the compiled reference itself, bit for bit.
"""
import numpy as np
import pytest

from util import MATRICES, matrix, sha, tvec


@pytest.mark.parametrize("name", list(MATRICES))
def test_port_kernels_match_golden(port, golden, name):
    A = matrix(name)
    n = len(A[0]) - 1
    e = golden["kernels"][name]
    assert e["matrix_sha"] == sha(*A), "generator drifted from the fixture"
    x, y = tvec(n), tvec(n, 1)
    assert sha(port.mv(0, A, x)) == e["mxy"]
    assert sha(port.mv(1, A, x, alpha=-1.75)) == e["amxy"]
    assert sha(port.mv(2, A, x, alpha=0.5, beta=-2.0, y=y)) == e["amxpby"]
    assert sha(port.mv(3, A, x, alpha=-1.0, beta=1.0, y=y)) == e["amxpbyz"]
    assert port.dot(x, y) == e["dot"]
    assert port.norm(x) == e["norm"]
    assert sha(port.axpby(1.25, x, -0.5, y)) == e["axpby"]
    assert sha(port.axpbyz(-3.0, x, 0.125, y)) == e["axpbyz"]


def test_known_answers_appendix_a2(port):
    # N = 100, v_i = sin(i)
    A = matrix("lap2d_100")
    v = np.sin(np.arange(10000, dtype=np.float64))
    Av = port.mv(0, A, v)
    assert abs(port.norm(Av) - 87.603662100238083) < 1e-12
    assert abs(port.dot(v, Av) - 6114.7266257793499) < 1e-9


@pytest.mark.parametrize("name", list(MATRICES))
def test_ref_kernels_match_golden(ref, golden, name):
    A = matrix(name)
    n = len(A[0]) - 1
    e = golden["kernels"][name]
    x, y = tvec(n), tvec(n, 1)
    assert sha(ref.mv(0, A, x)) == e["mxy"]
    assert sha(ref.mv(3, A, x, alpha=-1.0, beta=1.0, y=y)) == e["amxpbyz"]
    assert ref.dot(x, y) == e["dot"]


@pytest.mark.parametrize("name", ["lap2d_100", "cd3d_12", "powerlaw_4000"])
@pytest.mark.parametrize("tag,kw", [("iluk0", dict(kind="iluk", level=0)), ("iluk1", dict(kind="iluk", level=1)),
                                    ("ilut", dict(kind="ilut"))])
def test_port_trisolve_matches_golden(port, golden, name, tag, kw):
    # the factors come from the product's host set-up code, itself pinned in test_host.py
    from lssp_b200 import api
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, **kw)
    e = golden["factors"][name + "/" + tag]
    cache = port.tri_lower(L, tvec(n))
    assert sha(cache) == e["lower_sha"]
    assert sha(port.tri_upper(U, cache)) == e["apply_sha"]
    assert sha(port.ilu_apply(L, U, tvec(n))) == e["apply_sha"]


@pytest.mark.parametrize("key,solver,lu", [
    ("lap3d_32/cg/non", "cg", None), ("lap3d_32/cg/iluk_level0", "cg", dict(kind="iluk", level=0)),
    ("cd3d_32/bicgstab/non", "bicgstab", None), ("cd3d_32/bicgstab/iluk_level0", "bicgstab", dict(kind="iluk", level=0)),
    ("cd3d_32/bicgstab/iluk_level1", "bicgstab", dict(kind="iluk", level=1)),
    ("cd3d_32/bicgstab/ilut", "bicgstab", dict(kind="ilut"))])
def test_port_drivers_match_golden(port, golden, key, solver, lu):
    from lssp_b200 import api
    A = matrix(key.split("/")[0])
    n = len(A[0]) - 1
    LU = api.ilu_factor(A, **lu) if lu else None
    r = port.solve(solver, A, np.ones(n), LU=LU, maxit=3000, nhist=20)
    e = golden["solves"][key]
    assert r["nits"] == e["nits"]
    assert r["residual"] == e["residual"]
    h = golden["histories"][key]
    assert list(r["hist"][:len(h)]) == h      # full-precision, bit for bit


def test_port_bilu_apply(port):
    # block-ILU apply = lower sweep, D SpMV, upper sweep (src/pc-biluk.cxx:22-60) on synthetic factors
    from lssp_b200 import api
    A = matrix("cd3d_12")
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, "iluk", level=0)
    D = matrix("cd3d_12")
    rhs = tvec(n)
    y = port.tri_lower(L, rhs)
    z = port.mv(0, D, y)
    x = port.tri_upper(U, z)
    assert np.array_equal(port.bilu_apply(L, D, U, rhs), x)

"""CPU checks of the pencil schedule (lssp_b200/csrc/tri_pencil.cu): the image the device reads -- value stream, line
descriptors, ghost lanes, mailboxes -- is replayed on the host (lsspg_debug_tri_walk_pencil_host) and must reproduce the
serial sweeps of the reference (src/solver-tri.cxx:4-46) bit for bit.  No GPU involved."""
import os

import numpy as np
import pytest

from lssp_b200 import api
from lssp_b200 import generators as g
from util import matrix, tvec


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


CASES = [
    ("lap3d_32", 0, {}), ("lap3d_32", 1, {}), ("cd3d_32", 1, {}), ("lap2d_100", 0, {}), ("lap2d_100", 1, {}),
    ("lap2d_100", 2, {}), ("lap3d_32", 0, {"LSSPG_TRI_PENCIL": "16,16"}), ("cd3d_32", 1, {"LSSPG_TRI_PENCIL": "16,8"}),
    ("cd3d_32", 1, {"LSSPG_TRI_PENCIL": "8,4"}), ("lap2d_100", 1, {"LSSPG_TRI_PENCIL": "64,1"}),
]


@pytest.mark.parametrize("name,level,env", CASES)
def test_pencil_image_replays_the_serial_sweeps(checker, name, level, env):
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, "iluk", level=level)
    rhs = tvec(n, 5)
    for which, T, serial in ((0, L, checker.tri_lower), (1, U, checker.tri_upper)):
        x, info = _with_env(env, lambda: api.tri_walk_pencil_host(which, T, rhs))
        assert x is not None, "a stencil factor must get the pencil schedule"
        assert np.array_equal(x, serial(T, rhs))
        assert info["slots"] == max(np.diff(T[0])) - 1
        assert info["values_per_row"] == info["slots"] + (0 if which == 0 else 1)   # ILU(k): unit diagonal in L only


def test_pencil_rejects_what_is_not_a_lattice(checker):
    for A in (matrix("powerlaw_4000"), matrix("random_600")):
        L, U = api.ilu_factor(A, "iluk", level=0)
        n = len(A[0]) - 1
        assert api.tri_walk_pencil_host(0, L, tvec(n))[0] is None
    # a stencil factor with ONE wrap-around entry (column i - nx + 1 in a row with x = nx - 1 would be (0, y):
    # not the lattice neighbour the offset stands for) must be rejected, not silently mis-scheduled
    A = matrix("lap3d_32")
    L, _ = api.ilu_factor(A, "iluk", level=1)
    Lp, Lj, Lx = (np.array(a) for a in L)
    n = len(Lp) - 1
    i = 32 * 32 * 5 + 32 * 7 + 31            # x = 31: the offset -(nx - 1) cannot be a lattice neighbour here
    row = Lj[Lp[i]:Lp[i + 1]]
    assert (i - 31) not in row
    k = Lp[i] + int(np.searchsorted(row, i - 31))
    Lj2 = np.insert(Lj, k, i - 31)
    Lx2 = np.insert(Lx, k, 0.125)
    Lp2 = Lp.copy()
    Lp2[i + 1:] += 1
    bad = (Lp2.astype(np.int32), Lj2.astype(np.int32), Lx2)
    assert api.tri_walk_pencil_host(0, bad, tvec(n))[0] is None


def test_pencil_handles_non_unit_lower_diagonal_and_signed_zeros(checker):
    # a generic lower factor (lssp_pc_ilu_solve_lower_matrix accepts any diagonal) and right-hand sides holding -0.0:
    # the +0.0 * +0.0 padding of missing stencil entries must not flip the sign of a zero result
    A = matrix("lap3d_32")
    L, U = api.ilu_factor(A, "iluk", level=0)
    n = len(A[0]) - 1
    Lx = np.array(L[2])
    Lx[np.array(L[0][1:]) - 1] = 1.5 + 0.25 * np.cos(np.arange(n))     # diagonal is the last entry of every row
    L2 = (L[0], L[1], Lx)
    rhs = tvec(n, 2)
    rhs[::7] = -0.0
    rhs[:40] = -0.0
    x, info = api.tri_walk_pencil_host(0, L2, rhs)
    want = checker.tri_lower(L2, rhs)
    assert info["values_per_row"] == info["slots"] + 1
    assert np.array_equal(x, want) and np.array_equal(np.signbit(x), np.signbit(want))

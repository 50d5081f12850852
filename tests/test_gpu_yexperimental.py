"""GPU checks of the EXPERIMENTAL sweep variants that have only been verified on the host so far (ROADMAP.md).  They
are skipped unless LSSPG_TEST_EXPERIMENTAL=1, so that the regular `-m gpu` run only exercises verified code:

    LSSPG_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_yexperimental.py -q
"""
import os

import numpy as np
import pytest

from lssp_b200 import api
from util import matrix, tvec

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("LSSPG_TEST_EXPERIMENTAL") != "1", reason="experimental kernels: opt-in")]


def _apply_with(ctx, env, L, U, rhs):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        pc = api.Preconditioner(ctx, "ilu", len(rhs), L, U)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    x = pc.apply_host(rhs)
    y = pc.apply_host(rhs)       # a second application reuses counters / epochs
    pc.free()
    return x, y


@pytest.mark.parametrize("name,level", [("lap3d_32", 0), ("cd3d_32", 1), ("lap2d_100", 0), ("lap2d_100", 1)])
@pytest.mark.parametrize("env", [{"LSSPG_TRI_CHUNKS": "2"}, {"LSSPG_TRI_CHUNKS": "3"}, {"LSSPG_TRI_CHUNKS": "6"},
                                 {"LSSPG_TRI_SKEW_FORCE": "1,1,1"}, {"LSSPG_TRI_SKEW_FORCE": "1,1,1", "LSSPG_TRI_CHUNKS": "3"}])
def test_experimental_box_schedules_are_bit_exact(ctx, checker, name, level, env):
    A = matrix(name)
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, "iluk", level=level)
    rhs = tvec(n, 3)
    want = checker.tri_upper(U, checker.tri_lower(L, rhs))
    x, y = _apply_with(ctx, env, L, U, rhs)
    assert np.array_equal(x, want) and np.array_equal(y, want)


def test_skewed_boxes_for_wide_rows_are_bit_exact(ctx, checker):
    # ILU(2): more than 6 off-diagonals per row -- skewed boxes only with LSSPG_TRI_SKEW=2 until this has passed
    A = matrix("cd3d_32")
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, "iluk", level=2)
    rhs = tvec(n, 3)
    want = checker.tri_upper(U, checker.tri_lower(L, rhs))
    x, y = _apply_with(ctx, {"LSSPG_TRI_SKEW": "2"}, L, U, rhs)
    assert np.array_equal(x, want) and np.array_equal(y, want)

"""AMG set-up with its per-row phases on the GPU (lssp_b200/csrc/amg_gpu.cu): strong couplings, direct
interpolation, restriction and the Galerkin products run as count / scan / fill kernels over the row functions of
amg_rows.cuh (pinned on the CPU by tests/test_amg_rows.py); the Ruge-Stueben C/F splitting stays on the host.  The
hierarchy must equal the host set-up's array by array -- and then drives the same V-cycle."""
import time

import numpy as np
import pytest

from lssp_b200 import api
from lssp_b200 import generators as g
from test_amg_rows import same_hierarchy
from util import matrix

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cd3d_12", "lap3d_32", "lap2d_100", "random_600", "powerlaw_4000"])
def test_device_set_up_reproduces_the_host_hierarchy(ctx, name):
    A = matrix(name)
    same_hierarchy(api.AmgHierarchy(A, ctx=ctx), api.AmgHierarchy(A))


def test_other_parameters_and_a_solve(ctx):
    A = matrix("cd3d_12")
    for pars in (dict(strong_threshold=0.5, trunc_threshold=0.0), dict(cf_order=2, coarse_dof=10, max_levels=4)):
        same_hierarchy(api.AmgHierarchy(A, ctx=ctx, **pars), api.AmgHierarchy(A, **pars))
    A = matrix("lap3d_32")
    n = len(A[0]) - 1
    H = api.AmgHierarchy(A, ctx=ctx, zero_guess=1)
    pc = api.Preconditioner.sxamg(ctx, A, hierarchy=H)
    ref = api.Preconditioner.sxamg(ctx, A, zero_guess=1)
    dA = api.Csr(ctx, A)
    r1 = api.lssp_solver_solve(ctx, "cg", dA, pc, np.ones(n), np.zeros(n), maxit=100)
    r2 = api.lssp_solver_solve(ctx, "cg", dA, ref, np.ones(n), np.zeros(n), maxit=100)
    assert r1["nits"] == r2["nits"] and r1["residual"] == r2["residual"]


def test_set_up_time_next_to_the_host(ctx, record_property):
    A = g.lap3d(64)
    t0 = time.perf_counter()
    Hh = api.AmgHierarchy(A)
    t1 = time.perf_counter()
    Hd = api.AmgHierarchy(A, ctx=ctx)
    t2 = time.perf_counter()
    print("AMG set-up at 64^3: host %.3f s, per-row phases on the device %.3f s (C/F splitting on the host in both)" % (t1 - t0, t2 - t1))
    record_property("host_s", t1 - t0)
    record_property("device_s", t2 - t1)
    same_hierarchy(Hd, Hh)

"""GPU suite, part 5: block ILU(k) (LSSP_PC_BILUK) -- the host set-up's factors (tests/test_biluk.py) through
lsspg_pc_create_bilu and the Krylov drivers, against fixtures from the unmodified reference sources compiled with the
netlib reference dense kernels (tests/golden/make_biluk_golden.py)."""
import json
import os

import numpy as np
import pytest

from lssp_b200 import api
from util import matrix, sha, tvec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "biluk_golden.json")) as f:
    GOLD = json.load(f)


def _case(key):
    name, bs, k = key.split("/")[:3]
    return name, int(bs[2:]), int(k[1:])


# ---- GPU: the factors through lsspg_pc_create_bilu and the drivers ----------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(GOLD["factors"]))
def test_gpu_block_ilu_application_equals_the_reference(ctx, key):
    name, bs, level = _case(key)
    A = matrix(name)
    n = len(A[0]) - 1
    pc = api.Preconditioner.biluk(ctx, A, n // bs, level=level)
    assert sha(pc.apply_host(tvec(n, 2))) == GOLD["factors"][key]["apply_sha"]
    pc.free()


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(GOLD["solves"]))
def test_gpu_drivers_with_block_ilu_equal_the_reference_in_sequential_mode(key):
    name, bs, level = _case(key)
    solver = key.split("/")[3]
    e = GOLD["solves"][key]
    c = api.Context(0)
    c.set_option(api.OPT_REDUCE_SEQUENTIAL, 1)
    A = matrix(name)
    n = len(A[0]) - 1
    dA, pc = api.Csr(c, A), api.Preconditioner.biluk(c, A, n // bs, level=level)
    r = api.lssp_solver_solve(c, solver, dA, pc, np.ones(n), np.zeros(n), maxit=3000, restart=30)
    assert r["nits"] == e["nits"], (r["nits"], e["nits"])
    assert r["residual"] == e["residual"], (r["residual"], e["residual"])
    assert abs(np.linalg.norm(r["x"]) - e["xnorm"]) <= 1e-13 * e["xnorm"]
    pc.free()
    dA.free()
    c.close()


@pytest.mark.gpu
def test_exam_program_with_the_block_ilu_preconditioner():
    """LSSP_PC_BILUK through the C++ API (lssp_solver_create / s.num_blks / lssp_solver_assemble): the exam.cxx
    matrix with 2 x 2 blocks, level 1 -- the reference needs 32 CG iterations (fixture lap2d_100/bs2/k1/cg)."""
    import re
    import subprocess
    exam = os.path.join(ROOT, "examples", "exam")
    e = GOLD["solves"]["lap2d_100/bs2/k1/cg"]
    out = subprocess.run([exam, "100", "cg", "biluk", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"iterations: (\d+), solver residual: (\S+)", out.stdout)
    k = re.search(r"solution L2 norm: (\S+) residual: (\S+)", out.stdout)
    assert abs(int(m.group(1)) - e["nits"]) <= 1, out.stdout
    assert abs(float(k.group(1)) - e["xnorm"]) <= 1e-6 * e["xnorm"]
    assert abs(float(k.group(2)) - float(m.group(2))) <= 1e-3 * float(m.group(2)) + 1e-9

"""CPU replay of the device factorisations (lssp_b200/csrc/ilu_rows.cuh: the row recurrences the kernels of ilu_gpu.cu
run -- ILU(k) symbolic with the level-raising rule, numeric IKJ row, ILUT with the per-row column map and the reference's
quick-select order, L / U split).  Rows in ascending order on the host, no device: the factors must equal the host
set-up's (lsspg_ilu_factor), which is pinned against the unmodified reference (tests/test_host.py) and its golden factor
hashes (tests/golden/golden.json).  On the GPU the same functions run behind row waits (tests/test_gpu_setup.py)."""
import ctypes as C

import numpy as np
import pytest

from lssp_b200 import _lib, api
from util import matrix, sha


def replay(A, kind, level=0, p=-1, tol=1e-3, blk_size=0):
    L = _lib.lib()
    Ap, Aj, Ax = (np.ascontiguousarray(A[0], np.int32), np.ascontiguousarray(A[1], np.int32), np.ascontiguousarray(A[2], np.float64))
    n = len(Ap) - 1
    h, ok = C.c_void_p(), C.c_int()
    rc = L.lsspg_debug_ilu_gpu_replay_host(0 if kind == "iluk" else 1, n, Ap.ctypes.data_as(C.c_void_p), Aj.ctypes.data_as(C.c_void_p),
                                           Ax.ctypes.data_as(C.c_void_p), int(level), int(p), C.c_double(tol), int(blk_size),
                                           C.byref(ok), C.byref(h))
    _lib.check(rc)
    if not ok.value:
        return None
    return api._factors_out(h, n)


def same(F, G):
    return all(np.array_equal(a, b, equal_nan=(a.dtype == np.float64)) and a.dtype == b.dtype for X, Y in zip(F, G) for a, b in zip(X, Y))


@pytest.mark.parametrize("name", ["lap2d_100", "lap3d_32", "cd3d_32", "cd3d_12", "powerlaw_4000", "random_600"])
@pytest.mark.parametrize("level", [0, 1, 2, 3])
def test_iluk_row_functions_reproduce_the_host_factors(name, level):
    if name == "powerlaw_4000" and level > 2:
        pytest.skip("near-dense factors: minutes of sorted insertion")
    A = matrix(name)
    assert same(replay(A, "iluk", level=level), api.ilu_factor(A, "iluk", level=level)), (name, level)


@pytest.mark.parametrize("name", ["cd3d_12", "powerlaw_4000", "random_600"])
def test_iluk_row_functions_on_diagonal_blocks(name):
    A = matrix(name)
    n = len(A[0]) - 1
    for level, bs in ((0, (n + 3) // 4), (1, (n + 2) // 3), (1, 97), (2, 13)):
        assert same(replay(A, "iluk", level=level, blk_size=bs), api.ilu_factor(A, "iluk", level=level, blk_size=bs)), (name, level, bs)


@pytest.mark.parametrize("name", ["lap2d_100", "lap3d_32", "cd3d_32", "cd3d_12", "powerlaw_4000", "random_600"])
def test_ilut_row_function_reproduces_the_host_factors_and_their_stored_order(name):
    A = matrix(name)
    n = len(A[0]) - 1
    cases = [dict(), dict(p=9, tol=1e-4), dict(p=3, tol=1e-2), dict(p=4, tol=1e-2, blk_size=(n + 4) // 5), dict(p=1, tol=0.5)]
    if n <= 2000:
        cases.append(dict(p=50, tol=0.0))
    for kw in cases:
        assert same(replay(A, "ilut", **kw), api.ilu_factor(A, "ilut", **kw)), (name, kw)


def test_golden_factor_hashes(golden):
    """the replayed factors against the hashes taken from the unmodified reference"""
    for name in ("lap3d_32", "cd3d_32", "lap2d_100"):
        A = matrix(name)
        for tag, kw in (("iluk0", dict(kind="iluk", level=0)), ("iluk1", dict(kind="iluk", level=1)), ("ilut", dict(kind="ilut"))):
            e = golden["factors"].get(name + "/" + tag)
            if e is None:
                continue
            Lf, Uf = replay(A, kw["kind"], level=kw.get("level", 0))
            assert sha(Lf[0], Lf[1], Lf[2]) == e["L_sha"] and sha(Uf[0], Uf[1], Uf[2]) == e["U_sha"], (name, tag)


def test_unsorted_input_is_left_to_the_device_ingest():
    Ap = np.array([0, 2, 4], np.int32)
    Aj = np.array([1, 0, 0, 1], np.int32)
    assert replay((Ap, Aj, np.ones(4)), "iluk") is None


def test_fuzz_random_matrices_weak_pivots_blocks():
    """random patterns, repaired pivots (|d| < 1e-10 -> +-1e-3, src/pc-iluk.cxx:367-375), blocks, all levels / thresholds"""
    from lssp_b200 import generators as g
    rng = np.random.default_rng(31)
    for trial in range(80):
        n = int(rng.integers(5, 300))
        Ap, Aj, Ax = g.random_csr(n, avg=int(rng.integers(2, 9)), seed=int(rng.integers(0, 1 << 30)))
        if trial % 3 == 0:
            Ax = Ax.copy()
            for i in rng.integers(0, n, 3):
                k = Ap[i] + int(np.where(Aj[Ap[i]:Ap[i + 1]] == i)[0][0])
                Ax[k] = rng.choice([0.0, 1e-12, -1e-12, 1e-3])
        A = (Ap, Aj, Ax)
        lvl, bs = int(rng.integers(0, 4)), int(rng.choice([0, max(1, n // 3), 7]))
        assert same(replay(A, "iluk", level=lvl, blk_size=bs), api.ilu_factor(A, "iluk", level=lvl, blk_size=bs)), (trial, n, lvl, bs)
        p, tol = int(rng.choice([-1, 1, 2, 5, 20])), float(rng.choice([1e-3, 0.0, 1e-1, 1e-6]))
        assert same(replay(A, "ilut", p=p, tol=tol, blk_size=bs), api.ilu_factor(A, "ilut", p=p, tol=tol, blk_size=bs)), (trial, n, p, tol, bs)

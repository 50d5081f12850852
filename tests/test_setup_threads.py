"""The threaded host set-up (ilu_host.cpp, tri.cu, tri_tiled.cu over host_par.h) must produce exactly the bytes the
serial code produced: factors equal to the reference's, schedule images equal to the fingerprints pinned in
tests/golden/pack_golden.json (generated with LSSPG_HOST_THREADS=1 before the set-up was threaded), for any
number of host threads.  No GPU involved: lsspg_debug_tri_pack_host builds the image lsspg_tri_analyse uploads."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from lssp_b200 import api, generators as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_pack_golden as mpg  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "pack_golden.json")) as f:
    PACK = json.load(f)


fingerprints = mpg.fingerprints


@pytest.mark.parametrize("name", list(mpg.CASES))
def test_schedule_images_match_the_pinned_serial_ones(name):
    for key, got in fingerprints(name).items():
        assert got == PACK[key], key


@pytest.mark.parametrize("threads,chunk", [(1, 0), (3, 0), (7, 5), (4, 1000), (8, 1)])
def test_schedule_images_do_not_depend_on_the_thread_count(threads, chunk):
    # the thread count (and the RowPipeline chunk length) is read once per process: run the comparison in a child
    code = ("import json, sys; sys.path[:0] = [%r, %r]; import test_setup_threads as t; "
            "print(json.dumps({k: v for n in ('lap3d_48/iluk0_bj3', 'cd3d_40/ilut', 'cd3d_32/iluk1', 'cd3d_32/iluk1_slices', 'lap2d_300/iluk0') "
            "for k, v in t.fingerprints(n).items()}))" % (ROOT, os.path.join(ROOT, "tests")))
    env = dict(os.environ, LSSPG_HOST_THREADS=str(threads))
    if chunk:
        env["LSSPG_PIPE_CHUNK"] = str(chunk)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    assert got and all(PACK[k] == v for k, v in got.items())


def _same(F, G):
    return all(np.array_equal(a, b) for X, Y in zip(F, G) for a, b in zip(X, Y))


def test_threaded_factorisation_equals_the_reference_at_sizes_that_use_threads(ref):
    # n = 110 592 rows: every threaded loop of the set-up really runs on several threads
    A = g.cd3d(48)
    n = len(A[0]) - 1
    for kw in (dict(kind="iluk", level=0), dict(kind="iluk", level=1), dict(kind="iluk", level=2, blk_size=(n + 2) // 3),
               dict(kind="ilut", p=5, tol=1e-3, blk_size=(n + 1) // 2)):
        assert _same(api.ilu_factor(A, **kw), ref.ilu(A, **kw)), kw


def test_unsorted_rows_and_missing_diagonals_take_the_repair_path(ref):
    Ap, Aj, Ax = g.cd3d(40)
    Aj, Ax = Aj.copy(), Ax.copy()
    n = len(Ap) - 1
    rng = np.random.default_rng(5)
    for i in rng.choice(n, 2000, replace=False):      # shuffle some rows: sorted by the set-up (src/lssp.cxx:173)
        b, e = Ap[i], Ap[i + 1]
        p = rng.permutation(e - b)
        Aj[b:e], Ax[b:e] = Aj[b:e][p], Ax[b:e][p]
    B = (Ap, Aj, Ax)
    for kw in (dict(kind="iluk", level=0), dict(kind="iluk", level=1, blk_size=(n + 3) // 4)):
        assert _same(api.ilu_factor(B, **kw), ref.ilu(B, **kw)), kw
    # Missing diagonals: src/matrix-utils.cxx:483-587 inserts (i, tol = 1e-10) in sorted position.  The reference
    # leaves num_nnzs of the repaired matrix stale (:485 copies the struct, the count is never updated), so its own
    # factorisation of such a matrix reads short / crashes -- the pin is the documented semantics instead: the
    # factors equal those of the matrix with the entries written out.
    Ap, Aj, Ax = g.cd3d(40)
    drop = np.zeros(len(Aj), bool)
    rows = np.sort(rng.choice(n, 500, replace=False))
    for i in rows:
        b, e = Ap[i], Ap[i + 1]
        drop[b:e] |= Aj[b:e] == i
    cnt = np.add.reduceat((~drop).astype(np.int64), Ap[:-1])
    Bp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    B = (Bp, Aj[~drop], Ax[~drop])
    Cx = Ax.copy()
    Cx[drop] = 1e-10
    for kw in (dict(kind="iluk", level=0), dict(kind="iluk", level=1), dict(kind="ilut", p=5, tol=1e-3)):
        assert _same(api.ilu_factor(B, **kw), api.ilu_factor((Ap, Aj, Cx), **kw)), kw


@pytest.mark.parametrize("case", ["cd3d_32/iluk1", "cd3d_24/iluk2", "lap2d_150/iluk1", "cd3d_40/iluk1_bj2", "grid_20x48x12/iluk1",
                                  "grid_12x20x48/iluk2"])
def test_skewed_boxes_give_fill_factors_an_acyclic_box_schedule(case, port, monkeypatch):
    # ILU(1)/(2) fill couples axis-aligned neighbour boxes both ways (cyclic box graph -> slice schedule).  Boxes cut
    # along x + s1 y + t1 z, y + s2 z, z (the default; LSSPG_TRI_SKEW=0 disables, tri_tiled.cu) are coupled one way only: the factor gets
    # the completion-flag box schedule (kind 2), and walking it box by box reproduces the serial sweeps bit for bit.
    A, kw = {"cd3d_32/iluk1": (g.cd3d(32), dict(kind="iluk", level=1)),
             "cd3d_24/iluk2": (g.cd3d(24), dict(kind="iluk", level=2)),
             "lap2d_150/iluk1": (g.laplacian_5pt(150), dict(kind="iluk", level=1)),
             "cd3d_40/iluk1_bj2": (g.cd3d(40), dict(kind="iluk", level=1, blk_size=32000)),
             "grid_20x48x12/iluk1": (g.stencil_7pt_rows((20, 48, 12), 0, 11520, conv=(0.3, 0.2, 0.1)), dict(kind="iluk", level=1)),
             "grid_12x20x48/iluk2": (g.stencil_7pt_rows((12, 20, 48), 0, 11520, conv=(0.3, 0.2, 0.1)), dict(kind="iluk", level=2))}[case]
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, **kw)
    rhs = np.sin(np.arange(n) * 0.37) + 0.3
    monkeypatch.setenv("LSSPG_TRI_SKEW", "0")
    assert api.tri_pack_host(0, L)["kind"] == 0 and api.tri_walk_tiled_host(0, L, rhs)[1] is None
    monkeypatch.delenv("LSSPG_TRI_SKEW")
    # (rows wider than 6 entries, ILU(2), take skewed boxes too since that path ran bit-exact on a B200 in round 2)
    assert api.tri_pack_host(0, L)["kind"] == 2 and api.tri_pack_host(1, U)["kind"] == 2
    y, info = api.tri_walk_tiled_host(0, L, rhs)
    x, info_u = api.tri_walk_tiled_host(1, U, y)
    assert info["max_box_rows"] <= 512 and info["box_levels"] < info["row_levels"] / 3
    assert info["nx"] * info["ny"] * info["nz"] == n
    assert np.array_equal(x, port.ilu_apply(L, U, rhs))
    # the plain box grid of ILU(0) factors is the zero-skew case: same image as pinned
    L0, U0 = api.ilu_factor(g.lap3d(32), kind="iluk", level=0)
    r = api.tri_pack_host(0, L0)
    assert "%016x" % r["fingerprint"] == PACK["lap3d_32/iluk0/L"]["fingerprint"]


@pytest.mark.parametrize("case", ["lap3d_32/iluk0", "lap2d_300/iluk0", "cd3d_32/iluk1", "lap3d_48/iluk0_bj3"])
def test_packed_box_blobs_replay_to_the_serial_sweeps(case, port):
    """lsspg_debug_tri_walk_packed_host replays the BYTES the box kernel reads (blob sections, ELL columns, operand lists,
    gate rows), boxes advancing round-robin: the sweeps must be reproduced bit for bit, one hand-off per box level."""
    make, kw = mpg.CASES[case][:2]
    A = make()
    n = len(A[0]) - 1
    L, U = api.ilu_factor(A, **kw)
    rhs = np.sin(np.arange(n) * 0.37) + 0.3
    want = port.ilu_apply(L, U, rhs)
    y, info = api.tri_walk_packed_host(0, L, rhs)
    x, info_u = api.tri_walk_packed_host(1, U, y)
    assert np.array_equal(x, want)
    assert info["chunks"] == 1 and info["rounds"] == info["box_levels"]
    r = api.tri_pack_host(0, L)                                                # and the default image is untouched
    assert "%016x" % r["fingerprint"] == PACK[case + "/L"]["fingerprint"]

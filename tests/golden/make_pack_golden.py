"""Pins the device images of the triangular-factor schedules (tests/golden/pack_golden.json).

Generated with the serial set-up code (LSSPG_HOST_THREADS=1) before the set-up was threaded; the CPU test-suite
(tests/test_setup_threads.py) checks that the threaded code uploads exactly the same bytes.  The cd3d_32/iluk1
entries were regenerated when skewed boxes became the default for fill factors (their slice-schedule image stays
pinned as cd3d_32/iluk1_slices).

    LSSPG_HOST_THREADS=1 python tests/golden/make_pack_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from lssp_b200 import api, generators as g  # noqa: E402

CASES = {
    "lap3d_32/iluk0": (lambda: g.lap3d(32), dict(kind="iluk", level=0)),
    "lap3d_48/iluk0": (lambda: g.lap3d(48), dict(kind="iluk", level=0)),
    "lap3d_48/iluk0_bj3": (lambda: g.lap3d(48), dict(kind="iluk", level=0, blk_size=36864)),
    "cd3d_32/iluk1": (lambda: g.cd3d(32), dict(kind="iluk", level=1)),
    "cd3d_32/iluk1_slices": (lambda: g.cd3d(32), dict(kind="iluk", level=1), {"LSSPG_TRI_SKEW": "0"}),
    "cd3d_40/ilut": (lambda: g.cd3d(40), dict(kind="ilut", p=7, tol=1e-3)),
    "lap2d_300/iluk0": (lambda: g.laplacian_5pt(300), dict(kind="iluk", level=0)),
    "powerlaw_60000/iluk0": (lambda: g.powerlaw(60000, window=3000), dict(kind="iluk", level=0)),
}

def fingerprints(name):
    """{case/L, case/U: dict(kind, bytes, fingerprint)} of one case (its environment switches set for the call)"""
    make, kw = CASES[name][:2]
    env = CASES[name][2] if len(CASES[name]) > 2 else {}
    L, U = api.ilu_factor(make(), **kw)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        out = {}
        for which, T, tag in ((0, L, "L"), (1, U, "U")):
            r = api.tri_pack_host(which, T)
            out[name + "/" + tag] = dict(kind=r["kind"], bytes=r["bytes"], fingerprint="%016x" % r["fingerprint"])
        return out
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


if __name__ == "__main__":
    out = {}
    for name in CASES:
        for key, val in fingerprints(name).items():
            out[key] = val
            print(key, val)
    with open(os.path.join(HERE, "pack_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)

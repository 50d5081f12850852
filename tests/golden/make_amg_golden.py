#!/usr/bin/env python
"""Fixtures for the SX-AMG-style path (tests/golden/amg_golden.json).

libsxamg is not in the reference tree, so these vectors do NOT come from the reference: they pin
the specification of DESIGN.md "AMG" as implemented today -- the host set-up (hierarchy shapes and
SHA-256 of every level's arrays) and the restated cycle of oracle/amg_oracle.c (first entries, norm
and SHA-256 of one V-cycle, cycle counts of the stand-alone iteration, PCG counts).  PARITY WITH
libsxamg UNPINNED.  Regenerate only when the specification changes:  python tests/golden/make_amg_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from lssp_b200 import api  # noqa: E402  (host set-up only)
from util import matrix, tvec  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    port = oracle.Port()
    out = {}
    for name in ("lap2d_100", "lap3d_32", "cd3d_12"):
        for order in (1, 2):
            H = api.AmgHierarchy(matrix(name), cf_order=order)
            n = H.levels[0]["n"]
            m = port.amg(H.levels, coarse_inv=H.coarse_inv, cf_order=order)
            mz = port.amg(H.levels, coarse_inv=H.coarse_inv, cf_order=order, zero_guess=1)
            y = m.cycle(tvec(n), tvec(n, 3))
            sa = m.solve(np.ones(n), tol=1e-8, maxit=50)
            e = {"levels": [[L["n"], L["nc"], int(L["A"][0][-1])] for L in H.levels],
                 "level_sha": [sha(L["A"][0], L["A"][1], L["A"][2], L["cf"], L["rank"],
                                   *([] if L["P"] is None else [L["P"][0], L["P"][1], L["P"][2]])) for L in H.levels],
                 "coarse_inv_sha": sha(H.coarse_inv) if H.coarse_dense else None,
                 "cycle_sha": sha(y), "cycle_head": [float(v) for v in y[:4]], "cycle_norm": float(np.linalg.norm(y)),
                 "standalone_nits": sa["nits"], "standalone_residual": sa["residual"]}
            if name != "cd3d_12":
                r = port.solve("cg", matrix(name), np.ones(n), amg=mz, maxit=100)
                e["pcg_zero_guess_nits"], e["pcg_zero_guess_residual"] = r["nits"], r["residual"]
            out["%s/cf%d" % (name, order)] = e
    with open(os.path.join(ROOT, "tests", "golden", "amg_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()

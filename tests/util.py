import hashlib

import numpy as np

from lssp_b200 import generators as g

MATRICES = {
    "lap2d_100": lambda: g.laplacian_5pt(100),
    "lap3d_32": lambda: g.lap3d(32),
    "cd3d_32": lambda: g.cd3d(32),
    "cd3d_12": lambda: g.cd3d(12),
    "powerlaw_4000": lambda: g.powerlaw(4000, window=300),
    "random_600": lambda: g.random_csr(600, 5, seed=7),
}
_cache = {}


def matrix(name):
    if name not in _cache:
        _cache[name] = MATRICES[name]()
    return _cache[name]


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def tvec(n, k=0):
    i = np.arange(n, dtype=np.float64)
    return np.sin(i * (0.37 + 0.11 * k)) + 0.25 * np.cos(i * 1.3 + k)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    s = np.linalg.norm(b)
    return d / s if s > 0 else d

"""Host replay of the parallel reference-order summation (lssp_b200/csrc/exact_sum.cu, LSSPG_OPT_REDUCE_SEQUENTIAL = 2).

The reference adds the terms of a dot product one after the other (src/vector.cxx:129); the kernels reproduce that
result bit for bit without the n-step chain.  Here the SAME block functions the kernels use (exact_sum.cuh) are replayed
on the CPU and compared with a strictly sequential sum (np.add.accumulate adds in index order) on inputs chosen to hit
every branch: binade crossings, exact ties, sign changes, cancellation, non-finite terms, ragged tails."""
import ctypes as C

import numpy as np
import pytest

from lssp_b200 import _lib


def seq_sum(t):
    """s = 0; for i: s += t[i]  (IEEE double, round to nearest even)"""
    t = np.ascontiguousarray(t, dtype=np.float64)
    if t.size == 0:
        return 0.0
    with np.errstate(all="ignore"):
        return float(np.add.accumulate(np.concatenate([[0.0], t]))[-1])   # the sum starts at +0.0 (src/vector.cxx:127)


def exact(t):
    L = _lib.lib()
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = C.c_double(0.0)
    stats = (C.c_longlong * 4)()
    rc = L.lsspg_debug_exact_seq_sum_host(C.c_longlong(t.size), t.ctypes.data_as(C.POINTER(C.c_double)), C.byref(out), stats)
    assert rc == 0
    return out.value, tuple(stats)


def same(a, b):
    return np.float64(a).tobytes() == np.float64(b).tobytes() or (np.isnan(a) and np.isnan(b))


def check(t, what):
    got, stats = exact(t)
    want = seq_sum(t)
    assert same(got, want), (what, got, want, stats)
    return stats


@pytest.mark.parametrize("n", [0, 1, 2, 31, 255, 256, 257, 511, 512, 513, 1000, 4097, 65536, 262144 + 17])
def test_sizes_positive_and_mixed(n):
    rng = np.random.default_rng(n + 1)
    check(rng.random(n), "uniform")
    check(rng.standard_normal(n), "normal")
    check(-rng.random(n), "negative")
    check(rng.lognormal(0.0, 3.0, n), "lognormal")
    check(rng.standard_normal(n) * rng.lognormal(0.0, 6.0, n), "heavy tails, mixed sign")


def test_products_as_the_drivers_park_them():
    # terms of r.r, r.z, p.Ap-like sums: products rounded once, then added
    rng = np.random.default_rng(7)
    n = 1 << 20
    x = np.sin(np.arange(n) * 0.37) + 0.25
    y = rng.standard_normal(n)
    for t in (x * x, y * y, np.ones(n), np.full(n, 0.1), np.arange(n, dtype=np.float64)):
        stats = check(t, "products")
        assert stats[2] < 100 and stats[3] < 200, stats   # a few blocks per binade crossing, nothing else
    # mixed signs: the partial sums wander like a random walk, close to binade boundaries and with terms only ~2^8 below
    # the sum (a tie every ~256 terms): many more blocks are taken apart -- slower, never different
    stats = check(x * y, "mixed signs")
    assert stats[3] < stats[0] * 8 // 2, stats


def test_most_blocks_advance_as_one_integer_addition():
    rng = np.random.default_rng(3)
    n = 1 << 22
    stats = check(rng.random(n), "uniform 4M")
    nb, rounds, seq_blocks, seq_pieces = stats
    assert nb == n // 256 and seq_blocks < 120 and rounds < 400, stats


def test_exact_ties_everywhere():
    # s sits in [2^20, 2^21): u = 2^-32; every term is an odd multiple of 2^-33 -> every addition is a tie, broken by the
    # parity of s.  The blocks are unclean and must be added term by term.
    rng = np.random.default_rng(11)
    n = 5000
    t = np.empty(n)
    t[0] = 2.0 ** 20
    t[1:] = (2 * rng.integers(0, 1000, n - 1) + 1) * 2.0 ** -33
    stats = check(t, "ties")
    assert stats[3] >= n // 32 - 2   # (every piece)
    # a single tie in an otherwise clean stream
    t = rng.random(1 << 16)
    s_mid = seq_sum(t[: 1 << 15])
    u = 2.0 ** (np.floor(np.log2(s_mid)) - 52)
    t[1 << 15] = 1001 * u / 2
    check(t, "one tie")
    # dyadic terms: many additions are exact, some are ties
    t = rng.integers(1, 1 << 30, 1 << 16).astype(np.float64) * 2.0 ** -40
    t[0] = 2.0 ** 13
    check(t, "dyadic")


def test_binade_hovering_and_cancellation():
    rng = np.random.default_rng(5)
    n = 1 << 16
    t = np.empty(n)
    t[0] = 1.0
    t[1:] = np.where(np.arange(1, n) % 2 == 1, 1e-3, -1e-3) * (1 + 1e-9 * rng.standard_normal(n - 1))
    check(t, "hovering around 1.0")
    t = rng.standard_normal(n)
    t[n // 2] = -seq_sum(t[: n // 2])          # the running sum passes through 0 exactly
    check(t, "through zero")
    t = np.concatenate([np.full(3000, 1e10), np.full(3000, -1e10), rng.random(3000)])
    check(t, "cancellation")
    t = np.concatenate([rng.random(3000) * 1e-8, [1e8], rng.random(3000) * 1e-8, [-1e8], rng.random(3000)])
    check(t, "spikes")
    walk = rng.choice([-1.0, 1.0], n) * rng.random(n)
    check(walk, "random walk")


def test_non_finite_and_extreme_terms():
    rng = np.random.default_rng(9)
    t = rng.random(10000)
    for bad in (np.nan, np.inf, -np.inf):
        u = t.copy()
        u[5000] = bad
        check(u, "non-finite")
    u = t.copy()
    u[100] = np.inf
    u[9000] = -np.inf
    check(u, "inf - inf")
    check(t * 1e-310, "subnormal terms")
    check(t * 1e300, "huge terms")
    check(np.concatenate([t * 1e-300, t * 1e300]), "range")
    check(np.full(3000, 1.7e308), "overflow")
    check(np.zeros(5000), "zeros")
    check(-np.zeros(5000), "negative zeros")
    check(np.concatenate([np.zeros(300), -np.zeros(300), [5e-324] * 300]), "zeros and denormals")


def test_fuzz_small_blocks_of_everything():
    rng = np.random.default_rng(2024)
    for trial in range(300):
        n = int(rng.integers(1, 3000))
        kind = trial % 6
        if kind == 0:
            t = rng.random(n) * 10.0 ** rng.integers(-20, 20)
        elif kind == 1:
            t = rng.standard_normal(n) * 10.0 ** rng.integers(-5, 5, n)
        elif kind == 2:
            t = rng.integers(-1 << 20, 1 << 20, n).astype(np.float64) * 2.0 ** int(rng.integers(-60, 10))
        elif kind == 3:
            t = rng.random(n)
            t[rng.integers(0, n, 5)] *= -1e6
        elif kind == 4:
            t = np.ldexp(rng.integers(1, 4, n).astype(np.float64), rng.integers(-30, 30, n))
        else:
            t = np.cumsum(rng.standard_normal(n)) * 1e-3 + 1.0
        check(t, ("fuzz", trial, kind))


def test_baseline_size():
    # 256^3 terms, as the dot products of BASELINE.json configs[1]
    n = 1 << 24
    i = np.arange(n, dtype=np.float64)
    x = np.sin(i * 0.37) + 0.25 * np.cos(i * 1.3)
    stats = check(x * x, "16.8 M")
    assert stats[2] < 200, stats

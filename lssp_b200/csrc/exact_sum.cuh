// exact_sum.cuh -- the reference's SEQUENTIAL sum, computed in parallel, bit for bit.
//
// The reference adds the terms of every dot product / norm one after the other (`for (i = 0; i < n; i++) sum += x[i] *
// y[i]`, src/vector.cxx:129), so its result carries the rounding of n - 1 dependent additions.  A tree reduction is more
// accurate but DIFFERENT, and Krylov recurrences amplify the difference.  This file reproduces the
// sequential result exactly without the n-step dependency chain:
//
//   While the running sum s stays inside one binade [2^e, 2^(e+1)) its values are multiples of u = 2^(e-52), and
//       fl(s + t) = s + rn_u(t)          (rn_u: t rounded to the nearest multiple of u)
//   unless t lies exactly half-way between two multiples (the tie is broken by the parity of s).  Hence over a block of
//   kXsB terms that keeps s inside the binade and has no tie, the sequential additions amount to ONE exact integer
//   addition  s/u + sum_i rn_u(t_i)/u,  and the integers rn_u(t_i)/u are independent of s: they are computed and added in
//   parallel (int64, any order).
//
//   K1  per block: an approximate sum, the largest |prefix sum| and the sum of |t_i|        (parallel)
//   K2  approximate prefix sums -> the binade e_b each block is EXPECTED to run in          (one CTA; a prediction only)
//   K3  per block: D_b = sum_i rn_u(t_i) / u for u of e_b; ties mark the block "unclean"    (parallel)
//   K4  the walk: with the exact s at the start of a block, VERIFY that s - dev >= 2^e and s + dev < 2^(e+1), dev >= every
//       excursion of the partial sums inside the block; verified blocks advance s by D_b (exact), all others --
//       binade crossings, ties, a wrong prediction, non-finite terms -- are taken apart into pieces of kXsFine terms, rounded
//       for the binade s is in by then and advanced the same way; only pieces that fail again are added term by term.
//       Runs of verified blocks are checked 4096 at a time against a wrap-around prefix scan of D.
//
// Nothing in K1 / K2 has to be rigorous: only K4's verification, which uses the exact s, decides.  The functions below
// are shared by the kernels (exact_sum.cu) and their host replay (lsspg_debug_exact_seq_sum_host, CPU test-suite).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define LSSPG_HD __host__ __device__ __forceinline__
#else
#define LSSPG_HD inline
#endif

namespace lsspg {

constexpr int kXsB = 256;                    // terms per block
constexpr int kXsFine = 32;                  // terms per piece of a block that could not be advanced as a whole
constexpr int kXsChunk = 1024;               // threads of the walking CTA
constexpr int kXsPer = 4;                    // blocks verified per thread and round of the walk
constexpr int kXsUnclean = -(1 << 30);       // "binade" of a block that has to be added term by term

LSSPG_HD unsigned long long xs_bits(double v)
{
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b;
    memcpy(&b, &v, 8);
    return b;
#endif
}
LSSPG_HD double xs_from_bits(unsigned long long b)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double v;
    memcpy(&v, &b, 8);
    return v;
#endif
}

// e with 2^e <= |v| < 2^(e+1); kXsUnclean for zero, subnormal, non-finite values and exponents outside +-900 (so that
// 2^(e-53) and 2^(52-e) below are normal numbers)
LSSPG_HD int xs_exponent(double v)
{
    const int be = (int)((xs_bits(v) >> 52) & 0x7ffu);
    if (be == 0 || be == 0x7ff) return kXsUnclean;
    const int e = be - 1023;
    return (e < -900 || e > 900) ? kXsUnclean : e;
}
LSSPG_HD double xs_pow2(int e) { return xs_from_bits((unsigned long long)(e + 1023) << 52); }

// rn_u(t) / u as an integer for u = 2^(e-52), and whether t lies exactly half-way between two multiples of u (*tie).
// Scaling by a power of two is exact (a product that underflows is far below 1/2: no tie, rounds to 0), rint() rounds
// to nearest-even as the addition would, and q - rint(q) is exact.
LSSPG_HD long long xs_round(double t, int e, bool *tie)
{
    const double q = t * xs_pow2(52 - e);
    if (!(fabs(q) < 4.6e18)) { *tie = true; return 0; }   // (also NaN) such a block is never advanced as a whole
    const double r = rint(q);
    if (fabs(q - r) == 0.5) *tie = true;
    return (long long)r;
}

// Excursion bound of a block: `xmax` is the largest |t_0 + .. + t_i| over the block's prefixes and `absb` the sum of the
// |t_i|, both as computed in floating point (any order).  The real prefixes are within kXsB 2^-53 absb of the computed ones.
LSSPG_HD double xs_dev(double xmax, double absb) { return xmax * (1.0 + 9.5367431640625e-07) + absb * 9.094947017729282e-13; }

// K4: may a block be advanced as one integer addition from the exact sum s?  dev = xs_dev(..) + the roundings of the
// SEQUENTIAL additions (u / 2 each) + the rounding of the comparisons below bounds every |partial sum - s|.
// dev < 2^(e-2) is not needed for the lattice argument but keeps every |t_i| <= 2 dev far inside the range of the int64 sums.
LSSPG_HD bool xs_verify(double s, int e, double dev0)
{
    const double lo = xs_pow2(e), hi = xs_pow2(e + 1), u = xs_pow2(e - 52);
    const double dev = dev0 + (double)(kXsB + 8) * u;
    const double as = fabs(s);
    return (as - dev >= lo) && (as + dev < hi) && (dev < 0.25 * lo);   // false for NaN / inf
}

// s (a multiple of u = 2^(e-52) with 2^52 <= |s / u| < 2^53) -> s / u
LSSPG_HD long long xs_to_int(double s, int e) { return (long long)(s * xs_pow2(52 - e)); }
LSSPG_HD bool xs_int_in_binade(long long m)
{
    const long long am = m < 0 ? -m : m;
    return am >= (1ll << 52) && am < (1ll << 53);
}
LSSPG_HD double xs_from_int(long long m, int e) { return (double)m * xs_pow2(e - 52); }

}  // namespace lsspg

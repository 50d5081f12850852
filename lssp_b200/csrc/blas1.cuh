// blas1.cuh -- BLAS-1 vector kernels and the fused Krylov recurrences (K2-K4 of
// SURVEY.md 2.2) used by the drivers.  All operands are device pointers.
#pragma once
#include "common.cuh"
#include "scalars.cuh"

namespace lsspg {

// flag slots in ctx->d_flags
enum { FLAG_STOP = 0, FLAG_AUX = 1, FLAG_TRI_TIMEOUT = 8, FLAG_SETUP = 9 /* set-up kernels: unsorted row, overflow, bad index */ };

// Reduction target: sums -> d_scal[out_slot + k], then `fin` runs on the device.
struct RedOut {
    int out_slot = 0;
    bool guarded = false;   // skip the kernel when the device stop flag is set
    FinProg fin;
};

// reference-shaped primitives (operand order as src/vector.cxx)
int vec_set(lsspg_ctx *ctx, int n, double *x, double val, bool guarded = false);
int vec_copy(lsspg_ctx *ctx, int n, double *dst, const double *src);
int vec_axy(lsspg_ctx *ctx, int n, Coef a, const double *x, double *y);                          // y = x*a
int vec_axpby(lsspg_ctx *ctx, int n, Coef a, const double *x, Coef b, double *y);                // y = y*b + x*a
int vec_axpbyz(lsspg_ctx *ctx, int n, Coef a, const double *x, Coef b, const double *y, double *z);  // z = y*b + x*a
int vec_scale(lsspg_ctx *ctx, int n, double *x, Coef a);                                         // x *= a
int vec_scale_div(lsspg_ctx *ctx, int n, double *x, double d);                                   // x /= d (IEEE division, as `v.d[k] /= beta`)
// k dot products xs[i].ys[i] in one pass (k <= kMaxRedK)
int vec_multidot(lsspg_ctx *ctx, int n, int k, const double *const *xs, const double *const *ys, const RedOut &out);

// ---- fused recurrences ---------------------------------------------------
// CG (src/solver-cg.cxx:82-93): p = z (first) | p = z + beta*p
int cg_update_p(lsspg_ctx *ctx, int n, const double *z, double *p, Coef beta, bool first);
// CG (src/solver-cg.cxx:101-106): x = x + alpha*p; r = r - alpha*q; sums[0] = r.r ; sums[1] = r.z2 when z2 != NULL
int cg_update_xr(lsspg_ctx *ctx, int n, Coef alpha, const double *p, const double *q, double *x, double *r,
                 const RedOut &out);
// BiCGStab (src/solver-bicgstab.cxx:94-103): p = r (first) | p = r + beta*(p - omega*v)
int bicgstab_update_p(lsspg_ctx *ctx, int n, const double *r, double *p, const double *v, Coef beta, Coef omega,
                      bool first);
// BiCGStab (:113-117): s = r - alpha*v ; sums[0] = s.s
int bicgstab_update_s(lsspg_ctx *ctx, int n, const double *r, const double *v, Coef alpha, double *s,
                      const RedOut &out);
// BiCGStab (:136-141,:87): x = x + alpha*ph + omega*sh; r = s - omega*t; sums[0] = r.r, sums[1] = r.rh
int bicgstab_update_xr(lsspg_ctx *ctx, int n, Coef alpha, Coef omega, const double *ph, const double *sh,
                       const double *s, const double *t, const double *rh, double *x, double *r,
                       const RedOut &out);
// x = x + a*p  (BiCGStab breakdown branch :120-122 and friends)
int vec_xpay_inplace(lsspg_ctx *ctx, int n, Coef a, const double *p, double *x);

// sequential-order verification mode helpers (LSSPG_OPT_REDUCE_SEQUENTIAL)
int seq_prepare(lsspg_ctx *ctx, long long n);
int seq_finish(lsspg_ctx *ctx, long long n, int K, const RedOut &o);
// LSSPG_OPT_REDUCE_SEQUENTIAL = 2 (exact_sum.cu): the K sequential sums of the parked terms, computed in parallel
int exact_seq_sum(lsspg_ctx *ctx, long long n, int K, const RedOut &o);

// read scalars / flags back to the pinned mirrors (one sync)
int read_scalars(lsspg_ctx *ctx, int first, int count, bool with_flags);
int write_scalar(lsspg_ctx *ctx, int slot, double v);
int clear_flags(lsspg_ctx *ctx);

}  // namespace lsspg

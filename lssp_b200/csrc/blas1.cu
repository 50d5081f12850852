// blas1.cu -- BLAS-1 vector kernels (replaces src/vector.cxx) and the fused
// Krylov recurrences.  All kernels are HBM streams: grid = (#SM x 8) CTAs of 256
// threads, each CTA-iteration covers 1024 consecutive elements, 4 independent
// 8-byte loads per operand per thread in flight (fully coalesced 256 B / warp).
// Reductions are two-stage and fixed-order (common.cuh: grid_sum), so results
// are run-to-run reproducible.  Element-wise arithmetic keeps the reference's
// operand order and is compiled without FMA, hence bit-identical to the CPU.
#include "blas1.cuh"
#include "comm.cuh"

namespace lsspg {

constexpr int kU = 4;  // elements per thread per CTA-iteration

template <class Load, class Store>
__device__ __forceinline__ void ew_loop(long long n, Load load, Store store)
{
    const long long step = (long long)gridDim.x * (kBlock * kU);
    for (long long base = (long long)blockIdx.x * (kBlock * kU) + threadIdx.x; base < n; base += step) {
        if (base + (kU - 1) * kBlock < n) {
#pragma unroll
            for (int u = 0; u < kU; u++) load(base + u * kBlock, u);
#pragma unroll
            for (int u = 0; u < kU; u++) store(base + u * kBlock, u);
        }
        else {
            for (int u = 0; u < kU; u++) {
                const long long i = base + u * kBlock;
                if (i < n) {
                    load(i, 0);
                    store(i, 0);
                }
            }
        }
    }
}

struct RedArgs {
    double *scal;
    int *flags;
    double *partials;
    unsigned int *ticket;
    int out_slot;
    const int *stop;
    double *seq;          // sequential-order verification mode: per-element terms, [K][seq_n]
    long long seq_n;
    int defer_fin;        // multi-GPU: sums are combined across ranks before `fin` runs (comm.cu)
    FinProg fin;
};

static RedArgs red_args(lsspg_ctx *ctx, const RedOut &o)
{
    RedArgs r;
    r.scal = ctx->d_scal; r.flags = ctx->d_flags; r.partials = ctx->d_partials; r.ticket = ctx->d_ticket;
    r.out_slot = o.out_slot; r.fin = o.fin;
    r.stop = o.guarded ? ctx->d_flags + FLAG_STOP : nullptr;
    r.seq = ctx->opt_reduce_sequential ? ctx->d_seq : nullptr;
    r.seq_n = (long long)ctx->seq_len;
    r.defer_fin = distributed(ctx) ? 1 : 0;
    return r;
}

// one term of sum k at element i: accumulated per thread (fast mode) or parked for the
// single-thread in-order adder (sequential mode)
__device__ __forceinline__ void red_add(const RedArgs &ra, double &acc, int k, long long i, double term)
{
    if (ra.seq) ra.seq[(size_t)k * ra.seq_n + i] = term;
    else acc += term;
}

// Sequential-order adder: thread k adds the n parked terms of sum k exactly as the
// reference's `for (i = 0; i < n; i++) sum += x[i] * y[i]` (src/vector.cxx:129) does.
__global__ void k_seq_sum(long long n, int K, RedArgs ra)
{
    if (ra.stop && *ra.stop) return;
    if ((int)threadIdx.x < K) {
        const double *t = ra.seq + (size_t)threadIdx.x * ra.seq_n;
        double s = 0.0;
        for (long long i = 0; i < n; i++) s += t[i];
        ra.scal[ra.out_slot + threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0 && !ra.defer_fin) fin_run(ra.fin, ra.scal, ra.flags);
}

int seq_prepare(lsspg_ctx *ctx, long long n)
{
    if (!ctx->opt_reduce_sequential) return 0;
    return ensure_seq(ctx, (size_t)(n > 0 ? n : 1));
}

// called after every reducing kernel: sequential-order adder (verification mode) and, on
// several GPUs, the cross-rank combination of the sums followed by the deferred FinProg
int seq_finish(lsspg_ctx *ctx, long long n, int K, const RedOut &o)
{
    if (ctx->opt_reduce_sequential == 2) LSSPG_TRY(exact_seq_sum(ctx, n, K, o));   // the same sums, computed in parallel
    else if (ctx->opt_reduce_sequential) LSSPG_LAUNCH(ctx, k_seq_sum, 1, 32, 0, n, K, red_args(ctx, o));
    if (distributed(ctx)) return red_post(ctx, o.out_slot, K, o.fin, o.guarded);
    return 0;
}

template <int K>
__device__ __forceinline__ void finish(double (&acc)[K], const RedArgs &ra)
{
    if (ra.seq) return;   // sums are produced by k_seq_sum
    double *scal = ra.scal;
    int *flags = ra.flags;
    const int slot = ra.out_slot;
    const FinProg &fin = ra.fin;
    const int defer = ra.defer_fin;
    grid_sum<K>(acc, ra.partials, ra.ticket, [&](double(&s)[K]) {
#pragma unroll
        for (int k = 0; k < K; k++) scal[slot + k] = s[k];
        if (!defer) fin_run(fin, scal, flags);
    });
}

// ---- primitives -------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_set(long long n, double *__restrict__ x, double v, const int *stop)
{
    if (stop && *stop) return;
    ew_loop(n, [&](long long, int) {}, [&](long long i, int) { x[i] = v; });
}

__global__ void __launch_bounds__(kBlock) k_copy(long long n, double *__restrict__ d, const double *__restrict__ s)
{
    double a[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = s[i]; }, [&](long long i, int u) { d[i] = a[u]; });
}

__global__ void __launch_bounds__(kBlock) k_axy(long long n, Coef ca, const double *scal, const double *__restrict__ x,
                                                double *__restrict__ y)
{
    const double al = coef_get(ca, scal);
    double a[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = x[i]; }, [&](long long i, int u) { y[i] = a[u] * al; });
}

// y = y*b + x*a   (x may alias y: element-wise, same index)
__global__ void __launch_bounds__(kBlock) k_axpby(long long n, Coef ca, Coef cb, const double *scal, const double *x,
                                                  double *y)
{
    const double al = coef_get(ca, scal), be = coef_get(cb, scal);
    double a[kU], b[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = x[i]; b[u] = y[i]; },
            [&](long long i, int u) { y[i] = b[u] * be + a[u] * al; });
}

// z = y*b + x*a   (z may alias x or y)
__global__ void __launch_bounds__(kBlock) k_axpbyz(long long n, Coef ca, Coef cb, const double *scal, const double *x,
                                                   const double *y, double *z)
{
    const double al = coef_get(ca, scal), be = coef_get(cb, scal);
    double a[kU], b[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = x[i]; b[u] = y[i]; },
            [&](long long i, int u) { z[i] = b[u] * be + a[u] * al; });
}

__global__ void __launch_bounds__(kBlock) k_scale(long long n, Coef ca, const double *scal, double *x)
{
    const double al = coef_get(ca, scal);
    double a[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = x[i]; }, [&](long long i, int u) { x[i] = a[u] * al; });
}

__global__ void __launch_bounds__(kBlock) k_div(long long n, double d, double *x)
{
    double a[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = x[i]; }, [&](long long i, int u) { x[i] = a[u] / d; });
}

struct DotPtrs {
    const double *x[kMaxRedK];
    const double *y[kMaxRedK];
};

template <int K>
__global__ void __launch_bounds__(kBlock) k_multidot(long long n, DotPtrs p, RedArgs ra)
{
    if (ra.stop && *ra.stop) return;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = 0.0;
    double a[K][kU], b[K][kU];
    ew_loop(n,
            [&](long long i, int u) {
#pragma unroll
                for (int k = 0; k < K; k++) { a[k][u] = p.x[k][i]; b[k][u] = p.y[k][i]; }
            },
            [&](long long i, int u) {
#pragma unroll
                for (int k = 0; k < K; k++) red_add(ra, acc[k], k, i, a[k][u] * b[k][u]);
            });
    finish<K>(acc, ra);
}

// ---- fused recurrences --------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_cg_p(long long n, const double *__restrict__ z, double *__restrict__ p,
                                                 Coef cbeta, const double *scal, const int *flags, int first)
{
    if (flags[FLAG_STOP]) return;
    double a[kU], b[kU];
    if (first) {
        ew_loop(n, [&](long long i, int u) { a[u] = z[i]; }, [&](long long i, int u) { p[i] = a[u]; });
    }
    else {
        const double beta = coef_get(cbeta, scal);
        ew_loop(n, [&](long long i, int u) { a[u] = z[i]; b[u] = p[i]; },
                [&](long long i, int u) { p[i] = a[u] + beta * b[u]; });   // src/solver-cg.cxx:91
    }
}

__global__ void __launch_bounds__(kBlock) k_cg_xr(long long n, Coef calpha, const double *__restrict__ p,
                                                  const double *__restrict__ q, double *__restrict__ x,
                                                  double *__restrict__ r, RedArgs ra)
{
    if (ra.flags[FLAG_STOP]) return;
    const double alpha = coef_get(calpha, ra.scal);
    double acc[1] = {0.0};
    double vp[kU], vq[kU], vx[kU], vr[kU];
    ew_loop(n, [&](long long i, int u) { vp[u] = p[i]; vq[u] = q[i]; vx[u] = x[i]; vr[u] = r[i]; },
            [&](long long i, int u) {
                x[i] = vx[u] + alpha * vp[u];                 // src/solver-cg.cxx:102
                const double rn = vr[u] - alpha * vq[u];      // :103
                r[i] = rn;
                red_add(ra, acc[0], 0, i, rn * rn);           // :106 (norm = sqrt(dot(r,r)))
            });
    finish<1>(acc, ra);
}

__global__ void __launch_bounds__(kBlock) k_bicgstab_p(long long n, const double *__restrict__ r, double *__restrict__ p,
                                                       const double *__restrict__ v, Coef cbeta, Coef comega,
                                                       const double *scal, const int *flags, int first)
{
    if (flags[FLAG_STOP]) return;
    double a[kU], b[kU], c[kU];
    if (first) {
        ew_loop(n, [&](long long i, int u) { a[u] = r[i]; }, [&](long long i, int u) { p[i] = a[u]; });
    }
    else {
        const double beta = coef_get(cbeta, scal), omega = coef_get(comega, scal);
        ew_loop(n, [&](long long i, int u) { a[u] = r[i]; b[u] = p[i]; c[u] = v[i]; },
                [&](long long i, int u) { p[i] = a[u] + beta * (b[u] - omega * c[u]); });  // src/solver-bicgstab.cxx:101
    }
}

__global__ void __launch_bounds__(kBlock) k_bicgstab_s(long long n, const double *__restrict__ r,
                                                       const double *__restrict__ v, Coef calpha,
                                                       double *__restrict__ s, RedArgs ra)
{
    if (ra.flags[FLAG_STOP]) return;
    const double alpha = coef_get(calpha, ra.scal);
    double acc[1] = {0.0};
    double a[kU], b[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = r[i]; b[u] = v[i]; },
            [&](long long i, int u) {
                const double sn = a[u] - alpha * b[u];        // src/solver-bicgstab.cxx:114
                s[i] = sn;
                red_add(ra, acc[0], 0, i, sn * sn);
            });
    finish<1>(acc, ra);
}

__global__ void __launch_bounds__(kBlock) k_bicgstab_xr(long long n, Coef calpha, Coef comega,
                                                        const double *__restrict__ ph, const double *__restrict__ sh,
                                                        const double *__restrict__ s, const double *__restrict__ t,
                                                        const double *__restrict__ rh, double *__restrict__ x,
                                                        double *__restrict__ r, RedArgs ra)
{
    if (ra.flags[FLAG_STOP]) return;
    const double alpha = coef_get(calpha, ra.scal), omega = coef_get(comega, ra.scal);
    double acc[2] = {0.0, 0.0};
    double vph[kU], vsh[kU], vs[kU], vt[kU], vrh[kU], vx[kU];
    ew_loop(n,
            [&](long long i, int u) {
                vph[u] = ph[i]; vsh[u] = sh[i]; vs[u] = s[i]; vt[u] = t[i]; vrh[u] = rh[i]; vx[u] = x[i];
            },
            [&](long long i, int u) {
                x[i] = vx[u] + alpha * vph[u] + omega * vsh[u];   // src/solver-bicgstab.cxx:137
                const double rn = vs[u] - omega * vt[u];          // :138
                r[i] = rn;
                red_add(ra, acc[0], 0, i, rn * rn);               // :141
                red_add(ra, acc[1], 1, i, rn * vrh[u]);           // :87 of the next iteration
            });
    finish<2>(acc, ra);
}

__global__ void __launch_bounds__(kBlock) k_xpay(long long n, Coef ca, const double *scal, const double *__restrict__ p,
                                                 double *__restrict__ x)
{
    const double al = coef_get(ca, scal);
    double a[kU], b[kU];
    ew_loop(n, [&](long long i, int u) { a[u] = p[i]; b[u] = x[i]; },
            [&](long long i, int u) { x[i] = b[u] + al * a[u]; });
}

// ---- launchers ------------------------------------------------------------------
static inline int ew_grid(lsspg_ctx *ctx, int n) { return stream_grid(ctx, n, kBlock * kU); }

int vec_set(lsspg_ctx *ctx, int n, double *x, double val, bool guarded)
{
    if (n <= 0) return 0;
    LSSPG_LAUNCH(ctx, k_set, ew_grid(ctx, n), kBlock, 0, (long long)n, x, val,
                 guarded ? ctx->d_flags + FLAG_STOP : (const int *)nullptr);
    return 0;
}

int vec_copy(lsspg_ctx *ctx, int n, double *dst, const double *src)
{
    if (n <= 0 || dst == src) return 0;
    LSSPG_LAUNCH(ctx, k_copy, ew_grid(ctx, n), kBlock, 0, (long long)n, dst, src);
    return 0;
}

int vec_axy(lsspg_ctx *ctx, int n, Coef a, const double *x, double *y)
{
    if (n <= 0) return 0;
    if (x == y) return vec_scale(ctx, n, y, a);
    LSSPG_LAUNCH(ctx, k_axy, ew_grid(ctx, n), kBlock, 0, (long long)n, a, ctx->d_scal, x, y);
    return 0;
}

int vec_axpby(lsspg_ctx *ctx, int n, Coef a, const double *x, Coef b, double *y)
{
    if (n <= 0) return 0;
    LSSPG_LAUNCH(ctx, k_axpby, ew_grid(ctx, n), kBlock, 0, (long long)n, a, b, ctx->d_scal, x, y);
    return 0;
}

int vec_axpbyz(lsspg_ctx *ctx, int n, Coef a, const double *x, Coef b, const double *y, double *z)
{
    if (n <= 0) return 0;
    LSSPG_LAUNCH(ctx, k_axpbyz, ew_grid(ctx, n), kBlock, 0, (long long)n, a, b, ctx->d_scal, x, y, z);
    return 0;
}

int vec_scale(lsspg_ctx *ctx, int n, double *x, Coef a)
{
    if (n <= 0) return 0;
    LSSPG_LAUNCH(ctx, k_scale, ew_grid(ctx, n), kBlock, 0, (long long)n, a, ctx->d_scal, x);
    return 0;
}

int vec_scale_div(lsspg_ctx *ctx, int n, double *x, double d)
{
    if (n <= 0) return 0;
    LSSPG_LAUNCH(ctx, k_div, ew_grid(ctx, n), kBlock, 0, (long long)n, d, x);
    return 0;
}

int vec_multidot(lsspg_ctx *ctx, int n, int k, const double *const *xs, const double *const *ys, const RedOut &out)
{
    LSSPG_CHECK(k >= 1 && k <= kMaxRedK, "multidot: k=%d out of range", k);
    DotPtrs p;
    for (int i = 0; i < kMaxRedK; i++) {
        p.x[i] = xs[i < k ? i : 0];
        p.y[i] = ys[i < k ? i : 0];
    }
    LSSPG_TRY(seq_prepare(ctx, n));
    const RedArgs ra = red_args(ctx, out);
    const int grid = ew_grid(ctx, n > 0 ? n : 1);
    const long long nn = n > 0 ? n : 0;
    switch (k) {
        case 1: LSSPG_LAUNCH(ctx, k_multidot<1>, grid, kBlock, 0, nn, p, ra); break;
        case 2: LSSPG_LAUNCH(ctx, k_multidot<2>, grid, kBlock, 0, nn, p, ra); break;
        case 3: LSSPG_LAUNCH(ctx, k_multidot<3>, grid, kBlock, 0, nn, p, ra); break;
        case 4: LSSPG_LAUNCH(ctx, k_multidot<4>, grid, kBlock, 0, nn, p, ra); break;
        case 5: LSSPG_LAUNCH(ctx, k_multidot<5>, grid, kBlock, 0, nn, p, ra); break;
        case 6: LSSPG_LAUNCH(ctx, k_multidot<6>, grid, kBlock, 0, nn, p, ra); break;
        case 7: LSSPG_LAUNCH(ctx, k_multidot<7>, grid, kBlock, 0, nn, p, ra); break;
        default: LSSPG_LAUNCH(ctx, k_multidot<8>, grid, kBlock, 0, nn, p, ra); break;
    }
    return seq_finish(ctx, nn, k, out);
}

int cg_update_p(lsspg_ctx *ctx, int n, const double *z, double *p, Coef beta, bool first)
{
    LSSPG_LAUNCH(ctx, k_cg_p, ew_grid(ctx, n), kBlock, 0, (long long)n, z, p, beta, ctx->d_scal, ctx->d_flags, first ? 1 : 0);
    return 0;
}

int cg_update_xr(lsspg_ctx *ctx, int n, Coef alpha, const double *p, const double *q, double *x, double *r,
                 const RedOut &out)
{
    LSSPG_TRY(seq_prepare(ctx, n));
    LSSPG_LAUNCH(ctx, k_cg_xr, ew_grid(ctx, n), kBlock, 0, (long long)n, alpha, p, q, x, r, red_args(ctx, out));
    return seq_finish(ctx, n, 1, out);
}

int bicgstab_update_p(lsspg_ctx *ctx, int n, const double *r, double *p, const double *v, Coef beta, Coef omega,
                      bool first)
{
    LSSPG_LAUNCH(ctx, k_bicgstab_p, ew_grid(ctx, n), kBlock, 0, (long long)n, r, p, v, beta, omega, ctx->d_scal,
                 ctx->d_flags, first ? 1 : 0);
    return 0;
}

int bicgstab_update_s(lsspg_ctx *ctx, int n, const double *r, const double *v, Coef alpha, double *s,
                      const RedOut &out)
{
    LSSPG_TRY(seq_prepare(ctx, n));
    LSSPG_LAUNCH(ctx, k_bicgstab_s, ew_grid(ctx, n), kBlock, 0, (long long)n, r, v, alpha, s, red_args(ctx, out));
    return seq_finish(ctx, n, 1, out);
}

int bicgstab_update_xr(lsspg_ctx *ctx, int n, Coef alpha, Coef omega, const double *ph, const double *sh,
                       const double *s, const double *t, const double *rh, double *x, double *r,
                       const RedOut &out)
{
    LSSPG_TRY(seq_prepare(ctx, n));
    LSSPG_LAUNCH(ctx, k_bicgstab_xr, ew_grid(ctx, n), kBlock, 0, (long long)n, alpha, omega, ph, sh, s, t, rh, x, r,
                 red_args(ctx, out));
    return seq_finish(ctx, n, 2, out);
}

int vec_xpay_inplace(lsspg_ctx *ctx, int n, Coef a, const double *p, double *x)
{
    LSSPG_LAUNCH(ctx, k_xpay, ew_grid(ctx, n), kBlock, 0, (long long)n, a, ctx->d_scal, p, x);
    return 0;
}

int read_scalars(lsspg_ctx *ctx, int first, int count, bool with_flags)
{
    LSSPG_CUDA(cudaMemcpyAsync(ctx->h_scal + first, ctx->d_scal + first, sizeof(double) * count,
                               cudaMemcpyDeviceToHost, ctx->stream));
    if (with_flags)
        LSSPG_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(int) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (with_flags && ctx->h_flags[FLAG_TRI_TIMEOUT]) {
        cudaMemsetAsync(ctx->d_flags + FLAG_TRI_TIMEOUT, 0, sizeof(int), ctx->stream);
        ctx->tri_timeouts++;
        set_error("triangular solve watchdog: a row waited for a dependency that never arrived");
        return 1;
    }
    return 0;
}

int write_scalar(lsspg_ctx *ctx, int slot, double v)
{
    ctx->h_scal[slot] = v;
    LSSPG_CUDA(cudaMemcpyAsync(ctx->d_scal + slot, ctx->h_scal + slot, sizeof(double), cudaMemcpyHostToDevice,
                               ctx->stream));
    return 0;
}

int clear_flags(lsspg_ctx *ctx)
{
    LSSPG_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(int) * 8, ctx->stream));   // keeps the error flags
    return 0;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

// synchronises the stream, reads the device status flags and reports (and clears) a sweep aborted by the watchdog
int lsspg_check_flags(lsspg_ctx *ctx) { return read_scalars(ctx, 0, 1, true); }

int lsspg_vec_set(lsspg_ctx *ctx, int n, double *dx, double val) { return vec_set(ctx, n, dx, val, false); }
int lsspg_vec_copy(lsspg_ctx *ctx, int n, double *ddst, const double *dsrc) { return vec_copy(ctx, n, ddst, dsrc); }
int lsspg_vec_axy(lsspg_ctx *ctx, int n, double a, const double *dx, double *dy)
{ return vec_axy(ctx, n, coef_imm(a), dx, dy); }
int lsspg_vec_axpby(lsspg_ctx *ctx, int n, double a, const double *dx, double b, double *dy)
{ return vec_axpby(ctx, n, coef_imm(a), dx, coef_imm(b), dy); }
int lsspg_vec_axpbyz(lsspg_ctx *ctx, int n, double a, const double *dx, double b, const double *dy, double *dz)
{ return vec_axpbyz(ctx, n, coef_imm(a), dx, coef_imm(b), dy, dz); }
int lsspg_vec_scale(lsspg_ctx *ctx, int n, double *dx, double a) { return vec_scale(ctx, n, dx, coef_imm(a)); }

int lsspg_vec_multidot(lsspg_ctx *ctx, int n, int k, const double *const *dxs, const double *dy, double *h_out)
{
    LSSPG_CHECK(k >= 1 && k <= kMaxRedK, "lsspg_vec_multidot: k=%d out of range [1,%d]", k, kMaxRedK);
    const double *ys[kMaxRedK];
    for (int i = 0; i < k; i++) ys[i] = dy;
    RedOut out;
    out.out_slot = kNumScalars - kMaxRedK;  // scratch slots at the top of the slab
    LSSPG_TRY(vec_multidot(ctx, n, k, dxs, ys, out));
    LSSPG_TRY(read_scalars(ctx, out.out_slot, k, false));
    for (int i = 0; i < k; i++) h_out[i] = ctx->h_scal[out.out_slot + i];
    return 0;
}

int lsspg_vec_dot(lsspg_ctx *ctx, int n, const double *dx, const double *dy, double *h_out)
{
    const double *xs[1] = {dx};
    return lsspg_vec_multidot(ctx, n, 1, xs, dy, h_out);
}

int lsspg_vec_norm(lsspg_ctx *ctx, int n, const double *dx, double *h_out)
{
    double d = 0.0;
    LSSPG_TRY(lsspg_vec_dot(ctx, n, dx, dx, &d));
    *h_out = sqrt(d);   // src/vector.cxx:135-138
    return 0;
}

}  // extern "C"

// krylov_ops.cuh -- host-scalar toolkit of the drivers that follow the reference call by call.
// The reference interleaves BLAS-1 / SpMV / pc.solve calls with host arithmetic on the returned
// scalars; these helpers keep that shape: every call is one kernel on the library's stream, every
// dot or norm is one reduction kernel plus one 8-byte read-back.
#pragma once
#include "krylov.cuh"

namespace lsspg {

struct Ops {
    lsspg_ctx *ctx;
    int n;
    static constexpr int S_TMP = 96;   // scratch slots of the scalar slab
    int dots(int k, const double *const *xs, const double *const *ys, double *out)
    {
        RedOut o; o.out_slot = S_TMP;
        LSSPG_TRY(vec_multidot(ctx, n, k, xs, ys, o));
        LSSPG_TRY(read_scalars(ctx, S_TMP, k, false));
        for (int i = 0; i < k; i++) out[i] = ctx->h_scal[S_TMP + i];
        return 0;
    }
    int dot(const double *x, const double *y, double *out) { const double *xs[1] = {x}, *ys[1] = {y}; return dots(1, xs, ys, out); }
    int norm(const double *x, double *out) { double d; LSSPG_TRY(dot(x, x, &d)); *out = sqrt(d); return 0; }   // src/vector.cxx:135-138
    int axpby(double a, const double *x, double b, double *y) { return vec_axpby(ctx, n, coef_imm(a), x, coef_imm(b), y); }
    int axpbyz(double a, const double *x, double b, const double *y, double *z) { return vec_axpbyz(ctx, n, coef_imm(a), x, coef_imm(b), y, z); }
    int axy(double a, const double *x, double *y) { return vec_axy(ctx, n, coef_imm(a), x, y); }
    int scale(double *x, double a) { return vec_scale(ctx, n, x, coef_imm(a)); }
    int copy(double *d, const double *s) { return vec_copy(ctx, n, d, s); }
    int set(double *x, double v) { return vec_set(ctx, n, x, v); }
};

// one driver invocation: matrix, preconditioner, right-hand side, work vectors, result
struct Drv : Ops {
    KrylovArgs &k;
    const lsspg_solver_opts *raw;   // the caller's settings before the usual defaults (Lis-style drivers use them as is)
    Workspace W;
    const double *b;
    bool good = true;
    Drv(KrylovArgs &k_, const lsspg_solver_opts *raw_) : Ops{k_.ctx, k_.n}, k(k_), raw(raw_), W(k_.ctx, k_.nvec), b(k_.b) {}
    double *vec()
    {
        double *p = W.vec();
        if (!p) good = false;
        return p;
    }
    bool ok() const { return good; }
    int mxy(const double *x, double *y) { return spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), x, coef_imm(0.0), nullptr, y, nullptr); }
    int amxpbyz(double a, const double *x, double bb, const double *y, double *z)
    { return spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(a), x, coef_imm(bb), y, z, nullptr); }
    int resid(const double *x, double *r) { return amxpbyz(-1.0, x, 1.0, b, r); }   // r = b - A x
    int pc(double *out, const double *in) { return pc_apply(ctx, k.pc, out, in, false); }
    void report(const char *name, int iter, double nrm2, double ires, int minverb)
    {
        record(k, iter > 0 ? iter - 1 : 0, nrm2);
        if (k.verb >= minverb) log_printf("%s: itr: %5d, abs res: %.6e, rel res: %.6e\n", name, iter, nrm2, nrm2 / ires);
    }
    int finish(int nits, double residual)
    {
        k.info->nits = nits;
        k.info->residual = residual;
        return 0;
    }
};

int krylov_cgs(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_cr(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_crs(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_bicrstab(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_tfqmr(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_qmrcgstab(KrylovArgs &k);
int krylov_orthomin(KrylovArgs &k);
int krylov_bicgsafe(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_bicrsafe(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_gpbicg(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_gpbicr(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_bicgstabl(KrylovArgs &k, const lsspg_solver_opts *raw);
int krylov_rgmres(KrylovArgs &k);
int krylov_lgmres(KrylovArgs &k);
int krylov_rlgmres(KrylovArgs &k);

}  // namespace lsspg

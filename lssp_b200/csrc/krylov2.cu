// krylov2.cu -- GMRES(m) and IDR(s) drivers plus the host-scalar toolkit shared by the
// remaining drivers.
//
// GMRES keeps the reference's modified Gram-Schmidt (i+1 DEPENDENT reductions per inner step,
// src/solver-gmres.cxx:142-147) but fuses every "w -= h_j v_j" update with the NEXT dot product
// (k_axpby_dot: one pass, 32n bytes instead of 40n) and keeps the h_j on the device, so one inner
// step needs a single device->host read (the Hessenberg column).  The Givens rotations, the small
// triangular solve and IDR(s)'s s x s system stay on the host exactly as in the reference.
#include <algorithm>
#include <stdlib.h>
#include "comm.cuh"
#include "krylov.cuh"
#include "krylov_ops.cuh"

namespace lsspg {

constexpr int kU2 = 4;

struct RedArgs2 {
    double *scal;
    double *partials;
    unsigned int *ticket;
    int out_slot;
    double *seq;
    long long seq_n;
};

// y = y*b + x*a (coefficients may live on the device), then sum = y_new . z  (z == NULL: y_new . y_new)
__global__ void __launch_bounds__(kBlock) k_axpby_dot(long long n, Coef ca, Coef cb, const double *__restrict__ x,
                                                      double *__restrict__ y, const double *__restrict__ z, RedArgs2 ra)
{
    const double al = coef_get(ca, ra.scal), be = coef_get(cb, ra.scal);
    double acc[1] = {0.0};
    const long long step = (long long)gridDim.x * (kBlock * kU2);
    for (long long base = (long long)blockIdx.x * (kBlock * kU2) + threadIdx.x; base < n; base += step) {
        double vx[kU2], vy[kU2], vz[kU2];
#pragma unroll
        for (int u = 0; u < kU2; u++) {
            const long long i = base + u * kBlock;
            if (i < n) { vx[u] = x[i]; vy[u] = y[i]; vz[u] = z ? z[i] : 0.0; }
        }
#pragma unroll
        for (int u = 0; u < kU2; u++) {
            const long long i = base + u * kBlock;
            if (i < n) {
                const double yn = vy[u] * be + vx[u] * al;      // src/vector.cxx:105
                y[i] = yn;
                const double term = yn * (z ? vz[u] : yn);      // src/vector.cxx:129
                if (ra.seq) ra.seq[i] = term;
                else acc[0] += term;
            }
        }
    }
    if (ra.seq) return;
    double *scal = ra.scal;
    const int slot = ra.out_slot;
    grid_sum<1>(acc, ra.partials, ra.ticket, [&](double(&s)[1]) { scal[slot] = s[0]; });
}

// mode 0 (GMRES x-update, src/solver-gmres.cxx:196-204):  t = 0; t += V_i[j]*c_i (i ascending); out[j] += t
// mode 1 (IDR(s) combos,  src/solver-idrs.cxx:198-214):   h = s0*base[j]; h -= V_i[j]*c_i (i ascending); out[j] = h
// mode 2 (LGMRES update,   src/solver-lgmres.cxx:229-257): t = 0; t += V_i[j]*c_i; t += Z_i[j]*c2_i; out[j] += t; out2[j] = t
struct Lincomb2 {
    const double *Z;
    const double *coef2;
    int count2;
    double *out2;
};

__global__ void __launch_bounds__(kBlock) k_div_into(long long n, double d, const double *__restrict__ in, double *__restrict__ out)
{
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
        out[j] = in[j] / d;   // src/solver-lgmres.cxx:190-192
}

__global__ void __launch_bounds__(kBlock) k_lincomb(long long n, int mode, int count, const double *__restrict__ V,
                                                    long long stride, const double *__restrict__ coef, double s0,
                                                    const double *base, double *out, Lincomb2 ex)
{
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        if (mode == 2) {
            double t = 0;
            for (int i = 0; i < count; i++) t += V[i * stride + j] * coef[i];
            for (int i = 0; i < ex.count2; i++) t += ex.Z[i * stride + j] * ex.coef2[i];
            out[j] += t;
            ex.out2[j] = t;
        }
        else if (mode == 0) {
            double t = 0;
            for (int i = 0; i < count; i++) t += V[i * stride + j] * coef[i];
            out[j] += t;
        }
        else {
            double h = s0 * base[j];
            for (int i = 0; i < count; i++) h -= V[i * stride + j] * coef[i];
            out[j] = h;
        }
    }
}

static int axpby_dot(lsspg_ctx *ctx, int n, Coef a, const double *x, Coef b, double *y, const double *z, int out_slot)
{
    LSSPG_TRY(seq_prepare(ctx, n));
    RedArgs2 ra;
    ra.scal = ctx->d_scal; ra.partials = ctx->d_partials; ra.ticket = ctx->d_ticket; ra.out_slot = out_slot;
    ra.seq = ctx->opt_reduce_sequential ? ctx->d_seq : nullptr;
    ra.seq_n = (long long)ctx->seq_len;
    LSSPG_LAUNCH(ctx, k_axpby_dot, stream_grid(ctx, n, kBlock * kU2), kBlock, 0, (long long)n, a, b, x, y, z, ra);
    RedOut o; o.out_slot = out_slot;
    return seq_finish(ctx, n, 1, o);
}

static int upload_coefs(lsspg_ctx *ctx, const double *h, int count, int slot)
{
    for (int i = 0; i < count; i++) ctx->h_scal[slot + i] = h[i];
    LSSPG_CUDA(cudaMemcpyAsync(ctx->d_scal + slot, ctx->h_scal + slot, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// the preconditioner as the drivers see it: NON is a copy (src/pc.cxx:67-70)
static int apply_pc(KrylovArgs &k, double *out, const double *in) { return pc_apply(k.ctx, k.pc, out, in, false); }

// ---- GMRES(m), left preconditioning: src/solver-gmres.cxx:12-255 -----------------------------
int krylov_gmres(KrylovArgs &k)
{
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    int m = k.restart;
    if (m < 0) m = kDefRestart;                       // :44
    double tol_rb = k.tol_rb;
    if (tol_rb < 0) tol_rb = kDefRb;                  // :48
    constexpr int S_H = 128;                          // Hessenberg column of the current inner step, then ym
    LSSPG_CHECK(m >= 1 && S_H + m + 2 <= kNumScalars, "gmres: restart %d not supported (max %d)", m, kNumScalars - S_H - 2);
    Ops op{ctx, n};
    Workspace W(ctx, k.nvec);
    double *wj = W.vec(), *rg = W.vec();
    Workspace WV(ctx, (long long)k.nvec * m);                    // the Krylov basis, one contiguous block
    double *V = WV.vec();
    LSSPG_CHECK(wj && rg && V, "gmres: out of device memory");
    const long long ld = k.nvec;
    std::vector<double> Hg((size_t)(m + 1) * m, 0.0), gg(m + 1), ym(m), c(m), s(m);
    auto H = [&](int r, int col) -> double & { return Hg[(size_t)r * m + col]; };

    double b_norm, beta;
    LSSPG_TRY(op.norm(k.b, &b_norm));
    tol_rb *= b_norm;
    LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));   // :88
    LSSPG_TRY(op.norm(rg, &beta));
    int itr_inner = 0;
    if (beta <= k.tol_abs) {                          // :91-94
        k.info->nits = 0;
        k.info->residual = beta;
        return 0;
    }
    const double err_rel = beta;
    double tol = k.tol_rel * err_rel;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    const double rtol = tol / beta;
    double gstol = 0.;
    int cycle = 0;

    while (itr_inner < k.maxit) {
        int kk, i;
        double gs_norm = 0.;
        LSSPG_TRY(op.set(V, 0.));                                     // :112
        LSSPG_TRY(apply_pc(k, V, rg));                                // :113
        LSSPG_TRY(op.norm(V, &beta));
        gg[0] = beta;
        for (kk = 1; kk <= m; kk++) gg[kk] = 0;
        if (itr_inner == 0) gstol = rtol * beta * 0.5;                // :121-123
        std::fill(Hg.begin(), Hg.end(), 0.0);
        LSSPG_TRY(vec_scale_div(ctx, n, V, beta));                    // :129-131  v0[k] /= beta
        for (i = 0; i < m; i++) {
            itr_inner++;
            double *vi = V + (long long)i * ld;
            LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), vi, coef_imm(0.0), nullptr, rg, nullptr));   // :138
            LSSPG_TRY(op.set(wj, 0.));
            LSSPG_TRY(apply_pc(k, wj, rg));                           // :140
            // modified Gram-Schmidt, h_j kept on the device: dot, then (update fused with the next dot)
            {
                const double *xs[1] = {wj}, *ys[1] = {V};
                RedOut o; o.out_slot = S_H;
                LSSPG_TRY(vec_multidot(ctx, n, 1, xs, ys, o));        // h_0 = wj . v_0           :143
            }
            for (int j = 0; j <= i; j++) {
                const double *vj = V + (long long)j * ld;
                const double *next = (j < i) ? V + (long long)(j + 1) * ld : nullptr;      // last: ||wj||^2  :149
                LSSPG_TRY(axpby_dot(ctx, n, coef_slot(S_H + j, true), vj, coef_imm(1.0), wj, next, S_H + j + 1));   // :144
            }
            LSSPG_TRY(read_scalars(ctx, S_H, i + 2, false));
            for (int j = 0; j <= i; j++) H(j, i) = ctx->h_scal[S_H + j];
            double hij = sqrt(ctx->h_scal[S_H + i + 1]);
            H(i + 1, i) = hij;
            if (fabs(hij) <= kBreakdown) {                            // :152-155
                i--;
                break;
            }
            else if (i + 1 < m) {
                LSSPG_TRY(op.axy(1 / hij, wj, V + (long long)(i + 1) * ld));   // :157
            }
            for (int j = 0; j < i; j++) {                             // :160-166
                const double h1 = c[j] * H(j, i) + s[j] * H(j + 1, i);
                const double h2 = -s[j] * H(j, i) + c[j] * H(j + 1, i);
                H(j, i) = h1;
                H(j + 1, i) = h2;
            }
            double gma = sqrt(H(i, i) * H(i, i) + H(i + 1, i) * H(i + 1, i));
            if (fabs(gma) == 0.) gma = 1e-20;
            c[i] = H(i, i) / gma;
            s[i] = H(i + 1, i) / gma;
            gg[i + 1] = -s[i] * gg[i];
            gg[i] = c[i] * gg[i];
            H(i, i) = c[i] * H(i, i) + s[i] * H(i + 1, i);
            gs_norm = fabs(gg[i + 1]);
            if (gs_norm <= gstol) break;                              // :179-181 (goto solve)
        }
        kk = (i == m) ? m : i + 1;                                    // :185
        for (i = kk - 1; i >= 0; i--) {                               // :186-194
            ym[i] = gg[i] / H(i, i);
            for (int j = 0; j < i; j++) gg[j] = gg[j] - ym[i] * H(j, i);
        }
        if (kk > 0) {
            LSSPG_TRY(upload_coefs(ctx, ym.data(), kk, S_H));
            LSSPG_LAUNCH(ctx, k_lincomb, stream_grid(ctx, n, kBlock), kBlock, 0, (long long)n, 0, kk, V, ld,
                         ctx->d_scal + S_H, 0.0, (const double *)nullptr, k.x, Lincomb2{nullptr, nullptr, 0, nullptr});   // :196-204
        }
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));   // :206
        LSSPG_TRY(op.norm(rg, &beta));
        record(k, cycle++, beta);
        if (k.verb >= 1)
            log_printf("gmres: itr: %4d / %5d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", itr_inner, itr_inner, beta,
                   (err_rel == 0 ? 0 : beta / err_rel), (b_norm == 0 ? 0 : beta / b_norm));
        if (beta <= tol) break;
        gstol = rtol * gs_norm / (beta / err_rel) * 0.5;              // :220
    }
    k.info->nits = itr_inner;
    k.info->residual = beta;
    return 0;
}

// ---- IDR(s): src/solver-idrs.cxx:86-283 ----------------------------------------------------
static void small_solve(int n, const double *a, const double *b, double *x, double *w)   // :23-84
{
    for (int i = 0; i < n * n; i++) w[i] = a[i];
    if (n == 1) { x[0] = b[0] / w[0]; return; }
    if (n == 2) {
        w[0] = 1.0 / w[0];
        w[1] *= w[0];
        w[3] -= w[1] * w[2];
        w[3] = 1.0 / w[3];
        x[0] = b[0];
        x[1] = b[1] - w[1] * x[0];
        x[1] *= w[3];
        x[0] -= w[2] * x[1];
        x[0] *= w[0];
        return;
    }
    for (int kq = 0; kq < n; kq++) {
        w[kq + kq * n] = 1.0 / w[kq + kq * n];
        for (int i = kq + 1; i < n; i++) {
            const double t = w[i + kq * n] * w[kq + kq * n];
            for (int j = kq + 1; j < n; j++) w[i + j * n] -= t * w[kq + j * n];
            w[i + kq * n] = t;
        }
    }
    for (int i = 0; i < n; i++) {
        x[i] = b[i];
        for (int j = 0; j < i; j++) x[i] -= w[i + j * n] * x[j];
    }
    for (int i = n - 1; i >= 0; i--) {
        for (int j = i + 1; j < n; j++) x[i] -= w[i + j * n] * x[j];
        x[i] *= w[i + i * n];
    }
}

int krylov_idrs(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    int s = k.idrs;
    if (s <= 0) s = 4;                                                // :101
    // this driver reads the raw settings, without the usual defaults (:97-100, :121-130)
    const int maxiter = raw->maxit;
    const double tol_abs = raw->tol_abs, tol_rel = raw->tol_rel, tol_rbs = raw->tol_rb;
    constexpr int S_C = 128;
    LSSPG_CHECK(s <= 32, "idrs: s = %d not supported (max 32)", s);
    Ops op{ctx, n};
    Workspace W(ctx, k.nvec);
    double *r = W.vec(), *t = W.vec(), *v = W.vec(), *av = W.vec();
    Workspace WS(ctx, (long long)k.nvec * s);
    double *dX = WS.vec(), *dR = WS.vec(), *P = WS.vec();
    LSSPG_CHECK(r && t && v && av && dX && dR && P, "idrs: out of device memory");
    const long long ld = k.nvec;
    std::vector<double> m(s), c(s), M((size_t)s * s), MM((size_t)s * s);
    double om = 0, h, nrm2, ires, tol;
    int iter = 0, oldest;

    LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, r, nullptr));   // :118
    LSSPG_TRY(op.norm(r, &nrm2));
    ires = nrm2;
    if (nrm2 <= tol_abs) {
        k.info->nits = 0;
        k.info->residual = nrm2;
        return 0;
    }
    tol = nrm2 * tol_rel;
    LSSPG_TRY(op.norm(k.b, &h));
    h *= tol_rbs;
    if (tol < tol_abs) tol = tol_abs;
    if (tol < h) tol = h;
    {   // shadow space: glibc stream, k outer / i inner (:139-144), generated on the host.  Row-sharded: every
        // rank walks the SAME global stream and keeps the entries of its own rows, so the shadow vectors --
        // and with them the iteration -- are those of the serial run whatever the number of GPUs
        long long row0 = 0, n_global = n;
        if (distributed(ctx)) {
            int rank = 0, nranks = 1;
            LSSPG_TRY(lsspg_comm_size(ctx, &rank, &nranks));
            std::vector<double> cnt((size_t)nranks, 0.0);
            cnt[rank] = (double)n;
            double *d_cnt = nullptr;
            LSSPG_CUDA(cudaMalloc(&d_cnt, sizeof(double) * (size_t)nranks));
            LSSPG_CUDA(cudaMemcpyAsync(d_cnt, cnt.data(), sizeof(double) * (size_t)nranks, cudaMemcpyHostToDevice, ctx->stream));
            int rc = comm_allreduce(ctx, d_cnt, nranks);
            if (!rc && cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(double) * (size_t)nranks, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = 1;
            if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = 1;
            cudaFree(d_cnt);
            LSSPG_CHECK(rc == 0, "idrs: could not gather the row counts of the ranks");
            n_global = 0;
            for (int q = 0; q < nranks; q++) {
                if (q < rank) row0 += (long long)cnt[q];
                n_global += (long long)cnt[q];
            }
        }
        std::vector<double> hp((size_t)n);
        srand(0);
        for (int kq = 0; kq < s; kq++) {
            for (long long i = 0; i < n_global; i++) {
                const double val = (rand() * 1.) / (1. * RAND_MAX);
                if (i >= row0 && i < row0 + n) hp[(size_t)(i - row0)] = val;
            }
            LSSPG_CUDA(cudaMemcpyAsync(P + kq * ld, hp.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    for (int j = 0; j < s; j++) {                                     // idrs_orth :4-21
        double rn, d;
        LSSPG_TRY(op.norm(P + j * ld, &rn));
        rn = 1.0 / rn;
        LSSPG_TRY(op.scale(P + j * ld, rn));
        for (int i = j + 1; i < s; i++) {
            LSSPG_TRY(op.dot(P + j * ld, P + i * ld, &d));
            LSSPG_TRY(op.axpby(-d, P + j * ld, 1, P + i * ld));
        }
    }
    bool finished = false;
    for (int kq = 0; kq < s && !finished; kq++) {                     // :148-173
        double *dx = dX + kq * ld, *dr = dR + kq * ld;
        LSSPG_TRY(apply_pc(k, dx, r));
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), dx, coef_imm(0.0), nullptr, dr, nullptr));
        const double *xs[2] = {dr, dr}, *ys[2] = {dr, r};
        double d2[2];
        LSSPG_TRY(op.dots(2, xs, ys, d2));
        h = d2[0];
        om = d2[1] / h;
        LSSPG_TRY(op.scale(dx, om));
        LSSPG_TRY(op.scale(dr, -om));
        LSSPG_TRY(op.axpby(1, dx, 1, k.x));
        LSSPG_TRY(op.axpby(1, dr, 1, r));
        LSSPG_TRY(op.norm(r, &nrm2));
        if (tol >= nrm2) {
            iter = kq + 1;
            finished = true;
            break;
        }
        std::vector<const double *> px(s), py(s);
        for (int i = 0; i < s; i++) { px[i] = P + i * ld; py[i] = dr; }
        LSSPG_TRY(op.dots(s, px.data(), py.data(), &M[(size_t)kq * s]));
    }
    if (!finished) {
        iter = s;
        oldest = 0;
        {
            std::vector<const double *> px(s), py(s);
            for (int i = 0; i < s; i++) { px[i] = P + i * ld; py[i] = r; }
            LSSPG_TRY(op.dots(s, px.data(), py.data(), m.data()));
        }
        const int grid = stream_grid(ctx, n, kBlock);
        while (iter <= maxiter) {                                     // :181
            small_solve(s, M.data(), m.data(), c.data(), MM.data());
            LSSPG_TRY(op.copy(v, r));
            for (int j = 0; j < s; j++) LSSPG_TRY(op.axpby(-c[j], dR + j * ld, 1, v));      // :185-188
            LSSPG_TRY(upload_coefs(ctx, c.data(), s, S_C));
            double *dxo = dX + oldest * ld, *dro = dR + oldest * ld;
            if ((iter % (s + 1)) == s) {
                LSSPG_TRY(apply_pc(k, av, v));
                LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), av, coef_imm(0.0), nullptr, t, nullptr));
                const double *xs[2] = {t, t}, *ys[2] = {t, v};
                double d2[2];
                LSSPG_TRY(op.dots(2, xs, ys, d2));
                h = d2[0];
                om = d2[1] / h;
                LSSPG_LAUNCH(ctx, k_lincomb, grid, kBlock, 0, (long long)n, 1, s, dX, ld, ctx->d_scal + S_C, om, av, dxo, Lincomb2{nullptr, nullptr, 0, nullptr});    // :198-205
                LSSPG_LAUNCH(ctx, k_lincomb, grid, kBlock, 0, (long long)n, 1, s, dR, ld, ctx->d_scal + S_C, -om, t, dro, Lincomb2{nullptr, nullptr, 0, nullptr});    // :207-214
            }
            else {
                LSSPG_TRY(apply_pc(k, av, v));
                LSSPG_LAUNCH(ctx, k_lincomb, grid, kBlock, 0, (long long)n, 1, s, dX, ld, ctx->d_scal + S_C, om, av, dxo, Lincomb2{nullptr, nullptr, 0, nullptr});    // :219-226
                LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), dxo, coef_imm(0.0), nullptr, dro, nullptr));
                LSSPG_TRY(op.scale(dro, -1.));
            }
            LSSPG_TRY(op.axpby(1, dro, 1, r));
            LSSPG_TRY(op.axpby(1, dxo, 1, k.x));
            iter++;
            LSSPG_TRY(op.norm(r, &nrm2));
            record(k, iter - s - 1, nrm2);
            if (k.verb >= 1) log_printf("idrs: itr: %5d, abs res: %.6e, rel res: %.6e\n", iter, nrm2, nrm2 / ires);
            if (tol >= nrm2) break;
            std::vector<const double *> px(s), py(s);
            std::vector<double> hh(s);
            for (int i = 0; i < s; i++) { px[i] = P + i * ld; py[i] = dro; }
            LSSPG_TRY(op.dots(s, px.data(), py.data(), hh.data()));
            for (int i = 0; i < s; i++) {                             // :248-252
                m[i] += hh[i];
                M[(size_t)oldest * s + i] = hh[i];
            }
            oldest++;
            if (oldest == s) oldest = 0;
        }
    }
    k.info->nits = iter;
    k.info->residual = nrm2;
    return 0;
}

// ---- shared Arnoldi pieces of the GMRES family ------------------------------------------------------
// modified Gram-Schmidt of wj against v_0..v_i with the h_j kept on the device (see krylov_gmres);
// returns the Hessenberg column in col[0..i] and ||wj|| in *hnorm
static int mgs_column(lsspg_ctx *ctx, int n, double *wj, const double *V, long long ld, int i, int S_H, double *col,
                      double *hnorm)
{
    {
        const double *xs[1] = {wj}, *ys[1] = {V};
        RedOut o; o.out_slot = S_H;
        LSSPG_TRY(vec_multidot(ctx, n, 1, xs, ys, o));
    }
    for (int j = 0; j <= i; j++) {
        const double *vj = V + (long long)j * ld;
        const double *next = (j < i) ? V + (long long)(j + 1) * ld : nullptr;
        LSSPG_TRY(axpby_dot(ctx, n, coef_slot(S_H + j, true), vj, coef_imm(1.0), wj, next, S_H + j + 1));
    }
    LSSPG_TRY(read_scalars(ctx, S_H, i + 2, false));
    for (int j = 0; j <= i; j++) col[j] = ctx->h_scal[S_H + j];
    *hnorm = sqrt(ctx->h_scal[S_H + i + 1]);
    return 0;
}

struct Hess {
    int m;
    std::vector<double> Hg, gg, ym, c, s;
    explicit Hess(int m_) : m(m_), Hg((size_t)(m_ + 1) * m_, 0.0), gg(m_ + 1, 0.0), ym(m_, 0.0), c(m_, 0.0), s(m_, 0.0) {}
    double &H(int r, int col) { return Hg[(size_t)r * m + col]; }
    void clear(int mm)   // as the reference: only the (mm+1) x mm leading part of the current cycle
    {
        for (int kk = 0; kk <= mm; kk++)
            for (int i = 0; i < mm; i++) H(kk, i) = 0;
    }
    // Givens update of column i (src/solver-gmres.cxx:160-177); returns |gg[i+1]|
    double rotate(int i)
    {
        for (int j = 0; j < i; j++) {
            const double h1 = c[j] * H(j, i) + s[j] * H(j + 1, i);
            const double h2 = -s[j] * H(j, i) + c[j] * H(j + 1, i);
            H(j, i) = h1;
            H(j + 1, i) = h2;
        }
        double gma = sqrt(H(i, i) * H(i, i) + H(i + 1, i) * H(i + 1, i));
        if (fabs(gma) == 0.) gma = 1e-20;
        c[i] = H(i, i) / gma;
        s[i] = H(i + 1, i) / gma;
        gg[i + 1] = -s[i] * gg[i];
        gg[i] = c[i] * gg[i];
        H(i, i) = c[i] * H(i, i) + s[i] * H(i + 1, i);
        return fabs(gg[i + 1]);
    }
    void back_substitute(int kk)
    {
        for (int i = kk - 1; i >= 0; i--) {
            ym[i] = gg[i] / H(i, i);
            for (int j = 0; j < i; j++) gg[j] = gg[j] - ym[i] * H(j, i);
        }
    }
};

// ---- RGMRES(m), right preconditioning: src/solver-gmres.cxx:257-479 -----------------------------------
int krylov_rgmres(KrylovArgs &k)
{
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    int m = k.restart;
    if (m < 0) m = kDefRestart;
    double tol_rb = k.tol_rb;
    if (tol_rb < 0) tol_rb = kDefRb;
    constexpr int S_H = 128;
    LSSPG_CHECK(m >= 1 && S_H + m + 2 <= kNumScalars, "rgmres: restart %d not supported", m);
    Ops op{ctx, n};
    Workspace W(ctx, k.nvec);
    double *wj = W.vec(), *rg = W.vec();
    Workspace WV(ctx, (long long)k.nvec * m);
    double *V = WV.vec();
    LSSPG_CHECK(wj && rg && V, "rgmres: out of device memory");
    const long long ld = k.nvec;
    Hess h(m);
    double b_norm, beta;
    LSSPG_TRY(op.norm(k.b, &b_norm));
    tol_rb *= b_norm;
    LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));
    LSSPG_TRY(op.norm(rg, &beta));
    if (beta <= k.tol_abs) { k.info->nits = 0; k.info->residual = beta; return 0; }
    const double err_rel = beta;
    double tol = k.tol_rel * err_rel;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    int itr_inner = 0;
    std::vector<double> col(m + 1);
    while (itr_inner < k.maxit) {
        int i, kk;
        for (kk = 1; kk <= m; kk++) h.gg[kk] = 0;
        h.clear(m);
        LSSPG_TRY(op.norm(rg, &beta));
        h.gg[0] = beta;
        LSSPG_TRY(op.axy(1 / beta, rg, V));
        for (i = 0; i < m && itr_inner < k.maxit; i++) {
            double hij;
            itr_inner++;
            LSSPG_TRY(op.set(rg, 0.));
            LSSPG_TRY(apply_pc(k, rg, V + (long long)i * ld));
            LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), rg, coef_imm(0.0), nullptr, wj, nullptr));
            LSSPG_TRY(mgs_column(ctx, n, wj, V, ld, i, S_H, col.data(), &hij));
            for (int j = 0; j <= i; j++) h.H(j, i) = col[j];
            h.H(i + 1, i) = hij;
            if (fabs(hij) <= kBreakdown) { i -= 1; break; }
            else if (i + 1 < m) LSSPG_TRY(op.axy(1 / hij, wj, V + (long long)(i + 1) * ld));
            beta = h.rotate(i);
            record(k, itr_inner - 1, beta);
            if (k.verb >= 1)
                log_printf("rgmres: itr: %4d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", itr_inner, beta,
                       (err_rel == 0 ? 0 : beta / err_rel), (b_norm == 0 ? 0 : beta / b_norm));
            if (beta <= tol) break;
        }
        kk = (i == m) ? m : i + 1;
        h.back_substitute(kk);
        if (kk > 0) {                                            // :424-438
            LSSPG_TRY(op.axy(h.ym[kk - 1], V + (long long)(kk - 1) * ld, rg));
            for (i = kk - 2; i >= 0; i--) LSSPG_TRY(op.axpby(h.ym[i], V + (long long)i * ld, 1, rg));
            LSSPG_TRY(apply_pc(k, wj, rg));
            LSSPG_TRY(op.axpby(1, wj, 1, k.x));
        }
        if (beta <= tol) break;
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));
    }
    k.info->nits = itr_inner;
    k.info->residual = beta;
    return 0;
}

// ---- LGMRES(m, k), left preconditioning: src/solver-lgmres.cxx:12-311 ---------------------------------
int krylov_lgmres(KrylovArgs &k)
{
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    int mk = k.restart, auk = k.aug_k;
    if (mk < 0) mk = kDefRestart;
    if (auk < 0) auk = kDefAugK;
    double tol_rb = k.tol_rb;
    if (tol_rb < 0) tol_rb = kDefRb;
    int m = mk + auk;
    constexpr int S_H = 128;
    LSSPG_CHECK(mk >= 1 && auk >= 1 && S_H + m + 2 <= kNumScalars, "lgmres: restart %d + aug %d not supported", mk, auk);
    Ops op{ctx, n};
    Workspace W(ctx, k.nvec);
    double *wj = W.vec(), *rg = W.vec();
    Workspace WV(ctx, (long long)k.nvec * m), WZ(ctx, (long long)k.nvec * auk);
    double *V = WV.vec(), *Z = WZ.vec();
    LSSPG_CHECK(wj && rg && V && Z, "lgmres: out of device memory");
    const long long ld = k.nvec;
    Hess h(m);
    double b_norm, beta;
    LSSPG_TRY(op.norm(k.b, &b_norm));
    tol_rb *= b_norm;
    LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));
    LSSPG_TRY(op.norm(rg, &beta));
    if (beta <= k.tol_abs) { k.info->nits = 0; k.info->residual = beta; return 0; }
    const double err_rel = beta;
    double tol = k.tol_rel * err_rel;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    const double rtol = tol / beta;
    double gstol = 0.;
    int itr_outer = 0, itr_inner = 0;
    std::vector<double> col(m + 1);
    const int grid = stream_grid(ctx, n, kBlock);
    while (itr_inner < k.maxit) {
        int kk, i;
        double gs_norm = 0.;
        LSSPG_TRY(op.set(V, 0.));
        LSSPG_TRY(apply_pc(k, V, rg));
        LSSPG_TRY(op.norm(V, &beta));
        m = (itr_outer < auk) ? mk + itr_outer : mk + auk;                  // "tune m", :113-118
        h.gg[0] = beta;
        for (kk = 1; kk <= m; kk++) h.gg[kk] = 0;
        if (itr_outer == 0) gstol = rtol * beta * 0.5;
        {   // the reference clears Hg[0..m][0..m) of an array whose rows have mk+auk columns
            for (kk = 0; kk <= m; kk++)
                for (i = 0; i < m; i++) h.H(kk, i) = 0;
        }
        LSSPG_TRY(vec_scale_div(ctx, n, V, beta));
        for (i = 0; i < m; i++) {
            double hij;
            itr_inner++;
            const double *src = (i < mk) ? V + (long long)i * ld : Z + (long long)(i - mk) * ld;
            LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), src, coef_imm(0.0), nullptr, rg, nullptr));
            LSSPG_TRY(op.set(wj, 0.));
            LSSPG_TRY(apply_pc(k, wj, rg));
            LSSPG_TRY(mgs_column(ctx, n, wj, V, ld, i, S_H, col.data(), &hij));
            for (int j = 0; j <= i; j++) h.H(j, i) = col[j];
            h.H(i + 1, i) = hij;
            if (fabs(hij) <= kBreakdown) { i--; break; }
            else if (i + 1 < m)
                LSSPG_LAUNCH(ctx, k_div_into, grid, kBlock, 0, (long long)n, hij, wj, V + (long long)(i + 1) * ld);
            gs_norm = h.rotate(i);
            if (gs_norm <= gstol) break;
        }
        kk = i;                                                              // the reference's quirk, :214
        h.back_substitute(kk);
        const int zn = itr_outer % auk;
        {
            int cv, cz;
            if (kk <= mk) { cv = kk; cz = 0; }
            else { cv = mk; cz = (itr_outer <= auk) ? itr_outer : auk; }
            if (cz > auk) cz = auk;
            if (kk > 0 || true) {
                LSSPG_TRY(upload_coefs(ctx, h.ym.data(), std::max(1, std::min(m, mk + auk)), S_H));
                Lincomb2 ex{Z, ctx->d_scal + S_H + mk, cz, Z + (long long)zn * ld};
                LSSPG_LAUNCH(ctx, k_lincomb, grid, kBlock, 0, (long long)n, 2, cv > 0 ? cv : 0, V, ld, ctx->d_scal + S_H, 0.0,
                             (const double *)nullptr, k.x, ex);              // :229-257
            }
        }
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));
        LSSPG_TRY(op.norm(rg, &beta));
        record(k, itr_outer, beta);
        if (k.verb >= 1)
            log_printf("lgmres: itr: %4d / %5d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", itr_outer, itr_inner, beta,
                   (err_rel == 0 ? 0 : beta / err_rel), (b_norm == 0 ? 0 : beta / b_norm));
        if (beta <= tol) break;
        gstol = rtol * gs_norm / (beta / err_rel) * 0.5;
        itr_outer++;
    }
    k.info->nits = itr_inner;
    k.info->residual = beta;
    return 0;
}

// ---- RLGMRES(m, k), right preconditioning: src/solver-lgmres.cxx:313-604 ------------------------------
int krylov_rlgmres(KrylovArgs &k)
{
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    int mk = k.restart, auk = k.aug_k;
    if (mk < 0) mk = kDefRestart;
    if (auk < 0) auk = kDefAugK;
    double tol_rb = k.tol_rb;
    if (tol_rb < 0) tol_rb = kDefRb;
    int m = mk + auk;
    constexpr int S_H = 128;
    LSSPG_CHECK(mk >= 1 && auk >= 1 && S_H + m + 2 <= kNumScalars, "rlgmres: restart %d + aug %d not supported", mk, auk);
    Ops op{ctx, n};
    Workspace W(ctx, k.nvec);
    double *wj = W.vec(), *rg = W.vec();
    Workspace WV(ctx, (long long)k.nvec * m), WZ(ctx, (long long)k.nvec * auk);
    double *V = WV.vec(), *Z = WZ.vec();
    LSSPG_CHECK(wj && rg && V && Z, "rlgmres: out of device memory");
    const long long ld = k.nvec;
    Hess h(m);
    double b_norm, beta;
    LSSPG_TRY(op.norm(k.b, &b_norm));
    tol_rb *= b_norm;
    LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));
    LSSPG_TRY(op.norm(rg, &beta));
    if (beta <= k.tol_abs) { k.info->nits = 0; k.info->residual = beta; return 0; }
    const double err_rel = beta;
    double tol = k.tol_rel * err_rel;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    int itr_outer = 0, itr_inner = 0;
    std::vector<double> col(m + 1);
    while (itr_inner < k.maxit) {
        int kk, i;
        m = (itr_outer < auk) ? mk + itr_outer : mk + auk;
        for (kk = 1; kk <= m; kk++) h.gg[kk] = 0;
        for (kk = 0; kk <= m; kk++)
            for (i = 0; i < m; i++) h.H(kk, i) = 0;
        LSSPG_TRY(op.norm(rg, &beta));
        h.gg[0] = beta;
        LSSPG_TRY(op.axy(1 / beta, rg, V));
        for (i = 0; i < m && itr_inner < k.maxit; i++) {
            double hij;
            itr_inner++;
            LSSPG_TRY(op.set(rg, 0.));
            LSSPG_TRY(apply_pc(k, rg, (i < mk) ? V + (long long)i * ld : Z + (long long)(i - mk) * ld));
            LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), rg, coef_imm(0.0), nullptr, wj, nullptr));
            LSSPG_TRY(mgs_column(ctx, n, wj, V, ld, i, S_H, col.data(), &hij));
            for (int j = 0; j <= i; j++) h.H(j, i) = col[j];
            h.H(i + 1, i) = hij;
            if (fabs(hij) <= kBreakdown) { i--; break; }
            else if (i + 1 < m) LSSPG_TRY(op.axy(1 / hij, wj, V + (long long)(i + 1) * ld));
            beta = h.rotate(i);
            record(k, itr_inner - 1, beta);
            if (k.verb >= 1)
                log_printf("rlgmres: itr: %4d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", itr_inner, beta,
                       (err_rel == 0 ? 0 : beta / err_rel), (b_norm == 0 ? 0 : beta / b_norm));
            if (beta <= tol) break;
        }
        kk = i;                                                              // :506
        h.back_substitute(kk);
        if (kk > 0) {
            // after the back substitution the reference's loop index is -1, so its `if (i <= mk)` always
            // takes the first branch: only the v_i enter the correction (:519-524), also for kk > mk
            LSSPG_TRY(op.axy(h.ym[0], V, rg));
            for (i = 1; i < kk; i++) LSSPG_TRY(op.axpby(h.ym[i], V + (long long)i * ld, 1, rg));
            LSSPG_TRY(apply_pc(k, wj, rg));
            LSSPG_TRY(op.axpby(1, wj, 1, k.x));
            LSSPG_TRY(op.copy(Z + (long long)(itr_outer % auk) * ld, rg));
        }
        if (beta <= tol) break;
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, rg, nullptr));
        LSSPG_TRY(op.norm(rg, &beta));
        itr_outer++;
    }
    k.info->nits = itr_inner;
    k.info->residual = beta;
    return 0;
}

}  // namespace lsspg

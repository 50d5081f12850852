// krylov.cu -- Krylov driver entry points and the two headline drivers (CG,
// BiCGStab) as device-resident pipelines.
//
// CG / BiCGStab keep every scalar (rho, alpha, beta, omega) on the device: the
// reducing kernels derive them in their last CTA (scalars.cuh FinProg) with the
// same IEEE operations, in the same order, as the reference's host code.  One
// iteration is 4 (CG) / 6 (BiCGStab) fused kernels plus the preconditioner; the
// host reads the residual back once per `check_every` iterations, and a device
// stop flag turns the remaining kernels of a batch into no-ops so that
// iteration counts stay exactly those of the reference.
#include <algorithm>
#include "krylov.cuh"
#include "comm.cuh"
#include "krylov_ops.cuh"

namespace lsspg {

enum { FLAG_BRK = 2 };
constexpr int S_HIST = 64;      // residual of iteration j of the current batch
constexpr int kMaxBatch = 32;

// ---- CG: src/solver-cg.cxx:8-136 --------------------------------------------
int krylov_cg(KrylovArgs &k)
{
    enum { S_RHO = 0, S_PQ = 2, S_ALPHA = 3, S_BETA = 4, S_RES2 = 5, S_TOL = 6, S_B2 = 7 };
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    const bool non = (k.pc->kind == LSSPG_PC_NON);
    Workspace W(ctx, k.nvec);
    double *r = W.vec(), *p = W.vec(), *q = W.vec();
    double *z = non ? r : W.vec();   // pc NON: z is a bitwise copy of r (src/pc.cxx:67-70) -> alias
    LSSPG_CHECK(r && p && q && z, "cg: out of device memory");
    LSSPG_TRY(clear_flags(ctx));

    {   // b_norm (:56) and r = b - A x with ||r||^2 (:59-60)
        const double *xs[1] = {k.b}, *ys[1] = {k.b};
        RedOut o; o.out_slot = S_B2;
        LSSPG_TRY(vec_multidot(ctx, n, 1, xs, ys, o));
        SpmvDots d; d.ndot = 1; d.out_slot = S_RES2;
        if (non) d.fin.add(FIN_COPY, S_RHO + 0, S_RES2);   // rho_0 = z.r = r.r when z == r
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, r, &d));
        LSSPG_TRY(read_scalars(ctx, S_RES2, 3, false));
    }
    const double b_norm = sqrt(ctx->h_scal[S_B2]);
    double residual = sqrt(ctx->h_scal[S_RES2]);
    const double err_rel = residual;
    int it = 0;
    bool converged = false;
    if (residual <= k.tol_abs) {   // :61-64
        k.info->nits = 0;
        k.info->residual = residual;
        return 0;
    }
    const double tol = stop_tolerance(k, residual, b_norm);
    LSSPG_TRY(write_scalar(ctx, S_TOL, tol));
    const int batch = std::max(1, std::min(ctx->opt_check_every, kMaxBatch));

    // one batch of iterations [it0, it0 + nb): the launch train between two read-backs of the residuals
    auto launch_batch = [&](int it0, int nb) -> int {
        for (int j = 0; j < nb; j++) {
            const int i = it0 + j, cur = i & 1, nxt = cur ^ 1;
            if (!non) {
                LSSPG_TRY(pc_apply(ctx, k.pc, z, r, true));                        // :79
                const double *xs[1] = {z}, *ys[1] = {r};
                RedOut o; o.out_slot = S_RHO + cur; o.guarded = true;            // :80
                if (i > 0) o.fin.add(FIN_DIV, S_BETA, S_RHO + cur, S_RHO + nxt);  // :89
                LSSPG_TRY(vec_multidot(ctx, n, 1, xs, ys, o));
            }
            LSSPG_TRY(cg_update_p(ctx, n, z, p, coef_slot(S_BETA), i == 0));      // :82-93
            SpmvDots d; d.ndot = 1; d.w[0] = p; d.out_slot = S_PQ;                // :95-96
            d.fin.add(FIN_DIV, S_ALPHA, S_RHO + cur, S_PQ);                       // :98
            LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), p, coef_imm(0.0), nullptr, q, &d, true));
            RedOut o; o.out_slot = S_RES2; o.guarded = true;                      // :101-106
            o.fin.add(FIN_SQRT, S_HIST + j, S_RES2);
            if (non) {
                o.fin.add(FIN_COPY, S_RHO + nxt, S_RES2);
                o.fin.add(FIN_DIV, S_BETA, S_RHO + nxt, S_RHO + cur);
            }
            o.fin.add(FIN_FLAG_LE, FLAG_STOP, S_HIST + j, S_TOL);                 // :114
            LSSPG_TRY(cg_update_xr(ctx, n, coef_slot(S_ALPHA), p, q, k.x, r, o));
        }
        return 0;
    };
    // CUDA graph of the steady-state batch: with an even batch length every batch after the
    // first launches the same kernels with the same arguments -- coefficients are slot numbers, the scalars stay on the
    // device -- so the train is captured once and replayed; small problems are launch-bound (1000^2 CG: 3 launches per
    // 40 us iteration).  Not on several GPUs (the all-reduce kernel carries a sequence number) and not with a user
    // preconditioner (host callback).  LSSPG_OPT_GRAPHS = 0 turns it off.
    const bool use_graph = ctx->opt_graphs && !distributed(ctx) && batch % 2 == 0 && k.pc->kind != LSSPG_PC_USER &&
                           k.pc->kind != LSSPG_PC_AMG && !ctx->opt_reduce_sequential;
    cudaGraphExec_t gexec = nullptr;
    long long graph_launches = 0;
    struct GraphGuard {
        cudaGraphExec_t &g;
        ~GraphGuard() { if (g) cudaGraphExecDestroy(g); }
    } guard{gexec};

    while (it < k.maxit && !converged) {
        const int nb = std::min(batch, k.maxit - it);
        if (use_graph && it >= batch && nb == batch) {
            if (!gexec) {
                cudaGraph_t graph = nullptr;
                const long long l0 = ctx->launches;
                LSSPG_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                const int rc = launch_batch(it, nb);
                const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
                if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
                LSSPG_CUDA(ce);
                graph_launches = ctx->launches - l0;
                ctx->launches = l0;
                LSSPG_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
                cudaGraphDestroy(graph);
            }
            LSSPG_CUDA(cudaGraphLaunch(gexec, ctx->stream));
            ctx->launches += graph_launches;
        }
        else LSSPG_TRY(launch_batch(it, nb));
        LSSPG_TRY(read_scalars(ctx, S_HIST, nb, true));
        int j = 0;
        for (; j < nb; j++) {
            residual = ctx->h_scal[S_HIST + j];
            record(k, it + j, residual);
            if (k.verb >= 1)
                log_printf("cg: itr: %5d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", it + j, residual,
                       (err_rel == 0 ? 0 : residual / err_rel), (b_norm == 0 ? 0 : residual / b_norm));
            if (residual <= tol) {
                converged = true;
                break;
            }
        }
        it += converged ? j : nb;
    }
    if (converged) LSSPG_TRY(clear_flags(ctx));
    k.info->nits = (it < k.maxit) ? it + 1 : it;   // :117
    k.info->residual = residual;
    return 0;
}

// ---- BiCGStab: src/solver-bicgstab.cxx:10-175 -------------------------------
int krylov_bicgstab(KrylovArgs &k)
{
    enum { S_RHO = 0, S_RHV = 2, S_ALPHA = 3, S_BETA = 4, S_OMEGA = 5, S_TS = 6, S_TT = 7, S_S2 = 8, S_SN = 9,
           S_XR = 10 /* +0 r.r, +1 r.rh */, S_TOL = 12, S_BRK = 13, S_T1 = 14, S_T2 = 15, S_B2 = 16 };
    lsspg_ctx *ctx = k.ctx;
    const int n = k.n;
    const bool non = (k.pc->kind == LSSPG_PC_NON);
    Workspace W(ctx, k.nvec);
    double *r = W.vec(), *rh = W.vec(), *p = W.vec(), *s = W.vec(), *t = W.vec(), *v = W.vec();
    double *ph = non ? p : W.vec();   // pc NON: ph is a bitwise copy of p -> alias
    double *sh = non ? s : W.vec();
    LSSPG_CHECK(r && rh && p && s && t && v && ph && sh, "bicgstab: out of device memory");
    LSSPG_TRY(clear_flags(ctx));

    {   // r = b - A x (:70), rh = r (:71-74), b_norm, ||r||; rho_0 = r.rh = r.r
        SpmvDots d; d.ndot = 1; d.out_slot = S_XR;
        d.fin.add(FIN_COPY, S_RHO + 0, S_XR);
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, r, &d));
        LSSPG_TRY(vec_copy(ctx, n, rh, r));
        const double *xs[1] = {k.b}, *ys[1] = {k.b};
        RedOut o; o.out_slot = S_B2;
        LSSPG_TRY(vec_multidot(ctx, n, 1, xs, ys, o));
        LSSPG_TRY(read_scalars(ctx, S_XR, S_B2 - S_XR + 1, false));
    }
    const double b_norm = sqrt(ctx->h_scal[S_B2]);
    double residual = sqrt(ctx->h_scal[S_XR]);
    const double err_rel = residual;
    if (residual <= k.tol_abs) {
        k.info->nits = 0;
        k.info->residual = residual;
        return 0;
    }
    const double tol = stop_tolerance(k, residual, b_norm);
    LSSPG_TRY(write_scalar(ctx, S_TOL, tol));
    LSSPG_TRY(write_scalar(ctx, S_BRK, kBreakdown));
    const int batch = std::max(1, std::min(ctx->opt_check_every, kMaxBatch));

    int it = 0;
    bool done = false;
    while (it < k.maxit && !done) {
        const int nb = std::min(batch, k.maxit - it);
        for (int j = 0; j < nb; j++) {
            const int i = it + j, cur = i & 1, nxt = cur ^ 1;
            LSSPG_TRY(bicgstab_update_p(ctx, n, r, p, v, coef_slot(S_BETA), coef_slot(S_OMEGA), i == 0));   // :94-103
            if (!non) LSSPG_TRY(pc_apply(ctx, k.pc, ph, p, true));                                           // :107-108
            {
                SpmvDots d; d.ndot = 1; d.w[0] = rh; d.out_slot = S_RHV;                                     // :110-112
                d.fin.add(FIN_DIV, S_ALPHA, S_RHO + cur, S_RHV);
                LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), ph, coef_imm(0.0), nullptr, v, &d, true));
            }
            {
                RedOut o; o.out_slot = S_S2; o.guarded = true;                                               // :113-117
                o.fin.add(FIN_SQRT, S_SN, S_S2);
                o.fin.add(FIN_FLAG_LE, FLAG_BRK, S_SN, S_BRK, j + 1);
                o.fin.add(FIN_FLAG_LE, FLAG_STOP, S_SN, S_BRK);
                LSSPG_TRY(bicgstab_update_s(ctx, n, r, v, coef_slot(S_ALPHA), s, o));
            }
            if (!non) LSSPG_TRY(pc_apply(ctx, k.pc, sh, s, true));                                           // :130-131
            {
                SpmvDots d; d.ndot = 2; d.w[0] = s; d.w[1] = nullptr; d.out_slot = S_TS;                     // :133-135
                d.fin.add(FIN_DIV, S_OMEGA, S_TS, S_TT);
                LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, k.A, coef_imm(1.0), sh, coef_imm(0.0), nullptr, t, &d, true));
            }
            {
                RedOut o; o.out_slot = S_XR; o.guarded = true;                                               // :136-141
                o.fin.add(FIN_SQRT, S_HIST + j, S_XR);
                o.fin.add(FIN_COPY, S_RHO + nxt, S_XR + 1);                 // rho of the next iteration (:87)
                o.fin.add(FIN_MUL, S_T1, S_RHO + nxt, S_ALPHA);             // beta = (rho1*alpha)/(rho0*omega) (:99)
                o.fin.add(FIN_MUL, S_T2, S_RHO + cur, S_OMEGA);
                o.fin.add(FIN_DIV, S_BETA, S_T1, S_T2);
                o.fin.add(FIN_FLAG_LE, FLAG_STOP, S_HIST + j, S_TOL);       // :149
                o.fin.add(FIN_FLAG_EQ0, FLAG_AUX, S_RHO + nxt, 0, j + 1);   // :89-92
                o.fin.add(FIN_FLAG_EQ0, FLAG_STOP, S_RHO + nxt);
                LSSPG_TRY(bicgstab_update_xr(ctx, n, coef_slot(S_ALPHA), coef_slot(S_OMEGA), ph, sh, s, t, rh, k.x, r, o));
            }
        }
        LSSPG_TRY(read_scalars(ctx, S_HIST, nb, true));
        const int brk = ctx->h_flags[FLAG_BRK], aux = ctx->h_flags[FLAG_AUX];
        int j = 0;
        for (; j < nb; j++) {
            if (brk == j + 1) {   // ||s|| <= LSSP_BREAKDOWN  (:117-128)
                LSSPG_TRY(read_scalars(ctx, S_SN, 1, false));
                log_printf("bicgstab: ||s|| is too small: %f, terminated.\n", ctx->h_scal[S_SN]);
                LSSPG_TRY(clear_flags(ctx));
                LSSPG_TRY(vec_xpay_inplace(ctx, n, coef_slot(S_ALPHA), ph, k.x));
                SpmvDots d; d.ndot = 1; d.out_slot = S_XR;
                LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, k.A, coef_imm(-1.0), k.x, coef_imm(1.0), k.b, r, &d));
                LSSPG_TRY(read_scalars(ctx, S_XR, 1, false));
                residual = sqrt(ctx->h_scal[S_XR]);
                k.info->breakdown = 1;
                done = true;
                break;
            }
            residual = ctx->h_scal[S_HIST + j];
            record(k, it + j, residual);
            if (k.verb >= 1)
                log_printf("bicgstab: itr: %5d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", it + j, residual,
                       (err_rel == 0 ? 0 : residual / err_rel), (b_norm == 0 ? 0 : residual / b_norm));
            if (residual <= tol) {
                done = true;
                break;
            }
            if (aux == j + 1) {   // rho1 == 0 at the top of the next iteration (:89-92)
                if (it + j + 1 < k.maxit) {
                    log_printf("bicgstab: method failed.!\n");
                    k.info->breakdown = 1;
                    j++;          // the loop counter had already advanced when the reference breaks
                    done = true;
                }
                break;
            }
        }
        it += done ? j : nb;
        if (!done && (ctx->h_flags[FLAG_STOP] != 0)) done = true;   // defensive: never spin on a raised flag
    }
    LSSPG_TRY(clear_flags(ctx));
    k.info->nits = (it < k.maxit) ? it + 1 : it;   // :153
    k.info->residual = residual;
    return 0;
}

static int resolve(KrylovArgs &k, lsspg_ctx *ctx, const lsspg_csr *A, lsspg_pc *pc, const double *db, double *dx,
                   const lsspg_solver_opts *o, lsspg_solve_info *info)
{
    LSSPG_CHECK(ctx && A && pc && db && dx && o && info, "krylov: NULL argument");
    LSSPG_CHECK(A->num_rows == A->num_cols || A->halo, "krylov: matrix is not square");   // assert in every driver
    LSSPG_CHECK(pc->n == A->num_rows, "krylov: preconditioner size %d != matrix size %d", pc->n, A->num_rows);
    k.ctx = ctx; k.A = A; k.pc = pc; k.b = db; k.x = dx; k.n = A->num_rows;
    k.nvec = A->num_cols > A->num_rows ? A->num_cols : A->num_rows;
    // option resolution as at the top of every reference driver (e.g. src/solver-cg.cxx:36-38)
    k.maxit = o->maxit <= 0 ? kDefMaxit : o->maxit;
    k.tol_abs = o->tol_abs < 0 ? kDefAtol : o->tol_abs;
    k.tol_rel = o->tol_rel < 0 ? kDefRtol : o->tol_rel;
    k.tol_rb = o->tol_rb;
    k.restart = o->restart; k.aug_k = o->aug_k; k.bgsl = o->bgsl; k.idrs = o->idrs; k.verb = o->verb;
    k.hist = o->hist; k.hist_len = o->hist ? o->hist_len : 0;
    k.info = info;
    info->nits = 0; info->residual = 0.0; info->hist_used = 0; info->solve_ms = 0.0; info->launches = 0;
    info->breakdown = 0;
    return 0;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

int lsspg_solver_opts_default(lsspg_solver_opts *o)
{
    if (!o) return 1;
    o->tol_rel = kDefRtol; o->tol_abs = kDefAtol; o->tol_rb = kDefRb;
    o->maxit = kDefMaxit; o->restart = kDefRestart; o->aug_k = kDefAugK; o->bgsl = kDefBgsl; o->idrs = kDefIdrs;
    o->verb = 0; o->hist_len = 0; o->hist = nullptr;
    return 0;
}

int lsspg_solver_supported(int solver) { return solver >= LSSPG_GMRES && solver <= LSSPG_IDRS; }

int lsspg_krylov_solve(lsspg_ctx *ctx, int solver, const lsspg_csr *A, lsspg_pc *pc, const double *db, double *dx,
                       const lsspg_solver_opts *opts, lsspg_solve_info *info)
{
    KrylovArgs k;
    LSSPG_TRY(resolve(k, ctx, A, pc, db, dx, opts, info));
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const long long l0 = ctx->launches;
    LSSPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    int rc;
    switch (solver) {
        case LSSPG_CG: rc = krylov_cg(k); break;
        case LSSPG_BICGSTAB: rc = krylov_bicgstab(k); break;
        case LSSPG_GMRES: rc = krylov_gmres(k); break;
        case LSSPG_IDRS: rc = krylov_idrs(k, opts); break;
        case LSSPG_CGS: rc = krylov_cgs(k, opts); break;
        case LSSPG_CR: rc = krylov_cr(k, opts); break;
        case LSSPG_CRS: rc = krylov_crs(k, opts); break;
        case LSSPG_BICRSTAB: rc = krylov_bicrstab(k, opts); break;
        case LSSPG_TFQMR: rc = krylov_tfqmr(k, opts); break;
        case LSSPG_QMRCGSTAB: rc = krylov_qmrcgstab(k); break;
        case LSSPG_ORTHOMIN: rc = krylov_orthomin(k); break;
        case LSSPG_BICGSAFE: rc = krylov_bicgsafe(k, opts); break;
        case LSSPG_BICRSAFE: rc = krylov_bicrsafe(k, opts); break;
        case LSSPG_GPBICG: rc = krylov_gpbicg(k, opts); break;
        case LSSPG_GPBICR: rc = krylov_gpbicr(k, opts); break;
        case LSSPG_BICGSTABL: rc = krylov_bicgstabl(k, opts); break;
        case LSSPG_RGMRES: rc = krylov_rgmres(k); break;
        case LSSPG_LGMRES: rc = krylov_lgmres(k); break;
        case LSSPG_RLGMRES: rc = krylov_rlgmres(k); break;
        default:
            set_error("lsspg_krylov_solve: solver %d is not implemented", solver);
            return 1;
    }
    if (rc) return rc;
    LSSPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    LSSPG_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    LSSPG_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    info->solve_ms = ms;
    info->launches = ctx->launches - l0;
    return 0;
}

int lsspg_krylov_solve_host(lsspg_ctx *ctx, int solver, const lsspg_csr *A, lsspg_pc *pc, const double *hb,
                            double *hx, const lsspg_solver_opts *opts, lsspg_solve_info *info)
{
    LSSPG_CHECK(ctx && A && hb && hx, "lsspg_krylov_solve_host: NULL argument");
    const size_t n = A->num_rows;
    // x and b of the running solve get vectors of their OWN (from the work-vector pool), not the context's staging
    // buffers: a user preconditioner runs on the host inside the solve (pc.cu: LSSPG_PC_USER) and may call lssp_mv_*,
    // lssp_pc_ilu_solve, amg_solve -- all of which stage through ctx->stage[] (the reference allows such calls from
    // pc.solve, include/type-defs.h:104).
    Workspace ws(ctx, (long long)(A->num_cols > A->num_rows ? A->num_cols : A->num_rows));
    double *dx = ws.vec(), *db = ws.vec();
    LSSPG_CHECK(dx && db, "lsspg_krylov_solve_host: out of device memory");
    LSSPG_CUDA(cudaMemcpyAsync(dx, hx, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(db, hb, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_TRY(lsspg_krylov_solve(ctx, solver, A, pc, db, dx, opts, info));
    LSSPG_CUDA(cudaMemcpyAsync(hx, dx, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"

// pc.cu -- preconditioner application objects (replaces LSSP_PC.solve,
// include/type-defs.h:104,144): NON (src/pc.cxx:67-70), ILU = L then U sweep
// (src/solver-tri.cxx:48-60), block-ILU = L sweep, D SpMV, U sweep
// (src/pc-biluk.cxx:22-60).
#include "blas1.cuh"
#include "pc.cuh"
#include "spmv.cuh"
#include "tri.cuh"

namespace lsspg {

int pc_apply(lsspg_ctx *ctx, lsspg_pc *pc, double *dx, const double *drhs, bool guarded)
{
    LSSPG_CHECK(pc && dx && drhs, "pc_apply: NULL operand");
    switch (pc->kind) {
        case LSSPG_PC_NON:
            return vec_copy(ctx, pc->n, dx, drhs);
        case LSSPG_PC_ILU:
            LSSPG_TRY(tri_solve(ctx, pc->L, pc->cache, drhs, guarded));
            return tri_solve(ctx, pc->U, dx, pc->cache, guarded);
        case LSSPG_PC_BILU: {
            double *y = pc->cache, *z = pc->cache + pc->n;
            LSSPG_TRY(tri_solve(ctx, pc->L, y, drhs, guarded));
            LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, pc->D, coef_imm(1.0), y, coef_imm(0.0), nullptr, z, nullptr, guarded));
            return tri_solve(ctx, pc->U, dx, z, guarded);
        }
        case LSSPG_PC_AMG:
            return amg_cycle(ctx, pc->amg, dx, drhs, guarded);
        case LSSPG_PC_USER: {
            // a user-supplied pc.solve works on host vectors: round trip per application (documented slow path)
            const size_t nb = sizeof(double) * (size_t)pc->n;
            LSSPG_CUDA(cudaMemcpyAsync(pc->h_x, dx, nb, cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaMemcpyAsync(pc->h_rhs, drhs, nb, cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            pc->user_fn(pc->user, pc->h_x, pc->h_rhs, pc->n);
            LSSPG_CUDA(cudaMemcpyAsync(dx, pc->h_x, nb, cudaMemcpyHostToDevice, ctx->stream));
            return 0;
        }
        default:
            set_error("pc_apply: preconditioner kind %d is not implemented", pc->kind);
            return 1;
    }
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

int lsspg_pc_create_non(lsspg_ctx *ctx, int n, lsspg_pc **out)
{
    LSSPG_CHECK(ctx && out && n >= 0, "lsspg_pc_create_non: bad argument");
    lsspg_pc *pc = new lsspg_pc();
    pc->kind = LSSPG_PC_NON;
    pc->n = n;
    pc->bytes = 16.0 * n;
    *out = pc;
    return 0;
}

int lsspg_pc_create_ilu(lsspg_ctx *ctx, int n, const int *Lp, const int *Lj, const double *Lx, const int *Up,
                        const int *Uj, const double *Ux, lsspg_pc **out)
{
    LSSPG_CHECK(ctx && out && n >= 0, "lsspg_pc_create_ilu: bad argument");
    lsspg_pc *pc = new lsspg_pc();
    pc->kind = LSSPG_PC_ILU;
    pc->n = n;
    int rc = lsspg_tri_analyse(ctx, LSSPG_TRI_LOWER, n, Lp, Lj, Lx, &pc->L);
    if (!rc) rc = lsspg_tri_analyse(ctx, LSSPG_TRI_UPPER, n, Up, Uj, Ux, &pc->U);
    if (!rc && cudaMalloc(&pc->cache, sizeof(double) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) {
        set_error("lsspg_pc_create_ilu: out of device memory");
        rc = 1;
    }
    if (rc) {
        lsspg_pc_destroy(ctx, pc);
        return rc;
    }
    // 12 (nnz(L) + nnz(U)) + 40 n
    pc->bytes = 12.0 * ((double)Lp[n] + (double)Up[n]) + 40.0 * n;
    *out = pc;
    return 0;
}

int lsspg_pc_create_bilu(lsspg_ctx *ctx, int n, const int *Lp, const int *Lj, const double *Lx, const int *Dp,
                         const int *Dj, const double *Dx, const int *Up, const int *Uj, const double *Ux,
                         lsspg_pc **out)
{
    LSSPG_CHECK(ctx && out && n >= 0, "lsspg_pc_create_bilu: bad argument");
    lsspg_pc *pc = new lsspg_pc();
    pc->kind = LSSPG_PC_BILU;
    pc->n = n;
    int rc = lsspg_tri_analyse(ctx, LSSPG_TRI_LOWER, n, Lp, Lj, Lx, &pc->L);
    if (!rc) rc = lsspg_tri_analyse(ctx, LSSPG_TRI_UPPER, n, Up, Uj, Ux, &pc->U);
    if (!rc) rc = lsspg_csr_upload(ctx, n, n, Dp, Dj, Dx, &pc->D);
    if (!rc && cudaMalloc(&pc->cache, sizeof(double) * (size_t)(n > 0 ? 2 * (size_t)n : 1)) != cudaSuccess) {
        set_error("lsspg_pc_create_bilu: out of device memory");
        rc = 1;
    }
    if (rc) {
        lsspg_pc_destroy(ctx, pc);
        return rc;
    }
    pc->bytes = 12.0 * ((double)Lp[n] + (double)Up[n]) + 40.0 * n + 12.0 * Dp[n] + 20.0 * n;
    *out = pc;
    return 0;
}

int lsspg_pc_create_user(lsspg_ctx *ctx, int n, void (*fn)(void *, double *, const double *, int), void *user,
                         lsspg_pc **out)
{
    LSSPG_CHECK(ctx && out && fn && n >= 0, "lsspg_pc_create_user: bad argument");
    lsspg_pc *pc = new lsspg_pc();
    pc->kind = LSSPG_PC_USER;
    pc->n = n;
    pc->user_fn = fn;
    pc->user = user;
    LSSPG_CUDA(cudaMallocHost(&pc->h_x, sizeof(double) * (size_t)(n > 0 ? n : 1)));
    LSSPG_CUDA(cudaMallocHost(&pc->h_rhs, sizeof(double) * (size_t)(n > 0 ? n : 1)));
    pc->bytes = 32.0 * n;
    *out = pc;
    return 0;
}

int lsspg_pc_destroy(lsspg_ctx *ctx, lsspg_pc *pc)
{
    if (!pc) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (pc->h_x) cudaFreeHost(pc->h_x);
    if (pc->h_rhs) cudaFreeHost(pc->h_rhs);
    lsspg_tri_destroy(ctx, pc->L);
    lsspg_tri_destroy(ctx, pc->U);
    lsspg_csr_destroy(ctx, pc->D);
    amg_free(ctx, pc->amg);
    if (pc->cache) cudaFree(pc->cache);
    delete pc;
    return 0;
}

int lsspg_pc_kind(const lsspg_pc *pc) { return pc->kind; }
double lsspg_pc_bytes(const lsspg_pc *pc) { return pc->bytes; }

int lsspg_pc_info(const lsspg_pc *pc, int *levels_L, int *levels_U, long long *padded_L, long long *padded_U)
{
    if (levels_L) *levels_L = pc->L ? pc->L->num_levels : 0;
    if (levels_U) *levels_U = pc->U ? pc->U->num_levels : 0;
    if (padded_L) *padded_L = pc->L ? pc->L->padded_nnz : 0;
    if (padded_U) *padded_U = pc->U ? pc->U->padded_nnz : 0;
    return 0;
}

int lsspg_pc_apply(lsspg_ctx *ctx, lsspg_pc *pc, double *dx, const double *drhs)
{
    return pc_apply(ctx, pc, dx, drhs, false);
}

int lsspg_pc_apply_host(lsspg_ctx *ctx, lsspg_pc *pc, double *hx, const double *hrhs)
{
    LSSPG_CHECK(pc && hx && hrhs, "lsspg_pc_apply_host: NULL operand");
    const size_t n = pc->n;
    LSSPG_TRY(ensure_stage(ctx, n));
    // the incoming x is uploaded too: AMG uses it as the initial guess (src/pc-sxamg.cxx:58-64)
    LSSPG_CUDA(cudaMemcpyAsync(ctx->stage[0], hx, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(ctx->stage[1], hrhs, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_TRY(pc_apply(ctx, pc, ctx->stage[0], ctx->stage[1], false));
    LSSPG_CUDA(cudaMemcpyAsync(hx, ctx->stage[0], n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_scalars(ctx, 0, 1, true);   // syncs and surfaces device-side error flags
}

}  // extern "C"

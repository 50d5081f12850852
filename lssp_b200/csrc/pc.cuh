// pc.cuh -- preconditioner application object.
#pragma once
#include "common.cuh"

struct lsspg_tri;
struct lsspg_csr;
struct lsspg_amg;

struct lsspg_pc {
    int kind = 0;
    int n = 0;
    lsspg_tri *L = nullptr;
    lsspg_tri *U = nullptr;
    lsspg_csr *D = nullptr;     // block-ILU: block-diagonal of inverted pivot blocks
    lsspg_amg *amg = nullptr;   // AMG hierarchy (amg.cu)
    double *cache = nullptr;    // n doubles (ILU) / 2n doubles (block-ILU), as pc.cache in the reference
    double bytes = 0.0;         // algorithmic bytes of one application
    // LSSPG_PC_USER: host callback (the reference's LSSP_PC_USER hook, src/pc.cxx:219-227)
    void (*user_fn)(void *user, double *hx, const double *hrhs, int n) = nullptr;
    void *user = nullptr;
    double *h_x = nullptr, *h_rhs = nullptr;   // pinned staging for the callback
};

namespace lsspg {
// x = M^-1 rhs.  `guarded`: every kernel of the application is skipped when the
// device stop flag is set (used inside the Krylov drivers).
int pc_apply(lsspg_ctx *ctx, lsspg_pc *pc, double *dx, const double *drhs, bool guarded);
// one AMG cycle from the initial guess in dx (amg.cu)
int amg_cycle(lsspg_ctx *ctx, lsspg_amg *M, double *dx, const double *drhs, bool guarded);
void amg_free(lsspg_ctx *ctx, lsspg_amg *M);
}  // namespace lsspg

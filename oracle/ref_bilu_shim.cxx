/* TEST INFRASTRUCTURE (oracle) -- see ref_shim.cxx.  Reaches the reference's block ILU(k) set-up
 * (src/pc-biluk.cxx:62-431, compiled with USE_BLAS = USE_LAPACK = 1 against oracle/blas_standin.c) and returns
 * the three matrices its apply uses: L (unit diagonal last), D (block diagonal of the inverted pivot blocks), U (unit
 * diagonal first).  Built into oracle/_ref/liblssp_refb.so only. */
#include <string.h>
#include "lssp.h"

extern "C" {

typedef struct ref_bilu_ {
    LSSP_PC pc;
} ref_bilu;

/* as lssp_pc_biluk_assemble (src/pc-biluk.cxx:416-431): block size bs = n / num_blks; the input is column-sorted
 * first, as lssp_solver_assemble does (src/lssp.cxx:173) */
void *refb_bilu_create(int n, int *Ap, int *Aj, double *Ax, int num_blks, int level)
{
    ref_bilu *h = (ref_bilu *)calloc(1, sizeof(ref_bilu));
    lssp_mat_csr S;
    lssp_mat_bcsr B;

    S.num_rows = S.num_cols = n;
    S.num_nnzs = Ap[n];
    S.Ap = lssp_copy_on<int>(Ap, n + 1);
    S.Aj = lssp_copy_on<int>(Aj, S.num_nnzs);
    S.Ax = lssp_copy_on<double>(Ax, S.num_nnzs);
    if (!lssp_mat_csr_is_sorted(S)) lssp_mat_sort_column(S);

    h->pc.iluk_level = level;
    h->pc.verb = 0;
    B = lssp_mat_csr_to_bcsr(S, n / num_blks);
    lssp_pc_biluk_assemble_mat(h->pc, B);
    h->pc.assembled = true;
    lssp_mat_destroy(B);
    lssp_mat_destroy(S);
    return h;
}

void refb_bilu_sizes(void *hh, int *nnzL, int *nnzD, int *nnzU)
{
    ref_bilu *h = (ref_bilu *)hh;
    *nnzL = h->pc.L.num_nnzs; *nnzD = h->pc.D.num_nnzs; *nnzU = h->pc.U.num_nnzs;
}

static void get(const lssp_mat_csr &M, int *p, int *j, double *x)
{
    memcpy(p, M.Ap, sizeof(int) * (M.num_rows + 1));
    memcpy(j, M.Aj, sizeof(int) * M.num_nnzs);
    memcpy(x, M.Ax, sizeof(double) * M.num_nnzs);
}

void refb_bilu_get(void *hh, int *Lp, int *Lj, double *Lx, int *Dp, int *Dj, double *Dx, int *Up, int *Uj, double *Ux)
{
    ref_bilu *h = (ref_bilu *)hh;
    get(h->pc.L, Lp, Lj, Lx);
    get(h->pc.D, Dp, Dj, Dx);
    get(h->pc.U, Up, Uj, Ux);
}

/* x = U^-1 D L^-1 rhs through the reference's own apply (src/pc-biluk.cxx:22-60) */
void refb_bilu_apply(void *hh, int n, double *x, double *rhs)
{
    ref_bilu *h = (ref_bilu *)hh;
    lssp_vec vx, vr;
    vx.n = vr.n = n; vx.d = x; vr.d = rhs;
    h->pc.solve(&h->pc, vx, vr);
}

void refb_bilu_destroy(void *hh)
{
    ref_bilu *h = (ref_bilu *)hh;
    lssp_pc_biluk_destroy(&h->pc);
    free(h);
}

/* Whole solve with LSSP_PC_BILUK through the reference's public API (src/lssp.cxx:16-414, src/pc.cxx:124-135):
 * solver_type = the LSSP_SOLVER_TYPE value (the internal drivers come first, include/type-defs.h:156-174).
 * x: initial guess on entry, solution on return.  out[0] = solver.residual.  Returns the iteration count. */
int refb_solve_biluk(int solver_type, int n, int *Ap, int *Aj, double *Ax, double *b, double *x, int num_blks,
                     int level, double rtol, int maxit, int restart, double *out)
{
    LSSP_SOLVER s;
    LSSP_PC pc;
    lssp_mat_csr A;
    lssp_vec vx, vb;
    int nits;

    A.num_rows = A.num_cols = n;
    A.num_nnzs = Ap[n];
    A.Ap = Ap; A.Aj = Aj; A.Ax = Ax;
    vx.n = vb.n = n; vx.d = x; vb.d = b;
    lssp_verbosity = 0;
    lssp_solver_create(s, (LSSP_SOLVER_TYPE)solver_type, pc, LSSP_PC_BILUK);
    lssp_solver_set_rtol(s, rtol);
    lssp_solver_set_atol(s, rtol);
    lssp_solver_set_rbtol(s, rtol);
    lssp_solver_set_maxit(s, maxit);
    lssp_solver_set_restart(s, restart);
    lssp_pc_iluk_set_level(pc, level);
    s.num_blks = num_blks;
    lssp_solver_assemble(s, A, vx, vb, pc);
    nits = lssp_solver_solve(s, pc);
    out[0] = s.residual;
    lssp_solver_destroy(s, pc);
    return nits;
}

}

/* TEST INFRASTRUCTURE (oracle) -- never linked into, imported by, or executed
 * from the product path (lssp_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load the library
 * built from this file.
 *
 * extern "C" shim (new code) around the UNMODIFIED reference sources, which
 * are compiled where they lie under /root/reference by oracle/Makefile.  It
 * lets ctypes reach:
 *   - the reference kernels        (src/mvops.cxx, src/vector.cxx, src/solver-tri.cxx)
 *   - the reference factorisations (src/pc-iluk.cxx, src/pc-ilut.cxx), including
 *     their file-static block-diagonal drivers, reached by textually including
 *     the two .cxx files into this TU (SURVEY.md 8c "Block-Jacobi oracle")
 *   - whole solves through lssp_solver_create/assemble/solve/destroy
 *     (src/lssp.cxx:16-414), optionally with the block-Jacobi ILU installed
 *     through the LSSP_PC_USER hook (src/pc.cxx:219-227).
 */
#include "lssp.h"
#include "pc-iluk.cxx"   /* from $(REF)/src: gives access to lssp_pc_iluk_assemble_matrix */
#include "pc-ilut.cxx"   /* from $(REF)/src: gives access to lssp_pc_ilut_assemble_matrix */

extern "C" {

typedef struct ref_params_ {
    double rtol, atol, rbtol;
    int maxit, restart, augk, bgsl, idrs;
    int iluk_level, ilut_p;
    double ilut_tol;
    int blk_size;              /* >0: block-Jacobi ILU (blocks of blk_size rows) */
    int verb;
} ref_params;

typedef struct ref_ilu_ {
    lssp_mat_csr L, U;
    double *cache;
} ref_ilu;

static lssp_mat_csr view_csr(int n, int *Ap, int *Aj, double *Ax)
{
    lssp_mat_csr A;
    A.num_rows = A.num_cols = n;
    A.num_nnzs = Ap ? Ap[n] : 0;
    A.Ap = Ap; A.Aj = Aj; A.Ax = Ax;
    return A;
}

static lssp_vec view_vec(int n, double *d) { lssp_vec v; v.n = n; v.d = d; return v; }

/* ---- kernels ------------------------------------------------------------ */
/* kind: 0 mxy, 1 amxy, 2 amxpby (in-place on y), 3 amxpbyz */
void ref_mv(int kind, int n, int *Ap, int *Aj, double *Ax, double alpha, double *x,
            double beta, double *y, double *z)
{
    lssp_mat_csr A = view_csr(n, Ap, Aj, Ax);
    lssp_vec vx = view_vec(n, x), vy = view_vec(n, y), vz = view_vec(n, z);
    switch (kind) {
        case 0: lssp_mv_mxy(A, vx, vy); break;
        case 1: lssp_mv_amxy(alpha, A, vx, vy); break;
        case 2: lssp_mv_amxpby(alpha, A, vx, beta, vy); break;
        default: lssp_mv_amxpbyz(alpha, A, vx, beta, vy, vz); break;
    }
}

double ref_dot(int n, double *x, double *y) { return lssp_vec_dot(view_vec(n, x), view_vec(n, y)); }
double ref_norm(int n, double *x) { return lssp_vec_norm(view_vec(n, x)); }
void ref_axy(int n, double a, double *x, double *y) { lssp_vec_axy(a, view_vec(n, x), view_vec(n, y)); }
void ref_axpby(int n, double a, double *x, double b, double *y) { lssp_vec_axpby(a, view_vec(n, x), b, view_vec(n, y)); }
void ref_axpbyz(int n, double a, double *x, double b, double *y, double *z)
{ lssp_vec_axpbyz(a, view_vec(n, x), b, view_vec(n, y), view_vec(n, z)); }
void ref_scale(int n, double *x, double a) { lssp_vec_scale(view_vec(n, x), a); }

void ref_tri_lower(int n, int *Ap, int *Aj, double *Ax, double *x, double *rhs)
{ lssp_pc_ilu_solve_lower_matrix(view_csr(n, Ap, Aj, Ax), x, rhs); }
void ref_tri_upper(int n, int *Ap, int *Aj, double *Ax, double *x, double *rhs)
{ lssp_pc_ilu_solve_upper_matrix(view_csr(n, Ap, Aj, Ax), x, rhs); }

/* ---- factorisations ----------------------------------------------------- */
/* kind 0: ILU(k) with fill level `level`; kind 1: ILUT(p, tol) (p<=0 -> ceil(nnz/n)).
 * blk_size<=0 -> one block of n rows (the global factorisation). The input is
 * column-sorted first, exactly as lssp_solver_assemble does (src/lssp.cxx:173). */
void *ref_ilu_create(int kind, int n, int *Ap, int *Aj, double *Ax, int level, int p,
                     double tol, int blk_size)
{
    lssp_mat_csr V = view_csr(n, Ap, Aj, Ax), S, A;
    ref_ilu *h = (ref_ilu *)calloc(1, sizeof(ref_ilu));

    S = V;
    S.Ap = lssp_copy_on<int>(V.Ap, n + 1);
    S.Aj = lssp_copy_on<int>(V.Aj, V.num_nnzs);
    S.Ax = lssp_copy_on<double>(V.Ax, V.num_nnzs);
    if (!lssp_mat_csr_is_sorted(S)) lssp_mat_sort_column(S);

    A = lssp_mat_adjust_zero_diag(S, mat_zero_diag_tol);
    if (blk_size <= 0 || blk_size > n) blk_size = n;
    if (kind == 0) {
        lssp_pc_iluk_assemble_matrix(A, blk_size, level, h->L, h->U, 0);
    }
    else {
        if (p <= 0) p = (V.num_nnzs + n - 1) / n;
        lssp_pc_ilut_assemble_matrix(A, blk_size, h->L, h->U, tol, p, 0);
    }
    lssp_mat_destroy(A);
    lssp_mat_destroy(S);
    h->cache = lssp_malloc<double>(n);
    return h;
}

void ref_ilu_sizes(void *hh, int *nnzL, int *nnzU)
{ ref_ilu *h = (ref_ilu *)hh; *nnzL = h->L.num_nnzs; *nnzU = h->U.num_nnzs; }

void ref_ilu_get(void *hh, int *Lp, int *Lj, double *Lx, int *Up, int *Uj, double *Ux)
{
    ref_ilu *h = (ref_ilu *)hh;
    int n = h->L.num_rows;
    memcpy(Lp, h->L.Ap, sizeof(int) * (n + 1));
    memcpy(Lj, h->L.Aj, sizeof(int) * h->L.num_nnzs);
    memcpy(Lx, h->L.Ax, sizeof(double) * h->L.num_nnzs);
    memcpy(Up, h->U.Ap, sizeof(int) * (n + 1));
    memcpy(Uj, h->U.Aj, sizeof(int) * h->U.num_nnzs);
    memcpy(Ux, h->U.Ax, sizeof(double) * h->U.num_nnzs);
}

void ref_ilu_apply(void *hh, double *x, double *rhs)
{ ref_ilu *h = (ref_ilu *)hh; lssp_pc_ilu_solve_lu_matrix(h->L, h->U, x, rhs, h->cache); }

void ref_ilu_destroy(void *hh)
{
    ref_ilu *h = (ref_ilu *)hh;
    lssp_mat_destroy(h->L); lssp_mat_destroy(h->U); free(h->cache); free(h);
}

/* ---- whole solves ------------------------------------------------------- */
static int g_bj_blk = 0, g_bj_kind = 0;

static void bj_destroy(LSSP_PC *pc)
{
    lssp_mat_destroy(pc->L); lssp_mat_destroy(pc->U); lssp_free<double>(pc->cache);
}

/* LSSP_PC_USER assemble: the reference's own blocked factorisation with
 * blk_size < n (it is only ever called with blk_size = n inside the library). */
static void bj_assemble(LSSP_PC &pc, LSSP_SOLVER s)
{
    lssp_mat_csr A = lssp_mat_adjust_zero_diag(s.A, mat_zero_diag_tol);
    if (g_bj_kind == 0) {
        lssp_pc_iluk_assemble_matrix(A, g_bj_blk, pc.iluk_level, pc.L, pc.U, 0);
    }
    else {
        if (pc.ilut_p <= 0) pc.ilut_p = (s.A.num_nnzs + s.A.num_rows - 1) / s.A.num_rows;
        lssp_pc_ilut_assemble_matrix(A, g_bj_blk, pc.L, pc.U, pc.ilut_tol, pc.ilut_p, 0);
    }
    lssp_mat_destroy(A);
    pc.cache = lssp_malloc<double>(pc.L.num_rows);
    pc.solve = lssp_pc_ilu_solve;
    pc.destroy = bj_destroy;
}

/* solver_type / pc_type: the reference enum values with every USE_* = 0
 * (include/type-defs.h:156-174 and :64-98): pc 0 NON, 1 ILUK, 2 ILUT.
 * x is the initial guess on entry and the solution on return.
 * out[0] = solver.residual, out[1] = assemble seconds, out[2] = solve seconds. */
int ref_solve(int solver_type, int pc_type, int n, int *Ap, int *Aj, double *Ax,
              double *b, double *x, const ref_params *prm, double *out)
{
    LSSP_SOLVER s;
    LSSP_PC pc;
    lssp_mat_csr A = view_csr(n, Ap, Aj, Ax);
    int user = (prm->blk_size > 0 && prm->blk_size < n && (pc_type == 1 || pc_type == 2));
    int nits;
    double t0, t1, t2;

    lssp_verbosity = prm->verb;
    lssp_solver_create(s, (LSSP_SOLVER_TYPE)solver_type, pc,
                       user ? LSSP_PC_USER : (LSSP_PC_TYPE)pc_type);
    lssp_solver_set_rtol(s, prm->rtol);
    lssp_solver_set_atol(s, prm->atol);
    lssp_solver_set_rbtol(s, prm->rbtol);
    lssp_solver_set_maxit(s, prm->maxit);
    lssp_solver_set_restart(s, prm->restart);
    lssp_solver_set_augk(s, prm->augk);
    lssp_solver_set_bgsl(s, prm->bgsl);
    lssp_solver_set_idrs(s, prm->idrs);
    lssp_pc_iluk_set_level(pc, prm->iluk_level);
    lssp_pc_ilut_set_p(pc, prm->ilut_p);
    if (prm->ilut_tol >= 0) lssp_pc_ilut_set_drop_tol(pc, prm->ilut_tol);
    if (user) {
        g_bj_blk = prm->blk_size;
        g_bj_kind = pc_type - 1;
        pc.assemble = bj_assemble;
    }

    t0 = lssp_get_time();
    lssp_solver_assemble(s, A, view_vec(n, x), view_vec(n, b), pc);
    t1 = lssp_get_time();
    nits = lssp_solver_solve(s, pc);
    t2 = lssp_get_time();

    out[0] = s.residual;
    out[1] = t1 - t0;
    out[2] = t2 - t1;
    lssp_solver_destroy(s, pc);
    return nits;
}

} /* extern "C" */

/* ---- persistent solver (bench.py reference arm: time lssp_solver_solve only) ---- */
extern "C" {

typedef struct ref_session_ {
    LSSP_SOLVER s;
    LSSP_PC pc;
    lssp_vec x, b;
    int n;
} ref_session;

/* create + assemble once (src/lssp.cxx:16-189); the matrix is deep-copied by the reference */
void *ref_session_create(int solver_type, int pc_type, int n, int *Ap, int *Aj, double *Ax,
                         const ref_params *prm, double *assemble_seconds)
{
    ref_session *h = (ref_session *)calloc(1, sizeof(ref_session));
    lssp_mat_csr A = view_csr(n, Ap, Aj, Ax);
    double t0;
    lssp_verbosity = prm->verb;
    h->n = n;
    h->x = lssp_vec_create(n);
    h->b = lssp_vec_create(n);
    lssp_solver_create(h->s, (LSSP_SOLVER_TYPE)solver_type, h->pc, (LSSP_PC_TYPE)pc_type);
    lssp_solver_set_rtol(h->s, prm->rtol);
    lssp_solver_set_atol(h->s, prm->atol);
    lssp_solver_set_rbtol(h->s, prm->rbtol);
    lssp_solver_set_maxit(h->s, prm->maxit);
    lssp_solver_set_restart(h->s, prm->restart);
    lssp_solver_set_augk(h->s, prm->augk);
    lssp_solver_set_bgsl(h->s, prm->bgsl);
    lssp_solver_set_idrs(h->s, prm->idrs);
    lssp_pc_iluk_set_level(h->pc, prm->iluk_level);
    lssp_pc_ilut_set_p(h->pc, prm->ilut_p);
    if (prm->ilut_tol >= 0) lssp_pc_ilut_set_drop_tol(h->pc, prm->ilut_tol);
    t0 = lssp_get_time();
    lssp_solver_assemble(h->s, A, h->x, h->b, h->pc);
    if (assemble_seconds) *assemble_seconds = lssp_get_time() - t0;
    return h;
}

/* one lssp_solver_solve from the given x0 with the given maxit; returns nits */
int ref_session_solve(void *hh, const double *b, double *x, int maxit, double *residual, double *seconds)
{
    ref_session *h = (ref_session *)hh;
    double t0;
    int nits;
    memcpy(h->b.d, b, sizeof(double) * h->n);
    memcpy(h->x.d, x, sizeof(double) * h->n);
    lssp_solver_set_maxit(h->s, maxit);
    t0 = lssp_get_time();
    nits = lssp_solver_solve(h->s, h->pc);
    if (seconds) *seconds = lssp_get_time() - t0;
    memcpy(x, h->x.d, sizeof(double) * h->n);
    if (residual) *residual = h->s.residual;
    return nits;
}

/* one reference SpMV / ILU application on the session's own copies (micro-timings) */
double ref_session_time_mxy(void *hh, int reps)
{
    ref_session *h = (ref_session *)hh;
    lssp_vec y = lssp_vec_create(h->n);
    double t0 = lssp_get_time();
    for (int r = 0; r < reps; r++) lssp_mv_mxy(h->s.A, h->b, y);
    t0 = lssp_get_time() - t0;
    lssp_vec_destroy(y);
    return t0 / reps;
}

double ref_session_time_pc(void *hh, int reps)
{
    ref_session *h = (ref_session *)hh;
    lssp_vec y = lssp_vec_create(h->n);
    double t0 = lssp_get_time();
    for (int r = 0; r < reps; r++) h->pc.solve(&h->pc, y, h->b);
    t0 = lssp_get_time() - t0;
    lssp_vec_destroy(y);
    return t0 / reps;
}

void ref_session_destroy(void *hh)
{
    ref_session *h = (ref_session *)hh;
    lssp_solver_destroy(h->s, h->pc);
    lssp_vec_destroy(h->x);
    lssp_vec_destroy(h->b);
    free(h);
}

} /* extern "C" */

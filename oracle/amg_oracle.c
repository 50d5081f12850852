/* TEST INFRASTRUCTURE (oracle) -- plain-C CPU restatement of the SX-AMG-style cycle.
 *
 * PARITY UNPINNED.  The reference reaches its AMG through libsxamg (sx_solver_amg_solve,
 * src/pc-sxamg.cxx:64; sx_solver_amg, src/solver-sxamg.cxx:75), which is NOT part of the
 * reference tree, has no pinned version (configure.ac:1009-1031 only probes for an installed
 * copy) and is exercised by no reference test.  What is restated here is the published
 * algorithm of that library -- a classical Ruge-Stueben V-cycle with C/F-ordered Gauss-Seidel
 * smoothing (specification in DESIGN.md "AMG") -- as serial loops.  The GPU
 * cycle (lssp_b200/csrc/amg.cu) is checked against THIS file; agreement with libsxamg itself is
 * not verifiable here.  What the reference's own adapter code does guarantee is restated
 * faithfully: the cycle starts from the caller's x (src/pc-sxamg.cxx:58-64) and the stand-alone
 * solver returns (cycles, ||b - A x||) (src/solver-sxamg.cxx:96-98).
 *
 * The hierarchy (A_l, P_l, R_l, C/F marks, dense inverse of the last operator) is an INPUT: it
 * is set up by the product's host code and checked on its own in tests/test_amg_host.py.
 * Arithmetic contract as oracle.c: IEEE fp64, no FMA, sequential sums.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

void orc_mv(int kind, int n, const int *Ap, const int *Aj, const double *Ax, double alpha, const double *x,
            double beta, const double *y, double *z);
double orc_norm(int n, const double *x);

typedef struct orc_amg_level_ {
    int n, nc;
    const int *Ap, *Aj; const double *Ax;
    const int *Pp, *Pj; const double *Px;
    const int *Rp, *Rj; const double *Rx;
    const int *cf, *rank;
    double *x, *b, *r;
} orc_amg_level;

typedef struct orc_amg_ {
    int nl, pre, post, cf_order, coarse_dense, coarse_sweeps, zero_guess;
    const double *inv;      /* row-major inverse of the last operator */
    orc_amg_level *lv;
} orc_amg;

/* One Gauss-Seidel sweep in place, sequential: pre-smoothing visits the C points in ascending
 * order and then the F points, post-smoothing the F points and then the C points; cf == NULL is
 * the natural order.  Row update: t = b_i; t -= a_ij x_j for j != i in column order; x_i = t / a_ii. */
void orc_gs_sweep_ranked(int n, const int *Ap, const int *Aj, const double *Ax, const int *cf, const int *rank,
                         int post, const double *b, double *x)
{
    /* rank (a permutation of 0..n-1, NULL = identity) is the visiting order inside a block: the
     * multicolour variant (cf_order 2) visits a block colour by colour instead of by index */
    int pass, q, k;
    int *seq = NULL;
    if (rank != NULL) {
        seq = malloc(sizeof(int) * (n > 0 ? n : 1));
        for (q = 0; q < n; q++) seq[rank[q]] = q;
    }
    for (pass = 0; pass < 2; pass++) {
        const int want = post ? pass : 1 - pass;
        if (cf == NULL && pass == 1) break;
        for (q = 0; q < n; q++) {
            const int i = seq ? seq[q] : q;
            double t, d = 0.;
            if (cf != NULL && cf[i] != want) continue;
            t = b[i];
            for (k = Ap[i]; k < Ap[i + 1]; k++) {
                if (Aj[k] != i) t -= Ax[k] * x[Aj[k]];
                else d = Ax[k];
            }
            x[i] = t / d;
        }
    }
    free(seq);
}

void orc_gs_sweep(int n, const int *Ap, const int *Aj, const double *Ax, const int *cf, int post,
                  const double *b, double *x)
{
    orc_gs_sweep_ranked(n, Ap, Aj, Ax, cf, NULL, post, b, x);
}

orc_amg *orc_amg_create(int nl, int pre, int post, int cf_order, int coarse_dense, int coarse_sweeps,
                        int zero_guess, const double *inv)
{
    orc_amg *m = calloc(1, sizeof(orc_amg));
    m->nl = nl; m->pre = pre; m->post = post; m->cf_order = cf_order;
    m->coarse_dense = coarse_dense; m->coarse_sweeps = coarse_sweeps; m->zero_guess = zero_guess; m->inv = inv;
    m->lv = calloc(nl, sizeof(orc_amg_level));
    return m;
}

void orc_amg_set_level(orc_amg *m, int l, int n, int nc, const int *Ap, const int *Aj, const double *Ax,
                       const int *Pp, const int *Pj, const double *Px, const int *Rp, const int *Rj,
                       const double *Rx, const int *cf, const int *rank)
{
    orc_amg_level *L = &m->lv[l];
    L->n = n; L->nc = nc;
    L->Ap = Ap; L->Aj = Aj; L->Ax = Ax; L->Pp = Pp; L->Pj = Pj; L->Px = Px; L->Rp = Rp; L->Rj = Rj; L->Rx = Rx;
    L->cf = cf;
    L->rank = rank;
    L->x = calloc(n > 0 ? n : 1, sizeof(double));
    L->b = calloc(n > 0 ? n : 1, sizeof(double));
    L->r = calloc(n > 0 ? n : 1, sizeof(double));
}

void orc_amg_destroy(orc_amg *m)
{
    int l;
    if (m == NULL) return;
    for (l = 0; l < m->nl; l++) { free(m->lv[l].x); free(m->lv[l].b); free(m->lv[l].r); }
    free(m->lv);
    free(m);
}

/* one V-cycle from the initial guess in x */
void orc_amg_cycle(orc_amg *m, double *x, const double *rhs)
{
    int l, s, i, j;
    const int last = m->nl - 1;
    if (m->zero_guess) memset(x, 0, sizeof(double) * m->lv[0].n);
    for (l = 0; l < last; l++) {
        orc_amg_level *L = &m->lv[l], *C = &m->lv[l + 1];
        double *xl = l ? L->x : x;
        const double *bl = l ? L->b : rhs;
        const int *cf = m->cf_order ? L->cf : NULL;
        for (s = 0; s < m->pre; s++) orc_gs_sweep_ranked(L->n, L->Ap, L->Aj, L->Ax, cf, cf ? L->rank : NULL, 0, bl, xl);
        orc_mv(3, L->n, L->Ap, L->Aj, L->Ax, -1., xl, 1., bl, L->r);       /* r = b - A x */
        orc_mv(0, L->nc, L->Rp, L->Rj, L->Rx, 1., L->r, 0., NULL, C->b);   /* b_c = R r */
        for (i = 0; i < C->n; i++) C->x[i] = 0.;
    }
    {
        orc_amg_level *L = &m->lv[last];
        double *xl = last ? L->x : x;
        const double *bl = last ? L->b : rhs;
        if (m->coarse_dense) {
            for (i = 0; i < L->n; i++) {
                double sum = 0.;
                for (j = 0; j < L->n; j++) sum += m->inv[(size_t)i * L->n + j] * bl[j];
                L->r[i] = sum;
            }
            memcpy(xl, L->r, sizeof(double) * L->n);
        }
        else
            for (s = 0; s < m->coarse_sweeps; s++) orc_gs_sweep(L->n, L->Ap, L->Aj, L->Ax, NULL, 0, bl, xl);
    }
    for (l = last - 1; l >= 0; l--) {
        orc_amg_level *L = &m->lv[l], *C = &m->lv[l + 1];
        double *xl = l ? L->x : x;
        const double *bl = l ? L->b : rhs;
        const int *cf = m->cf_order ? L->cf : NULL;
        orc_mv(2, L->n, L->Pp, L->Pj, L->Px, 1., C->x, 1., xl, xl);        /* x += P x_c */
        for (s = 0; s < m->post; s++) orc_gs_sweep_ranked(L->n, L->Ap, L->Aj, L->Ax, cf, cf ? L->rank : NULL, 1, bl, xl);
    }
}

/* stand-alone iteration: cycles until ||b - A x|| / ||b|| <= tol; returns the number of cycles */
int orc_amg_solve(orc_amg *m, const double *b, double *x, double tol, int maxit, double *ares)
{
    orc_amg_level *L = &m->lv[0];
    double *r = malloc(sizeof(double) * (L->n > 0 ? L->n : 1));
    double bnorm = orc_norm(L->n, b), res, denom;
    int it = 0;
    denom = bnorm > 1e-20 ? bnorm : 1e-20;
    orc_mv(3, L->n, L->Ap, L->Aj, L->Ax, -1., x, 1., b, r);
    res = orc_norm(L->n, r);
    while (it < maxit && res / denom > tol) {
        const int keep = m->zero_guess;   /* the iteration always continues from the current x */
        m->zero_guess = 0;
        orc_amg_cycle(m, x, b);
        m->zero_guess = keep;
        orc_mv(3, L->n, L->Ap, L->Aj, L->Ax, -1., x, 1., b, r);
        res = orc_norm(L->n, r);
        it++;
    }
    free(r);
    *ares = res;
    return it;
}

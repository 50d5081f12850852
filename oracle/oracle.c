/* TEST INFRASTRUCTURE (oracle) -- plain-C CPU restatement of the LSSP solve-loop
 * hot path.  Never linked into, imported by or executed from the product path
 * (lssp_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so.
 *
 * PARITY PINNED: every function here is checked in tests/test_oracle.py against
 * (i) the unmodified reference compiled into oracle/_ref (when present) and
 * (ii) the golden fixtures under tests/golden/ that were generated from it by
 * tests/golden/make_golden.py, plus the known-answer tables of SURVEY.md App. A.
 *
 * Arithmetic contract: IEEE fp64, NO fused
 * multiply-add (build with -ffp-contract=off, no -march), sums strictly
 * sequential in storage order, epilogue operand order as in the reference.
 * Each function cites the reference lines (relative to /root/reference) that
 * it restates.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- mvops -- */
/* kind 0: y = A x            src/mvops.cxx:118-150
 * kind 1: y = a A x          src/mvops.cxx:81-115    (y = sum * a)
 * kind 2: y = y b + a A x    src/mvops.cxx:5-39      (y = sum * a + y * b)
 * kind 3: z = y b + a A x    src/mvops.cxx:42-78     (z = y * b + a * sum)
 * Ap == NULL is the reference's "zero matrix" branch. */
void orc_mv(int kind, int n, const int *Ap, const int *Aj, const double *Ax, double alpha,
            const double *x, double beta, const double *y, double *z)
{
    int i, jj;
    for (i = 0; i < n; i++) {
        double sum = 0.;
        if (Ap != NULL)
            for (jj = Ap[i]; jj < Ap[i + 1]; jj++) sum += x[Aj[jj]] * Ax[jj];
        switch (kind) {
            case 0: z[i] = sum; break;
            case 1: z[i] = sum * alpha; break;
            case 2: z[i] = (Ap != NULL) ? sum * alpha + y[i] * beta : y[i] * beta; break;
            default: z[i] = (Ap != NULL) ? y[i] * beta + alpha * sum : y[i] * beta; break;
        }
    }
}

/* --------------------------------------------------------------- vector -- */
double orc_dot(int n, const double *x, const double *y)           /* src/vector.cxx:123-133 */
{ double s = 0; int i; for (i = 0; i < n; i++) s += x[i] * y[i]; return s; }
double orc_norm(int n, const double *x) { return sqrt(orc_dot(n, x, x)); }   /* :135-138 */
void orc_axy(int n, double a, const double *x, double *y)          /* :86-96  */
{ int i; for (i = 0; i < n; i++) y[i] = x[i] * a; }
void orc_axpby(int n, double a, const double *x, double b, double *y)         /* :98-108 */
{ int i; for (i = 0; i < n; i++) y[i] = y[i] * b + x[i] * a; }
void orc_axpbyz(int n, double a, const double *x, double b, const double *y, double *z) /* :110-121 */
{ int i; for (i = 0; i < n; i++) z[i] = y[i] * b + x[i] * a; }
void orc_scale(int n, double *x, double a)                         /* :141-146 */
{ int i; for (i = 0; i < n; i++) x[i] *= a; }

/* ----------------------------------------------------------- solver-tri -- */
/* forward solve, diagonal stored LAST in each row: src/solver-tri.cxx:4-24 */
void orc_tri_lower(int n, const int *Ap, const int *Aj, const double *Ax, double *x, const double *rhs)
{
    int i, j;
    for (i = 0; i < n; i++) {
        int end = Ap[i + 1] - 1;
        double r = rhs[i];
        for (j = Ap[i]; j < end; j++) r = r - Ax[j] * x[Aj[j]];
        x[i] = r / Ax[end];
    }
}

/* backward solve, diagonal stored FIRST, off-diagonals visited in DESCENDING
 * storage order: src/solver-tri.cxx:26-46 */
void orc_tri_upper(int n, const int *Ap, const int *Aj, const double *Ax, double *x, const double *rhs)
{
    int i, j;
    for (i = n - 1; i >= 0; i--) {
        int end = Ap[i];
        double r = rhs[i];
        for (j = Ap[i + 1] - 1; j > end; j--) r = r - Ax[j] * x[Aj[j]];
        x[i] = r / Ax[end];
    }
}

/* x = U^-1 (L^-1 rhs): src/solver-tri.cxx:48-60 */
void orc_ilu_apply(int n, const int *Lp, const int *Lj, const double *Lx, const int *Up,
                   const int *Uj, const double *Ux, double *x, const double *rhs, double *cache)
{
    orc_tri_lower(n, Lp, Lj, Lx, cache, rhs);
    orc_tri_upper(n, Up, Uj, Ux, x, cache);
}

/* block-ILU apply: y = L^-1 rhs, z = D y (SpMV with the block-diagonal of
 * inverted pivot blocks), x = U^-1 z; L/U stored like the ILU factors (diag
 * last / diag first, upper applied in descending order):
 * src/pc-biluk.cxx:22-60.  cache holds 2n doubles. */
void orc_bilu_apply(int n, const int *Lp, const int *Lj, const double *Lx, const int *Dp,
                    const int *Dj, const double *Dx, const int *Up, const int *Uj,
                    const double *Ux, double *x, const double *rhs, double *cache)
{
    double *y = cache, *z = cache + n;
    orc_tri_lower(n, Lp, Lj, Lx, y, rhs);
    orc_mv(0, n, Dp, Dj, Dx, 1., y, 0., NULL, z);
    orc_tri_upper(n, Up, Uj, Ux, x, z);
}

/* ------------------------------------------------------ Krylov drivers --- */
/* Preconditioner handed to the drivers: kind 0 = none (memcpy, src/pc.cxx:67-70),
 * kind 1 = ILU apply with the given L/U (src/solver-tri.cxx:57-60), kind 2 = one AMG cycle
 * from the incoming x (src/pc-sxamg.cxx:42-73; amg_oracle.c, parity unpinned). */
typedef struct orc_pc_ {
    int kind;
    const int *Lp, *Lj; const double *Lx;
    const int *Up, *Uj; const double *Ux;
    double *cache;
    void *amg;
} orc_pc;
struct orc_amg_;
void orc_amg_cycle(struct orc_amg_ *m, double *x, const double *rhs);

static void pc_apply(const orc_pc *pc, int n, double *x, const double *rhs)
{
    if (pc == NULL || pc->kind == 0) memcpy(x, rhs, sizeof(double) * n);
    else if (pc->kind == 2) orc_amg_cycle((struct orc_amg_ *)pc->amg, x, rhs);
    else orc_ilu_apply(n, pc->Lp, pc->Lj, pc->Lx, pc->Up, pc->Uj, pc->Ux, x, rhs, pc->cache);
}

typedef struct orc_opts_ {
    double rtol, atol, rbtol;
    int maxit;
} orc_opts;

/* PCG: src/solver-cg.cxx:8-136.  hist[k] (k < nhist) receives ||r|| after
 * iteration k; *residual = solver.residual; returns nits. */
int orc_cg(int n, const int *Ap, const int *Aj, const double *Ax, const double *b, double *x,
           const orc_pc *pc, const orc_opts *o, double *residual, double *hist, int nhist)
{
    double *z = malloc(sizeof(double) * n), *r = malloc(sizeof(double) * n);
    double *p = malloc(sizeof(double) * n), *q = malloc(sizeof(double) * n);
    double rho0 = 0, rho1, beta, alpha, res, tol, bnorm, tol_rb;
    int it, i;

    bnorm = orc_norm(n, b);
    tol_rb = o->rbtol * bnorm;
    orc_mv(3, n, Ap, Aj, Ax, -1., x, 1., b, r);                 /* :59 */
    res = orc_norm(n, r);
    if (res <= o->atol) { it = 0; goto end; }                    /* :61-64 */
    tol = o->rtol * res;
    if (tol < o->atol) tol = o->atol;
    if (tol < tol_rb) tol = tol_rb;
    for (i = 0; i < n; i++) z[i] = 0;                            /* :72-74, once */

    for (it = 0; it < o->maxit; it++) {
        pc_apply(pc, n, z, r);                                   /* :79 */
        rho1 = orc_dot(n, z, r);
        if (it == 0) for (i = 0; i < n; i++) p[i] = z[i];
        else {
            beta = rho1 / rho0;
            for (i = 0; i < n; i++) p[i] = z[i] + beta * p[i];   /* :91 */
        }
        orc_mv(0, n, Ap, Aj, Ax, 1., p, 0., NULL, q);
        alpha = orc_dot(n, q, p);
        alpha = rho1 / alpha;
        rho0 = rho1;
        for (i = 0; i < n; i++) {                                /* :101-104 */
            x[i] = x[i] + alpha * p[i];
            r[i] = r[i] - alpha * q[i];
        }
        res = orc_norm(n, r);
        if (hist && it < nhist) hist[it] = res;
        if (res <= tol) break;
    }
    if (it < o->maxit) it += 1;                                  /* :117 */
end:
    *residual = res;
    free(z); free(r); free(p); free(q);
    return it;
}

/* BiCGStab: src/solver-bicgstab.cxx:10-175 */
int orc_bicgstab(int n, const int *Ap, const int *Aj, const double *Ax, const double *b, double *x,
                 const orc_pc *pc, const orc_opts *o, double *residual, double *hist, int nhist)
{
    double *r = malloc(sizeof(double) * n), *rh = malloc(sizeof(double) * n);
    double *p = malloc(sizeof(double) * n), *ph = malloc(sizeof(double) * n);
    double *s = malloc(sizeof(double) * n), *sh = malloc(sizeof(double) * n);
    double *t = malloc(sizeof(double) * n), *v = malloc(sizeof(double) * n);
    double rho0 = 0, rho1 = 0, alpha = 0, beta = 0, omega = 0, res, tol, bnorm, tol_rb;
    int it, i;

    orc_mv(3, n, Ap, Aj, Ax, -1., x, 1., b, r);                 /* :70 */
    for (i = 0; i < n; i++) { rh[i] = r[i]; sh[i] = ph[i] = 0.; }
    bnorm = orc_norm(n, b);
    tol_rb = o->rbtol * bnorm;
    res = orc_norm(n, r);
    if (res <= o->atol) { it = 0; goto end; }
    tol = res * o->rtol;
    if (tol < o->atol) tol = o->atol;
    if (tol < tol_rb) tol = tol_rb;

    for (it = 0; it < o->maxit; it++) {
        rho1 = orc_dot(n, r, rh);                                /* :87 */
        if (rho1 == 0) break;                                    /* :89-92 */
        if (it == 0) for (i = 0; i < n; i++) p[i] = r[i];
        else {
            beta = (rho1 * alpha) / (rho0 * omega);              /* :99 */
            for (i = 0; i < n; i++) p[i] = r[i] + beta * (p[i] - omega * v[i]);
        }
        rho0 = rho1;
        for (i = 0; i < n; i++) ph[i] = 0.;                      /* :107 */
        pc_apply(pc, n, ph, p);
        orc_mv(3, n, Ap, Aj, Ax, 1., ph, 0., p, v);              /* :110 (y*0 + 1*sum) */
        alpha = rho1 / orc_dot(n, rh, v);
        for (i = 0; i < n; i++) s[i] = r[i] - alpha * v[i];
        if (orc_norm(n, s) <= 1e-40) {                           /* :117-128, LSSP_BREAKDOWN */
            for (i = 0; i < n; i++) x[i] = x[i] + alpha * ph[i];
            orc_mv(3, n, Ap, Aj, Ax, -1., x, 1., b, r);
            res = orc_norm(n, r);
            break;
        }
        for (i = 0; i < n; i++) sh[i] = 0.;                      /* :130 */
        pc_apply(pc, n, sh, s);
        orc_mv(3, n, Ap, Aj, Ax, 1., sh, 0., p, t);              /* :133 */
        omega = orc_dot(n, t, s) / orc_dot(n, t, t);
        for (i = 0; i < n; i++) {                                /* :136-139 */
            x[i] = x[i] + alpha * ph[i] + omega * sh[i];
            r[i] = s[i] - omega * t[i];
        }
        res = orc_norm(n, r);
        if (hist && it < nhist) hist[it] = res;
        if (res <= tol) break;
    }
    if (it < o->maxit) it += 1;
end:
    *residual = res;
    free(r); free(rh); free(p); free(ph); free(s); free(sh); free(t); free(v);
    return it;
}

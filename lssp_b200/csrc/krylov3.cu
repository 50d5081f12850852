// krylov3.cu -- the remaining Lis-style drivers of the reference, first batch: CGS, CR, CRS,
// BiCRSTAB, TFQMR, QMRCGSTAB.  Each follows its reference file call by call (same BLAS-1
// sequence, same operand order, same aliasing of work vectors, same breakdown tests and
// iteration-count conventions -- SURVEY.md App. B.4/B.5/B.14); every call is one sm_100a kernel,
// the scalars between them are host doubles exactly as in the reference.  Work vectors are
// zero-initialised (the reference mallocs them; drivers that read a vector before writing it
// are compared against the zero-initialising oracle build).
#include "krylov.cuh"
#include "krylov_ops.cuh"

namespace lsspg {

// the stopping tolerance of the Lis-style drivers (e.g. src/solver-cgs.cxx:34-44): raw settings,
// no defaults
static int lis_tolerance(Drv &d, const double *r, double *nrm2, double *ires, double *tol, bool *done)
{
    LSSPG_TRY(d.norm(r, nrm2));
    *ires = *nrm2;
    *done = (*nrm2 <= d.raw->tol_abs);
    if (*done) return 0;
    double t = *nrm2 * d.raw->tol_rel, bn;
    LSSPG_TRY(d.norm(d.b, &bn));
    bn *= d.raw->tol_rb;
    if (t < d.raw->tol_abs) t = d.raw->tol_abs;
    if (t < bn) t = bn;
    *tol = t;
    return 0;
}

// ---- CGS: src/solver-cgs.cxx:4-133 --------------------------------------------------------------
int krylov_cgs(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *rtld = d.vec(), *p = d.vec(), *phat = d.vec(), *q = d.vec(), *qhat = d.vec(), *u = d.vec(),
           *uhat = d.vec();
    LSSPG_CHECK(d.ok(), "cgs: out of device memory");
    double *vhat = uhat;                                          // alias, :25
    double alpha = 1.0, beta, rho, rho_old = 1.0, tdot1, nrm2, ires, tol = 0;
    int iter = 0;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tolerance(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.copy(rtld, r));
    LSSPG_TRY(d.set(q, 0));
    LSSPG_TRY(d.set(p, 0));
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.dot(rtld, r, &rho));
        if (rho == 0.0) break;
        beta = (rho / rho_old);
        LSSPG_TRY(d.axpbyz(beta, q, 1, r, u));
        LSSPG_TRY(d.axpby(1, q, beta, p));
        LSSPG_TRY(d.axpby(1, u, beta, p));
        LSSPG_TRY(d.pc(phat, p));
        LSSPG_TRY(d.mxy(phat, vhat));
        LSSPG_TRY(d.dot(rtld, vhat, &tdot1));
        if (tdot1 == 0.0) break;
        alpha = rho / tdot1;
        LSSPG_TRY(d.axpbyz(-alpha, vhat, 1, u, q));
        LSSPG_TRY(d.axpbyz(1, u, 1, q, phat));
        LSSPG_TRY(d.pc(uhat, phat));
        LSSPG_TRY(d.axpby(alpha, uhat, 1, x));
        LSSPG_TRY(d.mxy(uhat, qhat));
        LSSPG_TRY(d.axpby(-alpha, qhat, 1, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report("cgs", iter, nrm2, ires, 1);
        if (tol >= nrm2) break;
        rho_old = rho;
    }
    return d.finish(iter, nrm2);
}

// ---- CR: src/solver-cr.cxx:4-115 ----------------------------------------------------------------
int krylov_cr(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *z = d.vec(), *p = d.vec(), *q = d.vec(), *qtld = d.vec(), *az = d.vec();
    LSSPG_CHECK(d.ok(), "cr: out of device memory");
    double alpha, beta, rho, dot_rq, dot_zq, nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tolerance(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.pc(p, r));
    LSSPG_TRY(d.mxy(p, q));
    LSSPG_TRY(d.copy(z, p));
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.pc(qtld, q));
        LSSPG_TRY(d.dot(qtld, q, &rho));
        if (rho == 0.0) break;
        LSSPG_TRY(d.dot(r, qtld, &dot_rq));
        alpha = dot_rq / rho;
        LSSPG_TRY(d.axpby(alpha, p, 1, x));
        LSSPG_TRY(d.axpby(-alpha, q, 1, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report("cr", iter, nrm2, ires, 1);
        if (tol >= nrm2) break;
        LSSPG_TRY(d.axpby(-alpha, qtld, 1, z));
        LSSPG_TRY(d.mxy(z, az));
        LSSPG_TRY(d.dot(az, qtld, &dot_zq));
        beta = -dot_zq / rho;
        LSSPG_TRY(d.axpby(1, z, beta, p));
        LSSPG_TRY(d.axpby(1, az, beta, q));
    }
    return d.finish(iter, nrm2);
}

// ---- CRS: src/solver-crs.cxx:4-109 (work vectors aliased as in the reference, :21-26) -------------
int krylov_crs(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *rtld = d.vec(), *p = d.vec(), *z = d.vec(), *q = d.vec(), *map = d.vec();
    LSSPG_CHECK(d.ok(), "crs: out of device memory");
    double *u = z, *uq = z, *ap = q, *auq = map;
    double alpha, beta, rho, rho_old, tdot1, nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tolerance(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.copy(p, r));
    LSSPG_TRY(d.mxy(p, rtld));
    rho_old = 1.0;
    LSSPG_TRY(d.set(q, 0.));
    LSSPG_TRY(d.set(p, 0.));
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.pc(z, r));
        LSSPG_TRY(d.dot(rtld, z, &rho));
        if (rho == 0.0) break;
        beta = rho / rho_old;
        LSSPG_TRY(d.axpbyz(beta, q, 1, z, u));
        LSSPG_TRY(d.axpby(1, q, beta, p));
        LSSPG_TRY(d.axpby(1, u, beta, p));
        LSSPG_TRY(d.mxy(p, ap));
        LSSPG_TRY(d.pc(map, ap));
        LSSPG_TRY(d.dot(rtld, map, &tdot1));
        if (tdot1 == 0.0) break;
        alpha = rho / tdot1;
        LSSPG_TRY(d.axpbyz(-alpha, map, 1, u, q));
        LSSPG_TRY(d.axpbyz(1, u, 1, q, uq));
        LSSPG_TRY(d.mxy(uq, auq));
        LSSPG_TRY(d.axpby(alpha, uq, 1, x));
        LSSPG_TRY(d.axpby(-alpha, auq, 1, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report("crs", iter, nrm2, ires, 1);
        if (tol >= nrm2) break;
        rho_old = rho;
    }
    return d.finish(iter, nrm2);
}

// ---- BiCRSTAB: src/solver-bicrstab.cxx:4-114 ------------------------------------------------------
int krylov_bicrstab(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *rtld = d.vec(), *r = d.vec(), *s = d.vec(), *ms = d.vec(), *ams = d.vec(), *p = d.vec(), *ap = d.vec(),
           *map = d.vec(), *z = d.vec();
    LSSPG_CHECK(d.ok(), "bicrstab: out of device memory");
    double alpha, beta, omega, rho, rho_old, tdot1, tdot2, nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tolerance(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.copy(p, r));
    LSSPG_TRY(d.mxy(p, rtld));
    LSSPG_TRY(d.pc(z, r));
    LSSPG_TRY(d.copy(p, z));
    LSSPG_TRY(d.dot(rtld, z, &rho_old));
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.mxy(p, ap));
        LSSPG_TRY(d.pc(map, ap));
        LSSPG_TRY(d.dot(rtld, map, &tdot1));
        alpha = rho_old / tdot1;
        LSSPG_TRY(d.axpbyz(-alpha, ap, 1, r, s));
        LSSPG_TRY(d.norm(s, &nrm2));
        if (nrm2 <= tol) {
            LSSPG_TRY(d.axpby(alpha, p, 1, x));
            break;
        }
        LSSPG_TRY(d.axpbyz(-alpha, map, 1, z, ms));
        LSSPG_TRY(d.mxy(ms, ams));
        LSSPG_TRY(d.dot(ams, s, &tdot1));
        LSSPG_TRY(d.dot(ams, ams, &tdot2));
        omega = tdot1 / tdot2;
        LSSPG_TRY(d.axpby(alpha, p, 1, x));
        LSSPG_TRY(d.axpby(omega, ms, 1, x));
        LSSPG_TRY(d.axpbyz(-omega, ams, 1, s, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report("bicrstab", iter, nrm2, ires, 2);
        if (tol >= nrm2) break;
        LSSPG_TRY(d.pc(z, r));
        LSSPG_TRY(d.dot(rtld, z, &rho));
        if (rho == 0.0) break;
        beta = (rho / rho_old) * (alpha / omega);
        LSSPG_TRY(d.axpby(-omega, map, 1, p));
        LSSPG_TRY(d.axpby(1, z, beta, p));
        rho_old = rho;
    }
    return d.finish(iter, nrm2);
}

// ---- TFQMR: src/solver-tfqmr.cxx:4-149 (the reported residual is the estimate tau*sqrt(1+m)) -------
int krylov_tfqmr(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *rtld = d.vec(), *u = d.vec(), *p = d.vec(), *dd = d.vec(), *t = d.vec(), *t1 = d.vec(),
           *q = d.vec(), *v = d.vec();
    LSSPG_CHECK(d.ok(), "tfqmr: out of device memory");
    double tau, rho, rhoold, theta, eta, beta, alpha, w, ww, wold, s, c, nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tolerance(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.copy(rtld, r));
    LSSPG_TRY(d.copy(p, r));
    LSSPG_TRY(d.copy(u, r));
    LSSPG_TRY(d.set(dd, 0.));
    LSSPG_TRY(d.pc(t, p));
    LSSPG_TRY(d.mxy(t, v));
    LSSPG_TRY(d.dot(r, rtld, &rhoold));
    LSSPG_TRY(d.norm(r, &tau));
    wold = tau;
    theta = 0.0;
    eta = 0.0;
    bool stop = false;
    while (iter <= maxiter && !stop) {
        LSSPG_TRY(d.dot(v, rtld, &s));
        if (fabs(s) == 0.0) break;
        alpha = rhoold / s;
        LSSPG_TRY(d.axpbyz(-alpha, v, 1, u, q));
        LSSPG_TRY(d.axpbyz(1, u, 1., q, t));
        LSSPG_TRY(d.pc(t1, t));
        LSSPG_TRY(d.mxy(t1, v));
        LSSPG_TRY(d.axpby(-alpha, v, 1, r));
        LSSPG_TRY(d.norm(r, &w));
        for (int m = 0; m < 2; m++) {
            if (m == 0) {
                ww = sqrt(w * wold);
                LSSPG_TRY(d.axpby(1, u, theta * theta * eta / alpha, dd));
            }
            else {
                ww = w;
                LSSPG_TRY(d.axpby(1, q, theta * theta * eta / alpha, dd));
            }
            theta = ww / tau;
            c = 1.0 / sqrt(1.0 + theta * theta);
            eta = c * c * alpha;
            tau = tau * theta * c;
            LSSPG_TRY(d.pc(t1, dd));
            LSSPG_TRY(d.axpby(eta, t1, 1, x));
            nrm2 = tau * sqrt(1.0 + m);
            d.report("tfqmr", iter, nrm2, ires, 1);
            if (tol >= nrm2) {
                stop = true;
                break;
            }
        }
        if (stop) break;
        LSSPG_TRY(d.dot(r, rtld, &rho));
        if (fabs(rho) == 0.0) break;
        beta = rho / rhoold;
        LSSPG_TRY(d.axpbyz(beta, q, 1, r, u));
        LSSPG_TRY(d.axpby(1, q, beta, p));
        LSSPG_TRY(d.axpby(1, u, beta, p));
        LSSPG_TRY(d.pc(t1, p));
        LSSPG_TRY(d.mxy(t1, v));
        rhoold = rho;
        wold = w;
        iter++;
    }
    return d.finish(iter, nrm2);
}

// ---- QMRCGSTAB: src/solver-qmrcgstab.cxx:9-186 (stops on the relative preconditioned residual) ------
int krylov_qmrcgstab(KrylovArgs &k)
{
    Drv d(k, nullptr);
    double *xk = k.x;
    double *rk = d.vec(), *br0 = d.vec(), *pk = d.vec(), *vk = d.vec(), *sk = d.vec(), *dk = d.vec(), *tk = d.vec(),
           *bdk = d.vec(), *bxk = d.vec(), *r = d.vec();
    LSSPG_CHECK(d.ok(), "qmrcgstab: out of device memory");
    double rho = 1, prho, alpha = 1, beta, omega = 1, theta = 0., btheta, b_eta, eta = 0., tau, btau;
    double residual, c, ires, rerror, b_norm, tol, tol_rb = k.tol_rb, t1, t2;
    int it;
    LSSPG_TRY(d.norm(d.b, &b_norm));
    tol_rb *= b_norm;
    LSSPG_TRY(d.resid(xk, tk));
    LSSPG_TRY(d.norm(tk, &residual));
    if (residual <= k.tol_abs) return d.finish(0, residual);
    tol = residual * k.tol_rel;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    tol = tol / residual;
    LSSPG_TRY(d.pc(rk, tk));
    LSSPG_TRY(d.copy(br0, rk));
    LSSPG_TRY(d.set(pk, 0));
    LSSPG_TRY(d.set(dk, 0));
    LSSPG_TRY(d.set(vk, 0));
    LSSPG_TRY(d.norm(rk, &tau));
    ires = tau;
    prho = rho;
    for (it = 0; it < k.maxit; it++) {
        LSSPG_TRY(d.dot(br0, rk, &rho));
        beta = rho * alpha / prho / omega;
        prho = rho;
        LSSPG_TRY(d.axpbyz(1, pk, -omega, vk, r));
        LSSPG_TRY(d.axpbyz(beta, r, 1, rk, pk));
        LSSPG_TRY(d.mxy(pk, r));
        LSSPG_TRY(d.pc(vk, r));
        LSSPG_TRY(d.dot(br0, vk, &t1));
        alpha = rho / t1;
        LSSPG_TRY(d.axpbyz(-alpha, vk, 1, rk, sk));
        LSSPG_TRY(d.norm(sk, &t1));
        btheta = t1 / tau;
        c = 1 / sqrt(1. + btheta * btheta);
        btau = tau * btheta * c;
        b_eta = c * c * alpha;
        LSSPG_TRY(d.axpbyz(1., pk, theta * theta * eta / alpha, dk, bdk));
        LSSPG_TRY(d.axpbyz(1, xk, b_eta, bdk, bxk));
        LSSPG_TRY(d.mxy(sk, r));
        LSSPG_TRY(d.pc(tk, r));
        LSSPG_TRY(d.dot(sk, tk, &t1));
        LSSPG_TRY(d.dot(tk, tk, &t2));
        omega = t1 / t2;
        LSSPG_TRY(d.axpbyz(1., sk, -omega, tk, rk));
        LSSPG_TRY(d.norm(rk, &t1));
        theta = t1 / btau;
        c = 1. / sqrt(1. + theta * theta);
        tau = btau * theta * c;
        eta = c * c * omega;
        LSSPG_TRY(d.axpbyz(1, sk, btheta * btheta * b_eta / omega, bdk, dk));
        LSSPG_TRY(d.axpbyz(1, bxk, eta, dk, xk));
        LSSPG_TRY(d.norm(rk, &t1));
        rerror = t1 / ires;
        record(k, it, rerror);
        if (k.verb >= 1) log_printf("qmrcgstab: itr: %4d, rel res: %.6e\n", it, rerror);
        if (rerror <= tol) {
            LSSPG_TRY(d.resid(xk, tk));
            LSSPG_TRY(d.norm(tk, &residual));
            break;
        }
    }
    if (it < k.maxit) it += 1;
    return d.finish(it, residual);
}

}  // namespace lsspg

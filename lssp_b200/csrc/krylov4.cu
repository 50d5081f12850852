// krylov4.cu -- second batch of reference drivers: ORTHOMIN(k), BiCGSafe, BiCRSafe, GPBiCG, GPBiCR,
// BiCGStab(l).  Same transcription rules as krylov3.cu.  The five-dot groups of the "safe" / GP
// variants (e.g. src/solver-bicgsafe.cxx:64-68) are ONE reduction kernel here (k_multidot<5>).
#include <vector>
#include "krylov.cuh"
#include "krylov_ops.cuh"

namespace lsspg {

static int lis_tol(Drv &d, const double *r, double *nrm2, double *ires, double *tol, bool *done)
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

// qsi / eta of the two-term minimisation shared by BiCGSafe, BiCRSafe, GPBiCG, GPBiCR
static void qsi_eta(int iter, const double *t, double *qsi, double *eta)
{
    if (iter == 1) {
        *qsi = t[1] / t[4];
        *eta = 0.0;
    }
    else {
        const double tmp = t[4] * t[0] - t[3] * t[3];
        *qsi = (t[0] * t[1] - t[2] * t[3]) / tmp;
        *eta = (t[4] * t[2] - t[3] * t[1]) / tmp;
    }
}

// ---- ORTHOMIN(k): src/solver-orthomin.cxx:12-180 ---------------------------------------------------
int krylov_orthomin(KrylovArgs &k)
{
    Drv d(k, nullptr);
    int kk = k.restart;
    if (kk < 0) kk = kDefRestart;
    LSSPG_CHECK(kk >= 1, "orthomin: k = %d", kk);
    double tol_rb = k.tol_rb;
    if (tol_rb < 0) tol_rb = kDefRb;
    double *x = k.x;
    double *z = d.vec(), *r = d.vec(), *s = d.vec(), *sd = d.vec();
    std::vector<double *> q(kk), p(kk);
    for (int i = 0; i < kk; i++) { q[i] = d.vec(); p[i] = d.vec(); }   // q zeroed (:70-72)
    LSSPG_CHECK(d.ok(), "orthomin: out of device memory");
    std::vector<double> b_j(kk), c_j(kk);
    double a_j, beta, b_norm, tol;
    int it;
    LSSPG_TRY(d.resid(x, z));
    LSSPG_TRY(d.set(r, 0.));
    LSSPG_TRY(d.pc(r, z));
    LSSPG_TRY(d.copy(p[0], r));
    LSSPG_TRY(d.copy(sd, r));
    LSSPG_TRY(d.norm(d.b, &b_norm));
    tol_rb *= b_norm;
    LSSPG_TRY(d.norm(z, &beta));
    if (beta <= k.tol_abs) return d.finish(0, beta);
    const double err_rel = beta;
    tol = k.tol_rel * err_rel;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    for (it = 0; it < k.maxit; it++) {
        LSSPG_TRY(d.mxy(sd, s));
        int j = it % kk;
        LSSPG_TRY(d.set(q[j], 0.));
        LSSPG_TRY(d.pc(q[j], s));
        {
            const double *xs[2] = {r, q[j]}, *ys[2] = {q[j], q[j]};
            double t2[2];
            LSSPG_TRY(d.dots(2, xs, ys, t2));
            a_j = t2[0];
            c_j[j] = t2[1];
        }
        if (fabs(c_j[j]) <= kBreakdown) break;
        a_j = a_j / c_j[j];
        LSSPG_TRY(d.axpby(a_j, p[j], 1, x));
        LSSPG_TRY(d.axpby(-a_j, q[j], 1, r));
        LSSPG_TRY(d.copy(sd, r));
        LSSPG_TRY(d.mxy(r, s));
        LSSPG_TRY(d.set(z, 0.));
        LSSPG_TRY(d.pc(z, s));
        const int upto = (it >= kk - 1) ? kk : it + 1;
        for (int i = 0; i < upto; i++) {
            LSSPG_TRY(d.dot(z, q[i], &beta));
            b_j[i] = -beta / c_j[i];
            LSSPG_TRY(d.axpby(b_j[i], p[i], 1, sd));
        }
        j = (it + 1) % kk;
        LSSPG_TRY(d.copy(p[j], sd));
        LSSPG_TRY(d.resid(x, z));
        LSSPG_TRY(d.norm(z, &beta));
        record(k, it, beta);
        if (k.verb >= 1)
            log_printf("orthomin: itr: %5d, abs res: %.6e, rel res: %.6e, rbn: %.6e\n", it, beta,
                   (err_rel == 0 ? 0 : beta / err_rel), (b_norm == 0 ? 0 : beta / b_norm));
        if (beta <= tol) break;
    }
    if (it < k.maxit) it += 1;
    return d.finish(it, beta);
}

// ---- BiCGSafe: src/solver-bicgsafe.cxx:4-155 --------------------------------------------------------
int krylov_bicgsafe(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *rtld = d.vec(), *mr = d.vec(), *amr = d.vec(), *t = d.vec(), *mt = d.vec(), *p = d.vec(),
           *ap = d.vec(), *y = d.vec(), *u = d.vec(), *au = d.vec(), *z = d.vec();   // y, u, z start at 0 (App. B.11)
    LSSPG_CHECK(d.ok(), "bicgsafe: out of device memory");
    double alpha, beta, rho, rho_old, qsi, eta, tdot[5], nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tol(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.copy(rtld, r));
    LSSPG_TRY(d.pc(mr, r));
    LSSPG_TRY(d.mxy(mr, amr));
    LSSPG_TRY(d.dot(rtld, r, &rho_old));
    LSSPG_TRY(d.copy(ap, amr));
    LSSPG_TRY(d.copy(p, mr));
    beta = 0.0;
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.dot(rtld, ap, &tdot[0]));
        alpha = rho_old / tdot[0];
        {
            const double *xs[5] = {y, amr, y, amr, amr}, *ys[5] = {y, r, r, y, amr};
            LSSPG_TRY(d.dots(5, xs, ys, tdot));
        }
        qsi_eta(iter, tdot, &qsi, &eta);
        LSSPG_TRY(d.copy(t, y));
        LSSPG_TRY(d.scale(t, eta));
        LSSPG_TRY(d.axpby(qsi, ap, 1, t));
        LSSPG_TRY(d.pc(mt, t));
        LSSPG_TRY(d.axpby(1, mt, eta * beta, u));
        LSSPG_TRY(d.mxy(u, au));
        LSSPG_TRY(d.scale(z, eta));
        LSSPG_TRY(d.axpby(qsi, mr, 1, z));
        LSSPG_TRY(d.axpby(-alpha, u, 1, z));
        LSSPG_TRY(d.scale(y, eta));
        LSSPG_TRY(d.axpby(qsi, amr, 1, y));
        LSSPG_TRY(d.axpby(-alpha, au, 1, y));
        LSSPG_TRY(d.axpby(alpha, p, 1, x));
        LSSPG_TRY(d.axpby(1, z, 1, x));
        LSSPG_TRY(d.axpby(-alpha, ap, 1, r));
        LSSPG_TRY(d.axpby(-1, y, 1.0, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report("bicgsafe", iter, nrm2, ires, 1);
        if (tol >= nrm2) break;
        LSSPG_TRY(d.dot(rtld, r, &rho));
        if (rho == 0.0) break;
        beta = (rho / rho_old) * (alpha / qsi);
        LSSPG_TRY(d.pc(mr, r));
        LSSPG_TRY(d.mxy(mr, amr));
        LSSPG_TRY(d.axpby(-1, u, 1.0, p));
        LSSPG_TRY(d.axpby(1, mr, beta, p));
        LSSPG_TRY(d.axpby(-1, au, 1.0, ap));
        LSSPG_TRY(d.axpby(1, amr, beta, ap));
        rho_old = rho;
    }
    return d.finish(iter, nrm2);
}

// ---- BiCRSafe: src/solver-bicrsafe.cxx:4-151 --------------------------------------------------------
int krylov_bicrsafe(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *rtld = d.vec(), *artld = d.vec(), *mr = d.vec(), *amr = d.vec(), *p = d.vec(), *ap = d.vec(),
           *map = d.vec(), *my = d.vec(), *y = d.vec(), *u = d.vec(), *au = d.vec(), *z = d.vec();
    LSSPG_CHECK(d.ok(), "bicrsafe: out of device memory");
    double alpha, beta, rho, rho_old, qsi, eta, tdot[5], nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tol(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    LSSPG_TRY(d.copy(rtld, r));
    LSSPG_TRY(d.mxy(rtld, artld));
    LSSPG_TRY(d.pc(mr, r));
    LSSPG_TRY(d.mxy(mr, amr));
    LSSPG_TRY(d.dot(rtld, amr, &rho_old));
    LSSPG_TRY(d.copy(ap, amr));
    LSSPG_TRY(d.copy(p, mr));
    beta = 0.0;
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.pc(map, ap));
        LSSPG_TRY(d.dot(artld, map, &tdot[0]));
        alpha = rho_old / tdot[0];
        {
            const double *xs[5] = {y, amr, y, amr, amr}, *ys[5] = {y, r, r, y, amr};
            LSSPG_TRY(d.dots(5, xs, ys, tdot));
        }
        qsi_eta(iter, tdot, &qsi, &eta);
        LSSPG_TRY(d.scale(u, eta * beta));
        LSSPG_TRY(d.axpby(qsi, map, 1, u));
        LSSPG_TRY(d.axpby(eta, my, 1, u));
        LSSPG_TRY(d.mxy(u, au));
        LSSPG_TRY(d.scale(z, eta));
        LSSPG_TRY(d.axpby(qsi, mr, 1, z));
        LSSPG_TRY(d.axpby(-alpha, u, 1, z));
        LSSPG_TRY(d.scale(y, eta));
        LSSPG_TRY(d.axpby(qsi, amr, 1, y));
        LSSPG_TRY(d.axpby(-alpha, au, 1, y));
        LSSPG_TRY(d.pc(my, y));
        LSSPG_TRY(d.axpby(alpha, p, 1, x));
        LSSPG_TRY(d.axpby(1, z, 1, x));
        LSSPG_TRY(d.axpby(-alpha, ap, 1, r));
        LSSPG_TRY(d.axpby(-1, y, 1, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report("bicrsafe", iter, nrm2, ires, 2);
        if (tol >= nrm2) break;
        LSSPG_TRY(d.axpby(-alpha, map, 1, mr));
        LSSPG_TRY(d.axpby(-1, my, 1, mr));
        LSSPG_TRY(d.mxy(mr, amr));
        LSSPG_TRY(d.dot(rtld, amr, &rho));
        if (rho == 0.0) break;
        beta = (rho / rho_old) * (alpha / qsi);
        LSSPG_TRY(d.axpby(-1, u, 1., p));
        LSSPG_TRY(d.axpby(1, mr, beta, p));
        LSSPG_TRY(d.axpby(-1, au, 1.0, ap));
        LSSPG_TRY(d.axpby(1, amr, beta, ap));
        rho_old = rho;
    }
    return d.finish(iter, nrm2);
}

// ---- GPBiCG / GPBiCR: src/solver-gpbicg.cxx:4-163, src/solver-gpbicr.cxx:4-164 (they differ in the
//      shadow vector and the two inner products that use it).  `mr` is read before it is written on
//      the first iteration: zero here, compared against the calloc oracle. -------
static int gpbic(KrylovArgs &k, const lsspg_solver_opts *raw, bool cr, const char *name)
{
    Drv d(k, raw);
    double *x = k.x;
    double *r = d.vec(), *rtld = d.vec(), *mr = d.vec(), *p = d.vec(), *ap = d.vec(), *map = d.vec(), *t = d.vec(),
           *mt = d.vec(), *amt = d.vec(), *mt_old = d.vec(), *u = d.vec(), *y = d.vec(), *z = d.vec(), *w = d.vec();
    LSSPG_CHECK(d.ok(), "%s: out of device memory", name);
    double alpha, beta, rho, rho_old, qsi, eta, tdot[5], nrm2, ires, tol = 0;
    int iter = 1;
    const int maxiter = raw->maxit;
    bool done;
    LSSPG_TRY(d.resid(x, r));
    LSSPG_TRY(lis_tol(d, r, &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    if (cr) {
        LSSPG_TRY(d.copy(p, r));
        LSSPG_TRY(d.mxy(p, rtld));
        LSSPG_TRY(d.pc(p, r));
        LSSPG_TRY(d.dot(rtld, p, &rho_old));
    }
    else {
        LSSPG_TRY(d.copy(rtld, r));
        LSSPG_TRY(d.pc(p, r));
        LSSPG_TRY(d.dot(rtld, r, &rho_old));
    }
    LSSPG_TRY(d.set(t, 0.));
    LSSPG_TRY(d.set(w, 0.));
    beta = 0.0;
    for (iter = 1; iter <= maxiter; iter++) {
        LSSPG_TRY(d.mxy(p, ap));
        LSSPG_TRY(d.pc(map, ap));
        LSSPG_TRY(d.dot(rtld, cr ? map : ap, &tdot[0]));
        if (cr ? (fabs(tdot[0]) == 0.0) : (tdot[0] == 0.0)) break;
        alpha = rho_old / tdot[0];
        LSSPG_TRY(d.axpbyz(-1, w, 1, ap, y));
        LSSPG_TRY(d.axpby(1, t, alpha, y));
        LSSPG_TRY(d.axpby(-1, r, 1, y));
        LSSPG_TRY(d.axpbyz(-alpha, ap, 1, r, t));
        LSSPG_TRY(d.norm(t, &nrm2));
        if (k.verb >= 1) log_printf("%s: itr: %5d, abs res: %.6e, rel res: %.6e\n", name, iter, nrm2, nrm2 / ires);
        if (nrm2 <= tol) {
            LSSPG_TRY(d.axpby(alpha, p, 1, x));
            break;
        }
        LSSPG_TRY(d.axpbyz(-alpha, map, 1, mr, mt));
        LSSPG_TRY(d.mxy(mt, amt));
        {
            const double *xs[5] = {y, amt, y, amt, amt}, *ys[5] = {y, t, t, y, amt};
            LSSPG_TRY(d.dots(5, xs, ys, tdot));
        }
        qsi_eta(iter, tdot, &qsi, &eta);
        LSSPG_TRY(d.axpby(1., mt_old, beta, u));
        LSSPG_TRY(d.axpby(-1, mr, 1, u));
        LSSPG_TRY(d.scale(u, eta));
        LSSPG_TRY(d.axpby(qsi, map, 1, u));
        LSSPG_TRY(d.scale(z, eta));
        LSSPG_TRY(d.axpby(qsi, mr, 1, z));
        LSSPG_TRY(d.axpby(-alpha, u, 1, z));
        LSSPG_TRY(d.axpby(alpha, p, 1, x));
        LSSPG_TRY(d.axpby(1, z, 1., x));
        LSSPG_TRY(d.axpbyz(-qsi, amt, 1, t, r));
        LSSPG_TRY(d.axpby(-eta, y, 1, r));
        LSSPG_TRY(d.norm(r, &nrm2));
        d.report(name, iter, nrm2, ires, 1);
        if (tol >= nrm2) break;
        LSSPG_TRY(d.pc(mr, r));
        LSSPG_TRY(d.dot(rtld, cr ? mr : r, &rho));
        if (rho == 0.0) break;
        beta = (rho / rho_old) * (alpha / qsi);
        LSSPG_TRY(d.axpbyz(beta, ap, 1, amt, w));
        LSSPG_TRY(d.axpby(-1, u, 1, p));
        LSSPG_TRY(d.axpby(1., mr, beta, p));
        LSSPG_TRY(d.copy(mt_old, mt));
        rho_old = rho;
    }
    return d.finish(iter, nrm2);
}

int krylov_gpbicg(KrylovArgs &k, const lsspg_solver_opts *raw) { return gpbic(k, raw, false, "gpbicg"); }
int krylov_gpbicr(KrylovArgs &k, const lsspg_solver_opts *raw) { return gpbic(k, raw, true, "gpbicr"); }

// ---- BiCGStab(l): src/solver-bicgstabl.cxx:4-217 ------------------------------------------------------
int krylov_bicgstabl(KrylovArgs &k, const lsspg_solver_opts *raw)
{
    Drv d(k, raw);
    double *x = k.x;
    int l = k.bgsl;
    if (l <= 0) l = 4;
    const int z_dim = l + 1;
    double *rtld = d.vec(), *xp = d.vec(), *bp = d.vec(), *t = d.vec();
    std::vector<double *> r(l + 1), u(l + 1);
    for (int i = 0; i <= l; i++) r[i] = d.vec();
    for (int i = 0; i <= l; i++) u[i] = d.vec();
    LSSPG_CHECK(d.ok(), "bicgstabl: out of device memory");
    std::vector<double> store((size_t)z_dim * (4 + l + 1), 0.0);
    double *tau = store.data(), *gamma = tau + z_dim * z_dim, *gamma1 = gamma + z_dim, *gamma2 = gamma1 + z_dim,
           *sigma = gamma2 + z_dim;
    double alpha, beta, omega, rho0, rho1, nu, nrm2, ires, tol = 0;
    int iter = 0;
    const int maxiter = raw->maxit;
    bool done;
    // every exit that has converged or broken down maps the iterate back: x = M^-1 x + xp (:… pc.solve(&pc,t,x))
    auto back = [&]() -> int {
        LSSPG_TRY(d.pc(t, x));
        LSSPG_TRY(d.copy(x, t));
        return d.axpby(1, xp, 1, x);
    };
    LSSPG_TRY(d.resid(x, r[0]));
    LSSPG_TRY(d.copy(rtld, r[0]));
    LSSPG_TRY(d.copy(bp, r[0]));
    LSSPG_TRY(d.copy(xp, x));
    LSSPG_TRY(d.set(u[0], 0.));
    LSSPG_TRY(lis_tol(d, r[0], &nrm2, &ires, &tol, &done));
    if (done) return d.finish(0, nrm2);
    alpha = 0.0;
    omega = 1.0;
    rho0 = 1.0;
    bool out = false;
    while (iter <= maxiter && !out) {
        rho0 = -omega * rho0;
        for (int j = 0; j < l; j++) {
            iter++;
            LSSPG_TRY(d.dot(rtld, r[j], &rho1));
            if (rho1 == 0.0) { LSSPG_TRY(back()); out = true; break; }
            beta = alpha * (rho1 / rho0);
            rho0 = rho1;
            for (int i = 0; i <= j; i++) LSSPG_TRY(d.axpby(1, r[i], -beta, u[i]));
            LSSPG_TRY(d.pc(t, u[j]));
            LSSPG_TRY(d.mxy(t, u[j + 1]));
            LSSPG_TRY(d.dot(rtld, u[j + 1], &nu));
            if (fabs(nu) == 0.0) { LSSPG_TRY(back()); out = true; break; }
            alpha = rho1 / nu;
            LSSPG_TRY(d.axpby(alpha, u[0], 1, x));
            for (int i = 0; i <= j; i++) LSSPG_TRY(d.axpby(-alpha, u[i + 1], 1, r[i]));
            LSSPG_TRY(d.norm(r[0], &nrm2));
            d.report("bicgstabl", iter, nrm2, ires, 1);
            if (nrm2 <= tol) { LSSPG_TRY(back()); out = true; break; }
            LSSPG_TRY(d.pc(t, r[j]));
            LSSPG_TRY(d.mxy(t, r[j + 1]));
        }
        if (out) break;
        for (int j = 1; j <= l; j++) {                     // MR part
            for (int i = 1; i <= j - 1; i++) {
                LSSPG_TRY(d.dot(r[j], r[i], &nu));
                nu = nu / sigma[i];
                tau[i * z_dim + j] = nu;
                LSSPG_TRY(d.axpby(-nu, r[i], 1, r[j]));
            }
            LSSPG_TRY(d.dot(r[j], r[j], &sigma[j]));
            LSSPG_TRY(d.dot(r[0], r[j], &nu));
            gamma1[j] = nu / sigma[j];
        }
        gamma[l] = gamma1[l];
        omega = gamma[l];
        for (int j = l - 1; j >= 1; j--) {
            nu = 0.0;
            for (int i = j + 1; i <= l; i++) nu += tau[j * z_dim + i] * gamma[i];
            gamma[j] = gamma1[j] - nu;
        }
        for (int j = 1; j <= l - 1; j++) {
            nu = 0.0;
            for (int i = j + 1; i <= l - 1; i++) nu += tau[j * z_dim + i] * gamma[i + 1];
            gamma2[j] = gamma[j + 1] + nu;
        }
        LSSPG_TRY(d.axpby(gamma[1], r[0], 1, x));        // update
        LSSPG_TRY(d.axpby(-gamma1[l], r[l], 1, r[0]));
        LSSPG_TRY(d.axpby(-gamma[l], u[l], 1, u[0]));
        for (int j = 1; j <= l - 1; j++) {
            LSSPG_TRY(d.axpby(-gamma[j], u[j], 1, u[0]));
            LSSPG_TRY(d.axpby(gamma2[j], r[j], 1, x));
            LSSPG_TRY(d.axpby(-gamma1[j], r[j], 1, r[0]));
        }
        LSSPG_TRY(d.norm(r[0], &nrm2));
        if (k.verb >= 1) log_printf("bicgstabl: itr: %5d, abs res: %.6e, rel res: %.6e\n", iter, nrm2, nrm2 / ires);
        if (nrm2 < tol) { LSSPG_TRY(back()); break; }
    }
    return d.finish(iter, nrm2);
}

}  // namespace lsspg

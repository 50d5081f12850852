// ilu_rows.cuh -- the row recurrences of the device factorisations (ilu_gpu.cu), written once for the kernels and for
// their host replay (lsspg_debug_ilu_gpu_replay_host, CPU test-suite).  A row function gets `wait(k)`: on the device it
// blocks until row k has published itself (false: the kernel is draining), in the replay rows run in ascending order and
// it is always true.  Data another row wrote is read through fac_ld* (L2 on the device: another SM wrote it).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define FAC_HD __host__ __device__ __forceinline__
#else
#define FAC_HD inline
#endif

namespace lsspg {

constexpr double kPivotTolG = 1e-10;    // mat_zero_diag_tol,   reference src/pc.cxx:7
constexpr double kPivotValueG = 1e-3;   // mat_zero_diag_value, reference src/pc.cxx:6

FAC_HD int fac_ldi(const int *p)
{
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}
FAC_HD double fac_ldd(const double *p)
{
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}
FAC_HD void fac_std(double *p, double v)
{
#ifdef __CUDA_ARCH__
    __stcg(p, v);
#else
    *p = v;
#endif
}

FAC_HD double repaired_pivot_dev(double d) { return (fabs(d) < kPivotTolG) ? (d > 0 ? kPivotValueG : -kPivotValueG) : d; }

// ---- ILU(k) symbolic (src/pc-iluk.cxx:22-135) ---------------------------------------------------------------------------------
// Row i starts as A's row (levels 0) in its slot of the pool, pc / pl [i cap ..), and stays sorted.  Pivots are the lower
// columns in ascending order -- fill lands behind the current pivot, so "the smallest lower column not used yet"
// (:62-75) is simply the next entry.  A candidate (c, lev(i,piv) + lev(piv,c) + 1) above `level` is ignored, an absent
// column is inserted, a present one has its level RAISED to the candidate's when that is larger (the reference's rule,
// :101); the diagonal is never a candidate.  dpos[i] = position of the diagonal.
// Returns 0, 1 (the row outgrew cap, or a wait was aborted) or 2 (no diagonal).
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class Wait>
FAC_HD int iluk_symbolic_row(int i, int level, int cap, const int *Ap, const int *Aj, int *pc, int *pl, int *plen, int *dpos,
                             Wait wait)
{
    int *c_ = pc + (size_t)i * cap, *l_ = pl + (size_t)i * cap;
    int len = 0;
    bool ok = true;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        if (len == cap) { ok = false; break; }
        c_[len] = Aj[k];
        l_[len] = 0;
        len++;
    }
    int t = 0;
    while (ok && t < len && c_[t] < i) {
        const int piv = c_[t], lt = l_[t];
        if (!wait(piv)) { ok = false; break; }
        const int pn = fac_ldi(plen + piv);
        const int *qc = pc + (size_t)piv * cap, *ql = pl + (size_t)piv * cap;
        int a = t + 1;
        for (int q = fac_ldi(dpos + piv) + 1; q < pn; q++) {
            const int c = fac_ldi(qc + q);
            const int cand = fac_ldi(ql + q) + lt + 1;
            if (cand > level || c == i) continue;
            while (a < len && c_[a] < c) a++;
            if (a < len && c_[a] == c) {
                if (l_[a] < cand) l_[a] = cand;
            }
            else {
                if (len == cap) { ok = false; break; }
                for (int z = len; z > a; z--) { c_[z] = c_[z - 1]; l_[z] = l_[z - 1]; }
                c_[a] = c;
                l_[a] = cand;
                len++;
            }
        }
        t++;
    }
    plen[i] = len;
    dpos[i] = t;
    if (!ok) return 1;
    return (t >= len || c_[t] != i) ? 2 : 0;
}

// pattern row out of the pool, with A's values where present and 0 on fill (src/pc-iluk.cxx:318-343)
FAC_HD void iluk_pattern_row(int i, int cap, const int *pc, const int *Ap, const int *Aj, const double *Ax, const int *Mp, int *Mj,
                             double *Mx)
{
    const int *c_ = pc + (size_t)i * cap;
    int a = Ap[i];
    const int ae = Ap[i + 1], o = Mp[i], len = Mp[i + 1] - o;
    for (int k = 0; k < len; k++) {
        const int c = c_[k];
        while (a < ae && Aj[a] < c) a++;
        Mj[o + k] = c;
        Mx[o + k] = (a < ae && Aj[a] == c) ? Ax[a] : 0.0;
    }
}

// ---- ILU numeric: the IKJ row of src/pc-iluk.cxx:347-409 (blocks of bs rows), in place ---------------------------------------------
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class Wait>
FAC_HD bool ilu_numeric_row(int i, int bs, const int *P, const int *C, double *X, double *inv, Wait wait)
{
    const int e = P[i + 1];
    int k = P[i];
    if (i % bs == 0) {
        // first row of a block: its leading entry is the pivot; the repaired value only enters the inverse, the stored
        // entry is left alone (as the host loop)
        fac_std(inv + i, 1. / repaired_pivot_dev(fac_ldd(X + k)));
        return true;
    }
    bool ok = true;
    for (; ok && k < e && C[k] < i; k++) {
        const int pr = C[k];
        if (!wait(pr)) { ok = false; break; }
        const double a_ik = fac_ldd(X + k) * fac_ldd(inv + pr);
        fac_std(X + k, a_ik);
        int pq = P[pr];
        const int pe = P[pr + 1];
        for (int q = k + 1; q < e; q++) {
            const int c = C[q];
            while (pq < pe && C[pq] < c) pq++;
            if (pq < pe && C[pq] == c) {
                const double w = fac_ldd(X + pq);
                if (w != 0.) fac_std(X + q, fac_ldd(X + q) - a_ik * w);
            }
        }
    }
    double d = kPivotValueG;
    if (k < e && C[k] == i) {
        double v = fac_ldd(X + k);
        if (fabs(v) < kPivotTolG) { v = kPivotValueG; fac_std(X + k, v); }
        d = v;
    }
    fac_std(inv + i, 1. / d);
    return ok;
}

// ---- ILUT (src/pc-ilut.cxx:51-286) -------------------------------------------------------------------------------------------
// Partial ordering with the reference's exact sequence of exchanges (:7-49): the kept entries are stored -- and later
// summed by the sweeps -- in the order this leaves them.
FAC_HD void select_largest_dev(double *a, int *ind, int n, int ncut)
{
    int lo = 0, hi = n - 1;
    if (ncut < lo || ncut >= hi) return;
    for (;;) {
        int mid = lo;
        const double key = fabs(a[mid]);
        for (int q = lo + 1; q <= hi; q++) {
            if (fabs(a[q]) > key) {
                ++mid;
                const double ta = a[mid]; a[mid] = a[q]; a[q] = ta;
                const int ti = ind[mid]; ind[mid] = ind[q]; ind[q] = ti;
            }
        }
        const double ta = a[mid]; a[mid] = a[lo]; a[lo] = ta;
        const int ti = ind[mid]; ind[mid] = ind[lo]; ind[lo] = ti;
        if (mid == ncut) return;
        if (mid > ncut) hi = mid - 1;
        else lo = mid + 1;
    }
}

// column -> position map of the work row: open addressing, hmask + 1 slots (a power of two, >= 4 wcap) of {column, row
// stamp, position}; a slot whose stamp is not the current row is empty, so nothing is ever cleared (and nothing is deleted
// inside a row: processed pivots are smaller than every later candidate and are never looked up again)
struct IlutSlot {
    int key, stamp, pos;
};
FAC_HD IlutSlot *ilut_find(IlutSlot *tab, int hmask, int row, int c)
{
    unsigned int h = ((unsigned int)c * 2654435761u) >> 7;
    for (;; h++) {
        IlutSlot *s = tab + (h & (unsigned int)hmask);
        if (s->stamp != row || s->key == c) return s;
    }
}

// The row recurrence of ilu_host.cpp: factor_ilut_rows, statement for statement: work row as two compact arrays (lower
// part jwl / wl, upper part jwu / wu, wcap entries each), `present?` answered by the column map, new fill below the drop
// threshold ignored, the p largest of each part kept in quick-select order.  Finished rows live in a pool
// ([kept lower | diagonal | kept upper], rcap entries each).  Returns false when a work array or the pool slot overflowed
// or a wait was aborted.
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class Wait>
FAC_HD bool ilut_row(int i, int bs, int p, double tau, const int *Bp, const int *Bj, const double *Bx, int rcap, int *rc, double *rv,
                     int *rlen, double *diag, int wcap, int *jwl, int *jwu, double *wl, double *wu, IlutSlot *tab, int hmask,
                     Wait wait)
{
    constexpr int kUp = 1 << 30;   // positions >= kUp: upper part
    const int b = Bp[i], e = Bp[i + 1];
    int *c_ = rc + (size_t)i * rcap;
    double *v_ = rv + (size_t)i * rcap;
    bool ok = true;
    if (i % bs == 0) {
        // first row of a block: copied verbatim; its leading entry is the pivot
        if (e - b > rcap) ok = false;
        for (int k = b; ok && k < e; k++) { c_[k - b] = Bj[k]; v_[k - b] = Bx[k]; }
        rlen[i] = ok ? e - b : 0;
        fac_std(diag + i, repaired_pivot_dev(Bx[b]));
        return ok;
    }
    double norm = 0.0;
    for (int k = b; k < e; k++) norm += fabs(Bx[k]);
    norm /= (double)(e - b);
    const double drop = tau * norm;
    int nl = 0, nu = 0;
    double wd = 0.0;
    for (int k = b; k < e && ok; k++) {
        const int c = Bj[k];
        if (c == i) { wd = Bx[k]; continue; }
        if ((c < i ? nl : nu) == wcap) { ok = false; break; }
        IlutSlot *s = ilut_find(tab, hmask, i, c);
        s->key = c; s->stamp = i;
        if (c < i) { s->pos = nl; jwl[nl] = c; wl[nl] = Bx[k]; nl++; }
        else { s->pos = kUp + nu; jwu[nu] = c; wu[nu] = Bx[k]; nu++; }
    }
    for (int tt = 0; ok && tt < nl; tt++) {
        int piv = jwl[tt], at = tt;
        for (int q = tt + 1; q < nl; q++)
            if (jwl[q] < piv) { piv = jwl[q]; at = q; }
        if (at != tt) {
            const int c = jwl[tt];
            jwl[tt] = jwl[at];
            jwl[at] = c;
            ilut_find(tab, hmask, i, c)->pos = at;
            const double tw = wl[tt]; wl[tt] = wl[at]; wl[at] = tw;
        }
        if (!wait(piv)) { ok = false; break; }
        const double a_ik = wl[tt] / fac_ldd(diag + piv);
        wl[tt] = a_ik;
        const int *pc = rc + (size_t)piv * rcap;
        const double *pv = rv + (size_t)piv * rcap;
        for (int q = 0, qe = fac_ldi(rlen + piv); q < qe; q++) {
            const int c = fac_ldi(pc + q);
            if (c <= piv) continue;
            const double mx = -a_ik * fac_ldd(pv + q);
            if (c == i) { wd += mx; continue; }
            IlutSlot *sl = ilut_find(tab, hmask, i, c);
            if (sl->stamp == i) {   // present
                if (sl->pos >= kUp) wu[sl->pos - kUp] += mx;
                else wl[sl->pos] += mx;
            }
            else if (!(fabs(mx) < drop)) {   // only NEW fill is dropped
                if (c < i) {
                    if (nl == wcap) { ok = false; break; }
                    sl->key = c; sl->stamp = i; sl->pos = nl;
                    jwl[nl] = c; wl[nl] = mx; nl++;
                }
                else {
                    if (nu == wcap) { ok = false; break; }
                    sl->key = c; sl->stamp = i; sl->pos = kUp + nu;
                    jwu[nu] = c; wu[nu] = mx; nu++;
                }
            }
        }
    }
    const double d = repaired_pivot_dev(wd);
    const int keepl = nl < p ? nl : p, keepu = nu < p ? nu : p;
    if (ok && keepl + 1 + keepu > rcap) ok = false;
    if (ok) {
        select_largest_dev(wl, jwl, nl, keepl);
        select_largest_dev(wu, jwu, nu, keepu);
        for (int q = 0; q < keepl; q++) { c_[q] = jwl[q]; v_[q] = wl[q]; }
        c_[keepl] = i;
        v_[keepl] = d;
        for (int q = 0; q < keepu; q++) { c_[keepl + 1 + q] = jwu[q]; v_[keepl + 1 + q] = wu[q]; }
    }
    rlen[i] = ok ? keepl + 1 + keepu : 0;
    fac_std(diag + i, d);
    return ok;
}

// ---- split: L = lower entries in stored order + unit diagonal LAST, U = diagonal FIRST + the rest (src/pc-iluk.cxx:501-532,
// src/pc-ilut.cxx:253-274) ---------------------------------------------------------------------------------------------------
FAC_HD void split_count_row(int i, const int *c_, int len, int *nl, int *nu)
{
    int a = 0, b = 0;
    for (int q = 0; q < len; q++) {
        a += (c_[q] <= i);
        b += (c_[q] >= i);
    }
    nl[i] = a;
    nu[i] = b;
}
FAC_HD void split_fill_row(int i, const int *c_, const double *v_, int len, const int *Lp, int *Lj, double *Lx, const int *Up, int *Uj,
                           double *Ux)
{
    int ol = Lp[i], ou = Up[i];
    for (int q = 0; q < len; q++) {
        const int c = c_[q];
        const double v = v_[q];
        if (c < i) { Lj[ol] = c; Lx[ol] = v; ol++; }
        else if (c == i) {
            Lj[ol] = i; Lx[ol] = 1; ol++;
            Uj[ou] = i; Ux[ou] = v; ou++;
        }
        else { Uj[ou] = c; Ux[ou] = v; ou++; }
    }
}

}  // namespace lsspg

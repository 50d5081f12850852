// amg_rows.cuh -- the per-row phases of the SX-AMG-style set-up (amg_host.cpp: strong couplings, direct interpolation,
// restriction = P^T, Galerkin products), written once as __host__ __device__ functions: the device set-up (amg_gpu.cu)
// runs them one row per thread, the CPU replay (lsspg_debug_amg_setup_replay_host) row after row; both must give the
// hierarchy of lsspg_amg_setup_host bit for bit (tests/test_amg_rows.py on the CPU, tests/test_gpu_amg_setup.py on the
// device).  The Ruge-Stueben C/F splitting is a serial greedy pass and stays on the host in every variant.
// Every function follows the statements of its host counterpart in the same order: sums are accumulated in the row's
// column order, products of the Galerkin triple product in the order of the row-by-row loops, no FMA.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define AMG_HD __host__ __device__ __forceinline__
#else
#define AMG_HD inline
#endif

namespace lsspg {

constexpr int kAmgCPT = 1;

// strong couplings of row i in the row's column order (amg_host.cpp: strong_couplings); out == NULL: count only
AMG_HD int amg_strong_row(int i, const int *Ap, const int *Aj, const double *Ax, double strong_threshold, double max_row_sum, int *out)
{
    double diag = 0.0, row_sum = 0.0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        if (Aj[k] == i) diag = Ax[k];
        row_sum += Ax[k];
    }
    const double s = diag < 0.0 ? -1.0 : 1.0;
    double most = 0.0;   // largest -s a_ij over the off-diagonals
    for (int k = Ap[i]; k < Ap[i + 1]; k++)
        if (Aj[k] != i) {
            const double v = -s * Ax[k];
            most = (most < v) ? v : most;   // std::max(most, v)
        }
    const bool dominated = max_row_sum < 1.0 && fabs(row_sum) > max_row_sum * fabs(diag);
    int cnt = 0;
    if (most > 0.0 && !dominated) {
        const double cut = strong_threshold * most;
        for (int k = Ap[i]; k < Ap[i + 1]; k++)
            if (Aj[k] != i && -s * Ax[k] >= cut) {
                if (out) out[cnt] = Aj[k];
                cnt++;
            }
    }
    return cnt;
}

// row i of the direct interpolation with truncation (amg_host.cpp: direct_interpolation); Pj == NULL: count only.
// cidx[j] = number of the C point j on the coarse level.  "j is a strong C neighbour of i" (the host's mark[j] == i) is
// answered by a search of i's strong list.
AMG_HD int amg_interp_row(int i, const int *Ap, const int *Aj, const double *Ax, const int *Sp, const int *Sj, const int *cf,
                          const int *cidx, double trunc_threshold, int *Pj, double *Px)
{
    if (cf[i] == kAmgCPT) {
        if (Pj) { Pj[0] = cidx[i]; Px[0] = 1.0; }
        return 1;
    }
    const int sb = Sp[i], se = Sp[i + 1];
    if (se <= sb) return 0;
    auto strong_c = [&](int j) {
        if (cf[j] != kAmgCPT) return false;
        for (int q = sb; q < se; q++)
            if (Sj[q] == j) return true;
        return false;
    };
    double diag = 0.0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++)
        if (Aj[k] == i) diag = Ax[k];
    const double s = diag < 0.0 ? -1.0 : 1.0;
    double all_neg = 0.0, all_pos = 0.0, c_neg = 0.0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        const int j = Aj[k];
        if (j == i) continue;
        const double a = s * Ax[k];
        if (a < 0.0) {
            all_neg += a;
            if (strong_c(j)) c_neg += a;
        }
        else all_pos += a;
    }
    // strong couplings are all of the "negative" kind: the others are lumped into the diagonal
    const double dd = s * diag + all_pos;
    if (!(c_neg != 0.0 && dd != 0.0)) return 0;
    const double alpha = all_neg / c_neg;
    double wmax = 0.0, total = 0.0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        const int j = Aj[k];
        if (j == i || !strong_c(j)) continue;
        const double v = -alpha * (s * Ax[k]) / dd;
        const double av = fabs(v);
        wmax = (wmax < av) ? av : wmax;
        total += v;
    }
    // truncation: drop the small weights, keep the row sum
    double kept = 0.0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        const int j = Aj[k];
        if (j == i || !strong_c(j)) continue;
        const double v = -alpha * (s * Ax[k]) / dd;
        if (fabs(v) >= trunc_threshold * wmax) kept += v;
    }
    const double scale = (kept != 0.0) ? total / kept : 1.0;
    int cnt = 0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        const int j = Aj[k];
        if (j == i || !strong_c(j)) continue;
        const double v = -alpha * (s * Ax[k]) / dd;
        if (fabs(v) >= trunc_threshold * wmax) {
            if (Pj) { Pj[cnt] = cidx[j]; Px[cnt] = v * scale; }
            cnt++;
        }
    }
    return cnt;
}

// row c of R = P^T (amg_host.cpp: transpose_csr -- the entries of a row in ascending order of the fine point).  f = the
// fine point that IS coarse point c; the other rows of P that can hold column c are the points that depend strongly on
// f, i.e. row f of the transposed strength graph T (ascending).  Rj == NULL: count only.
AMG_HD int amg_restrict_row(int c, int f, const int *Tp, const int *Tj, const int *Pp, const int *Pj, const double *Px, int *Rj, double *Rx)
{
    int cnt = 0;
    auto take = [&](int i) {
        for (int q = Pp[i]; q < Pp[i + 1]; q++)
            if (Pj[q] == c) {
                if (Rj) { Rj[cnt] = i; Rx[cnt] = Px[q]; }
                cnt++;
            }
    };
    bool self_done = false;
    for (int k = Tp[f]; k < Tp[f + 1]; k++) {
        const int i = Tj[k];
        if (!self_done && f < i) { take(f); self_done = true; }
        take(i);
    }
    if (!self_done) take(f);
    return cnt;
}

// Galerkin products (amg_host.cpp: spgemm): row i of C = A B with a per-thread accumulator table keyed by column (open
// addressing, slots stamped with the row so that nothing is cleared); a column's products are added in the order the
// row-by-row loops meet them.  cols[cap]: the row's distinct columns, sorted ascending at the end.  Cj == NULL: count
// only.  Returns the number of entries, or -1 when cols overflowed.
struct AmgSlot {
    int key, stamp;
    double acc;
};
AMG_HD AmgSlot *amg_slot(AmgSlot *tab, int hmask, int stamp, int c)
{
    unsigned int h = ((unsigned int)c * 2654435761u) >> 6;
    for (;; h++) {
        AmgSlot *s = tab + (h & (unsigned int)hmask);
        if (s->stamp != stamp || s->key == c) return s;
    }
}
AMG_HD int amg_spgemm_row(int i, int stamp, const int *Ap, const int *Aj, const double *Ax, const int *Bp, const int *Bj, const double *Bx,
                          AmgSlot *tab, int hmask, int *cols, int cap, int *Cj, double *Cx)
{
    int cnt = 0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) {
        const int m = Aj[k];
        const double a = Ax[k];
        for (int q = Bp[m]; q < Bp[m + 1]; q++) {
            const int c = Bj[q];
            AmgSlot *s = amg_slot(tab, hmask, stamp, c);
            if (s->stamp != stamp) {
                if (cnt == cap) return -1;
                s->key = c;
                s->stamp = stamp;
                s->acc = 0.0;
                cols[cnt++] = c;
            }
            s->acc += a * Bx[q];
        }
    }
    for (int u = 1; u < cnt; u++) {   // ascending columns
        const int c = cols[u];
        int v = u;
        for (; v > 0 && cols[v - 1] > c; v--) cols[v] = cols[v - 1];
        cols[v] = c;
    }
    if (Cj)
        for (int u = 0; u < cnt; u++) {
            Cj[u] = cols[u];
            Cx[u] = amg_slot(tab, hmask, stamp, cols[u])->acc;
        }
    return cnt;
}

// upper bound of the entries of row i of A B: the number of products
AMG_HD int amg_spgemm_bound(int i, const int *Ap, const int *Aj, const int *Bp)
{
    long long b = 0;
    for (int k = Ap[i]; k < Ap[i + 1]; k++) b += Bp[Aj[k] + 1] - Bp[Aj[k]];
    return b > 0x7fffffff ? 0x7fffffff : (int)b;
}

}  // namespace lsspg

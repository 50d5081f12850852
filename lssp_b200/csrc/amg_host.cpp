// amg_host.cpp -- set-up of the SX-AMG-style hierarchy (host; the GPU runs the cycle, amg.cu).
//
// The reference reaches its AMG through libsxamg (sx_amg_setup, src/pc-sxamg.cxx:107-110), which
// is not part of the reference tree; what is built here is the classical Ruge-Stueben method that
// library implements, from its published description.  DESIGN.md "AMG" is the
// specification; parity with libsxamg itself is UNPINNED.
//   strength     j strongly influences i  <=>  -s a_ij >= theta * max_k(-s a_ik), s = sign(a_ii);
//                rows with |sum_j a_ij| > max_row_sum |a_ii| have no strong couplings
//   C/F split    first Ruge-Stueben pass with bucket lists (measure = #points influenced), then
//                every F point left without a strong C neighbour is promoted to C
//   P            direct interpolation, w_ij = -alpha_i a_ij / a~_ii over the strong C neighbours,
//                truncated at trunc_threshold * max|w| with rescaling; C rows are unit rows
//   coarse grid  R = P^T, A_c = R A P (two row-wise sparse products, columns sorted)
//   last level   dense inverse by Gauss-Jordan with partial pivoting (<= coarse_dense_max rows)
// Also builds the dependency schedule of a Gauss-Seidel sweep (gs_build_host) and walks it on
// the CPU for the test-suite (lsspg_debug_amg_walk_gs_host) -- a layout check, not a fallback.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <thread>
#include <vector>
#include "amg_host.h"
#include "amg_rows.cuh"

namespace lsspg {
void set_error(const char *fmt, ...);
}
using namespace lsspg;

#define AMG_CHECK(cond, ...)               \
    do {                                   \
        if (!(cond)) {                     \
            lsspg::set_error(__VA_ARGS__); \
            return 1;                      \
        }                                  \
    } while (0)

namespace {

using Graph = AmgGraph;

// strong couplings of every row, in the row's column order
void strong_couplings(const AmgLevelHost &L, const lsspg_amg_pars &pr, Graph &S)
{
    const int n = L.n;
    S.p.assign(n + 1, 0);
    S.j.clear();
    S.j.reserve(L.Aj.size());
    for (int i = 0; i < n; i++) {
        double diag = 0.0, row_sum = 0.0;
        for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++) {
            if (L.Aj[k] == i) diag = L.Ax[k];
            row_sum += L.Ax[k];
        }
        const double s = diag < 0.0 ? -1.0 : 1.0;
        double most = 0.0;   // largest -s a_ij over the off-diagonals
        for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++)
            if (L.Aj[k] != i) most = std::max(most, -s * L.Ax[k]);
        const bool dominated = pr.max_row_sum < 1.0 && fabs(row_sum) > pr.max_row_sum * fabs(diag);
        if (most > 0.0 && !dominated) {
            const double cut = pr.strong_threshold * most;
            for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++)
                if (L.Aj[k] != i && -s * L.Ax[k] >= cut) S.j.push_back(L.Aj[k]);
        }
        S.p[i + 1] = (int)S.j.size();
    }
}

void transpose_graph(int n, const Graph &S, Graph &T)
{
    T.p.assign(n + 1, 0);
    T.j.resize(S.j.size());
    for (int v : S.j) T.p[v + 1]++;
    for (int i = 0; i < n; i++) T.p[i + 1] += T.p[i];
    std::vector<int> pos(T.p.begin(), T.p.end() - 1);
    for (int i = 0; i < n; i++)
        for (int k = S.p[i]; k < S.p[i + 1]; k++) T.j[pos[S.j[k]]++] = i;
}

// bucket lists keyed by the measure; the most recently touched point of a bucket is its head
struct Buckets {
    std::vector<int> head, next, prev, key;
    int top = 0;
    explicit Buckets(int n) : next(n, -1), prev(n, -1), key(n, -1) {}
    void insert(int i, int k)
    {
        if (k >= (int)head.size()) head.resize(k + 16, -1);
        key[i] = k;
        prev[i] = -1;
        next[i] = head[k];
        if (head[k] >= 0) prev[head[k]] = i;
        head[k] = i;
        top = std::max(top, k);
    }
    void remove(int i)
    {
        const int k = key[i];
        if (prev[i] >= 0) next[prev[i]] = next[i];
        else head[k] = next[i];
        if (next[i] >= 0) prev[next[i]] = prev[i];
        key[i] = -1;
    }
    void move(int i, int k)
    {
        remove(i);
        insert(i, k);
    }
    int pop_max()
    {
        while (top > 0 && (top >= (int)head.size() || head[top] < 0)) top--;
        return (top >= (int)head.size()) ? -1 : head[top];
    }
};

constexpr int UNDECIDED = -1, FPT = 0, CPT = 1;

int cf_split(int n, const Graph &S, const Graph &T, std::vector<int> &cf)
{
    cf.assign(n, UNDECIDED);
    std::vector<int> lambda(n);
    for (int i = 0; i < n; i++) lambda[i] = T.p[i + 1] - T.p[i];
    // points nobody depends on become F at once; their influencers gain weight
    for (int i = 0; i < n; i++) {
        if (S.p[i + 1] == S.p[i]) cf[i] = FPT;   // no strong coupling at all: smoothing alone treats it
        else if (T.p[i + 1] == T.p[i]) {
            cf[i] = FPT;
            for (int k = S.p[i]; k < S.p[i + 1]; k++) lambda[S.j[k]]++;
        }
    }
    Buckets B(n);
    for (int i = 0; i < n; i++)
        if (cf[i] == UNDECIDED) B.insert(i, lambda[i]);
    for (;;) {
        const int i = B.pop_max();
        if (i < 0 || B.top == 0) break;
        cf[i] = CPT;
        B.remove(i);
        for (int k = T.p[i]; k < T.p[i + 1]; k++) {
            const int j = T.j[k];   // j depends on i
            if (cf[j] != UNDECIDED) continue;
            cf[j] = FPT;
            B.remove(j);
            for (int q = S.p[j]; q < S.p[j + 1]; q++) {
                const int m = S.j[q];
                if (cf[m] == UNDECIDED) B.move(m, ++lambda[m]);
            }
        }
        for (int k = S.p[i]; k < S.p[i + 1]; k++) {
            const int j = S.j[k];   // i depends on j
            if (cf[j] != UNDECIDED) continue;
            if (lambda[j] > 0) lambda[j]--;
            B.move(j, lambda[j]);
        }
    }
    for (int i = 0; i < n; i++)
        if (cf[i] == UNDECIDED) cf[i] = FPT;
    // an F point with strong couplings but no strong C neighbour cannot be interpolated: promote it
    int nc = 0;
    for (int i = 0; i < n; i++) {
        if (cf[i] == FPT && S.p[i + 1] > S.p[i]) {
            bool has_c = false;
            for (int k = S.p[i]; k < S.p[i + 1] && !has_c; k++) has_c = (cf[S.j[k]] == CPT);
            if (!has_c) cf[i] = CPT;
        }
        nc += (cf[i] == CPT);
    }
    return nc;
}

void direct_interpolation(AmgLevelHost &L, const Graph &S, const lsspg_amg_pars &pr)
{
    const int n = L.n;
    std::vector<int> cidx(n, -1);
    int nc = 0;
    for (int i = 0; i < n; i++)
        if (L.cf[i] == CPT) cidx[i] = nc++;
    L.nc = nc;
    L.Pp.assign(n + 1, 0);
    L.Pj.clear();
    L.Px.clear();
    std::vector<int> mark(n, -1);
    std::vector<double> w;
    std::vector<int> wc;
    for (int i = 0; i < n; i++) {
        if (L.cf[i] == CPT) {
            L.Pj.push_back(cidx[i]);
            L.Px.push_back(1.0);
        }
        else if (S.p[i + 1] > S.p[i]) {
            for (int k = S.p[i]; k < S.p[i + 1]; k++)
                if (L.cf[S.j[k]] == CPT) mark[S.j[k]] = i;
            double diag = 0.0;
            for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++)
                if (L.Aj[k] == i) diag = L.Ax[k];
            const double s = diag < 0.0 ? -1.0 : 1.0;
            double all_neg = 0.0, all_pos = 0.0, c_neg = 0.0;
            for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++) {
                const int j = L.Aj[k];
                if (j == i) continue;
                const double a = s * L.Ax[k];
                if (a < 0.0) {
                    all_neg += a;
                    if (mark[j] == i) c_neg += a;
                }
                else all_pos += a;
            }
            // strong couplings are all of the "negative" kind: the others are lumped into the diagonal
            const double dd = s * diag + all_pos;
            w.clear();
            wc.clear();
            if (c_neg != 0.0 && dd != 0.0) {
                const double alpha = all_neg / c_neg;
                double wmax = 0.0, total = 0.0;
                for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++) {
                    const int j = L.Aj[k];
                    if (j == i || mark[j] != i) continue;
                    const double v = -alpha * (s * L.Ax[k]) / dd;
                    w.push_back(v);
                    wc.push_back(cidx[j]);
                    wmax = std::max(wmax, fabs(v));
                    total += v;
                }
                // truncation: drop the small weights, keep the row sum
                double kept = 0.0;
                for (double v : w)
                    if (fabs(v) >= pr.trunc_threshold * wmax) kept += v;
                const double scale = (kept != 0.0) ? total / kept : 1.0;
                for (size_t q = 0; q < w.size(); q++) {
                    if (fabs(w[q]) >= pr.trunc_threshold * wmax) {
                        L.Pj.push_back(wc[q]);
                        L.Px.push_back(w[q] * scale);
                    }
                }
            }
        }
        L.Pp[i + 1] = (int)L.Pj.size();
    }
}

// Visiting order of the points inside their C / F block.  cf_order 0 / 1: ascending index (the serial
// sweep of the library).  cf_order 2: MULTICOLOUR -- greedy colouring of each block's own coupling graph
// (a point takes the smallest colour none of its already coloured same-block neighbours has), points
// visited colour by colour, ascending index inside a colour.  Still a Gauss-Seidel sweep with exact
// sequential semantics, but its dependency chains are as long as the number of colours (tens) instead
// of the grid diameter (hundreds to thousands on the coarse levels), which is what a GPU needs.
void visiting_ranks(AmgLevelHost &L, int cf_order)
{
    const int n = L.n;
    L.rank.resize(n);
    if (cf_order != 2) {
        for (int i = 0; i < n; i++) L.rank[i] = i;
        return;
    }
    std::vector<int> colour(n, -1), used;
    int ncol = 0;
    for (int i = 0; i < n; i++) {
        used.assign(ncol + 1, 0);
        for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++) {
            const int c = L.Aj[k];
            if (c != i && L.cf[c] == L.cf[i] && colour[c] >= 0) used[colour[c]] = 1;
        }
        int col = 0;
        while (col < ncol && used[col]) col++;
        colour[i] = col;
        ncol = std::max(ncol, col + 1);
    }
    // a coupling a_ij without a_ji would let two neighbours share a colour seen from j's side only:
    // the schedule below never relies on colours being proper, only on ranks being a total order
    std::vector<int> start(ncol + 1, 0);
    for (int i = 0; i < n; i++) start[colour[i] + 1]++;
    for (int c = 0; c < ncol; c++) start[c + 1] += start[c];
    for (int i = 0; i < n; i++) L.rank[i] = start[colour[i]]++;
}

void transpose_csr(int nrows, int ncols, const std::vector<int> &p, const std::vector<int> &j,
                   const std::vector<double> &x, std::vector<int> &tp, std::vector<int> &tj, std::vector<double> &tx)
{
    tp.assign(ncols + 1, 0);
    tj.resize(j.size());
    tx.resize(x.size());
    for (int v : j) tp[v + 1]++;
    for (int i = 0; i < ncols; i++) tp[i + 1] += tp[i];
    std::vector<int> pos(tp.begin(), tp.end() - 1);
    for (int i = 0; i < nrows; i++)
        for (int k = p[i]; k < p[i + 1]; k++) {
            const int q = pos[j[k]]++;
            tj[q] = i;
            tx[q] = x[k];
        }
}

// C = A * B, row by row with a dense accumulator; columns of every row sorted ascending.  Rows are
// independent: contiguous row ranges are handed to host threads (each with its own accumulator) and
// the pieces concatenated, so the result does not depend on the number of threads.
int spgemm(int nrows, int ncolsB, const std::vector<int> &Ap, const std::vector<int> &Aj, const std::vector<double> &Ax,
           const std::vector<int> &Bp, const std::vector<int> &Bj, const std::vector<double> &Bx,
           std::vector<int> &Cp, std::vector<int> &Cj, std::vector<double> &Cx)
{
    int nt = (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, 16));
    if (nrows < 20000) nt = 1;
    struct Piece {
        std::vector<int> len, j;
        std::vector<double> x;
    };
    std::vector<Piece> pieces(nt);
    auto work = [&](int t) {
        const int r0 = (int)((long long)nrows * t / nt), r1 = (int)((long long)nrows * (t + 1) / nt);
        Piece &P = pieces[t];
        P.len.assign(r1 - r0, 0);
        {   // the number of products bounds the piece's entries: reserve once (untouched pages cost nothing)
            size_t bound = 0;
            for (int k = Ap[r0]; k < Ap[r1]; k++) bound += (size_t)(Bp[Aj[k] + 1] - Bp[Aj[k]]);
            P.j.reserve(bound);
            P.x.reserve(bound);
        }
        // accumulator of one row: open-addressed table keyed by column (grown when half full).  A column's
        // products are added in the order the row-by-row loop meets them, exactly as a dense accumulator
        // would -- without two ncolsB-sized arrays per thread to allocate and touch.
        size_t cap = 256;
        std::vector<int> keys(cap, -1), cols;
        std::vector<double> acc(cap, 0.0);
        std::vector<size_t> slots;
        auto slot_of = [&](int c) {
            size_t h = ((size_t)(unsigned)c * 2654435761u) & (cap - 1);
            while (keys[h] != -1 && keys[h] != c) h = (h + 1) & (cap - 1);
            return h;
        };
        for (int i = r0; i < r1; i++) {
            cols.clear();
            for (int k = Ap[i]; k < Ap[i + 1]; k++) {
                const int m = Aj[k];
                const double a = Ax[k];
                for (int q = Bp[m]; q < Bp[m + 1]; q++) {
                    const int c = Bj[q];
                    size_t h = slot_of(c);
                    if (keys[h] == -1) {
                        if (2 * (cols.size() + 1) > cap) {   // grow: move the partial sums, nothing is re-added
                            std::vector<int> okeys(2 * cap, -1);
                            std::vector<double> oacc(2 * cap, 0.0);
                            okeys.swap(keys);
                            oacc.swap(acc);
                            cap *= 2;
                            for (int cc : cols) {
                                size_t ho = ((size_t)(unsigned)cc * 2654435761u) & (cap / 2 - 1);
                                while (okeys[ho] != cc) ho = (ho + 1) & (cap / 2 - 1);
                                const size_t hn = slot_of(cc);
                                keys[hn] = cc;
                                acc[hn] = oacc[ho];
                            }
                            h = slot_of(c);
                        }
                        keys[h] = c;
                        acc[h] = 0.0;
                        cols.push_back(c);
                    }
                    acc[h] += a * Bx[q];
                }
            }
            std::sort(cols.begin(), cols.end());
            slots.clear();
            for (int c : cols) {
                const size_t h = slot_of(c);
                P.j.push_back(c);
                P.x.push_back(acc[h]);
                slots.push_back(h);
            }
            for (size_t h : slots) keys[h] = -1;   // only after every look-up of the row (probe chains stay intact)
            P.len[i - r0] = (int)cols.size();
        }
    };
    timespec ts0, ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(work, t);
        for (auto &q : th) q.join();
    }
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    if (getenv("LSSPG_SETUP_PROF") && atoi(getenv("LSSPG_SETUP_PROF")) != 0) fprintf(stderr, "[spgemm] rows %d threads part %.3f s\n", nrows, (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec));
    size_t total = 0;
    for (const Piece &P : pieces) total += P.j.size();
    AMG_CHECK(total < (size_t)0x7fffffff, "amg: coarse operator exceeds int32 indexing");
    Cp.assign(nrows + 1, 0);
    Cj.resize(total);
    Cx.resize(total);
    std::vector<size_t> off(nt + 1, 0);
    std::vector<int> row0(nt + 1, 0);
    for (int t = 0; t < nt; t++) {
        off[t + 1] = off[t] + pieces[t].j.size();
        row0[t + 1] = row0[t] + (int)pieces[t].len.size();
    }
    auto gather = [&](int t) {
        const Piece &P = pieces[t];
        std::copy(P.j.begin(), P.j.end(), Cj.begin() + off[t]);
        std::copy(P.x.begin(), P.x.end(), Cx.begin() + off[t]);
        size_t run = off[t];
        for (size_t q = 0; q < P.len.size(); q++) {
            run += (size_t)P.len[q];
            Cp[row0[t] + q + 1] = (int)run;
        }
    };
    if (nt == 1) gather(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back(gather, t);
        for (auto &q : th) q.join();
    }
    return 0;
}

template <class T>
void copy_out(T *dst, const std::vector<T> &src)
{
    if (dst && !src.empty()) memcpy(dst, src.data(), sizeof(T) * src.size());
}

// row-major inverse by Gauss-Jordan with partial pivoting
int dense_inverse(const AmgLevelHost &L, std::vector<double> &inv)
{
    const int n = L.n;
    std::vector<double> a((size_t)n * n, 0.0);
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) {
        for (int k = L.Ap[i]; k < L.Ap[i + 1]; k++) a[(size_t)i * n + L.Aj[k]] = L.Ax[k];
        inv[(size_t)i * n + i] = 1.0;
    }
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++)
            if (fabs(a[(size_t)r * n + c]) > fabs(a[(size_t)piv * n + c])) piv = r;
        AMG_CHECK(a[(size_t)piv * n + c] != 0.0, "amg: coarsest operator is singular (column %d)", c);
        if (piv != c) {
            for (int q = 0; q < n; q++) {
                std::swap(a[(size_t)piv * n + q], a[(size_t)c * n + q]);
                std::swap(inv[(size_t)piv * n + q], inv[(size_t)c * n + q]);
            }
        }
        const double d = a[(size_t)c * n + c];
        for (int q = 0; q < n; q++) {
            a[(size_t)c * n + q] /= d;
            inv[(size_t)c * n + q] /= d;
        }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            const double f = a[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int q = 0; q < n; q++) {
                a[(size_t)r * n + q] -= f * a[(size_t)c * n + q];
                inv[(size_t)r * n + q] -= f * inv[(size_t)c * n + q];
            }
        }
    }
    return 0;
}

}  // namespace

// ---- the set-up loop and its providers ----------------------------------------------------------------------------------
namespace {

struct HostPhases : AmgPhases {
    int strength(const AmgLevelHost &L, const lsspg_amg_pars &pr, AmgGraph &S) override
    {
        strong_couplings(L, pr, S);
        return 0;
    }
    int interpolation(AmgLevelHost &L, const AmgGraph &S, const AmgGraph &, const lsspg_amg_pars &pr) override
    {
        direct_interpolation(L, S, pr);
        return 0;
    }
    int restriction(AmgLevelHost &L, const AmgGraph &) override
    {
        transpose_csr(L.n, L.nc, L.Pp, L.Pj, L.Px, L.Rp, L.Rj, L.Rx);
        return 0;
    }
    int galerkin(const AmgLevelHost &L, AmgLevelHost &C) override
    {
        std::vector<int> Tp, Tj;
        std::vector<double> Tx;
        int rc = spgemm(L.n, L.nc, L.Ap, L.Aj, L.Ax, L.Pp, L.Pj, L.Px, Tp, Tj, Tx);
        if (!rc) rc = spgemm(L.nc, L.nc, L.Rp, L.Rj, L.Rx, Tp, Tj, Tx, C.Ap, C.Aj, C.Ax);
        return rc;
    }
    const char *name() const override { return "host"; }
};

// the row functions of amg_rows.cuh, one row after the other: count, scan, fill -- as the device kernels do
struct ReplayPhases : AmgPhases {
    static void scan(std::vector<int> &p)
    {
        int run = 0;
        for (size_t i = 0; i + 1 < p.size(); i++) { const int c = p[i]; p[i] = run; run += c; }
        p.back() = run;
    }
    int strength(const AmgLevelHost &L, const lsspg_amg_pars &pr, AmgGraph &S) override
    {
        S.p.assign((size_t)L.n + 1, 0);
        for (int i = 0; i < L.n; i++) S.p[i] = amg_strong_row(i, L.Ap.data(), L.Aj.data(), L.Ax.data(), pr.strong_threshold, pr.max_row_sum, nullptr);
        scan(S.p);
        S.j.resize((size_t)S.p[L.n]);
        for (int i = 0; i < L.n; i++)
            amg_strong_row(i, L.Ap.data(), L.Aj.data(), L.Ax.data(), pr.strong_threshold, pr.max_row_sum, S.j.data() + S.p[i]);
        return 0;
    }
    int interpolation(AmgLevelHost &L, const AmgGraph &S, const AmgGraph &, const lsspg_amg_pars &pr) override
    {
        const int n = L.n;
        std::vector<int> cidx((size_t)n + 1, 0);
        for (int i = 0; i < n; i++) cidx[i] = (L.cf[i] == kAmgCPT);
        scan(cidx);
        L.nc = cidx[n];
        L.Pp.assign((size_t)n + 1, 0);
        for (int i = 0; i < n; i++)
            L.Pp[i] = amg_interp_row(i, L.Ap.data(), L.Aj.data(), L.Ax.data(), S.p.data(), S.j.data(), L.cf.data(), cidx.data(), pr.trunc_threshold,
                                     nullptr, nullptr);
        scan(L.Pp);
        L.Pj.resize((size_t)L.Pp[n]);
        L.Px.resize((size_t)L.Pp[n]);
        for (int i = 0; i < n; i++)
            amg_interp_row(i, L.Ap.data(), L.Aj.data(), L.Ax.data(), S.p.data(), S.j.data(), L.cf.data(), cidx.data(), pr.trunc_threshold,
                           L.Pj.data() + L.Pp[i], L.Px.data() + L.Pp[i]);
        return 0;
    }
    int restriction(AmgLevelHost &L, const AmgGraph &T) override
    {
        const int n = L.n, nc = L.nc;
        std::vector<int> cpoint((size_t)nc);
        for (int i = 0, c = 0; i < n; i++)
            if (L.cf[i] == kAmgCPT) cpoint[c++] = i;
        L.Rp.assign((size_t)nc + 1, 0);
        for (int c = 0; c < nc; c++)
            L.Rp[c] = amg_restrict_row(c, cpoint[c], T.p.data(), T.j.data(), L.Pp.data(), L.Pj.data(), L.Px.data(), nullptr, nullptr);
        scan(L.Rp);
        L.Rj.resize((size_t)L.Rp[nc]);
        L.Rx.resize((size_t)L.Rp[nc]);
        for (int c = 0; c < nc; c++)
            amg_restrict_row(c, cpoint[c], T.p.data(), T.j.data(), L.Pp.data(), L.Pj.data(), L.Px.data(), L.Rj.data() + L.Rp[c], L.Rx.data() + L.Rp[c]);
        return 0;
    }
    static int product(int nrows, const std::vector<int> &Ap, const std::vector<int> &Aj, const std::vector<double> &Ax,
                       const std::vector<int> &Bp, const std::vector<int> &Bj, const std::vector<double> &Bx, std::vector<int> &Cp,
                       std::vector<int> &Cj, std::vector<double> &Cx)
    {
        int bound = 1;
        for (int i = 0; i < nrows; i++) bound = std::max(bound, amg_spgemm_bound(i, Ap.data(), Aj.data(), Bp.data()));
        int hsize = 64;
        while (hsize < 2 * bound) hsize *= 2;
        std::vector<AmgSlot> tab((size_t)hsize, AmgSlot{-1, -1, 0.0});
        std::vector<int> cols((size_t)bound);
        Cp.assign((size_t)nrows + 1, 0);
        for (int i = 0; i < nrows; i++) {
            Cp[i] = amg_spgemm_row(i, 2 * i, Ap.data(), Aj.data(), Ax.data(), Bp.data(), Bj.data(), Bx.data(), tab.data(), hsize - 1, cols.data(),
                                   bound, nullptr, nullptr);
            AMG_CHECK(Cp[i] >= 0, "amg replay: accumulator overflow in row %d", i);
        }
        scan(Cp);
        Cj.resize((size_t)Cp[nrows]);
        Cx.resize((size_t)Cp[nrows]);
        for (int i = 0; i < nrows; i++)
            amg_spgemm_row(i, 2 * i + 1, Ap.data(), Aj.data(), Ax.data(), Bp.data(), Bj.data(), Bx.data(), tab.data(), hsize - 1, cols.data(), bound,
                           Cj.data() + Cp[i], Cx.data() + Cp[i]);
        return 0;
    }
    int galerkin(const AmgLevelHost &L, AmgLevelHost &C) override
    {
        std::vector<int> Tp, Tj;
        std::vector<double> Tx;
        int rc = product(L.n, L.Ap, L.Aj, L.Ax, L.Pp, L.Pj, L.Px, Tp, Tj, Tx);
        if (!rc) rc = product(L.nc, L.Rp, L.Rj, L.Rx, Tp, Tj, Tx, C.Ap, C.Aj, C.Ax);
        return rc;
    }
    const char *name() const override { return "row functions (CPU replay)"; }
};

}  // namespace

namespace lsspg {

// the C/F splitting of a strength graph for the device set-up (amg_gpu.cu), which has S on the host at that point
int amg_cf_split_host(int n, const AmgGraph &S, AmgGraph &T, std::vector<int> &cf)
{
    transpose_graph(n, S, T);
    return cf_split(n, S, T, cf);
}

int amg_setup_with(AmgPhases &ph, int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                   lsspg_amg_host **out)
{
    AMG_CHECK(out && n > 0 && hAp && hAj && hAx, "lsspg_amg_setup: bad argument");
    lsspg_amg_pars pr;
    if (pars) pr = *pars;
    else lsspg_amg_pars_default(&pr);
    AMG_CHECK(pr.max_levels >= 1 && pr.coarse_dof >= 1 && pr.pre_iter >= 0 && pr.post_iter >= 0,
              "lsspg_amg_setup: bad parameters");
    for (int i = 0; i < n; i++)
        for (int k = hAp[i] + 1; k < hAp[i + 1]; k++)
            AMG_CHECK(hAj[k - 1] < hAj[k], "lsspg_amg_setup: columns of row %d are not sorted", i);
    lsspg_amg_host *H = new lsspg_amg_host();
    H->pars = pr;
    H->levels.emplace_back();
    {
        AmgLevelHost &L = H->levels[0];
        L.n = n;
        L.Ap.assign(hAp, hAp + n + 1);
        L.Aj.assign(hAj, hAj + hAp[n]);
        L.Ax.assign(hAx, hAx + hAp[n]);
    }
    int rc = ph.begin(H->levels[0]);
    const bool prof = getenv("LSSPG_SETUP_PROF") && atoi(getenv("LSSPG_SETUP_PROF")) != 0;
    auto now = [] {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    };
    double tp = now();
    auto PROF = [&](const char *what) {
        if (!prof) return;
        const double t = now();
        fprintf(stderr, "[amg setup, %s] level %d %-22s %.3f s\n", ph.name(), (int)H->levels.size() - 1, what, t - tp);
        tp = t;
    };
    while (!rc && (int)H->levels.size() < pr.max_levels && H->levels.back().n > pr.coarse_dof) {
        AmgLevelHost &L = H->levels.back();
        Graph S, T;
        rc = ph.strength(L, pr, S);
        if (rc) break;
        PROF("strong couplings");
        transpose_graph(L.n, S, T);
        PROF("transpose graph");
        const int nc = cf_split(L.n, S, T, L.cf);
        PROF("C/F split");
        if (nc == 0 || nc >= L.n) {   // coarsening stalled: this level is the last one
            L.cf.clear();
            break;
        }
        rc = ph.interpolation(L, S, T, pr);
        if (rc) break;
        PROF("interpolation");
        visiting_ranks(L, pr.cf_order);
        PROF("visiting ranks");
        rc = ph.restriction(L, T);
        if (rc) break;
        PROF("restriction");
        AmgLevelHost C;
        C.n = L.nc;
        rc = ph.galerkin(L, C);
        if (rc) break;
        PROF("Galerkin products");
        if (pr.verb > 0)
            printf("amg: level %d: n = %d, nnz = %d, C points = %d\n", (int)H->levels.size() - 1, L.n, L.Ap[L.n], L.nc);
        H->levels.push_back(std::move(C));
    }
    if (!rc) {
        AmgLevelHost &L = H->levels.back();
        L.nc = 0;
        L.cf.assign(L.n, 1);
        visiting_ranks(L, 0);   // the last level is swept in natural order (when it is swept at all)
        L.Pp.clear();
        L.Rp.clear();
        if (L.n <= pr.coarse_dense_max) {
            rc = dense_inverse(L, H->coarse_inv);
            H->coarse_dense = (rc == 0);
        }
        if (pr.verb > 0)
            printf("amg: level %d (last): n = %d, nnz = %d, %s\n", (int)H->levels.size() - 1, L.n, L.Ap[L.n],
                   H->coarse_dense ? "dense inverse" : "Gauss-Seidel sweeps");
    }
    if (rc) {
        delete H;
        return rc;
    }
    *out = H;
    return 0;
}

int gs_build_host(int n, const int *Ap, const int *Aj, const double *Ax, const int *cf, const int *rank, GsHost &G,
                  int mode)
{
    G.n = n;
    AMG_CHECK(n < (1 << 29), "amg: level too large for the smoother's column encoding");
    auto blk = [&](int i) { return cf ? cf[i] : 1; };
    auto rk = [&](int i) { return rank ? rank[i] : i; };
    // rows in visiting order of their block (ranks are distinct inside a block)
    std::vector<int> by_rank(n);
    for (int i = 0; i < n; i++) by_rank[i] = i;
    if (rank) {
        bool ascending = true;   // visiting by index (cf_order 0 / 1): already in order
        for (int i = 1; i < n && ascending; i++) ascending = rank[i - 1] <= rank[i];
        if (!ascending)
            std::sort(by_rank.begin(), by_rank.end(), [&](int a, int b) { return rank[a] < rank[b] || (rank[a] == rank[b] && a < b); });
    }
    std::vector<int> lev(n, 0);
    int nlev[2] = {0, 0};   // [0] F block, [1] C block
    for (int t = 0; t < n; t++) {
        const int i = by_rank[t];
        const int mine = blk(i);
        int l = 0;
        bool has_diag = false;
        for (int k = Ap[i]; k < Ap[i + 1]; k++) {
            const int c = Aj[k];
            AMG_CHECK(c >= 0 && c < n, "amg: row %d references column %d", i, c);
            if (c == i) {
                has_diag = Ax[k] != 0.0;
                continue;
            }
            if (blk(c) == mine && rk(c) < rk(i)) l = std::max(l, lev[c] + 1);
        }
        AMG_CHECK(has_diag, "amg: row %d has no (or a zero) diagonal entry", i);
        lev[i] = l;
        nlev[mine] = std::max(nlev[mine], l + 1);
    }
    G.levels_c = nlev[1];
    G.levels_f = nlev[0];
    // order: C block by level, then F block by level; ascending rank inside a level
    const int total_lev = nlev[1] + nlev[0];
    std::vector<int> start(total_lev + 1, 0);
    auto bucket = [&](int i) { return blk(i) ? lev[i] : nlev[1] + lev[i]; };
    for (int i = 0; i < n; i++) start[bucket(i) + 1]++;
    for (int l = 0; l < total_lev; l++) start[l + 1] += start[l];
    std::vector<int> order(n);
    {
        std::vector<int> pos(start.begin(), start.end() - 1);
        for (int t = 0; t < n; t++) order[pos[bucket(by_rank[t])]++] = by_rank[t];
    }
    auto encode = [&](int i, int c) { return (c << 2) | ((blk(c) == blk(i) && rk(c) < rk(i)) ? 2 : 0) | (blk(c) & 1); };
    // deep schedules of wide rows: one ticket per ROW (a warp works on it), entries contiguous
    // (wide rows in 32-row slices would walk the row in chunks, two memory latencies per chunk and level)
    if (mode < 0) mode = (n > 0 && (double)(Ap[n] - n) / n > 28.0) ? 1 : 0;
    G.mode = mode;
    if (mode == 1) {
        G.num_slices = n;
        G.slices_c = nlev[1] > 0 ? start[nlev[1]] : 0;
        G.perm = order;
        G.diag.assign(n, 1.0);
        G.slice_ptr.assign((size_t)n + 1, 0);
        G.offdiag_nnz = (long long)Ap[n] - n;
        G.padded_nnz = G.offdiag_nnz;
        G.col.resize((size_t)G.offdiag_nnz);
        G.val.resize((size_t)G.offdiag_nnz);
        {
            int k = 0;
            for (int p = 0; p < n; p++) {
                G.slice_ptr[p] = k;
                k += Ap[order[p] + 1] - Ap[order[p]] - 1;
            }
            G.slice_ptr[n] = k;
        }
        lsspg::parallel_ranges(n, [&](long long p0, long long p1, int) {
            for (int p = (int)p0; p < (int)p1; p++) {
                const int i = order[p];
                int k = G.slice_ptr[p];
                for (int q = Ap[i]; q < Ap[i + 1]; q++) {
                    const int c = Aj[q];
                    if (c == i) {
                        G.diag[p] = Ax[q];
                        continue;
                    }
                    G.col[k] = encode(i, c);
                    G.val[k] = Ax[q];
                    k++;
                }
            }
        });
        return 0;
    }
    long long nslices = 0, cslices = 0;
    G.level_ptr.assign((size_t)total_lev + 1, 0);
    for (int l = 0; l < total_lev; l++) {
        G.level_ptr[l] = (int)nslices;
        nslices += (start[l + 1] - start[l] + 31) / 32;
        if (l == nlev[1] - 1) cslices = nslices;
    }
    G.level_ptr[total_lev] = (int)nslices;
    AMG_CHECK(nslices * 32 < (1ll << 31), "amg: too many slices");
    G.num_slices = (int)nslices;
    G.slices_c = (int)cslices;
    G.perm.assign((size_t)nslices * 32, -1);
    G.diag.assign((size_t)nslices * 32, 1.0);
    G.slice_ptr.assign((size_t)nslices + 1, 0);
    long long s = 0, wsum = 0;
    G.offdiag_nnz = 0;
    for (int l = 0; l < total_lev; l++) {
        for (int r0 = start[l]; r0 < start[l + 1]; r0 += 32, s++) {
            const int cnt = std::min(32, start[l + 1] - r0);
            int w = 0;
            for (int q = 0; q < cnt; q++) {
                const int i = order[r0 + q];
                G.perm[s * 32 + q] = i;
                w = std::max(w, Ap[i + 1] - Ap[i] - 1);
                G.offdiag_nnz += Ap[i + 1] - Ap[i] - 1;
            }
            G.slice_ptr[s] = (int)wsum;
            wsum += w;
            AMG_CHECK(wsum < (1ll << 31) / 32, "amg: padded smoother layout too large for int32 offsets");
        }
    }
    G.slice_ptr[nslices] = (int)wsum;
    G.padded_nnz = wsum * 32;
    G.col.resize((size_t)G.padded_nnz);
    G.val.resize((size_t)G.padded_nnz);
    lsspg::parallel_ranges(nslices, [&](long long sl0, long long sl1, int) {
    for (long long sl = sl0; sl < sl1; sl++) {
        const long long base = (long long)G.slice_ptr[sl] * 32, wid = G.slice_ptr[sl + 1] - G.slice_ptr[sl];
        for (long long e = base; e < base + wid * 32; e++) { G.col[e] = -1; G.val[e] = 0.0; }
        for (int q = 0; q < 32; q++) {
            const int i = G.perm[sl * 32 + q];
            if (i < 0) continue;
            int k = 0;
            for (int p = Ap[i]; p < Ap[i + 1]; p++) {
                const int c = Aj[p];
                if (c == i) {
                    G.diag[sl * 32 + q] = Ax[p];
                    continue;
                }
                G.col[base + (long long)k * 32 + q] = encode(i, c);
                G.val[base + (long long)k * 32 + q] = Ax[p];
                k++;
            }
        }
    }
    }, 0, 256);
    return 0;
}

}  // namespace lsspg

extern "C" {

int lsspg_amg_pars_default(lsspg_amg_pars *p)
{
    AMG_CHECK(p, "lsspg_amg_pars_default: NULL");
    p->max_levels = 30;
    p->coarse_dof = 100;
    p->strong_threshold = 0.3;
    p->max_row_sum = 0.9;
    p->trunc_threshold = 0.2;
    p->pre_iter = 2;
    p->post_iter = 2;
    p->cf_order = 1;
    p->zero_guess = 0;
    p->coarse_dense_max = 4096;
    p->coarse_sweeps = 40;
    p->tol = 1e-8;
    p->maxit = 100;
    p->verb = 0;
    return 0;
}

int lsspg_amg_setup_host(int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                         lsspg_amg_host **out)
{
    HostPhases ph;
    return amg_setup_with(ph, n, hAp, hAj, hAx, pars, out);
}

/* CPU replay of the device set-up (amg_gpu.cu) for the test-suite: the SAME row functions (amg_rows.cuh), row after row.
 * The hierarchy must equal lsspg_amg_setup_host's array by array. */
int lsspg_debug_amg_setup_replay_host(int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                                      lsspg_amg_host **out)
{
    ReplayPhases ph;
    return amg_setup_with(ph, n, hAp, hAj, hAx, pars, out);
}

int lsspg_amg_host_levels(const lsspg_amg_host *H, int *num_levels, int *coarse_dense)
{
    AMG_CHECK(H, "lsspg_amg_host_levels: NULL");
    if (num_levels) *num_levels = (int)H->levels.size();
    if (coarse_dense) *coarse_dense = H->coarse_dense ? 1 : 0;
    return 0;
}

int lsspg_amg_host_level_sizes(const lsspg_amg_host *H, int l, int *n, int *nc, int *nnzA, int *nnzP, int *nnzR)
{
    AMG_CHECK(H && l >= 0 && l < (int)H->levels.size(), "lsspg_amg_host_level_sizes: bad level");
    const AmgLevelHost &L = H->levels[l];
    if (n) *n = L.n;
    if (nc) *nc = L.nc;
    if (nnzA) *nnzA = L.Ap[L.n];
    if (nnzP) *nnzP = L.Pp.empty() ? 0 : L.Pp[L.n];
    if (nnzR) *nnzR = L.Rp.empty() ? 0 : L.Rp[L.nc];
    return 0;
}

int lsspg_amg_host_level_get(const lsspg_amg_host *H, int l, int *Ap, int *Aj, double *Ax, int *Pp, int *Pj,
                             double *Px, int *Rp, int *Rj, double *Rx, int *cf)
{
    AMG_CHECK(H && l >= 0 && l < (int)H->levels.size(), "lsspg_amg_host_level_get: bad level");
    const AmgLevelHost &L = H->levels[l];
    copy_out(Ap, L.Ap); copy_out(Aj, L.Aj); copy_out(Ax, L.Ax);
    copy_out(Pp, L.Pp); copy_out(Pj, L.Pj); copy_out(Px, L.Px);
    copy_out(Rp, L.Rp); copy_out(Rj, L.Rj); copy_out(Rx, L.Rx);
    copy_out(cf, L.cf);
    return 0;
}

int lsspg_amg_host_level_rank(const lsspg_amg_host *H, int l, int *rank)
{
    AMG_CHECK(H && l >= 0 && l < (int)H->levels.size() && rank, "lsspg_amg_host_level_rank: bad argument");
    copy_out(rank, H->levels[l].rank);
    return 0;
}

int lsspg_amg_host_coarse_inverse(const lsspg_amg_host *H, double *inv)
{
    AMG_CHECK(H && inv && H->coarse_dense, "lsspg_amg_host_coarse_inverse: no dense inverse");
    copy_out(inv, H->coarse_inv);
    return 0;
}

int lsspg_amg_host_pars(const lsspg_amg_host *H, lsspg_amg_pars *pars)
{
    AMG_CHECK(H && pars, "lsspg_amg_host_pars: NULL");
    *pars = H->pars;
    return 0;
}

int lsspg_amg_host_destroy(lsspg_amg_host *H)
{
    delete H;
    return 0;
}

int lsspg_debug_amg_walk_gs_host(const lsspg_amg_host *H, int l, int post, const double *hb, const double *hx_old,
                                 double *hx_new, int *info)
{
    AMG_CHECK(H && l >= 0 && l < (int)H->levels.size() && hb && hx_old && hx_new, "lsspg_debug_amg_walk_gs_host: bad argument");
    const AmgLevelHost &L = H->levels[l];
    GsHost G;
    const bool cf_on = H->pars.cf_order && l + 1 < (int)H->levels.size();
    const int mode = (post >> 1) - 1;   // post bits 1..2: 0 = schedule chosen as on the device, 1 = slices, 2 = rows
    post &= 1;
    if (gs_build_host(L.n, L.Ap.data(), L.Aj.data(), L.Ax.data(), cf_on ? L.cf.data() : nullptr,
                      cf_on ? L.rank.data() : nullptr, G, mode)) return 1;
    std::vector<char> written(L.n, 0);
    const int nf = G.num_slices - G.slices_c;
    if (info) {
        info[0] = G.num_slices;
        info[1] = G.levels_c;
        info[2] = G.levels_f;
        info[3] = G.mode == 1 ? -1 : (int)(G.padded_nnz / 32);
    }
    if (G.mode == 1) {   // one ticket per row
        for (int t = 0; t < G.num_slices; t++) {
            const int p = post ? (t < nf ? G.slices_c + t : t - nf) : t;
            const bool row_c = p < G.slices_c;
            const int row = G.perm[p];
            double r = hb[row];
            for (int k = G.slice_ptr[p]; k < G.slice_ptr[p + 1]; k++) {
                const int enc = G.col[k], c = enc >> 2;
                const bool col_c = enc & 1;
                const bool is_new = (col_c == row_c) ? ((enc & 2) != 0) : (col_c == !post);
                if (is_new) AMG_CHECK(written[c], "amg walk: row %d (ticket %d) needs x[%d] before it is written", row, t, c);
                r = r - G.val[k] * (is_new ? hx_new[c] : hx_old[c]);
            }
            hx_new[row] = r / G.diag[p];
            written[row] = 1;
        }
        return 0;
    }
    for (int t = 0; t < G.num_slices; t++) {
        const int s = post ? (t < nf ? G.slices_c + t : t - nf) : t;
        const bool row_c = s < G.slices_c;
        const long long base = (long long)G.slice_ptr[s] * 32;
        const int w = G.slice_ptr[s + 1] - G.slice_ptr[s];
        double res[32];
        for (int q = 0; q < 32; q++) {
            const int row = G.perm[(long long)s * 32 + q];
            if (row < 0) continue;
            double r = hb[row];
            for (int k = 0; k < w; k++) {
                const int enc = G.col[base + (long long)k * 32 + q];
                if (enc < 0) continue;
                const int c = enc >> 2;
                const bool col_c = enc & 1;
                const bool is_new = (col_c == row_c) ? ((enc & 2) != 0) : (col_c == !post);
                if (is_new) AMG_CHECK(written[c], "amg walk: row %d (ticket %d) needs x[%d] before it is written", row, t, c);
                r = r - G.val[base + (long long)k * 32 + q] * (is_new ? hx_new[c] : hx_old[c]);
            }
            res[q] = r / G.diag[(long long)s * 32 + q];
        }
        for (int q = 0; q < 32; q++) {   // a slice publishes its rows together
            const int row = G.perm[(long long)s * 32 + q];
            if (row < 0) continue;
            hx_new[row] = res[q];
            written[row] = 1;
        }
    }
    return 0;
}

}  // extern "C"

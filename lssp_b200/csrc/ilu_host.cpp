// ilu_host.cpp -- host-side incomplete factorisations that feed L and U to the
// GPU triangular solves.  Setup code ("apply on GPU,
// setup on host in phase 1"), written fresh but arithmetically identical to the
// reference so that the factors -- and therefore every preconditioned iteration
// -- match bit for bit:
//   * missing-diagonal repair            (reference src/matrix-utils.cxx:483-587)
//   * uniform block-diagonal extraction  (src/matrix-utils.cxx:589-698) -- this is
//     also the definition of the block-Jacobi preconditioner used when the
//     matrix is row-sharded across GPUs
//   * ILU(k): level-of-fill symbolic phase with the reference's level-RAISING
//     rule (src/pc-iluk.cxx:22-135, esp. :84-86,:101), IKJ numeric phase with
//     its pivot repair (src/pc-iluk.cxx:347-409), L/U split with the unit
//     diagonal appended to L (src/pc-iluk.cxx:497-532)
//   * ILUT(p, tau): row-wise dual-threshold factorisation including the
//     quick-select ordering of the kept entries (src/pc-ilut.cxx:7-286) and the
//     same L/U split (src/pc-ilut.cxx:375-402).
// No FMA: built with -ffp-contract=off.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include <limits>
#include <numeric>
#include <vector>
#include "../../include/lsspg.h"
#include "host_par.h"

namespace lsspg {
void set_error(const char *fmt, ...);
}

namespace {

// LSSPG_SETUP_PROF=1: phase times of the set-up on stderr
static double prof_now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static bool prof_on()
{
    static const bool on = getenv("LSSPG_SETUP_PROF") && atoi(getenv("LSSPG_SETUP_PROF")) != 0;
    return on;
}
#define PROF_T0 double prof_t = prof_now()
#define PROF(what)                                                                   \
    do {                                                                             \
        if (prof_on()) {                                                             \
            const double t_ = prof_now();                                            \
            fprintf(stderr, "[setup] %-26s %.3f s\n", what, t_ - prof_t);            \
            prof_t = t_;                                                             \
        }                                                                            \
    } while (0)

constexpr double kPivotTol = 1e-10;   // mat_zero_diag_tol,   reference src/pc.cxx:7
constexpr double kPivotValue = 1e-3;  // mat_zero_diag_value, reference src/pc.cxx:6

using lsspg::DVec;
using lsspg::IVec;
using lsspg::parallel_copy;
using lsspg::parallel_exclusive_scan;
using lsspg::parallel_ranges;

struct Csr {
    int n = 0;
    IVec p, j;
    DVec x;
    int nnz() const { return p.empty() ? 0 : p[n]; }
};

// rows sorted by ascending column (what lssp_solver_assemble guarantees, src/lssp.cxx:173)
void sort_rows(Csr &A)
{
    std::vector<int> idx;
    std::vector<int> tj;
    std::vector<double> tx;
    for (int i = 0; i < A.n; i++) {
        const int b = A.p[i], e = A.p[i + 1];
        if (std::is_sorted(A.j.begin() + b, A.j.begin() + e)) continue;
        idx.resize(e - b);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int c) { return A.j[b + a] < A.j[b + c]; });
        tj.assign(A.j.begin() + b, A.j.begin() + e);
        tx.assign(A.x.begin() + b, A.x.begin() + e);
        for (int k = 0; k < e - b; k++) {
            A.j[b + k] = tj[idx[k]];
            A.x[b + k] = tx[idx[k]];
        }
    }
}

// A row without a stored diagonal receives (i, tol), slid into sorted position.
Csr with_diagonal(const Csr &A, double tol)
{
    Csr M;
    M.n = A.n;
    M.p.assign(A.n + 1, 0);
    std::vector<char> has(A.n, 0);
    for (int i = 0; i < A.n; i++) {
        for (int k = A.p[i]; k < A.p[i + 1]; k++)
            if (A.j[k] == i) has[i] = 1;
        M.p[i + 1] = M.p[i] + (A.p[i + 1] - A.p[i]) + (has[i] ? 0 : 1);
    }
    M.j.resize(M.p[A.n]);
    M.x.resize(M.p[A.n]);
    for (int i = 0; i < A.n; i++) {
        int o = M.p[i];
        for (int k = A.p[i]; k < A.p[i + 1]; k++, o++) {
            M.j[o] = A.j[k];
            M.x[o] = A.x[k];
        }
        if (!has[i]) {
            M.j[o] = i;
            M.x[o] = 1 * tol;
            // slide towards the front while the left neighbour has a larger column
            for (int q = o; q > M.p[i] && M.j[q - 1] > M.j[q]; q--) {
                std::swap(M.j[q - 1], M.j[q]);
                std::swap(M.x[q - 1], M.x[q]);
            }
        }
    }
    return M;
}

// Entries outside the row's own block are discarded; a row left empty becomes a unit row.
Csr block_diagonal(Csr &&A, int bs)
{
    if (bs >= A.n) return std::move(A);
    Csr M;
    const int n = A.n;
    M.n = n;
    M.p.resize((size_t)n + 1);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            const int lo = (i / bs) * bs, hi = std::min(n, lo + bs);
            int kept = 0;
            for (int k = A.p[i]; k < A.p[i + 1]; k++) kept += (A.j[k] >= lo && A.j[k] < hi);
            M.p[i] = kept ? kept : 1;
        }
    });
    const long long total = parallel_exclusive_scan(M.p.data(), n);
    M.p[n] = (int)total;
    M.j.resize((size_t)total);
    M.x.resize((size_t)total);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            const int lo = (i / bs) * bs, hi = std::min(n, lo + bs);
            int o = M.p[i];
            for (int k = A.p[i]; k < A.p[i + 1]; k++) {
                if (A.j[k] >= lo && A.j[k] < hi) {
                    M.j[o] = A.j[k];
                    M.x[o] = A.x[k];
                    o++;
                }
            }
            if (o == M.p[i]) {
                M.j[o] = i;
                M.x[o] = 1.0;
            }
        }
    });
    return M;
}

// The caller's matrix -> the matrix the factorisations work on: rows sorted by column (what
// lssp_solver_assemble guarantees, src/lssp.cxx:173), missing diagonals inserted (src/pc-iluk.cxx:573,
// src/pc-ilut.cxx:448).  The usual case -- sorted rows that all store their diagonal -- is ONE copy, made
// by the host threads; otherwise the serial repair path runs.
Csr ingest(int n, const int *Ap, const int *Aj, const double *Ax)
{
    const size_t nnz = (size_t)Ap[n];
    std::vector<char> bad(lsspg::host_threads() + 1, 0);
    parallel_ranges(n, [&](long long r0, long long r1, int piece) {
        char b = 0;
        for (int i = (int)r0; i < (int)r1 && !b; i++) {
            bool diag = false;
            for (int k = Ap[i]; k < Ap[i + 1]; k++) {
                if (Aj[k] == i) diag = true;
                if (k > Ap[i] && Aj[k - 1] > Aj[k]) b = 1;
            }
            if (!diag) b = 1;
        }
        bad[piece] = b;
    }, lsspg::host_threads());
    Csr A;
    A.n = n;
    A.p.resize((size_t)n + 1);
    A.j.resize(nnz);
    A.x.resize(nnz);
    parallel_copy(A.p.data(), Ap, sizeof(int) * ((size_t)n + 1));
    parallel_copy(A.j.data(), Aj, sizeof(int) * nnz);
    parallel_copy(A.x.data(), Ax, sizeof(double) * nnz);
    if (std::find(bad.begin(), bad.end(), (char)1) == bad.end()) return A;
    sort_rows(A);
    return with_diagonal(A, kPivotTol);
}

// ---- ILU(k) symbolic ----------------------------------------------------------
// Returns A's values scattered into the level-<=`level` fill pattern (zeros on
// fill), columns sorted.  Level rule: a candidate reached through pivot k gets
// lev(i,k) + lev(k,c) + 1; candidates above `level` are ignored; an entry that
// is ALREADY in the row has its level RAISED to the candidate's when that is
// larger (the reference's non-textbook rule).
Csr iluk_pattern_general(const Csr &A, int level)
{
    const int n = A.n;
    std::vector<std::vector<int>> ucol(n), ulev(n);   // strict upper part of every finished row
    std::vector<int> where(n, -1);                    // column -> index in the row's work lists
    std::vector<int> lc, ll, uc, ul;                  // lower cols/levels, upper cols/levels of row i
    Csr M;
    M.n = n;
    M.p.assign(n + 1, 0);
    std::vector<double> dense(n, 0.0);
    for (int i = 0; i < n; i++) {
        lc.clear(); ll.clear(); uc.clear(); ul.clear();
        for (int k = A.p[i]; k < A.p[i + 1]; k++) {
            const int c = A.j[k];
            if (c < i) { where[c] = (int)lc.size(); lc.push_back(c); ll.push_back(0); }
            else if (c > i) { where[c] = n + (int)uc.size(); uc.push_back(c); ul.push_back(0); }
        }
        for (size_t t = 0; t < lc.size(); t++) {
            // next pivot = smallest not-yet-used lower column (fill may have appended more)
            size_t m = t;
            for (size_t q = t + 1; q < lc.size(); q++)
                if (lc[q] < lc[m]) m = q;
            if (m != t) {
                std::swap(lc[t], lc[m]);
                std::swap(ll[t], ll[m]);
                where[lc[t]] = (int)t;
                where[lc[m]] = (int)m;
            }
            const int piv = lc[t];
            for (size_t q = 0; q < ucol[piv].size(); q++) {
                const int c = ucol[piv][q];
                const int cand = ulev[piv][q] + ll[t] + 1;
                if (cand > level) continue;
                const int w = where[c];
                if (w < 0) {
                    if (c < i) { where[c] = (int)lc.size(); lc.push_back(c); ll.push_back(cand); }
                    else if (c > i) { where[c] = n + (int)uc.size(); uc.push_back(c); ul.push_back(cand); }
                }
                else if (w >= n) { if (ul[w - n] < cand) ul[w - n] = cand; }
                else { if (ll[w] < cand) ll[w] = cand; }
            }
        }
        for (int c : lc) where[c] = -1;
        for (int c : uc) where[c] = -1;
        ucol[i] = uc;
        ulev[i] = ul;
        // emit the row: sorted pattern, A's values where present, 0 on fill
        for (int k = A.p[i]; k < A.p[i + 1]; k++) dense[A.j[k]] = A.x[k];
        std::vector<int> cols(lc);
        cols.push_back(i);
        cols.insert(cols.end(), uc.begin(), uc.end());
        std::sort(cols.begin(), cols.end());
        for (int c : cols) {
            M.j.push_back(c);
            M.x.push_back(dense[c]);
        }
        for (int k = A.p[i]; k < A.p[i + 1]; k++) dense[A.j[k]] = 0.0;
        M.p[i + 1] = (int)M.j.size();
    }
    return M;
}

// Chunk length for RowPipeline: the smallest sub-diagonal offset i - c > 1 that at least a quarter of the sampled
// rows have (the innermost grid dimension of a stencil matrix: a chunk is then one grid line, whose rows chain
// through i - 1 inside the chunk and meet the previous line one row behind), else 256.
int pipeline_chunk(const Csr &A)
{
    if (const char *e = getenv("LSSPG_PIPE_CHUNK"))
        if (atoi(e) > 0) return atoi(e);
    const int n = A.n, rows = std::min(n, 1 << 16), r0 = (n - rows) / 2;
    std::vector<std::pair<int, int>> hist;   // (offset, count)
    for (int i = r0; i < r0 + rows; i++) {
        for (int k = A.p[i]; k < A.p[i + 1]; k++) {
            const int d = i - A.j[k];
            if (d <= 1 || d > 8192) continue;
            size_t q = 0;
            for (; q < hist.size() && hist[q].first != d; q++) {}
            if (q == hist.size()) {
                if (hist.size() >= 64) continue;
                hist.emplace_back(d, 0);
            }
            hist[q].second++;
        }
    }
    int best = 0;
    for (auto &h : hist)
        if (h.second >= rows / 4 && (best == 0 || h.first < best)) best = h.first;
    return best ? std::max(best, 16) : 256;
}

// The same symbolic phase for the usual input (strictly ascending columns): the upper parts of the finished rows
// live in one flat array instead of two heap vectors per row, the rows of the pattern are appended to M.j directly,
// and the values (A's where present, 0 on fill) are merged in afterwards by the host threads.  The row recurrence
// itself -- pivots in ascending order, candidates in the order the pivot row discovered them, the level-raising
// rule -- is the one of iluk_pattern_general, statement for statement.
Csr iluk_pattern(const Csr &A, int level)
{
    const int n = A.n;
    {
        const int np = lsspg::host_threads();
        std::vector<char> loose(np, 0);
        parallel_ranges(n, [&](long long r0, long long r1, int p) {
            for (int i = (int)r0; i < (int)r1; i++)
                for (int k = A.p[i] + 1; k < A.p[i + 1]; k++)
                    if (A.j[k - 1] >= A.j[k]) { loose[p] = 1; return; }
        }, np);
        if (std::find(loose.begin(), loose.end(), (char)1) != loose.end()) return iluk_pattern_general(A, level);
    }
    // per finished row one arena record [sorted pattern (len) | upper columns (nu) | their levels (nu)], upper part in
    // discovery order; published through the pipeline's progress counters
    PROF_T0;
    std::vector<const int *> rec((size_t)n);
    IVec len((size_t)n + 1), nup((size_t)n);
    // every pipeline thread owns a column map of n ints: at most ~4 GB of them
    const int nt = (int)std::max<long long>(1, std::min<long long>(lsspg::host_threads(), (1ll << 30) / std::max(n, 1)));
    lsspg::RowPipeline pipe(n, pipeline_chunk(A));
    struct Scratch {
        std::vector<int> where, lc, ll, uc, ul, cols;
        lsspg::BlockArena<int> arena;
    };
    std::vector<Scratch> scratch(nt);
    pipe.run(nt, [&](int t, int i, int chunk_begin) {
        Scratch &S = scratch[t];
        if (S.where.empty()) S.where.assign(n, -1);   // column -> index in the row's work lists
        std::vector<int> &where = S.where, &lc = S.lc, &ll = S.ll, &uc = S.uc, &ul = S.ul, &cols = S.cols;
        lc.clear(); ll.clear(); uc.clear(); ul.clear();
        for (int k = A.p[i]; k < A.p[i + 1]; k++) {
            const int c = A.j[k];
            if (c < i) { where[c] = (int)lc.size(); lc.push_back(c); ll.push_back(0); }
            else if (c > i) { where[c] = n + (int)uc.size(); uc.push_back(c); ul.push_back(0); }
        }
        for (size_t tt = 0; tt < lc.size(); tt++) {
            // next pivot = smallest not-yet-used lower column (fill may have appended more)
            size_t m = tt;
            for (size_t q = tt + 1; q < lc.size(); q++)
                if (lc[q] < lc[m]) m = q;
            if (m != tt) {
                std::swap(lc[tt], lc[m]);
                std::swap(ll[tt], ll[m]);
                where[lc[tt]] = (int)tt;
                where[lc[m]] = (int)m;
            }
            const int piv = lc[tt], lt = ll[tt];
            pipe.wait(piv, chunk_begin);
            const int pn = nup[piv];
            const int *pc = rec[piv] + len[piv], *pl = pc + pn;
            for (int q = 0; q < pn; q++) {
                const int c = pc[q];
                const int cand = pl[q] + lt + 1;
                if (cand > level) continue;
                const int w = where[c];
                if (w < 0) {
                    if (c < i) { where[c] = (int)lc.size(); lc.push_back(c); ll.push_back(cand); }
                    else if (c > i) { where[c] = n + (int)uc.size(); uc.push_back(c); ul.push_back(cand); }
                }
                else if (w >= n) { if (ul[w - n] < cand) ul[w - n] = cand; }
                else { if (ll[w] < cand) ll[w] = cand; }
            }
        }
        for (int c : lc) where[c] = -1;
        for (int c : uc) where[c] = -1;
        cols.assign(lc.begin(), lc.end());
        cols.push_back(i);
        cols.insert(cols.end(), uc.begin(), uc.end());
        std::sort(cols.begin(), cols.end());
        int *r = S.arena.take(cols.size() + 2 * uc.size());
        std::copy(cols.begin(), cols.end(), r);
        std::copy(uc.begin(), uc.end(), r + cols.size());
        std::copy(ul.begin(), ul.end(), r + cols.size() + uc.size());
        len[i] = (int)cols.size();
        nup[i] = (int)uc.size();
        rec[i] = r;
    });
    PROF("  symbolic rows");
    Csr M;
    M.n = n;
    M.p.resize((size_t)n + 1);
    parallel_copy(M.p.data(), len.data(), sizeof(int) * (size_t)n);
    const long long total = parallel_exclusive_scan(M.p.data(), n);
    M.p[n] = (int)total;
    M.j.resize((size_t)total);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) std::copy(rec[i], rec[i] + len[i], M.j.data() + M.p[i]);
    });
    M.x.resize(M.j.size());
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            int a = A.p[i];
            const int ae = A.p[i + 1];
            for (int k = M.p[i]; k < M.p[i + 1]; k++) {
                const int c = M.j[k];
                while (a < ae && A.j[a] < c) a++;
                M.x[k] = (a < ae && A.j[a] == c) ? A.x[a] : 0.0;
            }
        }
    });
    return M;
}

inline double repaired_pivot_signed(double d)
{
    if (fabs(d) < kPivotTol) return d > 0 ? kPivotValue : -kPivotValue;
    return d;
}

// ---- ILU(0) numeric, in place on rows [r0, r1) of M (columns are global and
// confined to the block; sorted).  IKJ variant.
void ilu0_block(Csr &M, int r0, int r1, std::vector<double> &wk, std::vector<double> &inv)
{
    const int *P = M.p.data();
    int *C = M.j.data();
    double *X = M.x.data();
    // first row of the block: its leading entry is taken as the pivot; the
    // repaired value only enters the inverse, the stored entry is left alone
    inv[r0] = 1. / repaired_pivot_signed(X[P[r0]]);
    for (int i = r0 + 1; i < r1; i++) {
        const int e = P[i + 1];
        int k = P[i];
        for (; C[k] < i; k++) {
            const int pr = C[k];
            for (int q = P[pr]; q < P[pr + 1]; q++) wk[C[q]] = X[q];
            const double a_ik = X[k] = X[k] * inv[pr];
            for (int q = k + 1; q < e; q++)
                if (wk[C[q]] != 0.) X[q] = X[q] - a_ik * wk[C[q]];
            for (int q = P[pr]; q < P[pr + 1]; q++) wk[C[q]] = 0;
        }
        double d = kPivotValue;
        if (C[k] == i) {
            if (fabs(X[k]) < kPivotTol) X[k] = kPivotValue;
            d = X[k];
        }
        inv[i] = 1. / d;
    }
}

// The same factorisation, level-scheduled over the host threads: row i needs the FINISHED rows of its strictly
// lower columns, so rows are grouped by the level of that dependency graph (one serial O(nnz) pass) and the rows of
// a level are independent.  A row scales each lower entry by the pivot's inverse and subtracts a_ik * a_kj from
// the entries it shares with row k -- found by a two-pointer merge of the sorted column lists instead of the
// dense scatter of ilu0_block; same operations, same order, so the factors are bit-identical (the GPU numeric
// phase, ilu_gpu.cu, does the same).  Rows with repeated columns would not merge like the scatter: the serial loop
// handles such a matrix.
void ilu0_levels(Csr &M, int bs)
{
    const int n = M.n;
    const int *P = M.p.data();
    const int *C = M.j.data();
    double *X = M.x.data();
    const int np = lsspg::host_threads();
    std::vector<char> loose(np, 0);
    parallel_ranges(n, [&](long long r0, long long r1, int p) {
        for (int i = (int)r0; i < (int)r1; i++)
            for (int k = P[i] + 1; k < P[i + 1]; k++)
                if (C[k - 1] >= C[k]) { loose[p] = 1; return; }
    }, np);
    if (np <= 1 || std::find(loose.begin(), loose.end(), (char)1) != loose.end()) {
        std::vector<double> wk(n, 0.0), inv(n, 0.0);
        for (int r0 = 0; r0 < n; r0 += bs) ilu0_block(M, r0, std::min(n, r0 + bs), wk, inv);
        return;
    }
    IVec lev((size_t)n);
    std::vector<int> start(1, 0);
    for (int i = 0; i < n; i++) {
        int l = 0;
        for (int k = P[i]; C[k] < i; k++) l = std::max(l, lev[C[k]] + 1);
        lev[i] = l;
        if ((int)start.size() < l + 2) start.resize(l + 2, 0);
        start[l + 1]++;
    }
    const int nlev = (int)start.size() - 1;
    IVec order((size_t)n);
    for (int l = 0; l < nlev; l++) start[l + 1] += start[l];
    {
        std::vector<int> pos(start.begin(), start.end() - 1);
        for (int i = 0; i < n; i++) order[pos[lev[i]]++] = i;
    }
    DVec inv((size_t)n);
    lsspg::LevelTeam::run(nlev, start.data(), order.data(), [&](int i) {
        const int e = P[i + 1];
        int k = P[i];
        if (i % bs == 0) {   // first row of a block, as in ilu0_block
            inv[i] = 1. / repaired_pivot_signed(X[k]);
            return;
        }
        for (; C[k] < i; k++) {
            const int pr = C[k];
            const double a_ik = X[k] = X[k] * inv[pr];
            int pq = P[pr];
            const int pe = P[pr + 1];
            for (int q = k + 1; q < e; q++) {
                const int c = C[q];
                while (pq < pe && C[pq] < c) pq++;
                if (pq < pe && C[pq] == c) {
                    const double w = X[pq];
                    if (w != 0.) X[q] = X[q] - a_ik * w;
                }
            }
        }
        double d = kPivotValue;
        if (C[k] == i) {
            if (fabs(X[k]) < kPivotTol) X[k] = kPivotValue;
            d = X[k];
        }
        inv[i] = 1. / d;
    });
}

struct Factors {
    int n = 0;
    Csr L, U;
};

// rows of the factored matrix F -> L (strict lower + unit diagonal LAST) and U (diagonal FIRST + strict upper)
void split_row(const int *cj, const double *cx, int len, int row, Csr &L, Csr &U)
{
    for (int k = 0; k < len; k++) {
        if (cj[k] < row) { L.j.push_back(cj[k]); L.x.push_back(cx[k]); }
        else if (cj[k] == row) {
            L.j.push_back(row); L.x.push_back(1);
            U.j.push_back(row); U.x.push_back(cx[k]);
        }
        else { U.j.push_back(cj[k]); U.x.push_back(cx[k]); }
    }
    L.p.push_back((int)L.j.size());
    U.p.push_back((int)U.j.size());
}

// all rows of the factored matrix at once: sizes counted and scanned first, rows written by the host threads
void split_all(const Csr &M, Csr &L, Csr &U)
{
    const int n = M.n;
    L.n = U.n = n;
    L.p.resize((size_t)n + 1);
    U.p.resize((size_t)n + 1);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            int nl = 0, nu = 0;
            for (int k = M.p[i]; k < M.p[i + 1]; k++) {
                const int c = M.j[k];
                nl += (c <= i);
                nu += (c >= i);
            }
            L.p[i] = nl;
            U.p[i] = nu;
        }
    });
    const long long tl = parallel_exclusive_scan(L.p.data(), n), tu = parallel_exclusive_scan(U.p.data(), n);
    L.p[n] = (int)tl;
    U.p[n] = (int)tu;
    L.j.resize((size_t)tl); L.x.resize((size_t)tl);
    U.j.resize((size_t)tu); U.x.resize((size_t)tu);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            int ol = L.p[i], ou = U.p[i];
            for (int k = M.p[i]; k < M.p[i + 1]; k++) {
                const int c = M.j[k];
                const double v = M.x[k];
                if (c < i) { L.j[ol] = c; L.x[ol] = v; ol++; }
                else if (c == i) {
                    L.j[ol] = i; L.x[ol] = 1; ol++;
                    U.j[ou] = i; U.x[ou] = v; ou++;
                }
                else { U.j[ou] = c; U.x[ou] = v; ou++; }
            }
        }
    });
}

Factors factor_iluk(Csr &&A, int level, int bs)
{
    PROF_T0;
    const int n = A.n;
    Csr M = (level > 0) ? block_diagonal(iluk_pattern(A, level), bs) : block_diagonal(std::move(A), bs);
    PROF("pattern+block_diagonal");
    ilu0_levels(M, bs);
    PROF("numeric");
    Factors F;
    F.n = n;
    split_all(M, F.L, F.U);
    PROF("split");
    return F;
}

// ---- ILUT ---------------------------------------------------------------------
// Partial ordering: afterwards the `ncut` entries of largest magnitude occupy
// a[0..ncut).  The exact sequence of exchanges is part of the contract because
// the kept entries are stored -- and later summed by the triangular solves -- in
// the order this leaves them.
void select_largest(double *a, int *ind, int n, int ncut)
{
    int lo = 0, hi = n - 1;
    if (ncut < lo || ncut >= hi) return;
    for (;;) {
        int mid = lo;
        const double key = fabs(a[mid]);
        for (int q = lo + 1; q <= hi; q++) {
            if (fabs(a[q]) > key) {
                ++mid;
                std::swap(a[mid], a[q]);
                std::swap(ind[mid], ind[q]);
            }
        }
        std::swap(a[mid], a[lo]);
        std::swap(ind[mid], ind[lo]);
        if (mid == ncut) return;
        if (mid > ncut) hi = mid - 1;
        else lo = mid + 1;
    }
}

// One block: rows [r0, r1) of B (block-diagonal, global sorted columns).  Emits the
// factored rows straight into L/U.
void ilut_block(const Csr &B, int r0, int r1, double tau, int p, Csr &L, Csr &U)
{
    const int m = r1 - r0;
    // local factored matrix of the block, rows stored [kept lower | diagonal | kept upper], local columns
    std::vector<int> fp(1, 0), fj;
    std::vector<double> fx;
    std::vector<double> w(2 * (size_t)m + 2), diag(m);
    std::vector<int> jw(2 * (size_t)m + 2), jr(m, -1);
    std::vector<int> gcol;
    auto emit = [&](int li) {
        const int b = fp[li], e = fp[li + 1];
        gcol.resize(e - b);
        for (int k = b; k < e; k++) gcol[k - b] = fj[k] + r0;
        split_row(gcol.data(), &fx[b], e - b, r0 + li, L, U);
    };
    // first row: copied verbatim; its leading entry is the pivot
    for (int k = B.p[r0]; k < B.p[r0 + 1]; k++) {
        fj.push_back(B.j[k] - r0);
        fx.push_back(B.x[k]);
    }
    fp.push_back((int)fj.size());
    diag[0] = repaired_pivot_signed(B.x[B.p[r0]]);
    emit(0);
    for (int i = 1; i < m; i++) {
        const int b = B.p[r0 + i], e = B.p[r0 + i + 1];
        double norm = 0.0;
        for (int k = b; k < e; k++) norm += fabs(B.x[k]);
        norm /= (double)(e - b);
        const double drop = tau * norm;
        // work row: lower part at [0, nl), diagonal at i, upper part at (i, i + nu]
        int nl = 0, nu = 0;
        jw[i] = i;
        w[i] = 0.0;
        jr[i] = i;
        for (int k = b; k < e; k++) {
            const int c = B.j[k] - r0;
            if (c < i) { jr[c] = nl; jw[nl] = c; w[nl] = B.x[k]; nl++; }
            else if (c == i) w[i] = B.x[k];
            else { nu++; jr[c] = i + nu; jw[i + nu] = c; w[i + nu] = B.x[k]; }
        }
        for (int t = 0; t < nl; t++) {
            int piv = jw[t], at = t;
            for (int q = t + 1; q < nl; q++)
                if (jw[q] < piv) { piv = jw[q]; at = q; }
            if (at != t) {
                const int c = jw[t];
                jw[t] = jw[at];
                jw[at] = c;
                jr[piv] = t;
                jr[c] = at;
                std::swap(w[t], w[at]);
            }
            jr[piv] = -1;
            const double a_ik = w[t] = w[t] / diag[piv];
            for (int q = fp[piv]; q < fp[piv + 1]; q++) {
                const int c = fj[q];
                if (c <= piv) continue;
                const int at2 = jr[c];
                const double mx = -a_ik * fx[q];
                if (at2 == -1 && fabs(mx) < drop) continue;   // only NEW fill is dropped
                if (at2 != -1) w[at2] += mx;
                else if (c < i) { jw[nl] = c; jr[c] = nl; w[nl] = mx; nl++; }
                else { nu++; jw[i + nu] = c; jr[c] = i + nu; w[i + nu] = mx; }
            }
        }
        diag[i] = w[i];
        jr[i] = -1;
        for (int q = 0; q < nl; q++) jr[jw[q]] = -1;
        for (int q = 1; q <= nu; q++) jr[jw[i + q]] = -1;
        diag[i] = repaired_pivot_signed(diag[i]);
        int keep = std::min(nl, p);
        select_largest(w.data(), jw.data(), nl, keep);
        for (int q = 0; q < keep; q++) { fj.push_back(jw[q]); fx.push_back(w[q]); }
        fj.push_back(i);
        fx.push_back(diag[i]);
        keep = std::min(nu, p);
        select_largest(w.data() + i + 1, jw.data() + i + 1, nu, keep);
        for (int q = 0; q < keep; q++) { fj.push_back(jw[i + 1 + q]); fx.push_back(w[i + 1 + q]); }
        fp.push_back((int)fj.size());
        emit(i);
    }
}

// ILUT over the host threads.  Row i needs the finished rows of the lower entries it keeps while it is eliminated
// -- known only as the elimination proceeds -- so the rows go through a RowPipeline (host_par.h): each thread runs
// whole chunks of consecutive rows and waits for a pivot row of an earlier chunk until its owner has published it.
// A row performs exactly the statements of ilut_block in the same order (the work row is kept as two compact
// arrays, lower part and upper part, instead of one array indexed [0, nl) / (i, i + nu]; positions inside the two
// parts, which is all select_largest and the stored order depend on, are the same), so L and U are bit-identical
// for any number of threads.  Rows with repeated columns take the serial path.
Factors factor_ilut_rows(const Csr &B, double tau, int p, int bs)
{
    const int n = B.n;
    // every pipeline thread owns a column map of n ints: at most ~4 GB of them
    const int nt = (int)std::max<long long>(1, std::min<long long>(lsspg::host_threads(), (1ll << 30) / std::max(n, 1)));
    std::vector<const int *> rcol((size_t)n);
    std::vector<const double *> rval((size_t)n);
    IVec rlen((size_t)n);
    DVec diag((size_t)n);
    struct Scratch {
        std::vector<int> jr, jwl, jwu;
        std::vector<double> wl, wu;
        lsspg::BlockArena<int> ia;
        lsspg::BlockArena<double> da;
    };
    std::vector<Scratch> scratch(nt);
    constexpr int kUp = 1 << 30, kDiag = -2;   // jr: -1 absent, kDiag the diagonal, < kUp lower position, >= kUp upper position
    lsspg::RowPipeline pipe(n, pipeline_chunk(B));
    pipe.run(nt, [&](int t, int i, int chunk_begin) {
        Scratch &S = scratch[t];
        if (S.jr.empty()) S.jr.assign(n, -1);
        std::vector<int> &jr = S.jr, &jwl = S.jwl, &jwu = S.jwu;
        std::vector<double> &wl = S.wl, &wu = S.wu;
        const int b = B.p[i], e = B.p[i + 1];
        if (i % bs == 0) {
            // first row of a block: copied verbatim; its leading entry is the pivot
            int *c = S.ia.take(e - b);
            double *v = S.da.take(e - b);
            std::copy(B.j.begin() + b, B.j.begin() + e, c);
            std::copy(B.x.begin() + b, B.x.begin() + e, v);
            rcol[i] = c; rval[i] = v; rlen[i] = e - b;
            diag[i] = repaired_pivot_signed(B.x[b]);
            return;
        }
        double norm = 0.0;
        for (int k = b; k < e; k++) norm += fabs(B.x[k]);
        norm /= (double)(e - b);
        const double drop = tau * norm;
        int nl = 0, nu = 0;
        double wd = 0.0;
        jwl.clear(); wl.clear(); jwu.clear(); wu.clear();
        jr[i] = kDiag;
        for (int k = b; k < e; k++) {
            const int c = B.j[k];
            if (c < i) { jr[c] = nl; jwl.push_back(c); wl.push_back(B.x[k]); nl++; }
            else if (c == i) wd = B.x[k];
            else { jr[c] = kUp + nu; jwu.push_back(c); wu.push_back(B.x[k]); nu++; }
        }
        for (int tt = 0; tt < nl; tt++) {
            int piv = jwl[tt], at = tt;
            for (int q = tt + 1; q < nl; q++)
                if (jwl[q] < piv) { piv = jwl[q]; at = q; }
            if (at != tt) {
                const int c = jwl[tt];
                jwl[tt] = jwl[at];
                jwl[at] = c;
                jr[piv] = tt;
                jr[c] = at;
                std::swap(wl[tt], wl[at]);
            }
            jr[piv] = -1;
            pipe.wait(piv, chunk_begin);
            const double a_ik = wl[tt] = wl[tt] / diag[piv];
            const int *pc = rcol[piv];
            const double *pv = rval[piv];
            for (int q = 0, qe = rlen[piv]; q < qe; q++) {
                const int c = pc[q];
                if (c <= piv) continue;
                const int at2 = jr[c];
                const double mx = -a_ik * pv[q];
                if (at2 == -1 && fabs(mx) < drop) continue;   // only NEW fill is dropped
                if (at2 == kDiag) wd += mx;
                else if (at2 >= kUp) wu[at2 - kUp] += mx;
                else if (at2 >= 0) wl[at2] += mx;
                else if (c < i) { jr[c] = nl; jwl.push_back(c); wl.push_back(mx); nl++; }
                else { jr[c] = kUp + nu; jwu.push_back(c); wu.push_back(mx); nu++; }
            }
        }
        jr[i] = -1;
        for (int q = 0; q < nl; q++) jr[jwl[q]] = -1;
        for (int q = 0; q < nu; q++) jr[jwu[q]] = -1;
        const double d = repaired_pivot_signed(wd);
        const int keepl = std::min(nl, p), keepu = std::min(nu, p);
        select_largest(wl.data(), jwl.data(), nl, keepl);
        select_largest(wu.data(), jwu.data(), nu, keepu);
        const int len = keepl + 1 + keepu;
        int *c = S.ia.take(len);
        double *v = S.da.take(len);
        std::copy(jwl.begin(), jwl.begin() + keepl, c);
        std::copy(wl.begin(), wl.begin() + keepl, v);
        c[keepl] = i;
        v[keepl] = d;
        std::copy(jwu.begin(), jwu.begin() + keepu, c + keepl + 1);
        std::copy(wu.begin(), wu.begin() + keepu, v + keepl + 1);
        rcol[i] = c; rval[i] = v; rlen[i] = len;
        diag[i] = d;
    });
    // L (entries left of the diagonal in stored order, unit diagonal LAST) and U (diagonal FIRST, then the rest)
    Factors F;
    F.n = n;
    Csr &L = F.L, &U = F.U;
    L.n = U.n = n;
    L.p.resize((size_t)n + 1);
    U.p.resize((size_t)n + 1);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            int cl = 0, cu = 0;
            for (int q = 0; q < rlen[i]; q++) {
                cl += (rcol[i][q] <= i);
                cu += (rcol[i][q] >= i);
            }
            L.p[i] = cl;
            U.p[i] = cu;
        }
    });
    const long long tl = parallel_exclusive_scan(L.p.data(), n), tu = parallel_exclusive_scan(U.p.data(), n);
    L.p[n] = (int)tl;
    U.p[n] = (int)tu;
    L.j.resize((size_t)tl); L.x.resize((size_t)tl);
    U.j.resize((size_t)tu); U.x.resize((size_t)tu);
    parallel_ranges(n, [&](long long r0, long long r1, int) {
        for (int i = (int)r0; i < (int)r1; i++) {
            int ol = L.p[i], ou = U.p[i];
            for (int q = 0; q < rlen[i]; q++) {
                const int c = rcol[i][q];
                const double v = rval[i][q];
                if (c < i) { L.j[ol] = c; L.x[ol] = v; ol++; }
                else if (c == i) {
                    L.j[ol] = i; L.x[ol] = 1; ol++;
                    U.j[ou] = i; U.x[ou] = v; ou++;
                }
                else { U.j[ou] = c; U.x[ou] = v; ou++; }
            }
        }
    });
    return F;
}

Factors factor_ilut(Csr &&A, double tau, int p, int bs)
{
    PROF_T0;
    const int n = A.n;
    Csr B = block_diagonal(std::move(A), bs);
    {
        const int np = lsspg::host_threads();
        std::vector<char> loose(np, 0);
        parallel_ranges(n, [&](long long r0, long long r1, int q) {
            for (int i = (int)r0; i < (int)r1; i++)
                for (int k = B.p[i] + 1; k < B.p[i + 1]; k++)
                    if (B.j[k - 1] >= B.j[k]) { loose[q] = 1; return; }
        }, np);
        if (std::find(loose.begin(), loose.end(), (char)1) == loose.end()) {
            Factors F = factor_ilut_rows(B, tau, p, bs);
            PROF("ilut rows (pipelined)");
            return F;
        }
    }
    Factors F;
    F.n = n;
    F.L.n = F.U.n = n;
    F.L.p.push_back(0);
    F.U.p.push_back(0);
    for (int r0 = 0; r0 < n; r0 += bs) ilut_block(B, r0, std::min(n, r0 + bs), tau, p, F.L, F.U);
    PROF("ilut rows (serial)");
    return F;
}

// ---- block ILU(k) (reference src/pc-biluk.cxx:62-431, compiled there only with BLAS + LAPACK) ---------------
// Dense bs x bs blocks, column-major.  The three dense kernels restate the PUBLISHED netlib reference algorithms the
// reference reaches through dgemm_/dgetrf_/dgetri_ (src/pc-biluk.cxx:10-16): reference-BLAS DGEMM 'N','N', LAPACK's
// unblocked DGETF2 (partial pivoting) and DGETRI's unblocked path (DTRTI2, then the column sweep).  With an
// optimised BLAS the reference's own factors differ from these in the last bits -- the order of the dense sums is
// the BLAS's, not LSSP's.

// C = alpha A B + beta C  (src/pc-biluk.cxx:86-103, n == 1 handled as there)
void block_gemm(const double *A, const double *B, double alpha, double *C, double beta, int n)
{
    if (n == 1) {
        C[0] = A[0] * B[0] * alpha + beta * C[0];
        return;
    }
    for (int j = 0; j < n; j++) {
        if (beta == 0.0) for (int i = 0; i < n; i++) C[i + j * n] = 0.0;
        else if (beta != 1.0) for (int i = 0; i < n; i++) C[i + j * n] = beta * C[i + j * n];
        for (int l = 0; l < n; l++) {
            const double temp = alpha * B[l + j * n];
            for (int i = 0; i < n; i++) C[i + j * n] = C[i + j * n] + temp * A[i + l * n];
        }
    }
}

// A <- A^-1 in place (src/pc-biluk.cxx:62-84); != 0: singular
int block_inverse(double *A, int n, double *work, int *ipiv)
{
    if (n == 1) {
        if (A[0] == 0.0) return 1;
        A[0] = 1.0 / A[0];
        return 0;
    }
    int info = 0;
    for (int j = 0; j < n; j++) {   // DGETF2
        int jp = j;
        double big = fabs(A[j + j * n]);
        for (int i = j + 1; i < n; i++)
            if (fabs(A[i + j * n]) > big) { big = fabs(A[i + j * n]); jp = i; }
        ipiv[j] = jp;
        if (A[jp + j * n] != 0.0) {
            if (jp != j)
                for (int c = 0; c < n; c++) std::swap(A[j + c * n], A[jp + c * n]);
            if (j < n - 1) {
                if (fabs(A[j + j * n]) >= std::numeric_limits<double>::min()) {
                    const double r = 1.0 / A[j + j * n];
                    for (int i = j + 1; i < n; i++) A[i + j * n] = r * A[i + j * n];
                }
                else for (int i = j + 1; i < n; i++) A[i + j * n] = A[i + j * n] / A[j + j * n];
            }
        }
        else if (info == 0) info = j + 1;
        if (j < n - 1)
            for (int c = j + 1; c < n; c++) {
                if (A[j + c * n] == 0.0) continue;
                const double temp = -A[j + c * n];
                for (int i = j + 1; i < n; i++) A[i + c * n] = A[i + c * n] + A[i + j * n] * temp;
            }
    }
    if (info) return info;
    for (int j = 0; j < n; j++)
        if (A[j + j * n] == 0.0) return j + 1;
    for (int j = 0; j < n; j++) {   // DTRTI2: inverse of the upper triangle
        A[j + j * n] = 1.0 / A[j + j * n];
        const double ajj = -A[j + j * n];
        for (int c = 0; c < j; c++) {
            if (A[c + j * n] == 0.0) continue;
            const double temp = A[c + j * n];
            for (int i = 0; i < c; i++) A[i + j * n] = A[i + j * n] + temp * A[i + c * n];
            A[c + j * n] = A[c + j * n] * A[c + c * n];
        }
        for (int i = 0; i < j; i++) A[i + j * n] = ajj * A[i + j * n];
    }
    for (int j = n - 1; j >= 0; j--) {   // inv(A) L = inv(U)
        for (int i = j + 1; i < n; i++) { work[i] = A[i + j * n]; A[i + j * n] = 0.0; }
        for (int c = j + 1; c < n; c++) {
            const double temp = -work[c];
            for (int i = 0; i < n; i++) A[i + j * n] = A[i + j * n] + temp * A[i + c * n];
        }
    }
    for (int j = n - 2; j >= 0; j--)
        if (ipiv[j] != j)
            for (int i = 0; i < n; i++) std::swap(A[i + j * n], A[i + ipiv[j] * n]);
    return 0;
}

struct BFactors {
    int n = 0;
    Csr L, D, U;
};

// A: rows sorted by column.  bs x bs blocks; L = strict lower blocks + unit diagonal (last), D = inverted pivot
// blocks, U = inv(pivot) * upper blocks behind a unit diagonal (first): x = U^-1 D L^-1 rhs (src/pc-biluk.cxx:22-60).
int factor_biluk(const Csr &A, int bs, int level, BFactors &F)
{
    const int n = A.n, nb = n / bs, bs2 = bs * bs;
    // CSR -> BCSR (src/matrix-utils.cxx:62-162): block columns of a block row ascending, blocks column-major.
    // Block rows are independent: counted, scanned and filled by the host threads.
    Csr B;   // block graph; x unused
    B.n = nb;
    B.p.resize((size_t)nb + 1);
    auto block_cols = [&](int i, std::vector<int> &cols) {
        cols.clear();
        for (int r = i * bs; r < (i + 1) * bs; r++)
            for (int k = A.p[r]; k < A.p[r + 1]; k++) cols.push_back(A.j[k] / bs);
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    };
    const int np = lsspg::host_threads();
    std::vector<char> nodiag(np, 0);
    parallel_ranges(nb, [&](long long b0, long long b1, int piece) {
        std::vector<int> cols;
        for (int i = (int)b0; i < (int)b1; i++) {
            block_cols(i, cols);
            B.p[i] = (int)cols.size();
            if (!std::binary_search(cols.begin(), cols.end(), i)) nodiag[piece] = 1;
        }
    }, np, 1);
    if (std::find(nodiag.begin(), nodiag.end(), (char)1) != nodiag.end()) {
        for (int i = 0; i < nb; i++) {
            std::vector<int> cols;
            block_cols(i, cols);
            if (!std::binary_search(cols.begin(), cols.end(), i)) {
                lsspg::set_error("biluk: block row %d has no diagonal block", i);
                return 1;
            }
        }
    }
    const long long nblk = parallel_exclusive_scan(B.p.data(), nb);
    B.p[nb] = (int)nblk;
    B.j.resize((size_t)nblk);
    parallel_ranges(nb, [&](long long b0, long long b1, int) {
        std::vector<int> cols;
        for (int i = (int)b0; i < (int)b1; i++) {
            block_cols(i, cols);
            std::copy(cols.begin(), cols.end(), B.j.begin() + B.p[i]);
        }
    }, 0, 1024);
    // symbolic phase on the block graph (src/pc-biluk.cxx:316-375): level 0 keeps the pattern
    Csr T;
    if (level > 0) {
        B.x.resize(B.j.size());
        T = iluk_pattern(B, level);
    }
    else {
        T.n = nb;
        T.p.swap(B.p);
        T.j.swap(B.j);
    }
    const int *Tp = T.p.data(), *Tj = T.j.data();
    DVec X((size_t)Tp[nb] * bs2);
    parallel_ranges(nb, [&](long long b0, long long b1, int) {
        for (int i = (int)b0; i < (int)b1; i++) {
            std::fill(X.begin() + (size_t)Tp[i] * bs2, X.begin() + (size_t)Tp[i + 1] * bs2, 0.0);
            for (int r = i * bs; r < (i + 1) * bs; r++)
                for (int k = A.p[r]; k < A.p[r + 1]; k++) {
                    const int c = A.j[k];
                    const int at = (int)(std::lower_bound(Tj + Tp[i], Tj + Tp[i + 1], c / bs) - Tj);
                    X[(size_t)at * bs2 + (size_t)(c % bs) * bs + (r % bs)] = A.x[k];
                }
        }
    }, 0, 1024);
    // numeric phase (src/pc-biluk.cxx:198-277): block IKJ; inv[i] = inverse of the pivot block.  Block row i needs the
    // finished block rows of its lower block columns: rows grouped by the level of that dependency graph, the rows of a
    // level run by the host threads (as ilu0_levels); every block operation is the serial loop's.
    DVec inv((size_t)nb * bs2);
    IVec lev((size_t)nb), order((size_t)nb);
    std::vector<int> start(1, 0);
    for (int i = 0; i < nb; i++) {
        int l = 0;
        for (int k = Tp[i]; k < Tp[i + 1] && Tj[k] < i; k++) l = std::max(l, lev[Tj[k]] + 1);
        lev[i] = l;
        if ((int)start.size() < l + 2) start.resize(l + 2, 0);
        start[l + 1]++;
    }
    const int nlev = (int)start.size() - 1;
    for (int l = 0; l < nlev; l++) start[l + 1] += start[l];
    {
        std::vector<int> pos(start.begin(), start.end() - 1);
        for (int i = 0; i < nb; i++) order[pos[lev[i]]++] = i;
    }
    std::atomic<int> singular(0);
    lsspg::LevelTeam::run(nlev, start.data(), order.data(), [&](int i) {
        thread_local std::vector<double> blk, work;
        thread_local std::vector<int> ipiv;
        blk.resize(bs2); work.resize(10 * (size_t)bs); ipiv.resize(bs);
        const int e = Tp[i + 1];
        int k = Tp[i];
        for (; k < e && Tj[k] < i; k++) {
            const int pr = Tj[k];
            double *a_ik = &X[(size_t)k * bs2];
            std::copy(a_ik, a_ik + bs2, blk.begin());
            block_gemm(blk.data(), &inv[(size_t)pr * bs2], 1., a_ik, 0., bs);
            int pq = Tp[pr];
            const int pe = Tp[pr + 1];
            for (int j = k + 1; j < e; j++) {
                const int c = Tj[j];
                while (pq < pe && Tj[pq] < c) pq++;
                if (pq < pe && Tj[pq] == c) block_gemm(a_ik, &X[(size_t)pq * bs2], -1., &X[(size_t)j * bs2], 1., bs);
            }
        }
        std::copy(&X[(size_t)k * bs2], &X[(size_t)k * bs2] + bs2, &inv[(size_t)i * bs2]);   // k: the diagonal block
        if (block_inverse(&inv[(size_t)i * bs2], bs, work.data(), ipiv.data())) singular.store(1);
    });
    if (singular.load()) {
        lsspg::set_error("lssp: bilu(0) singular diagonal submatrix.");   // src/pc-biluk.cxx:262
        return 1;
    }
    // L, U, D as CSR (src/pc-biluk.cxx:105-196, :279-314): rows end up sorted by column; block rows written by the threads
    F.n = n;
    Csr &L = F.L, &U = F.U, &D = F.D;
    L.n = U.n = D.n = n;
    L.p.resize((size_t)n + 1);
    U.p.resize((size_t)n + 1);
    D.p.resize((size_t)n + 1);
    parallel_ranges(nb, [&](long long b0, long long b1, int) {
        for (int i = (int)b0; i < (int)b1; i++) {
            int nl = 0, nu = 0;
            for (int k = Tp[i]; k < Tp[i + 1]; k++) {
                nl += (Tj[k] < i);
                nu += (Tj[k] > i);
            }
            for (int a = 0; a < bs; a++) {
                L.p[i * bs + a] = nl * bs + 1;
                U.p[i * bs + a] = nu * bs + 1;
                D.p[i * bs + a] = bs;
            }
        }
    }, 0, 1024);
    const long long tl = parallel_exclusive_scan(L.p.data(), n), tu = parallel_exclusive_scan(U.p.data(), n),
                    td = parallel_exclusive_scan(D.p.data(), n);
    if (tl > 0x7fffffffll || tu > 0x7fffffffll) {
        lsspg::set_error("biluk: factors exceed int32 indexing");
        return 1;
    }
    L.p[n] = (int)tl; U.p[n] = (int)tu; D.p[n] = (int)td;
    L.j.resize((size_t)tl); L.x.resize((size_t)tl);
    U.j.resize((size_t)tu); U.x.resize((size_t)tu);
    D.j.resize((size_t)td); D.x.resize((size_t)td);
    parallel_ranges(nb, [&](long long b0, long long b1, int) {
        std::vector<double> cache(bs2, 0.0);
        std::vector<int> ol(bs), ou(bs);
        for (int i = (int)b0; i < (int)b1; i++) {
            for (int a = 0; a < bs; a++) {
                const int r = i * bs + a;
                ol[a] = L.p[r];
                ou[a] = U.p[r];
                U.j[ou[a]] = r; U.x[ou[a]] = 1.; ou[a]++;
                for (int b = 0; b < bs; b++) {
                    D.j[D.p[r] + b] = i * bs + b;
                    D.x[D.p[r] + b] = inv[(size_t)i * bs2 + (size_t)b * bs + a];
                }
            }
            for (int k = Tp[i]; k < Tp[i + 1]; k++) {
                const int c = Tj[k];
                const double *d = &X[(size_t)k * bs2];
                if (c < i) {
                    for (int a = 0; a < bs; a++)
                        for (int b = 0; b < bs; b++) { L.j[ol[a]] = c * bs + b; L.x[ol[a]] = d[b * bs + a]; ol[a]++; }
                }
                else if (c > i) {
                    block_gemm(&inv[(size_t)i * bs2], d, 1., cache.data(), 0., bs);
                    for (int a = 0; a < bs; a++)
                        for (int b = 0; b < bs; b++) { U.j[ou[a]] = c * bs + b; U.x[ou[a]] = cache[b * bs + a]; ou[a]++; }
                }
            }
            for (int a = 0; a < bs; a++) { L.j[ol[a]] = i * bs + a; L.x[ol[a]] = 1.; }
        }
    }, 0, 1024);
    return 0;
}

}  // namespace

struct lsspg_factors {
    Factors f;
};

struct lsspg_bfactors {
    BFactors f;
};

namespace lsspg {

// an empty factor pair of the given sizes for the device set-up (ilu_gpu.cu), which downloads into the arrays
lsspg_factors *factors_new(int n, size_t nnzL, size_t nnzU, int **Lp, int **Lj, double **Lx, int **Up, int **Uj, double **Ux)
{
    lsspg_factors *F = new lsspg_factors();
    F->f.n = n;
    Csr &L = F->f.L, &U = F->f.U;
    L.n = U.n = n;
    L.p.resize((size_t)n + 1); L.j.resize(nnzL); L.x.resize(nnzL);
    U.p.resize((size_t)n + 1); U.j.resize(nnzU); U.x.resize(nnzU);
    *Lp = L.p.data(); *Lj = L.j.data(); *Lx = L.x.data();
    *Up = U.p.data(); *Uj = U.j.data(); *Ux = U.x.data();
    return F;
}

}  // namespace lsspg

extern "C" {

int lsspg_ilu_factor(int kind, int n, const int *hAp, const int *hAj, const double *hAx, int level, int p,
                     double tol, int blk_size, lsspg_factors **out)
{
    if (!out || !hAp || !hAj || !hAx || n <= 0) {
        lsspg::set_error("lsspg_ilu_factor: bad argument");
        return 1;
    }
    if (kind != LSSPG_ILUK && kind != LSSPG_ILUT) {
        lsspg::set_error("lsspg_ilu_factor: unknown kind %d", kind);
        return 1;
    }
    PROF_T0;
    const int nnz = hAp[n];
    Csr Ad = ingest(n, hAp, hAj, hAx);
    PROF("ingest");
    const int bs = (blk_size <= 0 || blk_size > n) ? n : blk_size;
    lsspg_factors *F = new lsspg_factors();
    if (kind == LSSPG_ILUK) {
        if (level < 0) level = 0;               // reference src/pc-iluk.cxx:286-290
        F->f = factor_iluk(std::move(Ad), level, bs);
    }
    else {
        if (p <= 0) p = (nnz + n - 1) / n;      // reference src/pc-ilut.cxx:436-438
        if (tol < 0) tol = 1e-3;                // reference src/pc-ilut.cxx:440-442, src/pc.cxx:4
        F->f = factor_ilut(std::move(Ad), tol, p, bs);
    }
    *out = F;
    return 0;
}

int lsspg_factors_sizes(const lsspg_factors *F, int *n, int *nnzL, int *nnzU)
{
    if (n) *n = F->f.n;
    if (nnzL) *nnzL = F->f.L.nnz();
    if (nnzU) *nnzU = F->f.U.nnz();
    return 0;
}

int lsspg_factors_get(const lsspg_factors *F, int *Lp, int *Lj, double *Lx, int *Up, int *Uj, double *Ux)
{
    const Csr &L = F->f.L, &U = F->f.U;
    parallel_copy(Lp, L.p.data(), sizeof(int) * L.p.size());
    parallel_copy(Lj, L.j.data(), sizeof(int) * L.j.size());
    parallel_copy(Lx, L.x.data(), sizeof(double) * L.x.size());
    parallel_copy(Up, U.p.data(), sizeof(int) * U.p.size());
    parallel_copy(Uj, U.j.data(), sizeof(int) * U.j.size());
    parallel_copy(Ux, U.x.data(), sizeof(double) * U.x.size());
    return 0;
}

int lsspg_factors_destroy(lsspg_factors *F)
{
    delete F;
    return 0;
}

// Block ILU(k) set-up (reference lssp_pc_biluk_assemble, src/pc-biluk.cxx:416-431): blocks of n / num_blks rows.
int lsspg_bilu_factor(int n, const int *hAp, const int *hAj, const double *hAx, int num_blks, int level, lsspg_bfactors **out)
{
    if (!out || !hAp || !hAj || !hAx || n <= 0 || num_blks <= 0) {
        lsspg::set_error("lsspg_bilu_factor: bad argument");
        return 1;
    }
    if (n % num_blks != 0) {
        lsspg::set_error("lsspg_bilu_factor: num_rows %d is not a multiple of the number of blocks %d", n, num_blks);
        return 1;
    }
    Csr A;
    A.n = n;
    A.p.resize((size_t)n + 1);
    A.j.resize((size_t)hAp[n]);
    A.x.resize((size_t)hAp[n]);
    parallel_copy(A.p.data(), hAp, sizeof(int) * ((size_t)n + 1));
    parallel_copy(A.j.data(), hAj, sizeof(int) * (size_t)hAp[n]);
    parallel_copy(A.x.data(), hAx, sizeof(double) * (size_t)hAp[n]);
    sort_rows(A);   // src/lssp.cxx:173
    if (level < 0) level = 0;
    lsspg_bfactors *F = new lsspg_bfactors();
    if (factor_biluk(A, n / num_blks, level, F->f)) {
        delete F;
        return 1;
    }
    *out = F;
    return 0;
}

int lsspg_bfactors_sizes(const lsspg_bfactors *F, int *n, int *nnzL, int *nnzD, int *nnzU)
{
    if (n) *n = F->f.n;
    if (nnzL) *nnzL = F->f.L.nnz();
    if (nnzD) *nnzD = F->f.D.nnz();
    if (nnzU) *nnzU = F->f.U.nnz();
    return 0;
}

int lsspg_bfactors_get(const lsspg_bfactors *F, int *Lp, int *Lj, double *Lx, int *Dp, int *Dj, double *Dx, int *Up,
                       int *Uj, double *Ux)
{
    const Csr *M[3] = {&F->f.L, &F->f.D, &F->f.U};
    int *P[3] = {Lp, Dp, Up}, *J[3] = {Lj, Dj, Uj};
    double *X[3] = {Lx, Dx, Ux};
    for (int q = 0; q < 3; q++) {
        parallel_copy(P[q], M[q]->p.data(), sizeof(int) * M[q]->p.size());
        parallel_copy(J[q], M[q]->j.data(), sizeof(int) * M[q]->j.size());
        parallel_copy(X[q], M[q]->x.data(), sizeof(double) * M[q]->x.size());
    }
    return 0;
}

int lsspg_bfactors_destroy(lsspg_bfactors *F)
{
    delete F;
    return 0;
}

// host threads the set-up uses (LSSPG_HOST_THREADS; default: the cores of the affinity mask / LOCAL_WORLD_SIZE, <= 32)
int lsspg_host_threads(void) { return lsspg::host_threads(); }

}  // extern "C"

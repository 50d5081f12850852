// TEST HARNESS (tests/test_facade.py builds and runs it; CPU only).
//
// One binary caller, two libraries: the matrix utilities of the LSSP C++ API (include/matrix-utils.h) are fetched by
// their MANGLED names from the compiled reference (oracle/_ref/liblssp_ref.so) and from this repository's facade
// (lssp_b200/liblssp.so), each opened RTLD_LOCAL, called with the same by-value structs on the same matrices, and
// the results compared bit for bit.  That is the drop-in claim of SURVEY.md 8b for these entry points.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct csr { int num_rows, num_cols, num_nnzs; int *Ap, *Aj; double *Ax; };      // include/type-defs.h:15-24
struct coo { int num_rows, num_cols, num_nnzs; int *Ai, *Aj; double *Ax; };      // :26-35
struct bcsr { int num_rows, num_cols, num_nnzs, blk_size; int *Ap, *Aj; double *Ax; };   // :45-55

struct Api {
    bcsr (*csr_to_bcsr)(const csr, int);
    csr (*bcsr_to_csr)(const bcsr);
    coo (*csr_to_coo)(const csr);
    csr (*coo_to_csr)(const coo);
    bool (*csr_is_sorted)(const csr);
    bool (*bcsr_is_sorted)(const bcsr);
    void (*sort_column)(csr &);
    csr (*adjust_zero_diag)(const csr, double);
    csr (*get_block_diag)(const csr, int);
    csr (*transpose)(const csr);
};

static void *sym(void *h, const char *name)
{
    void *p = dlsym(h, name);
    if (!p) { fprintf(stderr, "missing symbol %s\n", name); exit(2); }
    return p;
}

static Api load(const char *path)
{
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); exit(2); }
    Api a;
    *(void **)&a.csr_to_bcsr = sym(h, "_Z20lssp_mat_csr_to_bcsr13lssp_mat_csr_i");
    *(void **)&a.bcsr_to_csr = sym(h, "_Z20lssp_mat_bcsr_to_csr14lssp_mat_bcsr_");
    *(void **)&a.csr_to_coo = sym(h, "_Z19lssp_mat_csr_to_coo13lssp_mat_csr_");
    *(void **)&a.coo_to_csr = sym(h, "_Z19lssp_mat_coo_to_csr13lssp_mat_coo_");
    *(void **)&a.csr_is_sorted = sym(h, "_Z22lssp_mat_csr_is_sorted13lssp_mat_csr_");
    *(void **)&a.bcsr_is_sorted = sym(h, "_Z23lssp_mat_bcsr_is_sorted14lssp_mat_bcsr_");
    *(void **)&a.sort_column = sym(h, "_Z20lssp_mat_sort_columnR13lssp_mat_csr_");
    *(void **)&a.adjust_zero_diag = sym(h, "_Z25lssp_mat_adjust_zero_diag13lssp_mat_csr_d");
    *(void **)&a.get_block_diag = sym(h, "_Z23lssp_mat_get_block_diag13lssp_mat_csr_i");
    *(void **)&a.transpose = sym(h, "_Z18lssp_mat_transpose13lssp_mat_csr_");
    return a;
}

static unsigned long long rng_state = 88172645463325252ull;
static unsigned int rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (unsigned int)(rng_state >> 11); }

// random square matrix: n rows, ~avg entries per row, columns unique per row, optionally unsorted / without some diagonals
static csr make(int n, int avg, bool sorted, bool all_diag, std::vector<int> &p, std::vector<int> &j, std::vector<double> &x)
{
    p.assign(1, 0); j.clear(); x.clear();
    std::vector<char> used(n, 0);
    for (int i = 0; i < n; i++) {
        std::vector<int> cols;
        if (all_diag || rnd() % 4) cols.push_back(i);
        const int want = 1 + rnd() % (2 * avg);
        for (int t = 0; t < want; t++) cols.push_back((int)(rnd() % n));
        std::vector<int> uniq;
        for (int c : cols) if (!used[c]) { used[c] = 1; uniq.push_back(c); }
        for (int c : uniq) used[c] = 0;
        if (sorted) for (size_t a = 0; a < uniq.size(); a++) for (size_t b = a + 1; b < uniq.size(); b++) if (uniq[b] < uniq[a]) { int t = uniq[a]; uniq[a] = uniq[b]; uniq[b] = t; }
        for (int c : uniq) { j.push_back(c); x.push_back((rnd() % 5 == 0) ? 0.0 : (double)(rnd() % 2001 - 1000) / 64.0); }
        p.push_back((int)j.size());
    }
    csr A = {n, n, (int)j.size(), p.data(), j.data(), x.data()};
    return A;
}

static int fails = 0;
static void check(bool ok, const char *what, int n, int arg)
{
    if (!ok) { fails++; printf("MISMATCH %s (n = %d, arg = %d)\n", what, n, arg); }
}
static bool same(const csr &a, const csr &b)
{
    if (a.num_rows != b.num_rows || a.num_cols != b.num_cols || a.num_nnzs != b.num_nnzs) return false;
    if (a.num_nnzs <= 0) return true;
    return !memcmp(a.Ap, b.Ap, sizeof(int) * (a.num_rows + 1)) && !memcmp(a.Aj, b.Aj, sizeof(int) * a.num_nnzs) &&
           !memcmp(a.Ax, b.Ax, sizeof(double) * a.num_nnzs);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s <reference .so> <facade .so>\n", argv[0]); return 2; }
    const Api R = load(argv[1]), F = load(argv[2]);
    std::vector<int> p, j;
    std::vector<double> x;
    int cases = 0;
    for (int n : {12, 60, 240, 1001}) {
        for (int variant = 0; variant < 4; variant++) {
            const bool sorted = variant & 1, all_diag = variant & 2;
            csr A = make(n, 4, sorted, all_diag, p, j, x);
            cases++;
            check(R.csr_is_sorted(A) == F.csr_is_sorted(A), "csr_is_sorted", n, variant);
            // COO round trip
            coo c1 = R.csr_to_coo(A), c2 = F.csr_to_coo(A);
            check(c1.num_nnzs == c2.num_nnzs && !memcmp(c1.Ai, c2.Ai, sizeof(int) * c1.num_nnzs) && !memcmp(c1.Aj, c2.Aj, sizeof(int) * c1.num_nnzs) &&
                  !memcmp(c1.Ax, c2.Ax, sizeof(double) * c1.num_nnzs), "csr_to_coo", n, variant);
            // shuffle the COO entries (same permutation for both), then back to CSR: row order of entries is kept
            for (int k = c1.num_nnzs - 1; k > 0; k--) {
                const int q = (int)(rnd() % (k + 1));
                std::swap(c1.Ai[k], c1.Ai[q]); std::swap(c1.Aj[k], c1.Aj[q]); std::swap(c1.Ax[k], c1.Ax[q]);
            }
            csr b1 = R.coo_to_csr(c1), b2 = F.coo_to_csr(c1);
            check(same(b1, b2), "coo_to_csr", n, variant);
            // transpose
            check(same(R.transpose(A), F.transpose(A)), "transpose", n, variant);
            // block diagonal, several block sizes incl. one that does not divide n and blk_size == n
            for (int bs : {1, 5, n / 3 + 1, n})
                check(same(R.get_block_diag(A, bs), F.get_block_diag(A, bs)), "get_block_diag", n, bs);
            // missing diagonals: compare everything but num_nnzs, which the reference leaves stale (src/matrix-utils.cxx:485)
            {
                csr m1 = R.adjust_zero_diag(A, 1e-10), m2 = F.adjust_zero_diag(A, 1e-10);
                const int nz = m2.Ap[n];
                check(m1.Ap[n] == nz && m2.num_nnzs == nz && !memcmp(m1.Ap, m2.Ap, sizeof(int) * (n + 1)) && !memcmp(m1.Aj, m2.Aj, sizeof(int) * nz) &&
                      !memcmp(m1.Ax, m2.Ax, sizeof(double) * nz), "adjust_zero_diag", n, variant);
            }
            // sort_column on copies (unique columns per row: the result is unique)
            {
                std::vector<int> j1(j), j2(j);
                std::vector<double> x1(x), x2(x);
                csr s1 = A, s2 = A;
                s1.Aj = j1.data(); s1.Ax = x1.data(); s2.Aj = j2.data(); s2.Ax = x2.data();
                R.sort_column(s1); F.sort_column(s2);
                check(j1 == j2 && !memcmp(x1.data(), x2.data(), sizeof(double) * x1.size()), "sort_column", n, variant);
            }
            // BCSR and back
            for (int bs : {1, 2, 3, 4, 6})
                if (n % bs == 0) {
                    bcsr g1 = R.csr_to_bcsr(A, bs), g2 = F.csr_to_bcsr(A, bs);
                    const bool eq = g1.num_rows == g2.num_rows && g1.num_nnzs == g2.num_nnzs && g1.blk_size == g2.blk_size &&
                                    !memcmp(g1.Ap, g2.Ap, sizeof(int) * (g1.num_rows + 1)) && !memcmp(g1.Aj, g2.Aj, sizeof(int) * g1.num_nnzs) &&
                                    !memcmp(g1.Ax, g2.Ax, sizeof(double) * g1.num_nnzs * bs * bs);
                    check(eq, "csr_to_bcsr", n, bs);
                    check(R.bcsr_is_sorted(g1) == F.bcsr_is_sorted(g1), "bcsr_is_sorted", n, bs);
                    if (eq) check(same(R.bcsr_to_csr(g1), F.bcsr_to_csr(g1)), "bcsr_to_csr", n, bs);
                }
        }
    }
    printf("%d matrices, %d mismatches\n", cases, fails);
    return fails ? 1 : 0;
}

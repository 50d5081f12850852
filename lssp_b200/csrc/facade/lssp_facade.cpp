// lssp_facade.cpp -- the LSSP C++ API (lssp.h / mvops.h / vector.h / pc*.h / solver-*.h) on top of
// the C ABI of liblsspg.so.  Written fresh for the B200 build; signatures, defaults, ownership and
// error behaviour follow the reference (cited per function), so that example/exam.cxx-style
// programs recompile and relink unchanged.  No arithmetic happens here: vectors and matrices are
// host objects as in the reference, every operation is forwarded to the sm_100a kernels.
#include <strings.h>
#include <sys/resource.h>
#include <sys/time.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../host_par.h"
#include "lssp.h"
#include "lsspg.h"

// ---- device handles without touching the reference's struct layout ------------------------------------------------
// LSSP_PC / LSSP_SOLVER are binary-identical to the reference's (include/type-defs.h:107-151, 225-304; checked by
// tests/cxx/struct_layout_check.cpp).  A built-in preconditioner keeps its device object in pc.data, which the reference
// sets to NULL in lssp_pc_create (src/pc.cxx:46) and never uses for its own preconditioners; a USER preconditioner owns
// pc.data, so its trampoline object lives in a side table keyed by the address of the LSSP_PC.  The solver's device matrix
// is keyed by the host deep copy's row-pointer array (s.A.Ap), which survives the by-value copies of LSSP_SOLVER that
// lssp_pc_assemble(LSSP_PC &, LSSP_SOLVER) makes.
#include <unordered_map>
static std::unordered_map<const void *, void *> &side_table()
{
    static std::unordered_map<const void *, void *> t;
    return t;
}
static void *side_get(const void *key)
{
    auto it = side_table().find(key);
    return it == side_table().end() ? NULL : it->second;
}
static void side_set(const void *key, void *v)
{
    if (v) side_table()[key] = v;
    else side_table().erase(key);
}
static inline void *pc_dev(const LSSP_PC *pc) { return pc->type == LSSP_PC_USER ? side_get(pc) : pc->data; }
static inline void pc_dev_set(LSSP_PC *pc, void *d)
{
    if (pc->type == LSSP_PC_USER) side_set(pc, d);
    else pc->data = d;
}
static inline void *solver_dev(const LSSP_SOLVER &s) { return s.A.Ap ? side_get(s.A.Ap) : NULL; }

// ---- globals: defaults of the reference (src/lssp.cxx:5-14, src/pc.cxx:3-7, src/utils.cxx:19-22) ----
int LSSP_RESTART = 50;
int LSSP_AUG_K = 3;
int LSSP_BGSL = 4;
int LSSP_IDRS = 4;
int LSSP_MAXIT = 1000;
double LSSP_ATOL = 1e-7;
double LSSP_RTOL = 1e-7;
double LSSP_RB = 1e-7;
double LSSP_BREAKDOWN = 1e-40;
int lssp_pc_iluk_level_default = 1;
double lssp_pc_ilut_tol = 1e-3;
double lssp_pc_ilut_p = -1;
int lssp_verbosity = 2;
static FILE *lssp_log_handle = NULL;

// ---- utils (reference src/utils.cxx) -----------------------------------------------------------
void lssp_set_log(FILE *io) { lssp_log_handle = io; }
int lssp_comp_int_asc(const void *p, const void *n) { return *(const int *)p - *(const int *)n; }
int lssp_comp_int_des(const void *p, const void *n) { return *(const int *)n - *(const int *)p; }

double lssp_get_time()
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

double lssp_get_mem_usage(double *peak)
{
    struct rusage ru;
    getrusage(RUSAGE_SELF, &ru);
    const double mb = ru.ru_maxrss / 1024.;
    if (peak) *peak = mb;
    return mb;
}

static int vprint(const char *prefix, const char *fmt, va_list ap)
{
    char buf[4096];
    vsnprintf(buf, sizeof(buf), fmt, ap);
    int r = fprintf(stdout, "%s%s", prefix, buf);
    fflush(stdout);
    if (lssp_log_handle) {
        fprintf(lssp_log_handle, "%s%s", prefix, buf);
        fflush(lssp_log_handle);
    }
    return r;
}

int lssp_printf(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    int r = vprint("", fmt, ap);
    va_end(ap);
    return r;
}

void lssp_error(int code, const char *fmt, ...)   // reference src/utils.cxx:114-135: print, then exit(code)
{
    va_list ap;
    va_start(ap, fmt);
    vprint("error: ", fmt, ap);
    va_end(ap);
    if (code != 0) exit(code);
}

void lssp_warning(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vprint("warning: ", fmt, ap);
    va_end(ap);
}

// ---- device context ------------------------------------------------------------------------------
static lsspg_ctx *g_ctx = NULL;

static void driver_print(const char *msg) { lssp_printf("%s", msg); }

static lsspg_ctx *ctx()
{
    if (!g_ctx) {
        const char *e = getenv("LSSP_GPU");
        if (lsspg_ctx_create(e ? atoi(e) : 0, &g_ctx)) lssp_error(1, "lssp: %s\n", lsspg_last_error());
        lsspg_set_printer(driver_print);   // per-iteration lines of the drivers go where lssp_printf sends them (log file too)
        // LSSP_REDUCE=reference: every dot product / norm is the reference's sequential sum (src/vector.cxx:127-131), bit
        // for bit, computed in parallel -> whole solves are bit-identical to the CPU library (exact_sum.cu); default: fixed
        // tree reductions (faster, more accurate, different in the last bits)
        if (const char *r = getenv("LSSP_REDUCE")) {
            const int mode = !strcmp(r, "reference") ? 2 : !strcmp(r, "sequential") ? 1 : 0;
            if (lsspg_ctx_set_option(g_ctx, LSSPG_OPT_REDUCE_SEQUENTIAL, mode)) lssp_error(1, "lssp: %s\n", lsspg_last_error());
        }
    }
    return g_ctx;
}

#define GPU(call)                                                   \
    do {                                                            \
        if (call) lssp_error(1, "lssp: %s\n", lsspg_last_error());  \
    } while (0)

// ---- matrix utils (reference src/matrix-utils.cxx:3-60, :249-279, :387-481) -----------------------
void lssp_mat_init(lssp_mat_csr &A) { bzero(&A, sizeof(A)); }
void lssp_mat_init(lssp_mat_coo &A) { bzero(&A, sizeof(A)); }
void lssp_mat_init(lssp_mat_bcsr &A) { bzero(&A, sizeof(A)); }
void lssp_mat_destroy(lssp_mat_csr &A) { lssp_free(A.Ap); lssp_free(A.Aj); lssp_free(A.Ax); lssp_mat_init(A); }
void lssp_mat_destroy(lssp_mat_coo &A) { lssp_free(A.Ai); lssp_free(A.Aj); lssp_free(A.Ax); lssp_mat_init(A); }
void lssp_mat_destroy(lssp_mat_bcsr &A) { lssp_free(A.Aj); lssp_free(A.Ap); lssp_free(A.Ax); lssp_mat_init(A); }

lssp_mat_csr lssp_mat_create(int nrows, int ncols, int *Ap, int *Aj, double *Ax)
{
    lssp_mat_csr A;
    assert(nrows > 0 && ncols > 0 && Ap != NULL);
    lssp_mat_init(A);
    A.num_rows = nrows;
    A.num_cols = ncols;
    A.num_nnzs = Ap[nrows];
    A.Ap = lssp_copy_on<int>(Ap, nrows + 1);
    A.Aj = lssp_copy_on<int>(Aj, Ap[nrows]);
    A.Ax = lssp_copy_on<double>(Ax, Ap[nrows]);
    return A;
}

bool lssp_mat_csr_is_sorted(const lssp_mat_csr A)
{
    for (int i = 0; i < A.num_rows; i++)
        for (int k = A.Ap[i] + 1; k < A.Ap[i + 1]; k++)
            if (A.Aj[k - 1] > A.Aj[k]) return false;
    return true;
}

void lssp_mat_sort_column(lssp_mat_csr &A)
{
    std::vector<int> idx, tj;
    std::vector<double> tx;
    for (int i = 0; i < A.num_rows; i++) {
        const int b = A.Ap[i], n = A.Ap[i + 1] - b;
        idx.resize(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int p, int q) { return A.Aj[b + p] < A.Aj[b + q]; });
        tj.assign(A.Aj + b, A.Aj + b + n);
        tx.assign(A.Ax + b, A.Ax + b + n);
        for (int k = 0; k < n; k++) { A.Aj[b + k] = tj[idx[k]]; A.Ax[b + k] = tx[idx[k]]; }
    }
}

bool lssp_mat_bcsr_is_sorted(const lssp_mat_bcsr A)   // src/matrix-utils.cxx:217-247
{
    assert(A.num_rows > 0 && A.num_cols > 0 && A.blk_size > 0);
    if (A.num_nnzs <= 0) return true;
    for (int i = 0; i < A.num_rows; i++)
        for (int k = A.Ap[i] + 1; k < A.Ap[i + 1]; k++)
            if (A.Aj[k - 1] > A.Aj[k]) return false;
    return true;
}

// ---- format converters and matrix utilities (reference src/matrix-utils.cxx:62-215, :281-380, :483-765): host-side
//      data formats either side of the solve loop, written fresh with the reference's results ------------------------
lssp_mat_coo lssp_mat_csr_to_coo(const lssp_mat_csr csr)   // :302-322
{
    lssp_mat_coo A;
    lssp_mat_init(A);
    A.num_rows = csr.num_rows;
    A.num_cols = csr.num_cols;
    A.num_nnzs = csr.num_nnzs;
    if (csr.num_nnzs <= 0) return A;
    A.Ai = lssp_malloc<int>(csr.num_nnzs);
    A.Aj = lssp_copy_on<int>(csr.Aj, csr.num_nnzs);
    A.Ax = lssp_copy_on<double>(csr.Ax, csr.num_nnzs);
    for (int i = 0; i < csr.num_rows; i++)
        for (int k = csr.Ap[i]; k < csr.Ap[i + 1]; k++) A.Ai[k] = i;
    return A;
}

lssp_mat_csr lssp_mat_coo_to_csr(const lssp_mat_coo A)   // :324-380: entries of a row keep their COO order
{
    lssp_mat_csr csr;
    lssp_mat_init(csr);
    csr.num_rows = A.num_rows;
    csr.num_cols = A.num_cols;
    csr.num_nnzs = A.num_nnzs;
    if (A.num_nnzs <= 0) return csr;
    csr.Ap = lssp_malloc<int>(csr.num_rows + 1);
    csr.Aj = lssp_malloc<int>(csr.num_nnzs);
    csr.Ax = lssp_malloc<double>(csr.num_nnzs);
    std::vector<int> pos(A.num_rows + 1, 0);
    for (int k = 0; k < A.num_nnzs; k++) pos[A.Ai[k] + 1]++;
    for (int i = 0; i < A.num_rows; i++) pos[i + 1] += pos[i];
    memcpy(csr.Ap, pos.data(), sizeof(int) * (A.num_rows + 1));
    for (int k = 0; k < A.num_nnzs; k++) {
        const int at = pos[A.Ai[k]]++;
        csr.Aj[at] = A.Aj[k];
        csr.Ax[at] = A.Ax[k];
    }
    return csr;
}

// :62-162.  Blocks column-major; the block columns of a block row ascending; entries absent from A are stored as 0.
lssp_mat_bcsr lssp_mat_csr_to_bcsr(const lssp_mat_csr A, int bs)
{
    assert(A.num_rows > 0 && A.num_rows == A.num_cols && A.num_nnzs > 0);
    assert(bs > 0);
    if (A.num_rows % bs != 0) lssp_error(1, "num_rows is not a multiple of block size: %d\n", bs);
    lssp_mat_bcsr B;
    lssp_mat_init(B);
    const int nb = A.num_rows / bs, bs2 = bs * bs;
    B.num_rows = B.num_cols = nb;
    B.blk_size = bs;
    B.Ap = lssp_malloc<int>(nb + 1);
    std::vector<int> mark(nb, -1), cols, all;
    B.Ap[0] = 0;
    for (int i = 0; i < nb; i++) {
        cols.clear();
        for (int r = i * bs; r < (i + 1) * bs; r++)
            for (int k = A.Ap[r]; k < A.Ap[r + 1]; k++) {
                const int c = A.Aj[k] / bs;
                if (mark[c] != i) { mark[c] = i; cols.push_back(c); }
            }
        std::sort(cols.begin(), cols.end());
        all.insert(all.end(), cols.begin(), cols.end());
        B.Ap[i + 1] = (int)all.size();
    }
    B.num_nnzs = B.Ap[nb];
    B.Aj = lssp_copy_on<int>(all.data(), B.num_nnzs);
    B.Ax = lssp_malloc<double>(B.num_nnzs * bs2);
    for (long long k = 0; k < (long long)B.num_nnzs * bs2; k++) B.Ax[k] = 0.;
    std::vector<int> where(nb, -1);
    for (int i = 0; i < nb; i++) {
        for (int k = B.Ap[i]; k < B.Ap[i + 1]; k++) where[B.Aj[k]] = k;
        for (int r = i * bs; r < (i + 1) * bs; r++)
            for (int k = A.Ap[r]; k < A.Ap[r + 1]; k++) {
                const int c = A.Aj[k];
                B.Ax[(size_t)where[c / bs] * bs2 + (size_t)(c % bs) * bs + (r % bs)] = A.Ax[k];
            }
    }
    return B;
}

// :164-215.  Stored zeros of the blocks are dropped; rows sorted by column.
lssp_mat_csr lssp_mat_bcsr_to_csr(const lssp_mat_bcsr A)
{
    const int bs = A.blk_size, bs2 = bs * bs, n = A.num_rows * bs;
    lssp_mat_csr B;
    lssp_mat_init(B);
    B.num_rows = n;
    B.num_cols = A.num_cols * bs;
    std::vector<int> cnt(n + 1, 0);
    for (int i = 0; i < A.num_rows; i++)
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++)
            for (int q = 0; q < bs2; q++)
                if (fabs(A.Ax[(size_t)k * bs2 + q]) > 0.) cnt[i * bs + q % bs + 1]++;
    for (int r = 0; r < n; r++) cnt[r + 1] += cnt[r];
    B.num_nnzs = cnt[n];
    if (B.num_nnzs <= 0) return B;   // as lssp_mat_coo_to_csr on an empty matrix (:369)
    B.Ap = lssp_copy_on<int>(cnt.data(), n + 1);
    B.Aj = lssp_malloc<int>(B.num_nnzs);
    B.Ax = lssp_malloc<double>(B.num_nnzs);
    for (int i = 0; i < A.num_rows; i++)
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++)
            for (int q = 0; q < bs2; q++) {
                const double v = A.Ax[(size_t)k * bs2 + q];
                if (fabs(v) > 0.) {
                    const int at = cnt[i * bs + q % bs]++;
                    B.Aj[at] = A.Aj[k] * bs + q / bs;
                    B.Ax[at] = v;
                }
            }
    lssp_mat_sort_column(B);
    return B;
}

// :483-587: a row without a stored diagonal receives (i, tol), slid into sorted position.  (The reference returns the
// result with the INPUT's num_nnzs, :485 -- its callers then under-allocate; here the count is the result's.)
lssp_mat_csr lssp_mat_adjust_zero_diag(const lssp_mat_csr A, double tol)
{
    const int n = A.num_rows;
    lssp_mat_csr M = A;
    std::vector<char> has(n, 0);
    int missing = 0;
    for (int i = 0; i < n; i++) {
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++)
            if (A.Aj[k] == i) has[i] = 1;
        missing += !has[i];
    }
    M.num_nnzs = A.Ap[n] + missing;
    M.Ap = lssp_malloc<int>(n + 1);
    M.Aj = lssp_malloc<int>(M.num_nnzs);
    M.Ax = lssp_malloc<double>(M.num_nnzs);
    M.Ap[0] = 0;
    for (int i = 0; i < n; i++) {
        int o = M.Ap[i];
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++, o++) { M.Aj[o] = A.Aj[k]; M.Ax[o] = A.Ax[k]; }
        if (!has[i]) {
            M.Aj[o] = i;
            M.Ax[o] = 1 * tol;
            for (int q = o; q > M.Ap[i] && M.Aj[q - 1] > M.Aj[q]; q--) {
                std::swap(M.Aj[q - 1], M.Aj[q]);
                std::swap(M.Ax[q - 1], M.Ax[q]);
            }
            o++;
        }
        M.Ap[i + 1] = o;
    }
    return M;
}

// :589-698: entries outside the row's own diagonal block (blocks of blk_size rows, the last one shorter) are
// discarded; a row left empty becomes the unit row.  This is the block-Jacobi restriction used when sharding.
lssp_mat_csr lssp_mat_get_block_diag(const lssp_mat_csr A, int blk_size)
{
    assert(A.num_nnzs > 0 && A.num_rows > 0 && A.num_cols > 0 && A.num_rows == A.num_cols && blk_size > 0);
    const int n = A.num_rows;
    lssp_mat_csr M = A;
    if (blk_size == n) {
        M.Ap = lssp_copy_on<int>(A.Ap, n + 1);
        M.Aj = lssp_copy_on<int>(A.Aj, A.num_nnzs);
        M.Ax = lssp_copy_on<double>(A.Ax, A.num_nnzs);
        return M;
    }
    M.Ap = lssp_malloc<int>(n + 1);
    M.Ap[0] = 0;
    for (int i = 0; i < n; i++) {
        const int lo = (i / blk_size) * blk_size, hi = std::min(n, lo + blk_size);
        int kept = 0;
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++) kept += (A.Aj[k] >= lo && A.Aj[k] < hi);
        M.Ap[i + 1] = M.Ap[i] + (kept ? kept : 1);
    }
    M.num_nnzs = M.Ap[n];
    M.Aj = lssp_malloc<int>(M.num_nnzs);
    M.Ax = lssp_malloc<double>(M.num_nnzs);
    for (int i = 0; i < n; i++) {
        const int lo = (i / blk_size) * blk_size, hi = std::min(n, lo + blk_size);
        int o = M.Ap[i];
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++)
            if (A.Aj[k] >= lo && A.Aj[k] < hi) { M.Aj[o] = A.Aj[k]; M.Ax[o] = A.Ax[k]; o++; }
        if (o == M.Ap[i]) { M.Aj[o] = i; M.Ax[o] = 1; }
    }
    return M;
}

lssp_mat_csr lssp_mat_transpose(const lssp_mat_csr A)   // :700-765: rows of the result ordered by source row
{
    assert(A.num_rows > 0 && A.num_cols > 0);
    lssp_mat_csr T;
    lssp_mat_init(T);
    T.num_rows = A.num_cols;
    T.num_cols = A.num_rows;
    T.num_nnzs = A.num_nnzs;
    if (A.num_nnzs <= 0) return T;
    const int nnz = A.Ap[A.num_rows];
    T.Ap = lssp_malloc<int>(T.num_rows + 1);
    T.Aj = lssp_malloc<int>(A.num_nnzs);
    T.Ax = lssp_malloc<double>(A.num_nnzs);
    std::vector<int> pos(T.num_rows + 1, 0);
    for (int k = 0; k < nnz; k++) pos[A.Aj[k] + 1]++;
    for (int c = 0; c < T.num_rows; c++) pos[c + 1] += pos[c];
    memcpy(T.Ap, pos.data(), sizeof(int) * (T.num_rows + 1));
    for (int i = 0; i < A.num_rows; i++)
        for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++) {
            const int at = pos[A.Aj[k]]++;
            T.Aj[at] = i;
            T.Ax[at] = A.Ax[k];
        }
    return T;
}

// ---- vector (reference src/vector.cxx) ------------------------------------------------------------
lssp_vec lssp_vec_create(int n)
{
    lssp_vec v;
    assert(n >= 0);            // n == 0 gives an empty vector, as the reference (src/vector.cxx:8-17)
    v.n = n;
    v.d = n > 0 ? lssp_malloc<double>(n) : NULL;
    return v;
}

void lssp_vec_destroy(lssp_vec &v) { lssp_free(v.d); v.n = 0; }
void lssp_vec_set_value(lssp_vec x, double val) { for (int i = 0; i < x.n; i++) x.d[i] = val; }        // plain stores, as :31-38
void lssp_vec_set_value_by_array(lssp_vec x, double *val) { memcpy(x.d, val, sizeof(double) * x.n); }
void lssp_vec_set_value_by_index(lssp_vec x, int i, double val) { assert(i >= 0 && i < x.n); x.d[i] = val; }
void lssp_vec_get_value(double *val, lssp_vec x) { memcpy(val, x.d, sizeof(double) * x.n); }
double lssp_vec_get_value_by_index(lssp_vec x, int i) { assert(i >= 0 && i < x.n); return x.d[i]; }
void lssp_vec_copy(lssp_vec des, const lssp_vec src) { assert(des.n == src.n); memcpy(des.d, src.d, sizeof(double) * src.n); }

// arithmetic on host vectors: staged through device memory, computed by the BLAS-1 kernels
struct DevVec {
    double *d = NULL;
    int n;
    DevVec(const lssp_vec &h, bool upload = true) : n(h.n)
    {
        GPU(lsspg_malloc(ctx(), sizeof(double) * (size_t)n, (void **)&d));
        if (upload) GPU(lsspg_h2d(ctx(), d, h.d, sizeof(double) * (size_t)n));
    }
    void download(lssp_vec h) { GPU(lsspg_d2h(ctx(), h.d, d, sizeof(double) * (size_t)n)); }
    ~DevVec() { lsspg_free(ctx(), d); }
};

void lssp_vec_axy(double alpha, const lssp_vec x, lssp_vec y)
{
    assert(x.n == y.n);
    DevVec dx(x), dy(y, false);
    GPU(lsspg_vec_axy(ctx(), x.n, alpha, dx.d, dy.d));
    dy.download(y);
}

void lssp_vec_axpby(double alpha, const lssp_vec x, double beta, lssp_vec y)
{
    assert(x.n == y.n);
    DevVec dx(x), dy(y);
    GPU(lsspg_vec_axpby(ctx(), x.n, alpha, dx.d, beta, dy.d));
    dy.download(y);
}

void lssp_vec_axpbyz(double alpha, const lssp_vec x, double beta, lssp_vec y, lssp_vec z)
{
    assert(x.n == y.n && z.n == y.n);
    DevVec dx(x), dy(y), dz(z, false);
    GPU(lsspg_vec_axpbyz(ctx(), x.n, alpha, dx.d, beta, dy.d, dz.d));
    dz.download(z);
}

double lssp_vec_dot(const lssp_vec x, const lssp_vec y)
{
    assert(x.n == y.n);
    DevVec dx(x), dy(y);
    double r = 0;
    GPU(lsspg_vec_dot(ctx(), x.n, dx.d, dy.d, &r));
    return r;
}

double lssp_vec_norm(const lssp_vec x) { return sqrt(lssp_vec_dot(x, x)); }   // :135-138

void lssp_vec_scale(lssp_vec x, double a)
{
    DevVec dx(x);
    GPU(lsspg_vec_scale(ctx(), x.n, dx.d, a));
    dx.download(x);
}

// ---- mvops (reference src/mvops.cxx) ---------------------------------------------------------------
// Standalone calls upload the matrix for the call; inside the solvers the matrix is resident.
static void mv_host(int kind, const lssp_mat_csr &A, double alpha, const lssp_vec &x, double beta, const double *y, lssp_vec &z)
{
    assert(x.n == A.num_cols);
    lsspg_csr *dA = NULL;
    GPU(lsspg_csr_upload(ctx(), A.num_rows, A.num_cols, A.Ap, A.Aj, A.Ax, &dA));
    GPU(lsspg_mv_host(ctx(), kind, dA, alpha, x.d, beta, y, z.d));
    lsspg_csr_destroy(ctx(), dA);
}

void lssp_mv_amxpby(double alpha, const lssp_mat_csr A, const lssp_vec x, double beta, lssp_vec y)
{
    assert(x.n && y.n);
    mv_host(LSSPG_MV_AMXPBY, A, alpha, x, beta, y.d, y);
}

void lssp_mv_amxpbyz(double alpha, const lssp_mat_csr A, const lssp_vec x, double beta, const lssp_vec y, lssp_vec z)
{
    assert(x.n == y.n && y.n == z.n);
    mv_host(LSSPG_MV_AMXPBYZ, A, alpha, x, beta, y.d, z);
}

void lssp_mv_amxy(double a, const lssp_mat_csr A, const lssp_vec x, lssp_vec y)
{
    assert(x.n == y.n);
    mv_host(LSSPG_MV_AMXY, A, a, x, 0., NULL, y);
}

void lssp_mv_mxy(const lssp_mat_csr A, const lssp_vec x, lssp_vec y)
{
    assert(x.n == y.n);
    mv_host(LSSPG_MV_MXY, A, 1., x, 0., NULL, y);
}

// ---- triangular sweeps (reference src/solver-tri.cxx) ------------------------------------------------
static void tri_host(int which, const lssp_mat_csr &T, double *x, const double *rhs)
{
    assert(x != NULL && rhs != NULL);
    lsspg_tri *dT = NULL;
    const size_t nb = sizeof(double) * (size_t)T.num_rows;
    double *dx = NULL, *dr = NULL;
    GPU(lsspg_tri_analyse(ctx(), which, T.num_rows, T.Ap, T.Aj, T.Ax, &dT));
    GPU(lsspg_malloc(ctx(), nb, (void **)&dx));
    GPU(lsspg_malloc(ctx(), nb, (void **)&dr));
    GPU(lsspg_h2d(ctx(), dr, rhs, nb));
    GPU(lsspg_tri_solve(ctx(), dT, dx, dr));
    GPU(lsspg_check_flags(ctx()));   // a sweep aborted by the watchdog is an error here, not a silently wrong x
    GPU(lsspg_d2h(ctx(), x, dx, nb));
    lsspg_free(ctx(), dx);
    lsspg_free(ctx(), dr);
    lsspg_tri_destroy(ctx(), dT);
}

void lssp_pc_ilu_solve_lower_matrix(lssp_mat_csr L, double *x, double *rhs) { tri_host(LSSPG_TRI_LOWER, L, x, rhs); }
void lssp_pc_ilu_solve_upper_matrix(lssp_mat_csr U, double *x, double *rhs) { tri_host(LSSPG_TRI_UPPER, U, x, rhs); }

void lssp_pc_ilu_solve_lu_matrix(lssp_mat_csr L, lssp_mat_csr U, double *x, double *rhs, double *cache)
{
    assert(x != NULL && rhs != NULL && cache != NULL);
    lssp_pc_ilu_solve_lower_matrix(L, cache, rhs);
    lssp_pc_ilu_solve_upper_matrix(U, x, cache);
}

// pc.solve of ILUK / ILUT (reference src/solver-tri.cxx:57-60): the factors are resident on the device
void lssp_pc_ilu_solve(LSSP_PC *pc, lssp_vec x, lssp_vec rhs)
{
    assert(pc_dev(pc) != NULL);
    GPU(lsspg_pc_apply_host(ctx(), (lsspg_pc *)pc_dev(pc), x.d, rhs.d));
}

// ---- preconditioners (reference src/pc.cxx, src/pc-iluk.cxx:554-592, src/pc-ilut.cxx:423-466) ----------
void lssp_pc_create(LSSP_PC &pc, LSSP_PC_TYPE type)
{
    pc.type = type;
    pc.iluk_level = lssp_pc_iluk_level_default;
    pc.ilut_tol = lssp_pc_ilut_tol;
    pc.ilut_p = (int)lssp_pc_ilut_p;
    pc.cache = NULL;
    pc.solve = NULL;
    pc.destroy = NULL;
    pc.data = NULL;
    side_set(&pc, NULL);
    lssp_mat_init(pc.L);
    lssp_mat_init(pc.U);
    lssp_mat_init(pc.D);
    pc.verb = lssp_verbosity - 1;
    pc.log = NULL;
    pc.assembled = false;
#if USE_SXAMG
    pc.sxamg = NULL;
    if (type == LSSP_PC_SXAMG) lssp_pc_sxamg_create(pc);   // src/pc.cxx:36-40
#endif
}

#if USE_SXAMG
// ---- SX-AMG adapters (reference src/pc-sxamg.cxx, src/solver-sxamg.cxx).  libsxamg itself is not in the
// reference tree: the hierarchy and the cycle are this build's own (lsspg_amg_*, DESIGN.md "AMG").
struct SX_DATA_ {
    SX_AMG_PARS pars;
};
struct SXAMG_DATA_ {
    SX_AMG_PARS pars;
};

void sx_amg_pars_init(SX_AMG_PARS *p)
{
    lsspg_amg_pars d;
    lsspg_amg_pars_default(&d);
    bzero(p, sizeof(*p));
    p->verb = 0;
    p->cycle_itr = 1;
    p->tol = d.tol;
    p->ctol = 1e-7;
    p->maxit = d.maxit;
    p->cs_type = 1;
    p->interp_type = 1;
    p->max_levels = d.max_levels;
    p->max_coarsest_dof = d.coarse_dof;
    p->strong_threshold = d.strong_threshold;
    p->max_row_sum = d.max_row_sum;
    p->trunc_threshold = d.trunc_threshold;
    p->smoother = 1;
    p->relaxation = 1.0;
    p->cf_order = d.cf_order;
    p->pre_iter = d.pre_iter;
    p->post_iter = d.post_iter;
    const char *e = getenv("LSSP_SXAMG_ZERO_GUESS");
    p->zero_guess = e ? atoi(e) : d.zero_guess;
}

static lsspg_amg_pars to_native(const SX_AMG_PARS &p)
{
    lsspg_amg_pars d;
    lsspg_amg_pars_default(&d);
    d.max_levels = p.max_levels;
    d.coarse_dof = p.max_coarsest_dof;
    d.strong_threshold = p.strong_threshold;
    d.max_row_sum = p.max_row_sum;
    d.trunc_threshold = p.trunc_threshold;
    d.pre_iter = p.pre_iter;
    d.post_iter = p.post_iter;
    d.cf_order = p.cf_order;
    d.zero_guess = p.zero_guess;
    d.tol = p.tol;
    d.maxit = p.maxit;
    d.verb = p.verb > 0 ? p.verb : 0;
    return d;
}

static lsspg_pc *build_device_amg(const lssp_mat_csr &A, const lsspg_csr *dA, const SX_AMG_PARS &pars)
{
    const lsspg_amg_pars np = to_native(pars);
    lsspg_amg_host *H = NULL;
    GPU(lsspg_amg_setup_host(A.num_rows, A.Ap, A.Aj, A.Ax, &np, &H));
    lsspg_pc *d = NULL;
    GPU(lsspg_pc_create_amg(ctx(), H, dA, &d));
    lsspg_amg_host_destroy(H);
    return d;
}

void lssp_pc_sxamg_create(LSSP_PC &pc)   // src/pc-sxamg.cxx:12-25
{
    pc.sxamg = lssp_malloc<SX_DATA_>(1);
    sx_amg_pars_init(&pc.sxamg->pars);
    pc.sxamg->pars.maxit = 1;
    pc.sxamg->pars.verb = pc.verb;
}

static void lssp_pc_sxamg_destroy(LSSP_PC *pc)   // src/pc-sxamg.cxx:27-40
{
    if (pc == NULL || pc->sxamg == NULL) return;
    lssp_free(pc->sxamg);
    pc->sxamg = NULL;
}

// pc.solve: one cycle from the incoming x (src/pc-sxamg.cxx:42-73)
static void amg_solve(LSSP_PC *pc, lssp_vec x, lssp_vec rhs)
{
    assert(pc != NULL && pc_dev(pc) != NULL);
    GPU(lsspg_pc_apply_host(ctx(), (lsspg_pc *)pc_dev(pc), x.d, rhs.d));
}

void lssp_pc_sxamg_assemble(LSSP_PC &pc, LSSP_SOLVER s)   // src/pc-sxamg.cxx:75-126
{
    assert(pc.sxamg != NULL);
    pc.data = build_device_amg(s.A, (const lsspg_csr *)solver_dev(s), pc.sxamg->pars);
    pc.solve = amg_solve;
    pc.destroy = lssp_pc_sxamg_destroy;
}

void lssp_pc_sxamg_set_pars(LSSP_PC &pc, SX_AMG_PARS *pars)   // src/pc-sxamg.cxx:128-131
{
    if (pars != NULL && pc.sxamg != NULL) pc.sxamg->pars = *pars;
}

void lssp_solver_sxamg_create(LSSP_SOLVER &s)   // src/solver-sxamg.cxx:11-15
{
    s.sxamg = lssp_malloc<SXAMG_DATA_>(1);
    sx_amg_pars_init(&s.sxamg->pars);
}

void lssp_solver_sxamg_destroy(LSSP_SOLVER &s)   // src/solver-sxamg.cxx:17-23
{
    if (s.sxamg != NULL) {
        lssp_free(s.sxamg);
        s.sxamg = NULL;
    }
}

// stand-alone AMG iteration (src/solver-sxamg.cxx:25-99): set-up and cycles in one call, as sx_solver_amg
int lssp_solver_sxamg(LSSP_SOLVER *solver)
{
    solver->sxamg->pars.maxit = solver->maxit;
    solver->sxamg->pars.verb = solver->verb;
    solver->sxamg->pars.tol = solver->tol_rel;
    lsspg_pc *d = build_device_amg(solver->A, (const lsspg_csr *)solver_dev(*solver), solver->sxamg->pars);
    int nits = 0;
    double ares = 0.;
    GPU(lsspg_amg_solve_host(ctx(), d, solver->rhs.d, solver->x.d, solver->tol_rel, solver->maxit, &nits, &ares));
    lsspg_pc_destroy(ctx(), d);
    solver->residual = ares;   // :96
    solver->nits = nits;
    return nits;
}

void lssp_solver_sxamg_set_pars(LSSP_SOLVER *solver, SX_AMG_PARS *pars)   // src/solver-sxamg.cxx:102-105
{
    if (pars != NULL && solver->sxamg != NULL) solver->sxamg->pars = *pars;
}
#endif

static void release_device_pc(LSSP_PC *pc)
{
    if (pc_dev(pc)) lsspg_pc_destroy(ctx(), (lsspg_pc *)pc_dev(pc));
    pc_dev_set(pc, NULL);
}

void lssp_pc_destroy(LSSP_PC &pc)
{
    assert(pc.assembled);
    if (pc.destroy != NULL) (*pc.destroy)(&pc);
    release_device_pc(&pc);
    pc.assembled = false;
}

static void non_solve(LSSP_PC *pc, lssp_vec x, lssp_vec rhs) { lssp_memcpy_on(x.d, rhs.d, pc->A.num_rows); }   // src/pc.cxx:67-70

static void adopt_factors(LSSP_PC &pc, lsspg_factors *F)
{
    int n, nl, nu;
    lsspg_factors_sizes(F, &n, &nl, &nu);
    pc.L.num_rows = pc.L.num_cols = pc.U.num_rows = pc.U.num_cols = n;
    pc.L.num_nnzs = nl;
    pc.U.num_nnzs = nu;
    pc.L.Ap = lssp_malloc<int>(n + 1); pc.L.Aj = lssp_malloc<int>(nl); pc.L.Ax = lssp_malloc<double>(nl);
    pc.U.Ap = lssp_malloc<int>(n + 1); pc.U.Aj = lssp_malloc<int>(nu); pc.U.Ax = lssp_malloc<double>(nu);
    lsspg_factors_get(F, pc.L.Ap, pc.L.Aj, pc.L.Ax, pc.U.Ap, pc.U.Aj, pc.U.Ax);
    lsspg_factors_destroy(F);
    lsspg_pc *d = NULL;
    GPU(lsspg_pc_create_ilu(ctx(), n, pc.L.Ap, pc.L.Aj, pc.L.Ax, pc.U.Ap, pc.U.Aj, pc.U.Ax, &d));
    pc_dev_set(&pc, d);
    pc.cache = lssp_malloc<double>(n);
    pc.solve = lssp_pc_ilu_solve;
}

void lssp_pc_iluk_destroy(LSSP_PC *pc)
{
    assert(pc->assembled);
    lssp_mat_destroy(pc->L);
    lssp_mat_destroy(pc->U);
    lssp_free<double>(pc->cache);
    release_device_pc(pc);
    pc->assembled = false;
}

void lssp_pc_iluk_assemble(LSSP_PC &pc, LSSP_SOLVER s)
{
    assert(s.A.num_rows == s.A.num_cols && s.A.num_rows > 0 && s.A.num_nnzs > 0);
    lsspg_factors *F = NULL;
    GPU(lsspg_ilu_factor(LSSPG_ILUK, s.A.num_rows, s.A.Ap, s.A.Aj, s.A.Ax, pc.iluk_level, 0, 0., 0, &F));
    adopt_factors(pc, F);
    pc.destroy = lssp_pc_iluk_destroy;
}

void lssp_pc_iluk_set_level(LSSP_PC &pc, int level)
{
    if (level < 0) {
        lssp_warning("pc: level is too small, set it to %d!\n", lssp_pc_iluk_level_default);
        pc.iluk_level = lssp_pc_iluk_level_default;
    }
    else pc.iluk_level = level;
}

void lssp_pc_ilut_destroy(LSSP_PC *pc) { lssp_pc_iluk_destroy(pc); }

void lssp_pc_ilut_assemble(LSSP_PC &pc, LSSP_SOLVER s)
{
    assert(s.A.num_rows == s.A.num_cols && s.A.num_rows > 0 && s.A.num_nnzs > 0);
    if (pc.ilut_p <= 0) pc.ilut_p = (s.A.num_nnzs + s.A.num_rows - 1) / s.A.num_rows;   // written back, as :436-438
    if (pc.ilut_tol < 0) pc.ilut_tol = lssp_pc_ilut_tol;
    if (pc.verb > 1) lssp_printf("pc: ilut, tol: %f, p: %d\n", pc.ilut_tol, pc.ilut_p);
    lsspg_factors *F = NULL;
    GPU(lsspg_ilu_factor(LSSPG_ILUT, s.A.num_rows, s.A.Ap, s.A.Aj, s.A.Ax, 0, pc.ilut_p, pc.ilut_tol, 0, &F));
    adopt_factors(pc, F);
    pc.destroy = lssp_pc_ilut_destroy;
}

void lssp_pc_ilut_set_drop_tol(LSSP_PC &pc, double tol) { pc.ilut_tol = fabs(tol); }
void lssp_pc_ilut_set_p(LSSP_PC &pc, int p) { pc.ilut_p = p; }

// ---- block ILU(k) (reference src/pc-biluk.cxx) ------------------------------------------------------------
// pc.solve: x = U^-1 D L^-1 rhs (src/pc-biluk.cxx:22-60) with the factors resident on the device
void lssp_pc_bilu_solve(LSSP_PC *pc, lssp_vec x, lssp_vec rhs)
{
    assert(pc_dev(pc) != NULL);
    GPU(lsspg_pc_apply_host(ctx(), (lsspg_pc *)pc_dev(pc), x.d, rhs.d));
}

void lssp_pc_biluk_destroy(LSSP_PC *pc)   // src/pc-biluk.cxx:303-313
{
    assert(pc->assembled);
    lssp_mat_destroy(pc->L);
    lssp_mat_destroy(pc->U);
    lssp_mat_destroy(pc->D);
    lssp_free<double>(pc->cache);
    release_device_pc(pc);
    pc->assembled = false;
}

static void adopt_block_factors(LSSP_PC &pc, lsspg_bfactors *F)
{
    int n, nl, nd, nu;
    lsspg_bfactors_sizes(F, &n, &nl, &nd, &nu);
    lssp_mat_csr *M[3] = {&pc.L, &pc.D, &pc.U};
    const int nz[3] = {nl, nd, nu};
    for (int q = 0; q < 3; q++) {
        M[q]->num_rows = M[q]->num_cols = n;
        M[q]->num_nnzs = nz[q];
        M[q]->Ap = lssp_malloc<int>(n + 1);
        M[q]->Aj = lssp_malloc<int>(nz[q]);
        M[q]->Ax = lssp_malloc<double>(nz[q]);
    }
    lsspg_bfactors_get(F, pc.L.Ap, pc.L.Aj, pc.L.Ax, pc.D.Ap, pc.D.Aj, pc.D.Ax, pc.U.Ap, pc.U.Aj, pc.U.Ax);
    lsspg_bfactors_destroy(F);
    lsspg_pc *d = NULL;
    GPU(lsspg_pc_create_bilu(ctx(), n, pc.L.Ap, pc.L.Aj, pc.L.Ax, pc.D.Ap, pc.D.Aj, pc.D.Ax, pc.U.Ap, pc.U.Aj, pc.U.Ax, &d));
    pc_dev_set(&pc, d);
    pc.cache = lssp_malloc<double>(2 * n);   // src/pc-biluk.cxx:406
    pc.solve = lssp_pc_bilu_solve;
    pc.destroy = lssp_pc_biluk_destroy;
}

// src/pc-biluk.cxx:377-414.  The block matrix is expanded to CSR (every entry of every block, stored zeros included:
// the factors keep them, src/pc-biluk.cxx:129-171) and handed to the same set-up as lssp_pc_biluk_assemble; block
// rows need not be sorted.
void lssp_pc_biluk_assemble_mat(LSSP_PC &pc, lssp_mat_bcsr A)
{
    double time = lssp_get_time();
    const int bs = A.blk_size, bs2 = bs * bs;
    assert(bs > 0 && A.num_rows == A.num_cols && A.num_rows > 0);
    const int n = A.num_rows * bs;
    const int nnz = A.Ap[A.num_rows] * bs2;
    int *Ap = lssp_malloc<int>(n + 1), *Aj = lssp_malloc<int>(nnz);
    double *Ax = lssp_malloc<double>(nnz);
    int o = 0;
    Ap[0] = 0;
    for (int i = 0; i < A.num_rows; i++)
        for (int a = 0; a < bs; a++) {
            for (int k = A.Ap[i]; k < A.Ap[i + 1]; k++)
                for (int b = 0; b < bs; b++, o++) {
                    Aj[o] = A.Aj[k] * bs + b;
                    Ax[o] = A.Ax[(size_t)k * bs2 + b * bs + a];
                }
            Ap[i * bs + a + 1] = o;
        }
    lsspg_bfactors *F = NULL;
    GPU(lsspg_bilu_factor(n, Ap, Aj, Ax, A.num_rows, pc.iluk_level, &F));
    lssp_free(Ap); lssp_free(Aj); lssp_free(Ax);
    adopt_block_factors(pc, F);
    if (pc.verb > 0) lssp_printf("pc: BILUK assemble time: %f\n", lssp_get_time() - time);
}

void lssp_pc_biluk_assemble(LSSP_PC &pc, LSSP_SOLVER s)   // src/pc-biluk.cxx:416-431
{
    assert(s.A.num_rows == s.A.num_cols);
    assert(s.A.num_rows > 0 && s.A.num_nnzs > 0);
    assert(s.num_blks > 0);
    assert(s.A.num_rows % s.num_blks == 0);
    double time = lssp_get_time();
    lsspg_bfactors *F = NULL;
    GPU(lsspg_bilu_factor(s.A.num_rows, s.A.Ap, s.A.Aj, s.A.Ax, s.num_blks, pc.iluk_level, &F));
    adopt_block_factors(pc, F);
    if (pc.verb > 0) lssp_printf("pc: BILUK assemble time: %f\n", lssp_get_time() - time);
}

// user-defined preconditioner: its pc.solve works on host vectors
static void user_pc_trampoline(void *user, double *hx, const double *hrhs, int n)
{
    LSSP_PC *pc = (LSSP_PC *)user;
    lssp_vec x, rhs;
    x.n = rhs.n = n;
    x.d = hx;
    rhs.d = const_cast<double *>(hrhs);
    pc->solve(pc, x, rhs);
}

void lssp_pc_assemble(LSSP_PC &pc, LSSP_SOLVER s)
{
    double tm = 0;
    if (!s.assembled) lssp_error(1, "pc: solver hasn't been assembled, call lssp_assemble first!\n");
    if (s.verb > 0) tm = lssp_get_time();
    pc.verb = s.verb - 1;
    pc.log = s.log;
    lssp_mat_init(pc.L);
    lssp_mat_init(pc.U);
    lssp_mat_init(pc.D);
    pc.A = s.A;
    if (pc.type != LSSP_PC_USER) pc.data = NULL;
    switch (pc.type) {
        case LSSP_PC_NON: {
            pc.cache = NULL;
            pc.solve = non_solve;
            pc.destroy = NULL;
            lsspg_pc *d = NULL;
            GPU(lsspg_pc_create_non(ctx(), s.A.num_rows, &d));
            pc_dev_set(&pc, d);
            break;
        }
        case LSSP_PC_ILUK:
            if (pc.verb >= 0) lssp_printf("pc: type: ILUK\n");
            lssp_pc_iluk_assemble(pc, s);
            break;
        case LSSP_PC_ILUT:
            if (pc.verb >= 0) lssp_printf("pc: type: ILUT\n");
            lssp_pc_ilut_assemble(pc, s);
            break;
        case LSSP_PC_BILUK:   // src/pc.cxx:124-135
            if (pc.verb >= 0) lssp_printf("pc: type: block version ILUK\n");
            lssp_pc_biluk_assemble(pc, s);
            break;
#if USE_SXAMG
        case LSSP_PC_SXAMG:   // src/pc.cxx:208-217
            if (pc.verb >= 0) lssp_printf("pc: type: SX-AMG\n");
            lssp_pc_sxamg_assemble(pc, s);
            break;
#endif
        case LSSP_PC_USER: {
            if (pc.verb >= 0) lssp_printf("pc: type: user defined\n");
            assert(pc.assemble != NULL);
            pc.assemble(pc, s);
            assert(pc.solve != NULL);
            lsspg_pc *d = NULL;
            GPU(lsspg_pc_create_user(ctx(), s.A.num_rows, user_pc_trampoline, &pc, &d));
            pc_dev_set(&pc, d);
            break;
        }
        default:
            lssp_error(1, "pc: incorrect preconditioner type !\n");
            break;
    }
    if (pc.verb > 0) lssp_printf("pc: time for pc assemble: %f s\n", lssp_get_time() - tm);
    pc.assembled = true;
}

// ---- solver life cycle (reference src/lssp.cxx:16-249, :416-535) ----------------------------------------
void lssp_solver_create(LSSP_SOLVER &s, LSSP_SOLVER_TYPE s_type, LSSP_PC &pc, LSSP_PC_TYPE p_type)
{
    bzero(&s, sizeof(LSSP_SOLVER));
    s.type = s_type;
    s.residual = 0;
    s.tol_rel = LSSP_RTOL;
    s.tol_abs = LSSP_ATOL;
    s.tol_rb = LSSP_RB;
    s.restart = LSSP_RESTART;
    s.aug_k = LSSP_AUG_K;
    s.maxit = LSSP_MAXIT;
    s.nits = 0;
    s.bgsl = LSSP_BGSL;
    s.idrs = LSSP_IDRS;
    lssp_mat_init(s.A);
    s.num_blks = -1;
    s.blk_size = NULL;
    s.log = NULL;
    s.verb = lssp_verbosity;
#if USE_SXAMG
    if (s_type == LSSP_SOLVER_SXAMG) {   // src/lssp.cxx:131-136
        p_type = LSSP_PC_NON;
        lssp_solver_sxamg_create(s);
    }
#endif
    lssp_pc_create(pc, p_type);
    s.assembled = false;
}

void lssp_solver_assemble(LSSP_SOLVER &s, lssp_mat_csr &Ax, lssp_vec x, lssp_vec b, LSSP_PC &pc)
{
    double t = 0.;
    if (Ax.num_rows <= 0) lssp_error(1, "solver: wrong input matrix, number of rows should be greater than 0\n");
    else if (Ax.num_rows != Ax.num_cols) lssp_error(1, "solver: wrong input matrix, number of rows != number of columns\n");
    if (Ax.num_nnzs < Ax.num_rows) lssp_error(1, "solver: wrong input matrix, singular\n");
    if (s.verb > 1) t = lssp_get_time();
    lssp_mat_csr A;
    A.num_rows = Ax.num_rows;
    A.num_cols = Ax.num_cols;
    A.num_nnzs = Ax.num_nnzs;
    A.Ap = lssp_malloc<int>(Ax.num_rows + 1);                    // deep copy, :169-171 (by the host threads)
    A.Aj = lssp_malloc<int>(Ax.num_nnzs);
    A.Ax = lssp_malloc<double>(Ax.num_nnzs);
    lsspg::parallel_copy(A.Ap, Ax.Ap, sizeof(int) * ((size_t)Ax.num_rows + 1));
    lsspg::parallel_copy(A.Aj, Ax.Aj, sizeof(int) * (size_t)Ax.num_nnzs);
    lsspg::parallel_copy(A.Ax, Ax.Ax, sizeof(double) * (size_t)Ax.num_nnzs);
    if (!lssp_mat_csr_is_sorted(A)) lssp_mat_sort_column(A);    // :173
    s.rhs = b;                                                  // aliases, :175-176
    s.x = x;
    s.A = A;
    lsspg_csr *dA = NULL;
    GPU(lsspg_csr_upload(ctx(), A.num_rows, A.num_cols, A.Ap, A.Aj, A.Ax, &dA));
    side_set(s.A.Ap, dA);
    if (s.verb > 1) lssp_printf("solver: assemble time: %g\n", lssp_get_time() - t);
    s.assembled = true;
    lssp_pc_assemble(pc, s);
}

void lssp_solver_destroy(LSSP_SOLVER &s, LSSP_PC &pc)
{
    assert(s.assembled);
    lssp_mat_destroy(s.A);
#if USE_SXAMG
    if (s.type == LSSP_SOLVER_SXAMG) lssp_solver_sxamg_destroy(s);   // src/lssp.cxx:240-244
#endif
    lssp_pc_destroy(pc);   // before the matrix: an AMG preconditioner shares the device copy of A
    if (solver_dev(s)) {
        lsspg_csr_destroy(ctx(), (lsspg_csr *)solver_dev(s));
        side_set(s.A.Ap, NULL);
    }
    s.assembled = false;
}

void lssp_solver_reset_rhs(LSSP_SOLVER &s, lssp_vec rhs) { assert(s.assembled); s.rhs = rhs; }
void lssp_solver_reset_unknown(LSSP_SOLVER &s, lssp_vec x) { assert(s.assembled); s.x = x; }
void lssp_solver_reset_type(LSSP_SOLVER &s, LSSP_SOLVER_TYPE type) { s.type = type; }

#define SETTER(name, field, type, cond, msg)            \
    void name(LSSP_SOLVER &s, type v)                   \
    {                                                   \
        if (cond) lssp_warning(msg);                    \
        else s.field = v;                               \
    }
SETTER(lssp_solver_set_rtol, tol_rel, double, v < 0, "solver: tol is less than zero!\n")
SETTER(lssp_solver_set_atol, tol_abs, double, v < 0, "solver: tol is less than zero!\n")
SETTER(lssp_solver_set_rbtol, tol_rb, double, v < 0, "solver: tol is less than zero!\n")
SETTER(lssp_solver_set_maxit, maxit, int, v <= 0, "solver: maxit is less or equal to zero!\n")
SETTER(lssp_solver_set_restart, restart, int, v <= 0, "solver: restart is too small!\n")
SETTER(lssp_solver_set_augk, aug_k, int, v <= 0, "solver: aug_k is less or equal to zero!\n")
SETTER(lssp_solver_set_bgsl, bgsl, int, v <= 0, "solver: bgsl is less or equal to zero!\n")
SETTER(lssp_solver_set_idrs, idrs, int, v <= 0, "solver: idrs is less or equal to zero!\n")
#undef SETTER

void lssp_solver_reset_verbosity(LSSP_SOLVER &s, int v) { s.verb = v; }
double lssp_solver_get_residual(LSSP_SOLVER s) { return s.residual; }
int lssp_solver_get_nits(LSSP_SOLVER s) { return s.nits; }

void lssp_solver_set_log(LSSP_SOLVER &s, FILE *io)
{
    assert(io != NULL);
    s.log = io;
    lssp_set_log(io);
}

// ---- Krylov drivers: int lssp_solver_<m>(LSSP_SOLVER&, LSSP_PC&) (reference src/solver-*.cxx) -----------
static int drive(LSSP_SOLVER &solver, LSSP_PC &pc, int kind, const char *name)
{
    assert(solver.assembled);
    assert(pc.assembled);
    if (!lsspg_solver_supported(kind))
        lssp_error(1, "%s: this driver has no GPU implementation yet in this build\n", name);
    const double t0 = lssp_get_time();
    lsspg_solver_opts o;
    lsspg_solver_opts_default(&o);
    o.tol_rel = solver.tol_rel; o.tol_abs = solver.tol_abs; o.tol_rb = solver.tol_rb;
    o.maxit = solver.maxit; o.restart = solver.restart; o.aug_k = solver.aug_k; o.bgsl = solver.bgsl; o.idrs = solver.idrs;
    o.verb = solver.verb;
    if (solver.verb >= 2) {
        lssp_printf("%s: maximal iteration: %d\n", name, solver.maxit <= 0 ? LSSP_MAXIT : solver.maxit);
        lssp_printf("%s: tolerance abs: %g\n", name, solver.tol_abs < 0 ? LSSP_ATOL : solver.tol_abs);
        lssp_printf("%s: tolerance rel: %g\n", name, solver.tol_rel < 0 ? LSSP_RTOL : solver.tol_rel);
        lssp_printf("%s: tolerance rbn: %g\n", name, solver.tol_rb);
    }
    lsspg_solve_info info;
    GPU(lsspg_krylov_solve_host(ctx(), kind, (lsspg_csr *)solver_dev(solver), (lsspg_pc *)pc_dev(&pc), solver.rhs.d, solver.x.d, &o, &info));
    solver.residual = info.residual;
    solver.nits = info.nits;
    if (solver.verb >= 2) {
        lssp_printf("%s: total iteration: %d\n", name, info.nits);
        lssp_printf("%s: total time: %g\n", name, lssp_get_time() - t0);
    }
    return info.nits;
}

#define DRIVER(fn, kind, name) \
    int fn(LSSP_SOLVER &solver, LSSP_PC &pc) { return drive(solver, pc, kind, name); }
DRIVER(lssp_solver_gmres, LSSPG_GMRES, "gmres")
DRIVER(lssp_solver_gmres_r, LSSPG_RGMRES, "rgmres")
DRIVER(lssp_solver_lgmres, LSSPG_LGMRES, "lgmres")
DRIVER(lssp_solver_lgmres_r, LSSPG_RLGMRES, "rlgmres")
DRIVER(lssp_solver_bicgstab, LSSPG_BICGSTAB, "bicgstab")
DRIVER(lssp_solver_bicgstabl, LSSPG_BICGSTABL, "bicgstabl")
DRIVER(lssp_solver_bicgsafe, LSSPG_BICGSAFE, "bicgsafe")
DRIVER(lssp_solver_cg, LSSPG_CG, "cg")
DRIVER(lssp_solver_cgs, LSSPG_CGS, "cgs")
DRIVER(lssp_solver_gpbicg, LSSPG_GPBICG, "gpbicg")
DRIVER(lssp_solver_cr, LSSPG_CR, "cr")
DRIVER(lssp_solver_crs, LSSPG_CRS, "crs")
DRIVER(lssp_solver_bicrstab, LSSPG_BICRSTAB, "bicrstab")
DRIVER(lssp_solver_bicrsafe, LSSPG_BICRSAFE, "bicrsafe")
DRIVER(lssp_solver_gpbicr, LSSPG_GPBICR, "gpbicr")
DRIVER(lssp_solver_qmrcgstab, LSSPG_QMRCGSTAB, "qmrcgstab")
DRIVER(lssp_solver_tfqmr, LSSPG_TFQMR, "tfqmr")
DRIVER(lssp_solver_orthomin, LSSPG_ORTHOMIN, "orthomin")
DRIVER(lssp_solver_idrs, LSSPG_IDRS, "idrs")
#undef DRIVER

int lssp_solver_solve(LSSP_SOLVER &solver, LSSP_PC &pc)   // dispatch switch, reference src/lssp.cxx:250-414
{
    assert(solver.assembled);
    assert(pc.assembled);
    switch (solver.type) {
        case LSSP_SOLVER_GMRES: return lssp_solver_gmres(solver, pc);
        case LSSP_SOLVER_LGMRES: return lssp_solver_lgmres(solver, pc);
        case LSSP_SOLVER_RGMRES: return lssp_solver_gmres_r(solver, pc);
        case LSSP_SOLVER_RLGMRES: return lssp_solver_lgmres_r(solver, pc);
        case LSSP_SOLVER_BICGSTAB: return lssp_solver_bicgstab(solver, pc);
        case LSSP_SOLVER_BICGSTABL: return lssp_solver_bicgstabl(solver, pc);
        case LSSP_SOLVER_BICGSAFE: return lssp_solver_bicgsafe(solver, pc);
        case LSSP_SOLVER_CG: return lssp_solver_cg(solver, pc);
        case LSSP_SOLVER_CGS: return lssp_solver_cgs(solver, pc);
        case LSSP_SOLVER_GPBICG: return lssp_solver_gpbicg(solver, pc);
        case LSSP_SOLVER_CR: return lssp_solver_cr(solver, pc);
        case LSSP_SOLVER_CRS: return lssp_solver_crs(solver, pc);
        case LSSP_SOLVER_BICRSTAB: return lssp_solver_bicrstab(solver, pc);
        case LSSP_SOLVER_BICRSAFE: return lssp_solver_bicrsafe(solver, pc);
        case LSSP_SOLVER_GPBICR: return lssp_solver_gpbicr(solver, pc);
        case LSSP_SOLVER_QMRCGSTAB: return lssp_solver_qmrcgstab(solver, pc);
        case LSSP_SOLVER_TFQMR: return lssp_solver_tfqmr(solver, pc);
        case LSSP_SOLVER_ORTHOMIN: return lssp_solver_orthomin(solver, pc);
        case LSSP_SOLVER_IDRS: return lssp_solver_idrs(solver, pc);
#if USE_SXAMG
        case LSSP_SOLVER_SXAMG: return lssp_solver_sxamg(&solver);   // src/lssp.cxx:404-408
#endif
        default:
            lssp_error(0, "solver: unsupported solver type!\n");
            return -1;
    }
}

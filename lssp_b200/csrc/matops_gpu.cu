// matops_gpu.cu -- device-side ingest and matrix utilities: what lssp_solver_assemble and the
// preconditioner set-up do to the caller's matrix before the solve loop sees it -- deep copy (src/lssp.cxx:169-171),
// lssp_mat_sort_column (:173, src/matrix-utils.cxx:387-481), lssp_mat_adjust_zero_diag (:483-587),
// lssp_mat_get_block_diag (:589-698), CSR -> BCSR (:62-162) -- and the synthetic stencil generators of SURVEY.md 8d, on
// matrices that live in device memory (lsspg_dmat).  Every result is byte-identical to the host utilities of the facade
// (lssp_facade.cpp) / lssp_b200/generators.py, which are pinned against the reference (tests/cxx/mat_utils_abi_check.cpp);
// tests/test_gpu_setup.py compares them array by array.
#include <algorithm>
#include <vector>
#include "setup_gpu.cuh"
#include "blas1.cuh"
#include "spmv.cuh"

namespace lsspg {

// ---- container ---------------------------------------------------------------------------------------------------
int dmat_alloc(lsspg_ctx *ctx, int n, int m, long long nnz, int bs, bool with_p, lsspg_dmat **out)
{
    lsspg_dmat *M = new lsspg_dmat();
    M->n = n; M->m = m; M->bs = bs; M->nnz = nnz;
    const size_t vals = (size_t)nnz * bs * bs;
    if ((with_p && cudaMalloc(&M->p, sizeof(int) * ((size_t)n + 1 + 8)) != cudaSuccess) ||
        cudaMalloc(&M->j, sizeof(int) * ((size_t)nnz + 16)) != cudaSuccess ||
        cudaMalloc(&M->x, sizeof(double) * (vals + 16)) != cudaSuccess) {
        cudaGetLastError();
        dmat_free(M);
        set_error("device matrix: out of device memory (n %d, nnz %lld)", n, nnz);
        return 1;
    }
    if (with_p) cudaMemsetAsync(M->p + n + 1, 0, sizeof(int) * 8, ctx->stream);
    cudaMemsetAsync(M->j + nnz, 0, sizeof(int) * 16, ctx->stream);
    cudaMemsetAsync(M->x + vals, 0, sizeof(double) * 16, ctx->stream);
    *out = M;
    return 0;
}

void dmat_free(lsspg_dmat *M)
{
    if (!M) return;
    cudaFree(M->p);
    cudaFree(M->j);
    cudaFree(M->x);
    delete M;
}

// ---- exclusive scan (ints; three kernels, 4096 elements per CTA) ---------------------------------------------------
constexpr int kScanPer = 4;

__global__ void __launch_bounds__(1024) k_scan_local(int *d, long long n, long long *sums)
{
    __shared__ long long s_w[32];
    const long long base = ((long long)blockIdx.x * 1024 + threadIdx.x) * kScanPer;
    int v[kScanPer];
    long long loc = 0;
#pragma unroll
    for (int q = 0; q < kScanPer; q++) {
        v[q] = (base + q < n) ? d[base + q] : 0;
        loc += v[q];
    }
    long long total;
    long long run = cta_scan_1024<long long>(loc, s_w, &total) - loc;
#pragma unroll
    for (int q = 0; q < kScanPer; q++) {
        if (base + q < n) d[base + q] = (int)run;   // (chunk-relative: < 2^31 is checked on the grand total)
        run += v[q];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(long long *sums, long long nblocks)
{
    __shared__ long long s_w[32];
    long long carry = 0;
    for (long long b0 = 0; b0 < nblocks; b0 += 1024) {
        const long long b = b0 + threadIdx.x;
        const long long v = (b < nblocks) ? sums[b] : 0;
        long long total;
        const long long inc = cta_scan_1024<long long>(v, s_w, &total);
        if (b < nblocks) sums[b] = carry + inc - v;
        carry += total;
    }
    if (threadIdx.x == 0) sums[nblocks] = carry;
}

__global__ void __launch_bounds__(1024) k_scan_add(int *d, long long n, const long long *sums, long long nblocks)
{
    const long long off = sums[blockIdx.x];
    const long long base = ((long long)blockIdx.x * 1024 + threadIdx.x) * kScanPer;
#pragma unroll
    for (int q = 0; q < kScanPer; q++)
        if (base + q < n) d[base + q] = (int)(d[base + q] + off);
    if (blockIdx.x == 0 && threadIdx.x == 0) d[n] = (int)sums[nblocks];
}

int dev_exclusive_scan(lsspg_ctx *ctx, int *d, long long n, long long *total)
{
    const long long nblocks = std::max<long long>(1, (n + 1024 * kScanPer - 1) / (1024 * kScanPer));
    long long *sums = nullptr;
    LSSPG_CUDA(cudaMalloc(&sums, sizeof(long long) * (size_t)(nblocks + 1)));
    LSSPG_LAUNCH(ctx, k_scan_local, (unsigned int)nblocks, 1024, 0, d, n, sums);
    LSSPG_LAUNCH(ctx, k_scan_sums, 1, 1024, 0, sums, nblocks);
    LSSPG_LAUNCH(ctx, k_scan_add, (unsigned int)nblocks, 1024, 0, d, n, sums, nblocks);
    long long t = 0;
    LSSPG_CUDA(cudaMemcpyAsync(&t, sums + nblocks, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(sums);
    LSSPG_CHECK(t < (1ll << 31), "device matrix: %lld entries exceed the int32 CSR of the reference (include/type-defs.h:15-24)", t);
    if (total) *total = t;
    return 0;
}

// ---- lssp_mat_csr_is_sorted / lssp_mat_sort_column ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rows_sorted(int n, const int *__restrict__ p, const int *__restrict__ j, int *unsorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = p[i] + 1; k < p[i + 1]; k++)
        if (j[k - 1] > j[k]) { *unsorted = 1; return; }
}

// One warp per row: the position of entry e after a STABLE sort by column is the number of entries (c, k) with
// (c, k) < (col_e, e) -- what std::stable_sort produces in the facade (and the reference's qsort for rows without
// duplicate columns).  Rows that are sorted already are copied.  O(len^2 / 32) per row: set-up code for rows of
// tens to a few thousand entries.
__global__ void __launch_bounds__(256) k_sort_rows(int n, const int *__restrict__ p, const int *__restrict__ j,
                                                  const double *__restrict__ x, int *__restrict__ oj, double *__restrict__ ox)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
        const int b = p[i], len = p[i + 1] - b;
        bool bad = false;
        for (int e = lane + 1; e < len; e += 32) bad |= (j[b + e - 1] > j[b + e]);
        bad = __any_sync(0xffffffffu, bad);
        for (int e = lane; e < len; e += 32) {
            const int c = j[b + e];
            int rank = e;
            if (bad) {
                rank = 0;
                for (int k = 0; k < len; k++) {
                    const int ck = j[b + k];
                    rank += (ck < c) || (ck == c && k < e);
                }
            }
            oj[b + rank] = c;
            ox[b + rank] = x[b + e];
        }
    }
}

// ---- lssp_mat_adjust_zero_diag ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_diag_count(int n, const int *__restrict__ p, const int *__restrict__ j, int *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool has = false;
    for (int k = p[i]; k < p[i + 1]; k++) has |= (j[k] == i);
    cnt[i] = p[i + 1] - p[i] + (has ? 0 : 1);
}

// A row without a stored diagonal receives (i, tol) appended and slid towards the front while its left neighbour has a
// larger column (lssp_facade.cpp: lssp_mat_adjust_zero_diag): it ends up in front of the trailing run of columns > i.
__global__ void __launch_bounds__(256) k_diag_fill(int n, const int *__restrict__ p, const int *__restrict__ j,
                                                  const double *__restrict__ x, const int *__restrict__ op,
                                                  int *__restrict__ oj, double *__restrict__ ox, double tol)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = p[i], len = p[i + 1] - b, o = op[i];
    if (op[i + 1] - o == len) {
        for (int k = 0; k < len; k++) { oj[o + k] = j[b + k]; ox[o + k] = x[b + k]; }
        return;
    }
    int at = len;
    while (at > 0 && j[b + at - 1] > i) at--;
    for (int k = 0; k < at; k++) { oj[o + k] = j[b + k]; ox[o + k] = x[b + k]; }
    oj[o + at] = i;
    ox[o + at] = 1 * tol;
    for (int k = at; k < len; k++) { oj[o + k + 1] = j[b + k]; ox[o + k + 1] = x[b + k]; }
}

// ---- lssp_mat_get_block_diag -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_blockdiag_count(int n, int bs, const int *__restrict__ p, const int *__restrict__ j, int *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lo = (i / bs) * bs, hi = min(n, lo + bs);
    int kept = 0;
    for (int k = p[i]; k < p[i + 1]; k++) kept += (j[k] >= lo && j[k] < hi);
    cnt[i] = kept ? kept : 1;
}

__global__ void __launch_bounds__(256) k_blockdiag_fill(int n, int bs, const int *__restrict__ p, const int *__restrict__ j,
                                                       const double *__restrict__ x, const int *__restrict__ op,
                                                       int *__restrict__ oj, double *__restrict__ ox)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lo = (i / bs) * bs, hi = min(n, lo + bs);
    int o = op[i];
    for (int k = p[i]; k < p[i + 1]; k++)
        if (j[k] >= lo && j[k] < hi) { oj[o] = j[k]; ox[o] = x[k]; o++; }
    if (o == op[i]) { oj[o] = i; ox[o] = 1.0; }   // a row left empty becomes the unit row
}

// ---- CSR -> BCSR --------------------------------------------------------------------------------------------------------
// Block row i of B covers rows i bs .. (i + 1) bs - 1 of A, which are contiguous in A: its candidate block columns are
// Aj[k] / bs for k in [Ap[i bs], Ap[(i + 1) bs]).  One warp per block row: an entry is the FIRST occurrence of its
// block column when no earlier candidate has it; first occurrences are ranked by block column (= sorted, unique).
__global__ void __launch_bounds__(256) k_bcsr_count(int nb, int bs, const int *__restrict__ p, const int *__restrict__ j, int *cnt,
                                                   int *__restrict__ firsts)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nb; i += warps) {
        const int b = p[i * bs], len = p[(i + 1) * bs] - b;
        int mine = 0;
        for (int e = lane; e < len; e += 32) {
            const int c = j[b + e] / bs;
            bool first = true;
            for (int k = 0; k < e && first; k++) first = (j[b + k] / bs != c);
            firsts[b + e] = first ? 1 : 0;
            mine += first;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0) cnt[i] = mine;
    }
}

__global__ void __launch_bounds__(256) k_bcsr_cols(int nb, int bs, const int *__restrict__ p, const int *__restrict__ j,
                                                  const int *__restrict__ firsts, const int *__restrict__ bp, int *__restrict__ bj)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < nb; i += warps) {
        const int b = p[i * bs], len = p[(i + 1) * bs] - b;
        for (int e = lane; e < len; e += 32) {
            if (!firsts[b + e]) continue;
            const int c = j[b + e] / bs;
            int rank = 0;   // distinct block columns smaller than c = first occurrences with a smaller block column
            for (int k = 0; k < len; k++) rank += (firsts[b + k] && j[b + k] / bs < c);
            bj[bp[i] + rank] = c;
        }
    }
}

// values: one thread per ROW of A walks its entries in storage order (a duplicate entry overwrites the earlier one, as
// the host loop does); blocks are column-major: B.Ax[k bs^2 + (c % bs) bs + (r % bs)]
__global__ void __launch_bounds__(256) k_bcsr_vals(int n, int bs, const int *__restrict__ p, const int *__restrict__ j,
                                                  const double *__restrict__ x, const int *__restrict__ bp,
                                                  const int *__restrict__ bj, double *__restrict__ bx)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = r / bs, b0 = bp[i], b1 = bp[i + 1];
    for (int k = p[r]; k < p[r + 1]; k++) {
        const int c = j[k], cb = c / bs;
        int lo = b0, hi = b1 - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (bj[mid] < cb) lo = mid + 1;
            else hi = mid;
        }
        bx[(size_t)lo * bs * bs + (size_t)(c % bs) * bs + (r % bs)] = x[k];
    }
}

// ---- generators (lssp_b200/generators.py: stencil_7pt, stencil_7pt_rows, laplacian_5pt) -----------------
// Rows [r0, r1) of the 7-point operator on an nx x ny x nz grid, natural order, entries in ascending column order
// (-nx ny, -nx, -1, 0, +1, +nx, +nx ny) with Dirichlet truncation.  A 2-D 5-point operator is the case nz == 1.
struct Stencil7 {
    double v[7];
};

__global__ void __launch_bounds__(256) k_stencil_count(long long r0, long long rows, int nx, int ny, int nz, int *cnt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    const long long i = r0 + t;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((long long)nx * ny));
    cnt[t] = 1 + (z > 0) + (y > 0) + (x > 0) + (x < nx - 1) + (y < ny - 1) + (z < nz - 1);
}

__global__ void __launch_bounds__(256) k_stencil_fill(long long r0, long long rows, int nx, int ny, int nz, Stencil7 s,
                                                     long long col_shift, const int *__restrict__ p, int *__restrict__ oj,
                                                     double *__restrict__ ox)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows) return;
    const long long i = r0 + t, pl = (long long)nx * ny;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / pl);
    int o = p[t];
    const long long c = i - col_shift;
    if (z > 0) { oj[o] = (int)(c - pl); ox[o] = s.v[0]; o++; }
    if (y > 0) { oj[o] = (int)(c - nx); ox[o] = s.v[1]; o++; }
    if (x > 0) { oj[o] = (int)(c - 1); ox[o] = s.v[2]; o++; }
    oj[o] = (int)c; ox[o] = s.v[3]; o++;
    if (x < nx - 1) { oj[o] = (int)(c + 1); ox[o] = s.v[4]; o++; }
    if (y < ny - 1) { oj[o] = (int)(c + nx); ox[o] = s.v[5]; o++; }
    if (z < nz - 1) { oj[o] = (int)(c + pl); ox[o] = s.v[6]; o++; }
}

// row-count array that becomes the row pointer of a result: n + 1 entries plus the zeroed slack of dmat_alloc
static int count_array(lsspg_ctx *ctx, long long n, int **cnt)
{
    LSSPG_CUDA(cudaMalloc(cnt, sizeof(int) * ((size_t)n + 1 + 8)));
    LSSPG_CUDA(cudaMemsetAsync(*cnt + n, 0, sizeof(int) * 9, ctx->stream));
    return 0;
}


// The common shape of the utilities: count the entries of every output row, scan, allocate the result, fill it.
// The row-count array becomes the result's row pointer; on any failure everything allocated here is released.
template <class Count, class Fill>
static int count_scan_fill(lsspg_ctx *ctx, long long rows, int n, int m, int bs, Count count, Fill fill, lsspg_dmat **out)
{
    int *cnt = nullptr;
    LSSPG_TRY(count_array(ctx, rows, &cnt));
    lsspg_dmat *M = nullptr;
    auto body = [&]() -> int {
        LSSPG_TRY(count(cnt));
        long long total = 0;
        LSSPG_TRY(dev_exclusive_scan(ctx, cnt, rows, &total));
        LSSPG_TRY(dmat_alloc(ctx, n, m, total, bs, false, &M));
        M->p = cnt;
        cnt = nullptr;
        return fill(M);
    };
    const int rc = body();
    if (rc) {
        cudaFree(cnt);
        dmat_free(M);
        return rc;
    }
    *out = M;
    return 0;
}

static inline unsigned int grid_for(long long items, int block) { return (unsigned int)std::max<long long>(1, (items + block - 1) / block); }
static inline unsigned int warp_grid(const lsspg_ctx *ctx, long long rows)
{
    return (unsigned int)std::max<long long>(1, std::min<long long>((rows + 7) / 8, (long long)ctx->num_sms * 16));
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

int lsspg_dmat_upload(lsspg_ctx *ctx, int num_rows, int num_cols, const int *hAp, const int *hAj, const double *hAx, lsspg_dmat **out)
{
    LSSPG_CHECK(ctx && out && hAp && num_rows >= 0 && num_cols >= 0, "lsspg_dmat_upload: bad argument");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const long long nnz = hAp[num_rows];
    LSSPG_CHECK(hAp[0] == 0 && nnz >= 0 && (nnz == 0 || (hAj && hAx)), "lsspg_dmat_upload: malformed CSR arrays");
    lsspg_dmat *M = nullptr;
    LSSPG_TRY(dmat_alloc(ctx, num_rows, num_cols, nnz, 1, true, &M));
    cudaMemcpyAsync(M->p, hAp, sizeof(int) * ((size_t)num_rows + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (nnz) {
        cudaMemcpyAsync(M->j, hAj, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(M->x, hAx, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { dmat_free(M); return cuda_fail(cudaGetLastError(), "lsspg_dmat_upload", __FILE__, __LINE__); }
    *out = M;
    return 0;
}

int lsspg_dmat_dims(const lsspg_dmat *M, int *num_rows, int *num_cols, long long *num_nnzs, int *blk_size)
{
    LSSPG_CHECK(M, "lsspg_dmat_dims: NULL matrix");
    if (num_rows) *num_rows = M->n;
    if (num_cols) *num_cols = M->m;
    if (num_nnzs) *num_nnzs = M->nnz;
    if (blk_size) *blk_size = M->bs;
    return 0;
}

int lsspg_dmat_download(lsspg_ctx *ctx, const lsspg_dmat *M, int *hAp, int *hAj, double *hAx)
{
    LSSPG_CHECK(ctx && M, "lsspg_dmat_download: NULL argument");
    if (hAp) LSSPG_CUDA(cudaMemcpyAsync(hAp, M->p, sizeof(int) * ((size_t)M->n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (hAj && M->nnz) LSSPG_CUDA(cudaMemcpyAsync(hAj, M->j, sizeof(int) * (size_t)M->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    if (hAx && M->nnz)
        LSSPG_CUDA(cudaMemcpyAsync(hAx, M->x, sizeof(double) * (size_t)M->nnz * M->bs * M->bs, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lsspg_dmat_destroy(lsspg_ctx *ctx, lsspg_dmat *M)
{
    if (ctx) cudaStreamSynchronize(ctx->stream);
    dmat_free(M);
    return 0;
}

/* the deep copy of lssp_solver_assemble (src/lssp.cxx:169-171), device to device */
int lsspg_dmat_copy(lsspg_ctx *ctx, const lsspg_dmat *A, lsspg_dmat **out)
{
    LSSPG_CHECK(ctx && A && out, "lsspg_dmat_copy: NULL argument");
    lsspg_dmat *M = nullptr;
    LSSPG_TRY(dmat_alloc(ctx, A->n, A->m, A->nnz, A->bs, true, &M));
    if (cudaMemcpyAsync(M->p, A->p, sizeof(int) * ((size_t)A->n + 1), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(M->j, A->j, sizeof(int) * (size_t)A->nnz, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess ||
        cudaMemcpyAsync(M->x, A->x, sizeof(double) * (size_t)A->nnz * A->bs * A->bs, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
        dmat_free(M);
        return cuda_fail(cudaGetLastError(), "lsspg_dmat_copy", __FILE__, __LINE__);
    }
    *out = M;
    return 0;
}

/* lssp_mat_csr_is_sorted (src/matrix-utils.cxx:249-279) */
int lsspg_dmat_is_sorted(lsspg_ctx *ctx, const lsspg_dmat *A, int *sorted)
{
    LSSPG_CHECK(ctx && A && sorted, "lsspg_dmat_is_sorted: NULL argument");
    int *flag = ctx->d_flags + FLAG_SETUP;
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    if (A->n > 0) LSSPG_LAUNCH(ctx, k_rows_sorted, grid_for(A->n, 256), 256, 0, A->n, A->p, A->j, flag);
    int h = 0;
    LSSPG_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    *sorted = !h;
    return 0;
}

/* lssp_mat_sort_column (src/matrix-utils.cxx:387-481), in place (the arrays are replaced) */
int lsspg_dmat_sort_columns(lsspg_ctx *ctx, lsspg_dmat *A)
{
    LSSPG_CHECK(ctx && A && A->bs == 1, "lsspg_dmat_sort_columns: needs a CSR matrix");
    if (A->n == 0 || A->nnz == 0) return 0;
    lsspg_dmat *T = nullptr;
    LSSPG_TRY(dmat_alloc(ctx, A->n, A->m, A->nnz, 1, false, &T));
    k_sort_rows<<<warp_grid(ctx, A->n), 256, 0, ctx->stream>>>(A->n, A->p, A->j, A->x, T->j, T->x);
    ctx->launches++;
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        dmat_free(T);
        return cuda_fail(e, "k_sort_rows", __FILE__, __LINE__);
    }
    std::swap(A->j, T->j);
    std::swap(A->x, T->x);
    dmat_free(T);
    return 0;
}

/* lssp_mat_adjust_zero_diag (src/matrix-utils.cxx:483-587) */
int lsspg_dmat_adjust_zero_diag(lsspg_ctx *ctx, const lsspg_dmat *A, double tol, lsspg_dmat **out)
{
    LSSPG_CHECK(ctx && A && out && A->bs == 1 && A->n == A->m, "lsspg_dmat_adjust_zero_diag: needs a square CSR matrix");
    return count_scan_fill(
        ctx, A->n, A->n, A->m, 1,
        [&](int *cnt) -> int {
            LSSPG_LAUNCH(ctx, k_diag_count, grid_for(A->n, 256), 256, 0, A->n, A->p, A->j, cnt);
            return 0;
        },
        [&](lsspg_dmat *M) -> int {
            LSSPG_LAUNCH(ctx, k_diag_fill, grid_for(A->n, 256), 256, 0, A->n, A->p, A->j, A->x, M->p, M->j, M->x, tol);
            return 0;
        },
        out);
}

/* lssp_mat_get_block_diag (src/matrix-utils.cxx:589-698) */
int lsspg_dmat_get_block_diag(lsspg_ctx *ctx, const lsspg_dmat *A, int blk_size, lsspg_dmat **out)
{
    LSSPG_CHECK(ctx && A && out && A->bs == 1 && A->n == A->m && A->n > 0 && blk_size > 0, "lsspg_dmat_get_block_diag: bad argument");
    if (blk_size >= A->n) return lsspg_dmat_copy(ctx, A, out);
    return count_scan_fill(
        ctx, A->n, A->n, A->m, 1,
        [&](int *cnt) -> int {
            LSSPG_LAUNCH(ctx, k_blockdiag_count, grid_for(A->n, 256), 256, 0, A->n, blk_size, A->p, A->j, cnt);
            return 0;
        },
        [&](lsspg_dmat *M) -> int {
            LSSPG_LAUNCH(ctx, k_blockdiag_fill, grid_for(A->n, 256), 256, 0, A->n, blk_size, A->p, A->j, A->x, M->p, M->j, M->x);
            return 0;
        },
        out);
}

/* lssp_mat_csr_to_bcsr (src/matrix-utils.cxx:62-162) */
int lsspg_dmat_to_bcsr(lsspg_ctx *ctx, const lsspg_dmat *A, int bs, lsspg_dmat **out)
{
    LSSPG_CHECK(ctx && A && out && A->bs == 1 && A->n == A->m && A->n > 0 && bs > 0, "lsspg_dmat_to_bcsr: bad argument");
    LSSPG_CHECK(A->n % bs == 0, "num_rows is not a multiple of block size: %d", bs);
    const int nb = A->n / bs;
    int *firsts = nullptr;
    LSSPG_CUDA(cudaMalloc(&firsts, sizeof(int) * (size_t)std::max<long long>(A->nnz, 1)));
    const int rc = count_scan_fill(
        ctx, nb, nb, nb, bs,
        [&](int *cnt) -> int {
            LSSPG_LAUNCH(ctx, k_bcsr_count, warp_grid(ctx, nb), 256, 0, nb, bs, A->p, A->j, cnt, firsts);
            return 0;
        },
        [&](lsspg_dmat *B) -> int {
            LSSPG_CUDA(cudaMemsetAsync(B->x, 0, sizeof(double) * (size_t)B->nnz * bs * bs, ctx->stream));
            LSSPG_LAUNCH(ctx, k_bcsr_cols, warp_grid(ctx, nb), 256, 0, nb, bs, A->p, A->j, firsts, B->p, B->j);
            LSSPG_LAUNCH(ctx, k_bcsr_vals, grid_for(A->n, 256), 256, 0, A->n, bs, A->p, A->j, A->x, B->p, B->j, B->x);
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));   // `firsts` is released below
            return 0;
        },
        out);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(firsts);
    return rc;
}

/* rows [r0, r1) of the 7-point operator on an nx x ny x nz grid (nz == 1: 5-point, 2-D); columns are global minus
 * col_shift (a rank that renumbers its own block to start at 0 passes r0); stencil[7] = values at -nx ny, -nx, -1, 0, +1,
 * +nx, +nx ny */
int lsspg_dmat_gen_stencil(lsspg_ctx *ctx, int nx, int ny, int nz, long long r0, long long r1, long long col_shift,
                           const double *stencil, lsspg_dmat **out)
{
    LSSPG_CHECK(ctx && out && stencil && nx > 0 && ny > 0 && nz > 0, "lsspg_dmat_gen_stencil: bad argument");
    const long long n = (long long)nx * ny * nz, rows = r1 - r0;
    LSSPG_CHECK(r0 >= 0 && r1 <= n && rows > 0 && rows < (1ll << 31) && n - col_shift < (1ll << 31), "lsspg_dmat_gen_stencil: row range out of bounds");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    Stencil7 s;
    for (int k = 0; k < 7; k++) s.v[k] = stencil[k];
    return count_scan_fill(
        ctx, rows, (int)rows, (int)std::min<long long>(n, (1ll << 31) - 1), 1,
        [&](int *cnt) -> int {
            LSSPG_LAUNCH(ctx, k_stencil_count, grid_for(rows, 256), 256, 0, r0, rows, nx, ny, nz, cnt);
            return 0;
        },
        [&](lsspg_dmat *M) -> int {
            LSSPG_LAUNCH(ctx, k_stencil_fill, grid_for(rows, 256), 256, 0, r0, rows, nx, ny, nz, s, col_shift, M->p, M->j, M->x);
            return 0;
        },
        out);
}

}  // extern "C"

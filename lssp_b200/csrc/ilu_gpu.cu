// ilu_gpu.cu -- ILU(k) set-up on the GPU: symbolic level-of-fill phase (src/pc-iluk.cxx:22-135,
// :279-345), block restriction (:441-446), numeric IKJ phase (:347-409) and the split into L / U (:501-532), all on a
// matrix that lives in device memory (setup_gpu.cuh).  Only the finished factors travel to the host, where the
// triangular-sweep schedules are analysed.
//
// Both factorisation phases have the dependency graph of the forward sweep: row i needs the FINISHED rows of its
// strictly lower columns.  They run as ONE persistent kernel each: a thread owns a row, executes the reference's serial
// row recurrence statement for statement, and before it uses pivot row k it waits for done[k] (the row publishes itself
// with a fence + flag).  Rows are handed out in ascending order (warp tickets), so a waiting thread only ever waits for
// rows that are finished, running or about to be issued: deadlock-free on a grid of resident CTAs.  Consecutive rows
// almost always depend on each other, so the 32 rows of a warp are kStride apart (lane l of warp ticket t owns row
// (t / kStride) 32 kStride + l kStride + t % kStride): chains run ACROSS warps and 32 chains advance per warp.
// No FMA, same operations in the same order: the factors are bit-identical to ilu_host.cpp's and hence to the
// reference's (tests/test_gpu_setup.py, tests/test_gpu_kernels.py).
#include <algorithm>
#include <vector>
#include "blas1.cuh"
#include "host_par.h"
#include "setup_gpu.cuh"

struct lsspg_factors;

extern "C" {
int lsspg_dmat_upload(lsspg_ctx *ctx, int num_rows, int num_cols, const int *hAp, const int *hAj, const double *hAx, lsspg_dmat **out);
int lsspg_dmat_destroy(lsspg_ctx *ctx, lsspg_dmat *M);
int lsspg_dmat_copy(lsspg_ctx *ctx, const lsspg_dmat *A, lsspg_dmat **out);
int lsspg_dmat_is_sorted(lsspg_ctx *ctx, const lsspg_dmat *A, int *sorted);
int lsspg_dmat_sort_columns(lsspg_ctx *ctx, lsspg_dmat *A);
int lsspg_dmat_adjust_zero_diag(lsspg_ctx *ctx, const lsspg_dmat *A, double tol, lsspg_dmat **out);
int lsspg_dmat_get_block_diag(lsspg_ctx *ctx, const lsspg_dmat *A, int blk_size, lsspg_dmat **out);
}

namespace lsspg {

lsspg_factors *factors_new(int n, size_t nnzL, size_t nnzU, int **Lp, int **Lj, double **Lx, int **Up, int **Uj, double **Ux);

constexpr double kPivotTolG = 1e-10;    // mat_zero_diag_tol,   reference src/pc.cxx:7
constexpr double kPivotValueG = 1e-3;   // mat_zero_diag_value, reference src/pc.cxx:6
constexpr int kStride = 257;            // distance between the rows of a warp (not a divisor of the usual grid offsets)
constexpr int kFacBlock = 128;

// rows of a warp ticket; all 32 lanes call (the ticket is fetched by lane 0)
__device__ __forceinline__ long long fac_next_row(unsigned int *ticket, int lane, long long n, bool *more)
{
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    const long long sc = t / kStride, w = t % kStride;
    *more = sc * 32 * kStride < n;
    return sc * 32 * kStride + (long long)lane * kStride + w;
}

__device__ __forceinline__ int ld_flag(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until row k has published itself; false when the kernel was aborted (overflow of a pattern row)
__device__ __forceinline__ bool fac_wait(const int *done, int k, const int *abort_flag)
{
    int spins = 0;
    while (ld_flag(done + k) == 0) {
        if (++spins > 16) {
            __nanosleep(spins > 256 ? 400 : 60);
            if ((spins & 63) == 0 && ld_flag(abort_flag)) return false;
        }
    }
    return true;
}

// ---- symbolic phase ---------------------------------------------------------------------------------------------------
// Row i starts as A's row (levels 0) in its slot of the pool, pc / pl [i cap ..), and stays sorted.  Pivots are the lower
// columns in ascending order -- fill lands behind the current pivot, so "the smallest lower column not used yet"
// (src/pc-iluk.cxx:62-75) is simply the next entry.  A candidate (c, lev(i,piv) + lev(piv,c) + 1) above `level` is
// ignored, an absent column is inserted, a present one has its level RAISED to the candidate's when that is larger
// (the reference's rule, :101); the diagonal is never a candidate.  dpos[i] = position of the diagonal.
__global__ void __launch_bounds__(kFacBlock) k_iluk_symbolic(int n, int level, int cap, const int *__restrict__ Ap,
                                                            const int *__restrict__ Aj, int *pc, int *pl, int *plen, int *dpos,
                                                            int *done, unsigned int *ticket, int *flags)
{
    const int lane = threadIdx.x & 31;
    int *overflow = flags + FLAG_SETUP;
    for (;;) {
        bool more;
        const long long row = fac_next_row(ticket, lane, n, &more);
        if (!more) break;
        if (row >= n) continue;
        const int i = (int)row;
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
            if (!fac_wait(done, piv, overflow)) { ok = false; break; }
            const int pn = __ldcg(plen + piv);
            const int *qc = pc + (size_t)piv * cap, *ql = pl + (size_t)piv * cap;
            int a = t + 1;
            for (int q = __ldcg(dpos + piv) + 1; q < pn; q++) {
                const int c = __ldcg(qc + q);
                const int cand = __ldcg(ql + q) + lt + 1;
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
        if (!ok || t >= len || c_[t] != i) atomicExch(overflow, (!ok) ? 1 : 2);   // 1: a row outgrew cap, 2: no diagonal
        plen[i] = len;
        dpos[i] = t;
        __threadfence();
        st_flag(done + i, 1);
    }
}

__global__ void __launch_bounds__(256) k_max_row(int n, const int *__restrict__ p, int *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = (i < n) ? p[i + 1] - p[i] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// strictly ascending columns and a stored diagonal in every row (what the phases above assume)
__global__ void __launch_bounds__(256) k_check_rows(int n, const int *__restrict__ p, const int *__restrict__ j, int *bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool diag = false;
    for (int k = p[i]; k < p[i + 1]; k++) {
        diag |= (j[k] == i);
        if (k > p[i] && j[k - 1] >= j[k]) { *bad = 1; return; }
    }
    if (!diag) *bad = 1;
}

// pattern rows out of the pool, with A's values where present and 0 on fill (src/pc-iluk.cxx:318-343)
__global__ void __launch_bounds__(256) k_pattern_rows(int n, int cap, const int *__restrict__ pc, const int *__restrict__ Ap,
                                                     const int *__restrict__ Aj, const double *__restrict__ Ax,
                                                     const int *__restrict__ Mp, int *__restrict__ Mj, double *__restrict__ Mx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
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

// ---- numeric phase: the IKJ loop of src/pc-iluk.cxx:347-409 per block of bs rows, in place --------------------------------
__global__ void __launch_bounds__(kFacBlock) k_ilu_numeric(int n, int bs, const int *__restrict__ P, const int *__restrict__ C,
                                                          double *X, double *inv, int *done, unsigned int *ticket, int *flags)
{
    const int lane = threadIdx.x & 31;
    int *abort_flag = flags + FLAG_SETUP;
    for (;;) {
        bool more;
        const long long row = fac_next_row(ticket, lane, n, &more);
        if (!more) break;
        if (row >= n) continue;
        const int i = (int)row;
        const int e = P[i + 1];
        int k = P[i];
        if (i % bs == 0) {
            // first row of a block: its leading entry is the pivot; the repaired value only enters the inverse, the
            // stored entry is left alone (as the host loop)
            const double d = __ldcg(X + k);
            __stcg(inv + i, 1. / (fabs(d) < kPivotTolG ? (d > 0 ? kPivotValueG : -kPivotValueG) : d));
        }
        else {
            bool ok = true;
            for (; ok && k < e && C[k] < i; k++) {
                const int pr = C[k];
                if (!fac_wait(done, pr, abort_flag)) { ok = false; break; }
                const double a_ik = __ldcg(X + k) * __ldcg(inv + pr);
                __stcg(X + k, a_ik);
                int pq = P[pr];
                const int pe = P[pr + 1];
                for (int q = k + 1; q < e; q++) {
                    const int c = C[q];
                    while (pq < pe && C[pq] < c) pq++;
                    if (pq < pe && C[pq] == c) {
                        const double w = __ldcg(X + pq);
                        if (w != 0.) __stcg(X + q, __ldcg(X + q) - a_ik * w);
                    }
                }
            }
            double d = kPivotValueG;
            if (k < e && C[k] == i) {
                double v = __ldcg(X + k);
                if (fabs(v) < kPivotTolG) { v = kPivotValueG; __stcg(X + k, v); }
                d = v;
            }
            __stcg(inv + i, 1. / d);
        }
        __threadfence();
        st_flag(done + i, 1);
    }
}

// ---- split: L = strict lower + unit diagonal LAST, U = diagonal FIRST + strict upper (src/pc-iluk.cxx:501-532) ----------------
__global__ void __launch_bounds__(256) k_split_count(int n, const int *__restrict__ P, const int *__restrict__ C, int *nl, int *nu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = 0, b = 0;
    for (int k = P[i]; k < P[i + 1]; k++) {
        a += (C[k] <= i);
        b += (C[k] >= i);
    }
    nl[i] = a;
    nu[i] = b;
}

__global__ void __launch_bounds__(256) k_split_fill(int n, const int *__restrict__ P, const int *__restrict__ C,
                                                   const double *__restrict__ X, const int *__restrict__ Lp, int *__restrict__ Lj,
                                                   double *__restrict__ Lx, const int *__restrict__ Up, int *__restrict__ Uj,
                                                   double *__restrict__ Ux)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ol = Lp[i], ou = Up[i];
    for (int k = P[i]; k < P[i + 1]; k++) {
        const int c = C[k];
        if (c < i) { Lj[ol] = c; Lx[ol] = X[k]; ol++; }
        else if (c == i) {
            Lj[ol] = i; Lx[ol] = 1; ol++;
            Uj[ou] = i; Ux[ou] = X[k]; ou++;
        }
        else { Uj[ou] = c; Ux[ou] = X[k]; ou++; }
    }
}

static inline unsigned int rows_grid(long long n) { return (unsigned int)std::max<long long>(1, (n + 255) / 256); }

// grid of a persistent factorisation kernel: every CTA resident, at least kStride + 32 warps in flight
template <class K>
static int fac_grid(lsspg_ctx *ctx, K kernel, int *grid)
{
    int occ = 0;
    LSSPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kFacBlock, 0));
    LSSPG_CHECK(occ >= 1, "ilu_gpu: the factorisation kernel does not fit on an SM");
    *grid = ctx->num_sms * std::min(occ, 16);
    LSSPG_CHECK((long long)*grid * (kFacBlock / 32) >= kStride + 32, "ilu_gpu: %d resident warps are too few for the row interleave", *grid * (kFacBlock / 32));
    return 0;
}

// symbolic phase on the device: M = pattern of level <= `level` with A's values (A: strictly ascending columns, diagonals stored)
static int iluk_symbolic_gpu(lsspg_ctx *ctx, const lsspg_dmat *A, int level, lsspg_dmat **out)
{
    const int n = A->n;
    int *flag = ctx->d_flags + FLAG_SETUP;
    int hmax = 0;
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LSSPG_LAUNCH(ctx, k_max_row, rows_grid(n), 256, 0, n, A->p, flag);
    LSSPG_CUDA(cudaMemcpyAsync(&hmax, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    long long cap = std::min<long long>(1024, (long long)std::max(hmax, 1) * (level + 1) * (level + 1));
    cap = std::max<long long>(32, (cap + 31) / 32 * 32);
    int grid = 0;
    LSSPG_TRY(fac_grid(ctx, k_iluk_symbolic, &grid));
    for (;; cap *= 2) {
        LSSPG_CHECK(cap <= 4096, "ilu_gpu: pattern rows longer than 4096 entries (level %d): use the host set-up", level);
        int *pc = nullptr, *pl = nullptr, *meta = nullptr;
        unsigned int *ticket = nullptr;
        const size_t pool = (size_t)n * (size_t)cap;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        LSSPG_CHECK(pool * 8 + (size_t)n * 16 < free_b, "ilu_gpu: the symbolic phase needs %zu MB of device memory (rows of up to %lld entries)", pool * 8 >> 20, cap);
        int rc = 0, hflag = 0;
        lsspg_dmat *M = nullptr;
        auto body = [&]() -> int {
            LSSPG_CUDA(cudaMalloc(&pc, sizeof(int) * pool));
            LSSPG_CUDA(cudaMalloc(&pl, sizeof(int) * pool));
            LSSPG_CUDA(cudaMalloc(&meta, sizeof(int) * ((size_t)n * 3 + 1 + 8 + 4)));   // plen (-> row pointer, + slack), dpos, done, ticket
            int *plen = meta, *dpos = meta + n + 1 + 8, *done = dpos + n;
            ticket = reinterpret_cast<unsigned int *>(done + n);
            LSSPG_CUDA(cudaMemsetAsync(meta, 0, sizeof(int) * ((size_t)n * 3 + 1 + 8 + 4), ctx->stream));
            LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
            LSSPG_LAUNCH(ctx, k_iluk_symbolic, grid, kFacBlock, 0, n, level, (int)cap, A->p, A->j, pc, pl, plen, dpos, done, ticket, ctx->d_flags);
            LSSPG_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            if (hflag) return 0;
            long long total = 0;
            int *Mp = nullptr;
            LSSPG_CUDA(cudaMalloc(&Mp, sizeof(int) * ((size_t)n + 1 + 8)));
            LSSPG_CUDA(cudaMemcpyAsync(Mp, plen, sizeof(int) * ((size_t)n + 1 + 8), cudaMemcpyDeviceToDevice, ctx->stream));
            if (dev_exclusive_scan(ctx, Mp, n, &total)) { cudaFree(Mp); return 1; }
            if (dmat_alloc(ctx, n, A->m, total, 1, false, &M)) { cudaFree(Mp); return 1; }
            M->p = Mp;
            LSSPG_LAUNCH(ctx, k_pattern_rows, rows_grid(n), 256, 0, n, (int)cap, pc, A->p, A->j, A->x, M->p, M->j, M->x);
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            return 0;
        };
        rc = body();
        cudaFree(pc); cudaFree(pl); cudaFree(meta);
        LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        if (rc) { dmat_free(M); return rc; }
        LSSPG_CHECK(hflag != 2, "ilu_gpu: a row without a stored diagonal reached the symbolic phase");
        if (!hflag) { *out = M; return 0; }
    }
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

/* ILU(k) of a device-resident matrix (rows need not be sorted, diagonals need not be stored: lssp_mat_sort_column and
 * lssp_mat_adjust_zero_diag run first when needed, src/lssp.cxx:173, src/pc-iluk.cxx:573).  blk_size <= 0 or >= n: one
 * block.  The factors come back in the reference's L / U layout on the host. */
int lsspg_ilu_factor_dmat(lsspg_ctx *ctx, const lsspg_dmat *A_in, int level, int blk_size, lsspg_factors **out)
{
    LSSPG_CHECK(ctx && A_in && out && A_in->bs == 1 && A_in->n > 0 && A_in->n == A_in->m && A_in->nnz > 0, "lsspg_ilu_factor_dmat: needs a square CSR matrix on the device");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const int n = A_in->n;
    if (level < 0) level = 0;
    const int bs = (blk_size <= 0 || blk_size > n) ? n : blk_size;
    int *flag = ctx->d_flags + FLAG_SETUP;
    lsspg_dmat *A = nullptr, *M = nullptr, *B = nullptr;
    int *cl = nullptr, *cu = nullptr, *done = nullptr;
    double *inv = nullptr;
    lsspg_dmat *L = nullptr, *U = nullptr;
    auto body = [&]() -> int {
        // ingest: strictly ascending columns with the diagonal stored (ilu_host.cpp: ingest)
        int bad = 0;
        LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        LSSPG_LAUNCH(ctx, k_check_rows, rows_grid(n), 256, 0, n, A_in->p, A_in->j, flag);
        LSSPG_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
        const lsspg_dmat *Ad = A_in;
        if (bad) {
            lsspg_dmat *S = nullptr;
            LSSPG_TRY(lsspg_dmat_copy(ctx, A_in, &S));
            int rc = lsspg_dmat_sort_columns(ctx, S);
            if (!rc) rc = lsspg_dmat_adjust_zero_diag(ctx, S, kPivotTolG, &A);
            dmat_free(S);
            LSSPG_TRY(rc);
            LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
            LSSPG_LAUNCH(ctx, k_check_rows, rows_grid(n), 256, 0, n, A->p, A->j, flag);
            LSSPG_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            LSSPG_CHECK(!bad, "lsspg_ilu_factor_dmat: rows with repeated columns are not supported on the device (use lsspg_ilu_factor)");
            Ad = A;
        }
        // symbolic phase, then the block restriction (src/pc-iluk.cxx:432-446: in this order)
        const lsspg_dmat *Mfull = Ad;
        if (level > 0) {
            LSSPG_TRY(iluk_symbolic_gpu(ctx, Ad, level, &M));
            Mfull = M;
        }
        if (bs < n) LSSPG_TRY(lsspg_dmat_get_block_diag(ctx, Mfull, bs, &B));
        else LSSPG_TRY(lsspg_dmat_copy(ctx, Mfull, &B));
        // numeric phase, in place on B
        int grid = 0;
        LSSPG_TRY(fac_grid(ctx, k_ilu_numeric, &grid));
        LSSPG_CUDA(cudaMalloc(&done, sizeof(int) * ((size_t)n + 4)));
        LSSPG_CUDA(cudaMalloc(&inv, sizeof(double) * (size_t)n));
        LSSPG_CUDA(cudaMemsetAsync(done, 0, sizeof(int) * ((size_t)n + 4), ctx->stream));
        LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        LSSPG_LAUNCH(ctx, k_ilu_numeric, grid, kFacBlock, 0, n, bs, B->p, B->j, B->x, inv, done, reinterpret_cast<unsigned int *>(done + n), ctx->d_flags);
        // split
        LSSPG_CUDA(cudaMalloc(&cl, sizeof(int) * ((size_t)n + 1 + 8)));
        LSSPG_CUDA(cudaMalloc(&cu, sizeof(int) * ((size_t)n + 1 + 8)));
        LSSPG_CUDA(cudaMemsetAsync(cl + n, 0, sizeof(int) * 9, ctx->stream));
        LSSPG_CUDA(cudaMemsetAsync(cu + n, 0, sizeof(int) * 9, ctx->stream));
        LSSPG_LAUNCH(ctx, k_split_count, rows_grid(n), 256, 0, n, B->p, B->j, cl, cu);
        long long tl = 0, tu = 0;
        LSSPG_TRY(dev_exclusive_scan(ctx, cl, n, &tl));
        LSSPG_TRY(dev_exclusive_scan(ctx, cu, n, &tu));
        LSSPG_TRY(dmat_alloc(ctx, n, n, tl, 1, false, &L));
        LSSPG_TRY(dmat_alloc(ctx, n, n, tu, 1, false, &U));
        L->p = cl; cl = nullptr;
        U->p = cu; cu = nullptr;
        LSSPG_LAUNCH(ctx, k_split_fill, rows_grid(n), 256, 0, n, B->p, B->j, B->x, L->p, L->j, L->x, U->p, U->j, U->x);
        int *Lp, *Lj, *Up, *Uj;
        double *Lx, *Ux;
        lsspg_factors *F = factors_new(n, (size_t)tl, (size_t)tu, &Lp, &Lj, &Lx, &Up, &Uj, &Ux);
        LSSPG_CUDA(cudaMemcpyAsync(Lp, L->p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(Lj, L->j, sizeof(int) * (size_t)tl, cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(Lx, L->x, sizeof(double) * (size_t)tl, cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(Up, U->p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(Uj, U->j, sizeof(int) * (size_t)tu, cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(Ux, U->x, sizeof(double) * (size_t)tu, cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
        *out = F;
        return 0;
    };
    const int rc = body();
    cudaStreamSynchronize(ctx->stream);
    cudaFree(cl); cudaFree(cu); cudaFree(done); cudaFree(inv);
    dmat_free(A); dmat_free(M); dmat_free(B); dmat_free(L); dmat_free(U);
    return rc;
}

/* host arrays in, host factors out: upload + lsspg_ilu_factor_dmat (the entry point api.ilu_factor(ctx=...) and the
 * facade's lssp_pc_iluk_assemble use) */
int lsspg_ilu_factor_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx, int level,
                            int blk_size, lsspg_factors **out)
{
    LSSPG_CHECK(ctx && out && hAp && hAj && hAx && n > 0, "lsspg_ilu_factor_device: bad argument");
    lsspg_dmat *A = nullptr;
    LSSPG_TRY(lsspg_dmat_upload(ctx, n, n, hAp, hAj, hAx, &A));
    const int rc = lsspg_ilu_factor_dmat(ctx, A, level, blk_size, out);
    lsspg_dmat_destroy(ctx, A);
    return rc;
}

}  // extern "C"

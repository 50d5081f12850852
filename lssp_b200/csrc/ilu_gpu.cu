// ilu_gpu.cu -- ILU(k) and ILUT set-up on the GPU: symbolic level-of-fill phase
// (src/pc-iluk.cxx:22-135, :279-345), block restriction (:441-446), numeric IKJ phase (:347-409), the dual-threshold ILUT
// recurrence (src/pc-ilut.cxx:51-286) and the split into L / U (:501-532), all on a matrix that lives in device memory
// (setup_gpu.cuh).  Only the finished factors travel to the host, where the triangular-sweep schedules are analysed.
//
// The factorisation phases have the dependency graph of the forward sweep: row i needs the FINISHED rows of its strictly
// lower columns.  Each runs as ONE persistent kernel: a thread owns a row, executes the reference's serial row
// recurrence statement for statement (ilu_rows.cuh -- the same functions the CPU replay at the end of this file runs),
// and before it uses pivot row k it waits for done[k] (the row publishes itself with a fence + flag).  Rows are handed
// out in ascending order (warp tickets), so a waiting thread only ever waits for rows that are finished, running or about
// to be issued: deadlock-free on a grid of resident CTAs.  Consecutive rows almost always depend on each other, so the
// 32 rows of a warp are kStride apart (lane l of warp ticket t owns row (t / kStride) 32 kStride + l kStride +
// t % kStride): chains run ACROSS warps and 32 chains advance per warp.  A watchdog turns a dependency that is never
// published into an error instead of a hung device.
// No FMA, same operations in the same order: the factors are bit-identical to ilu_host.cpp's and hence to the
// reference's (tests/test_gpu_setup.py on the device, tests/test_ilu_rows.py for the row functions on the CPU).
#include <time.h>
#include <algorithm>
#include <vector>
#include "blas1.cuh"
#include "host_par.h"
#include "ilu_rows.cuh"
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

// LSSPG_SETUP_PROF=1: phase times of the device set-up on stderr (each phase is synchronised first)
static double gprof_now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static bool gprof_on()
{
    static const bool on = getenv("LSSPG_SETUP_PROF") && atoi(getenv("LSSPG_SETUP_PROF")) != 0;
    return on;
}
#define GPROF_T0 double gprof_t = gprof_now()
#define GPROF(ctx, what)                                                              \
    do {                                                                              \
        if (gprof_on()) {                                                             \
            cudaStreamSynchronize((ctx)->stream);                                     \
            const double t_ = gprof_now();                                            \
            fprintf(stderr, "[setup/gpu] %-26s %.3f s\n", what, t_ - gprof_t);        \
            gprof_t = t_;                                                             \
        }                                                                             \
    } while (0)

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
// wait until row k has published itself; false when the kernel was aborted (a pattern row outgrew its pool slot, or
// the watchdog fired: no row waits longer than a fraction of a second on a healthy run, so 20 s without progress mean
// a dependency that will never be published -- the kernel then drains instead of hanging the device)
__device__ __forceinline__ bool fac_wait(const int *done, int k, int *abort_flag)
{
    int spins = 0;
    unsigned long long t0 = 0;
    while (ld_flag(done + k) == 0) {
        if (++spins > 16) {
            __nanosleep(spins > 256 ? 400 : 60);
            if ((spins & 63) == 0) {
                if (ld_flag(abort_flag)) return false;
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 20000000000ull) {
                    atomicExch(abort_flag, 3);
                    return false;
                }
            }
        }
    }
    return true;
}

struct FacWaitDev {
    const int *done;
    int *abort_flag;
    __device__ __forceinline__ bool operator()(int k) const { return fac_wait(done, k, abort_flag); }
};
struct FacWaitNone {   // host replay: rows run in ascending order
    bool operator()(int) const { return true; }
};

// ---- symbolic phase (row recurrence: ilu_rows.cuh) --------------------------------------------------------------------------
__global__ void __launch_bounds__(kFacBlock) k_iluk_symbolic(int n, int level, int cap, const int *__restrict__ Ap,
                                                            const int *__restrict__ Aj, int *pc, int *pl, int *plen, int *dpos,
                                                            int *done, unsigned int *ticket, int *flags)
{
    const int lane = threadIdx.x & 31;
    int *overflow = flags + FLAG_SETUP;
    const FacWaitDev wait{done, overflow};
    for (;;) {
        bool more;
        const long long row = fac_next_row(ticket, lane, n, &more);
        if (!more) break;
        if (row >= n) continue;
        const int i = (int)row;
        const int st = iluk_symbolic_row(i, level, cap, Ap, Aj, pc, pl, plen, dpos, wait);
        if (st == 1) atomicCAS(overflow, 0, 1);      // 1: a row outgrew cap (or the kernel is draining)
        else if (st == 2) atomicExch(overflow, 2);   // 2: no diagonal
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
        if (j[k] < 0 || j[k] >= n) { *bad = 2; return; }          // out of range: nothing below may run
        if (k > p[i] && j[k - 1] >= j[k]) atomicCAS(bad, 0, 1);
    }
    if (!diag) atomicCAS(bad, 0, 1);
}

__global__ void __launch_bounds__(256) k_pattern_rows(int n, int cap, const int *__restrict__ pc, const int *__restrict__ Ap,
                                                     const int *__restrict__ Aj, const double *__restrict__ Ax,
                                                     const int *__restrict__ Mp, int *__restrict__ Mj, double *__restrict__ Mx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) iluk_pattern_row(i, cap, pc, Ap, Aj, Ax, Mp, Mj, Mx);
}

// ---- numeric phase (row recurrence: ilu_rows.cuh) ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kFacBlock) k_ilu_numeric(int n, int bs, const int *__restrict__ P, const int *__restrict__ C,
                                                          double *X, double *inv, int *done, unsigned int *ticket, int *flags)
{
    const int lane = threadIdx.x & 31;
    const FacWaitDev wait{done, flags + FLAG_SETUP};
    for (;;) {
        bool more;
        const long long row = fac_next_row(ticket, lane, n, &more);
        if (!more) break;
        if (row >= n) continue;
        const int i = (int)row;
        ilu_numeric_row(i, bs, P, C, X, inv, wait);   // (an aborted wait has raised the flag already)
        __threadfence();
        st_flag(done + i, 1);
    }
}

// ---- split (ilu_rows.cuh) ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_split_count(int n, const int *__restrict__ P, const int *__restrict__ C, int *nl, int *nu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_count_row(i, C + P[i], P[i + 1] - P[i], nl, nu);
}

__global__ void __launch_bounds__(256) k_split_fill(int n, const int *__restrict__ P, const int *__restrict__ C,
                                                   const double *__restrict__ X, const int *__restrict__ Lp, int *__restrict__ Lj,
                                                   double *__restrict__ Lx, const int *__restrict__ Up, int *__restrict__ Uj,
                                                   double *__restrict__ Ux)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_fill_row(i, C + P[i], X + P[i], P[i + 1] - P[i], Lp, Lj, Lx, Up, Uj, Ux);
}

// ---- ILUT (row recurrence: ilu_rows.cuh) ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFacBlock) k_ilut_rows(int n, int bs, int p, double tau, const int *__restrict__ Bp,
                                                        const int *__restrict__ Bj, const double *__restrict__ Bx, int rcap,
                                                        int *rc, double *rv, int *rlen, double *diag, int wcap, int *wj,
                                                        double *wx, IlutSlot *tabs, int hmask, int *done, unsigned int *ticket,
                                                        int *flags)
{
    const int lane = threadIdx.x & 31;
    int *abort_flag = flags + FLAG_SETUP;
    const FacWaitDev wait{done, abort_flag};
    const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int *jwl = wj + gt * 2 * wcap, *jwu = jwl + wcap;
    double *wl = wx + gt * 2 * wcap, *wu = wl + wcap;
    IlutSlot *tab = tabs + gt * ((size_t)hmask + 1);
    for (;;) {
        bool more;
        const long long row = fac_next_row(ticket, lane, n, &more);
        if (!more) break;
        if (row >= n) continue;
        const int i = (int)row;
        if (!ilut_row(i, bs, p, tau, Bp, Bj, Bx, rcap, rc, rv, rlen, diag, wcap, jwl, jwu, wl, wu, tab, hmask, wait))
            atomicCAS(abort_flag, 0, 1);
        __threadfence();
        st_flag(done + i, 1);
    }
}

// split of the pool rows (stored order kept: the sweeps add in this order, src/pc-ilut.cxx:253-274)
__global__ void __launch_bounds__(256) k_pool_split_count(int n, int rcap, const int *__restrict__ rc, const int *__restrict__ rlen,
                                                         int *nl, int *nu)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_count_row(i, rc + (size_t)i * rcap, rlen[i], nl, nu);
}

__global__ void __launch_bounds__(256) k_pool_split_fill(int n, int rcap, const int *__restrict__ rc, const double *__restrict__ rv,
                                                        const int *__restrict__ rlen, const int *__restrict__ Lp,
                                                        int *__restrict__ Lj, double *__restrict__ Lx, const int *__restrict__ Up,
                                                        int *__restrict__ Uj, double *__restrict__ Ux)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) split_fill_row(i, rc + (size_t)i * rcap, rv + (size_t)i * rcap, rlen[i], Lp, Lj, Lx, Up, Uj, Ux);
}

static inline unsigned int rows_grid(long long n) { return (unsigned int)std::max<long long>(1, (n + 255) / 256); }

// grid of a persistent factorisation kernel: every CTA resident, at least kStride + 32 warps in flight (the row
// interleave needs them, see the header), no more CTAs than there are warp tickets
template <class K>
static int fac_grid(lsspg_ctx *ctx, K kernel, long long n, int *grid)
{
    int occ = 0;
    LSSPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kFacBlock, 0));
    LSSPG_CHECK(occ >= 1, "ilu_gpu: the factorisation kernel does not fit on an SM");
    const long long resident = (long long)ctx->num_sms * std::min(occ, 16), warps_per_cta = kFacBlock / 32;
    const long long tickets = ((n + 32ll * kStride - 1) / (32ll * kStride)) * kStride;
    const long long need = std::max<long long>((kStride + 32 + warps_per_cta - 1) / warps_per_cta, (tickets + warps_per_cta - 1) / warps_per_cta);
    *grid = (int)std::min(resident, need);
    LSSPG_CHECK((long long)*grid * warps_per_cta >= kStride + 32, "ilu_gpu: %lld resident warps are too few for the row interleave", (long long)*grid * warps_per_cta);
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
    LSSPG_TRY(fac_grid(ctx, k_iluk_symbolic, n, &grid));
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
        LSSPG_CHECK(hflag != 3, "ilu_gpu: the symbolic phase made no progress for 20 s (watchdog)");
        if (!hflag) { *out = M; return 0; }
    }
}


// download a device L / U pair into a host factor object
static int factors_download(lsspg_ctx *ctx, int n, const lsspg_dmat *L, const lsspg_dmat *U, lsspg_factors **out)
{
    int *Lp, *Lj, *Up, *Uj;
    double *Lx, *Ux;
    lsspg_factors *F = factors_new(n, (size_t)L->nnz, (size_t)U->nnz, &Lp, &Lj, &Lx, &Up, &Uj, &Ux);
    LSSPG_CUDA(cudaMemcpyAsync(Lp, L->p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(Lj, L->j, sizeof(int) * (size_t)L->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(Lx, L->x, sizeof(double) * (size_t)L->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(Up, U->p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(Uj, U->j, sizeof(int) * (size_t)U->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(Ux, U->x, sizeof(double) * (size_t)U->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = F;
    return 0;
}

// ingest of ilu_host.cpp on the device: strictly ascending columns with the diagonal stored.  *owned receives a repaired
// copy when A_in needed one (else NULL and A_in itself is used).
static int ingest_gpu(lsspg_ctx *ctx, const lsspg_dmat *A_in, lsspg_dmat **owned, const char *who)
{
    const int n = A_in->n;
    int *flag = ctx->d_flags + FLAG_SETUP;
    int bad = 0;
    *owned = nullptr;
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LSSPG_LAUNCH(ctx, k_check_rows, rows_grid(n), 256, 0, n, A_in->p, A_in->j, flag);
    LSSPG_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!bad) return 0;
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LSSPG_CHECK(bad != 2, "%s: column index out of range", who);
    lsspg_dmat *S = nullptr;
    LSSPG_TRY(lsspg_dmat_copy(ctx, A_in, &S));
    int rc = lsspg_dmat_sort_columns(ctx, S);
    if (!rc) rc = lsspg_dmat_adjust_zero_diag(ctx, S, kPivotTolG, owned);
    dmat_free(S);
    LSSPG_TRY(rc);
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LSSPG_LAUNCH(ctx, k_check_rows, rows_grid(n), 256, 0, n, (*owned)->p, (*owned)->j, flag);
    LSSPG_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LSSPG_CHECK(!bad, "%s: rows with repeated columns are not supported on the device (use lsspg_ilu_factor)", who);
    return 0;
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
        GPROF_T0;
        LSSPG_TRY(ingest_gpu(ctx, A_in, &A, "lsspg_ilu_factor_dmat"));
        const lsspg_dmat *Ad = A ? A : A_in;
        GPROF(ctx, "ingest");
        // symbolic phase, then the block restriction (src/pc-iluk.cxx:432-446: in this order)
        const lsspg_dmat *Mfull = Ad;
        if (level > 0) {
            LSSPG_TRY(iluk_symbolic_gpu(ctx, Ad, level, &M));
            Mfull = M;
        }
        GPROF(ctx, "symbolic");
        if (bs < n) LSSPG_TRY(lsspg_dmat_get_block_diag(ctx, Mfull, bs, &B));
        else LSSPG_TRY(lsspg_dmat_copy(ctx, Mfull, &B));
        GPROF(ctx, "block restriction / copy");
        // numeric phase, in place on B
        int grid = 0;
        LSSPG_TRY(fac_grid(ctx, k_ilu_numeric, n, &grid));
        LSSPG_CUDA(cudaMalloc(&done, sizeof(int) * ((size_t)n + 4)));
        LSSPG_CUDA(cudaMalloc(&inv, sizeof(double) * (size_t)n));
        LSSPG_CUDA(cudaMemsetAsync(done, 0, sizeof(int) * ((size_t)n + 4), ctx->stream));
        LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        LSSPG_LAUNCH(ctx, k_ilu_numeric, grid, kFacBlock, 0, n, bs, B->p, B->j, B->x, inv, done, reinterpret_cast<unsigned int *>(done + n), ctx->d_flags);
        {
            int hflag = 0;
            LSSPG_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
            LSSPG_CHECK(!hflag, "lsspg_ilu_factor_dmat: the numeric phase made no progress for 20 s (watchdog)");
        }
        GPROF(ctx, "numeric");
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
        GPROF(ctx, "split");
        LSSPG_TRY(factors_download(ctx, n, L, U, out));
        GPROF(ctx, "download of L, U");
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

/* ILUT(p, tol) of a device-resident matrix (src/pc-ilut.cxx:51-286, :429-456): p <= 0 -> ceil(nnz / n) (:436-438), tol < 0 ->
 * 1e-3 (:440-442); blk_size as lsspg_ilu_factor_dmat.  Rows keep the reference's unsorted storage order. */
int lsspg_ilut_factor_dmat(lsspg_ctx *ctx, const lsspg_dmat *A_in, int p, double tol, int blk_size, lsspg_factors **out)
{
    LSSPG_CHECK(ctx && A_in && out && A_in->bs == 1 && A_in->n > 0 && A_in->n == A_in->m && A_in->nnz > 0, "lsspg_ilut_factor_dmat: needs a square CSR matrix on the device");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const int n = A_in->n;
    if (p <= 0) p = (int)((A_in->nnz + n - 1) / n);
    if (tol < 0) tol = 1e-3;
    const int bs = (blk_size <= 0 || blk_size > n) ? n : blk_size;
    int *flag = ctx->d_flags + FLAG_SETUP;
    lsspg_dmat *A = nullptr, *B = nullptr, *L = nullptr, *U = nullptr;
    int *rc_ = nullptr, *meta = nullptr, *wj = nullptr, *cl = nullptr, *cu = nullptr;
    double *rv = nullptr, *diag = nullptr, *wx = nullptr;
    IlutSlot *tabs = nullptr;
    auto body = [&]() -> int {
        GPROF_T0;
        LSSPG_TRY(ingest_gpu(ctx, A_in, &A, "lsspg_ilut_factor_dmat"));
        const lsspg_dmat *Ad = A ? A : A_in;
        if (bs < n) LSSPG_TRY(lsspg_dmat_get_block_diag(ctx, Ad, bs, &B));
        const lsspg_dmat *Bd = B ? B : Ad;
        int hmax = 0;
        LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        LSSPG_LAUNCH(ctx, k_max_row, rows_grid(n), 256, 0, n, Bd->p, flag);
        LSSPG_CUDA(cudaMemcpyAsync(&hmax, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
        const int rcap = std::max(2 * p + 1, hmax);
        int grid = 0;
        LSSPG_TRY(fac_grid(ctx, k_ilut_rows, n, &grid));
        const size_t pool = (size_t)n * rcap;
        LSSPG_CUDA(cudaMalloc(&rc_, sizeof(int) * pool));
        LSSPG_CUDA(cudaMalloc(&rv, sizeof(double) * pool));
        LSSPG_CUDA(cudaMalloc(&diag, sizeof(double) * (size_t)n));
        LSSPG_CUDA(cudaMalloc(&meta, sizeof(int) * ((size_t)n * 2 + 4)));   // rlen, done, ticket
        int *rlen = meta, *done = meta + n;
        for (int wcap = std::max(64, 4 * (hmax + p));; wcap *= 2) {
            LSSPG_CHECK(wcap <= 65536, "lsspg_ilut_factor_dmat: work rows longer than 65536 entries: use the host set-up");
            int hsize = 256;
            while (hsize < 4 * wcap) hsize *= 2;
            // scratch per thread: two work arrays of wcap (column, value) and the column map; fewer CTAs when it does not fit
            const size_t per_thread = (size_t)2 * wcap * 12 + (size_t)hsize * sizeof(IlutSlot);
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            int g = grid;
            while (g > 1 && (size_t)g * kFacBlock * per_thread > free_b / 2) g = (g + 1) / 2;
            LSSPG_CHECK((long long)g * (kFacBlock / 32) >= kStride + 32 && (size_t)g * kFacBlock * per_thread <= free_b / 2,
                        "lsspg_ilut_factor_dmat: work rows of %d entries need more device memory than is free: use the host set-up", wcap);
            const size_t threads = (size_t)g * kFacBlock;
            LSSPG_CUDA(cudaMalloc(&wj, sizeof(int) * threads * 2 * wcap));
            LSSPG_CUDA(cudaMalloc(&wx, sizeof(double) * threads * 2 * wcap));
            LSSPG_CUDA(cudaMalloc(&tabs, sizeof(IlutSlot) * threads * hsize));
            LSSPG_CUDA(cudaMemsetAsync(tabs, 0xff, sizeof(IlutSlot) * threads * hsize, ctx->stream));   // stamp -1: empty
            LSSPG_CUDA(cudaMemsetAsync(meta, 0, sizeof(int) * ((size_t)n * 2 + 4), ctx->stream));
            LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
            LSSPG_LAUNCH(ctx, k_ilut_rows, g, kFacBlock, 0, n, bs, p, tol, Bd->p, Bd->j, Bd->x, rcap, rc_, rv, rlen, diag, wcap, wj, wx, tabs,
                         hsize - 1, done, reinterpret_cast<unsigned int *>(done + n), ctx->d_flags);
            int hflag = 0;
            LSSPG_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            cudaFree(wj); cudaFree(wx); cudaFree(tabs);
            wj = nullptr; wx = nullptr; tabs = nullptr;
            LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
            LSSPG_CHECK(hflag != 3, "lsspg_ilut_factor_dmat: no progress for 20 s (watchdog)");
            if (!hflag) break;
        }
        GPROF(ctx, "ilut rows");
        LSSPG_CUDA(cudaMalloc(&cl, sizeof(int) * ((size_t)n + 1 + 8)));
        LSSPG_CUDA(cudaMalloc(&cu, sizeof(int) * ((size_t)n + 1 + 8)));
        LSSPG_CUDA(cudaMemsetAsync(cl + n, 0, sizeof(int) * 9, ctx->stream));
        LSSPG_CUDA(cudaMemsetAsync(cu + n, 0, sizeof(int) * 9, ctx->stream));
        LSSPG_LAUNCH(ctx, k_pool_split_count, rows_grid(n), 256, 0, n, rcap, rc_, rlen, cl, cu);
        long long tl = 0, tu = 0;
        LSSPG_TRY(dev_exclusive_scan(ctx, cl, n, &tl));
        LSSPG_TRY(dev_exclusive_scan(ctx, cu, n, &tu));
        LSSPG_TRY(dmat_alloc(ctx, n, n, tl, 1, false, &L));
        LSSPG_TRY(dmat_alloc(ctx, n, n, tu, 1, false, &U));
        L->p = cl; cl = nullptr;
        U->p = cu; cu = nullptr;
        LSSPG_LAUNCH(ctx, k_pool_split_fill, rows_grid(n), 256, 0, n, rcap, rc_, rv, rlen, L->p, L->j, L->x, U->p, U->j, U->x);
        GPROF(ctx, "split");
        LSSPG_TRY(factors_download(ctx, n, L, U, out));
        GPROF(ctx, "download of L, U");
        return 0;
    };
    const int rc = body();
    cudaStreamSynchronize(ctx->stream);
    cudaFree(rc_); cudaFree(rv); cudaFree(diag); cudaFree(meta); cudaFree(wj); cudaFree(wx); cudaFree(tabs); cudaFree(cl); cudaFree(cu);
    dmat_free(A); dmat_free(B); dmat_free(L); dmat_free(U);
    return rc;
}

int lsspg_ilut_factor_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx, int p, double tol,
                             int blk_size, lsspg_factors **out)
{
    LSSPG_CHECK(ctx && out && hAp && hAj && hAx && n > 0, "lsspg_ilut_factor_device: bad argument");
    lsspg_dmat *A = nullptr;
    LSSPG_TRY(lsspg_dmat_upload(ctx, n, n, hAp, hAj, hAx, &A));
    const int rc = lsspg_ilut_factor_dmat(ctx, A, p, tol, blk_size, out);
    lsspg_dmat_destroy(ctx, A);
    return rc;
}

/* CPU replay of the device factorisations for the test-suite (never called by a product path): the SAME row functions
 * the kernels run (ilu_rows.cuh), rows in ascending order, no device.  kind 0: ILU(k) (symbolic with a growing pool,
 * pattern with A's values, block restriction, numeric, split), kind 1: ILUT.  The input must have strictly ascending
 * columns and stored diagonals (the device path sorts and repairs first); otherwise *applicable = 0.  The factors must
 * equal lsspg_ilu_factor's bit for bit. */
int lsspg_debug_ilu_gpu_replay_host(int kind, int n, const int *Ap, const int *Aj, const double *Ax, int level, int p, double tol,
                                    int blk_size, int *applicable, lsspg_factors **out)
{
    LSSPG_CHECK(n > 0 && Ap && Aj && Ax && out && applicable, "lsspg_debug_ilu_gpu_replay_host: bad argument");
    *applicable = 0;
    for (int i = 0; i < n; i++) {
        bool diag = false;
        for (int k = Ap[i]; k < Ap[i + 1]; k++) {
            if (Aj[k] < 0 || Aj[k] >= n || (k > Ap[i] && Aj[k - 1] >= Aj[k])) return 0;
            diag |= (Aj[k] == i);
        }
        if (!diag) return 0;
    }
    *applicable = 1;
    const int bs = (blk_size <= 0 || blk_size > n) ? n : blk_size;
    const FacWaitNone wait;
    // block restriction of a sorted matrix with diagonals (k_blockdiag_count / k_blockdiag_fill)
    auto restrict_blocks = [&](const std::vector<int> &Mp, const std::vector<int> &Mj, const std::vector<double> &Mx, std::vector<int> &Bp,
                               std::vector<int> &Bj, std::vector<double> &Bx) {
        Bp.assign((size_t)n + 1, 0);
        Bj.clear();
        Bx.clear();
        for (int i = 0; i < n; i++) {
            const int lo = (i / bs) * bs, hi = std::min(n, lo + bs);
            const size_t before = Bj.size();
            for (int k = Mp[i]; k < Mp[i + 1]; k++)
                if (Mj[k] >= lo && Mj[k] < hi) { Bj.push_back(Mj[k]); Bx.push_back(Mx[k]); }
            if (Bj.size() == before) { Bj.push_back(i); Bx.push_back(1.0); }
            Bp[i + 1] = (int)Bj.size();
        }
    };
    std::vector<int> nl((size_t)n + 1, 0), nu((size_t)n + 1, 0);
    auto scan = [&](std::vector<int> &v) {
        int run = 0;
        for (int i = 0; i < n; i++) { const int c = v[i]; v[i] = run; run += c; }
        v[n] = run;
    };
    int *Lp, *Lj, *Up, *Uj;
    double *Lx, *Ux;
    if (kind == 0) {
        if (level < 0) level = 0;
        std::vector<int> Mp(Ap, Ap + n + 1), Mj(Aj, Aj + Ap[n]);
        std::vector<double> Mx(Ax, Ax + Ap[n]);
        if (level > 0) {
            int hmax = 1;
            for (int i = 0; i < n; i++) hmax = std::max(hmax, Ap[i + 1] - Ap[i]);
            long long cap = std::max<long long>(32, (std::min<long long>(1024, (long long)hmax * (level + 1) * (level + 1)) + 31) / 32 * 32);
            std::vector<int> pc, pl, plen((size_t)n + 1), dpos((size_t)n);
            for (;; cap *= 2) {
                LSSPG_CHECK(cap <= 4096, "replay: pattern rows longer than 4096 entries");
                pc.assign((size_t)n * cap, 0);
                pl.assign((size_t)n * cap, 0);
                int st = 0;
                for (int i = 0; i < n && st == 0; i++)
                    st = iluk_symbolic_row(i, level, (int)cap, Ap, Aj, pc.data(), pl.data(), plen.data(), dpos.data(), wait);
                LSSPG_CHECK(st != 2, "replay: a row without a stored diagonal");
                if (st == 0) break;
            }
            Mp.assign(plen.begin(), plen.end());
            scan(Mp);
            Mj.resize((size_t)Mp[n]);
            Mx.resize((size_t)Mp[n]);
            for (int i = 0; i < n; i++) iluk_pattern_row(i, (int)cap, pc.data(), Ap, Aj, Ax, Mp.data(), Mj.data(), Mx.data());
        }
        if (bs < n) {
            std::vector<int> Bp, Bj;
            std::vector<double> Bx;
            restrict_blocks(Mp, Mj, Mx, Bp, Bj, Bx);
            Mp.swap(Bp); Mj.swap(Bj); Mx.swap(Bx);
        }
        std::vector<double> inv((size_t)n);
        for (int i = 0; i < n; i++) ilu_numeric_row(i, bs, Mp.data(), Mj.data(), Mx.data(), inv.data(), wait);
        for (int i = 0; i < n; i++) split_count_row(i, Mj.data() + Mp[i], Mp[i + 1] - Mp[i], nl.data(), nu.data());
        scan(nl);
        scan(nu);
        lsspg_factors *F = factors_new(n, (size_t)nl[n], (size_t)nu[n], &Lp, &Lj, &Lx, &Up, &Uj, &Ux);
        std::copy(nl.begin(), nl.end(), Lp);
        std::copy(nu.begin(), nu.end(), Up);
        for (int i = 0; i < n; i++) split_fill_row(i, Mj.data() + Mp[i], Mx.data() + Mp[i], Mp[i + 1] - Mp[i], Lp, Lj, Lx, Up, Uj, Ux);
        *out = F;
        return 0;
    }
    // ILUT
    if (p <= 0) p = (int)(((long long)Ap[n] + n - 1) / n);
    if (tol < 0) tol = 1e-3;
    std::vector<int> Bp(Ap, Ap + n + 1), Bj(Aj, Aj + Ap[n]);
    std::vector<double> Bx(Ax, Ax + Ap[n]);
    if (bs < n) {
        std::vector<int> Mp(Bp), Mj(Bj);
        std::vector<double> Mx(Bx);
        restrict_blocks(Mp, Mj, Mx, Bp, Bj, Bx);
    }
    int hmax = 1;
    for (int i = 0; i < n; i++) hmax = std::max(hmax, Bp[i + 1] - Bp[i]);
    const int rcap = std::max(2 * p + 1, hmax);
    std::vector<int> rc((size_t)n * rcap), rlen((size_t)n);
    std::vector<double> rv((size_t)n * rcap), diag((size_t)n);
    for (int wcap = std::max(64, 4 * (hmax + p));; wcap *= 2) {
        LSSPG_CHECK(wcap <= 65536, "replay: work rows longer than 65536 entries");
        int hsize = 256;
        while (hsize < 4 * wcap) hsize *= 2;
        std::vector<int> wj((size_t)2 * wcap);
        std::vector<double> wx((size_t)2 * wcap);
        std::vector<IlutSlot> tab((size_t)hsize, IlutSlot{-1, -1, -1});
        bool ok = true;
        for (int i = 0; i < n && ok; i++)
            ok = ilut_row(i, bs, p, tol, Bp.data(), Bj.data(), Bx.data(), rcap, rc.data(), rv.data(), rlen.data(), diag.data(), wcap, wj.data(),
                          wj.data() + wcap, wx.data(), wx.data() + wcap, tab.data(), hsize - 1, wait);
        if (ok) break;
    }
    for (int i = 0; i < n; i++) split_count_row(i, rc.data() + (size_t)i * rcap, rlen[i], nl.data(), nu.data());
    scan(nl);
    scan(nu);
    lsspg_factors *F = factors_new(n, (size_t)nl[n], (size_t)nu[n], &Lp, &Lj, &Lx, &Up, &Uj, &Ux);
    std::copy(nl.begin(), nl.end(), Lp);
    std::copy(nu.begin(), nu.end(), Up);
    for (int i = 0; i < n; i++)
        split_fill_row(i, rc.data() + (size_t)i * rcap, rv.data() + (size_t)i * rcap, rlen[i], Lp, Lj, Lx, Up, Uj, Ux);
    *out = F;
    return 0;
}

}  // extern "C"

// amg_gpu.cu -- the per-row phases of the SX-AMG-style set-up on the GPU: strong couplings, direct
// interpolation with truncation, restriction R = P^T and the two Galerkin products of every level run as count / scan /
// fill kernels, one row per thread, over the row functions of amg_rows.cuh -- the functions the CPU replay
// (lsspg_debug_amg_setup_replay_host) pins against the host set-up.  The level operators stay on the device from one
// level to the next; what the serial phases need (strength graph for the Ruge-Stueben C/F splitting, which stays on the
// host) and what the hierarchy object keeps (A, P, R of every level: the smoother layouts are built from them) is copied
// back.  The set-up loop itself is amg_setup_with (amg_host.cpp), shared with the host and replay variants, so the
// hierarchy is the same array by array (tests/test_gpu_amg_setup.py).
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <vector>
#include "amg_host.h"
#include "amg_rows.cuh"
#include "blas1.cuh"
#include "setup_gpu.cuh"

extern "C" {
int lsspg_dmat_upload(lsspg_ctx *ctx, int num_rows, int num_cols, const int *hAp, const int *hAj, const double *hAx, lsspg_dmat **out);
}

namespace lsspg {

static inline unsigned int rows_grid(long long n) { return (unsigned int)std::max<long long>(1, (n + 255) / 256); }

// ---- kernels: one row per thread ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_amg_strong_count(int n, const int *__restrict__ Ap, const int *__restrict__ Aj,
                                                         const double *__restrict__ Ax, double st, double mrs, int *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = amg_strong_row(i, Ap, Aj, Ax, st, mrs, nullptr);
}
__global__ void __launch_bounds__(256) k_amg_strong_fill(int n, const int *__restrict__ Ap, const int *__restrict__ Aj,
                                                        const double *__restrict__ Ax, double st, double mrs,
                                                        const int *__restrict__ Sp, int *__restrict__ Sj)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) amg_strong_row(i, Ap, Aj, Ax, st, mrs, Sj + Sp[i]);
}

__global__ void __launch_bounds__(256) k_amg_cflag(int n, const int *__restrict__ cf, int *flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (cf[i] == kAmgCPT);
}
__global__ void __launch_bounds__(256) k_amg_cpoint(int n, const int *__restrict__ cf, const int *__restrict__ cidx, int *cpoint)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && cf[i] == kAmgCPT) cpoint[cidx[i]] = i;
}

__global__ void __launch_bounds__(256) k_amg_interp_count(int n, const int *__restrict__ Ap, const int *__restrict__ Aj,
                                                         const double *__restrict__ Ax, const int *__restrict__ Sp,
                                                         const int *__restrict__ Sj, const int *__restrict__ cf,
                                                         const int *__restrict__ cidx, double trunc, int *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = amg_interp_row(i, Ap, Aj, Ax, Sp, Sj, cf, cidx, trunc, nullptr, nullptr);
}
__global__ void __launch_bounds__(256) k_amg_interp_fill(int n, const int *__restrict__ Ap, const int *__restrict__ Aj,
                                                        const double *__restrict__ Ax, const int *__restrict__ Sp,
                                                        const int *__restrict__ Sj, const int *__restrict__ cf,
                                                        const int *__restrict__ cidx, double trunc, const int *__restrict__ Pp,
                                                        int *__restrict__ Pj, double *__restrict__ Px)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) amg_interp_row(i, Ap, Aj, Ax, Sp, Sj, cf, cidx, trunc, Pj + Pp[i], Px + Pp[i]);
}

__global__ void __launch_bounds__(256) k_amg_restrict_count(int nc, const int *__restrict__ cpoint, const int *__restrict__ Tp,
                                                           const int *__restrict__ Tj, const int *__restrict__ Pp,
                                                           const int *__restrict__ Pj, const double *__restrict__ Px, int *cnt)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nc) cnt[c] = amg_restrict_row(c, cpoint[c], Tp, Tj, Pp, Pj, Px, nullptr, nullptr);
}
__global__ void __launch_bounds__(256) k_amg_restrict_fill(int nc, const int *__restrict__ cpoint, const int *__restrict__ Tp,
                                                          const int *__restrict__ Tj, const int *__restrict__ Pp,
                                                          const int *__restrict__ Pj, const double *__restrict__ Px,
                                                          const int *__restrict__ Rp, int *__restrict__ Rj, double *__restrict__ Rx)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nc) amg_restrict_row(c, cpoint[c], Tp, Tj, Pp, Pj, Px, Rj + Rp[c], Rx + Rp[c]);
}

__global__ void __launch_bounds__(256) k_amg_bound(int n, const int *__restrict__ Ap, const int *__restrict__ Aj,
                                                  const int *__restrict__ Bp, int *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = (i < n) ? amg_spgemm_bound(i, Ap, Aj, Bp) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// Galerkin products: a fixed number of threads, each with its own accumulator table and column list, strides over the
// rows (pass 0: count, pass 1: fill -- different stamps, so that the sums of the count pass are not seen by the fill pass)
__global__ void __launch_bounds__(128) k_amg_spgemm(int nrows, int pass, const int *__restrict__ Ap, const int *__restrict__ Aj,
                                                   const double *__restrict__ Ax, const int *__restrict__ Bp,
                                                   const int *__restrict__ Bj, const double *__restrict__ Bx, AmgSlot *tabs,
                                                   int hmask, int *colbuf, int cap, int *Cp, int *Cj, double *Cx, int *overflow)
{
    const size_t gt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    AmgSlot *tab = tabs + gt * ((size_t)hmask + 1);
    int *cols = colbuf + gt * (size_t)cap;
    for (long long i = (long long)gt; i < nrows; i += (long long)gridDim.x * blockDim.x) {
        if (pass == 0) {
            const int c = amg_spgemm_row((int)i, 2 * (int)i, Ap, Aj, Ax, Bp, Bj, Bx, tab, hmask, cols, cap, nullptr, nullptr);
            if (c < 0) { *overflow = 1; Cp[i] = 0; }
            else Cp[i] = c;
        }
        else amg_spgemm_row((int)i, 2 * (int)i + 1, Ap, Aj, Ax, Bp, Bj, Bx, tab, hmask, cols, cap, Cj + Cp[i], Cx + Cp[i]);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
namespace {

// count -> scan -> allocate -> fill; the count array becomes the row pointer of the result (no values when !with_x ... the
// value array is allocated anyway: dmat_alloc, a strength graph simply leaves it untouched)
template <class Count, class Fill>
int csf(lsspg_ctx *ctx, long long rows, int n, int m, Count count, Fill fill, lsspg_dmat **out)
{
    int *cnt = nullptr;
    LSSPG_CUDA(cudaMalloc(&cnt, sizeof(int) * ((size_t)rows + 1 + 8)));
    lsspg_dmat *M = nullptr;
    auto body = [&]() -> int {
        LSSPG_CUDA(cudaMemsetAsync(cnt + rows, 0, sizeof(int) * 9, ctx->stream));
        LSSPG_TRY(count(cnt));
        long long total = 0;
        LSSPG_TRY(dev_exclusive_scan(ctx, cnt, rows, &total));
        LSSPG_TRY(dmat_alloc(ctx, n, m, total, 1, false, &M));
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

int download(lsspg_ctx *ctx, const lsspg_dmat *M, std::vector<int> &p, std::vector<int> &j, std::vector<double> *x)
{
    p.resize((size_t)M->n + 1);
    j.resize((size_t)M->nnz);
    LSSPG_CUDA(cudaMemcpyAsync(p.data(), M->p, sizeof(int) * ((size_t)M->n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    if (M->nnz) LSSPG_CUDA(cudaMemcpyAsync(j.data(), M->j, sizeof(int) * (size_t)M->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    if (x) {
        x->resize((size_t)M->nnz);
        if (M->nnz) LSSPG_CUDA(cudaMemcpyAsync(x->data(), M->x, sizeof(double) * (size_t)M->nnz, cudaMemcpyDeviceToHost, ctx->stream));
    }
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int upload_ints(lsspg_ctx *ctx, const std::vector<int> &h, int **d)
{
    LSSPG_CUDA(cudaMalloc(d, sizeof(int) * std::max<size_t>(h.size(), 1)));
    if (!h.empty()) LSSPG_CUDA(cudaMemcpyAsync(*d, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// C = A B on the device (amg_host.cpp: spgemm)
int spgemm_gpu(lsspg_ctx *ctx, const lsspg_dmat *A, const lsspg_dmat *B, lsspg_dmat **out)
{
    const int nrows = A->n;
    int *flag = ctx->d_flags + FLAG_SETUP;
    int bound = 0;
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    LSSPG_LAUNCH(ctx, k_amg_bound, rows_grid(nrows), 256, 0, nrows, A->p, A->j, B->p, flag);
    LSSPG_CUDA(cudaMemcpyAsync(&bound, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    bound = std::max(bound, 1);
    int hsize = 64;
    while (hsize < 2 * bound) hsize *= 2;
    // threads: enough to fill the device, fewer when the per-thread scratch (table + column list) would not fit
    const size_t per_thread = (size_t)hsize * sizeof(AmgSlot) + (size_t)bound * sizeof(int);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    long long ctas = std::min<long long>((long long)ctx->num_sms * 8, ((long long)nrows + 127) / 128);
    while (ctas > 1 && (size_t)ctas * 128 * per_thread > free_b / 4) ctas = (ctas + 1) / 2;
    LSSPG_CHECK((size_t)ctas * 128 * per_thread <= free_b / 2, "amg set-up on the device: rows of up to %d products need more memory than is free", bound);
    const size_t threads = (size_t)ctas * 128;
    AmgSlot *tabs = nullptr;
    int *cols = nullptr;
    LSSPG_CUDA(cudaMalloc(&tabs, sizeof(AmgSlot) * threads * hsize));
    if (cudaMalloc(&cols, sizeof(int) * threads * bound) != cudaSuccess) { cudaFree(tabs); return cuda_fail(cudaGetLastError(), "spgemm_gpu", __FILE__, __LINE__); }
    int hflag = 0;
    const int rc = csf(
        ctx, nrows, nrows, B->m,
        [&](int *cnt) -> int {
            LSSPG_CUDA(cudaMemsetAsync(tabs, 0xff, sizeof(AmgSlot) * threads * hsize, ctx->stream));   // stamp -1: empty
            LSSPG_LAUNCH(ctx, k_amg_spgemm, (unsigned int)ctas, 128, 0, nrows, 0, A->p, A->j, A->x, B->p, B->j, B->x, tabs, hsize - 1, cols, bound, cnt,
                         nullptr, nullptr, flag);
            LSSPG_CUDA(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            LSSPG_CHECK(!hflag, "amg set-up on the device: accumulator overflow in a Galerkin product");
            return 0;
        },
        [&](lsspg_dmat *C) -> int {
            LSSPG_LAUNCH(ctx, k_amg_spgemm, (unsigned int)ctas, 128, 0, nrows, 1, A->p, A->j, A->x, B->p, B->j, B->x, tabs, hsize - 1, cols, bound, C->p,
                         C->j, C->x, flag);
            LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
            return 0;
        },
        out);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(tabs);
    cudaFree(cols);
    return rc;
}

struct DevicePhases : AmgPhases {
    lsspg_ctx *ctx;
    lsspg_dmat *dA = nullptr, *dS = nullptr, *dP = nullptr, *dR = nullptr;
    int *d_cf = nullptr, *d_cidx = nullptr, *d_cpoint = nullptr;
    explicit DevicePhases(lsspg_ctx *c) : ctx(c) {}
    ~DevicePhases() override
    {
        cudaStreamSynchronize(ctx->stream);
        drop_level();
        dmat_free(dA);
    }
    void drop_level()
    {
        dmat_free(dS); dmat_free(dP); dmat_free(dR);
        dS = dP = dR = nullptr;
        cudaFree(d_cf); cudaFree(d_cidx); cudaFree(d_cpoint);
        d_cf = d_cidx = d_cpoint = nullptr;
    }
    int begin(const AmgLevelHost &L0) override
    {
        LSSPG_CUDA(cudaSetDevice(ctx->device));
        return lsspg_dmat_upload(ctx, L0.n, L0.n, L0.Ap.data(), L0.Aj.data(), L0.Ax.data(), &dA);
    }
    int strength(const AmgLevelHost &L, const lsspg_amg_pars &pr, AmgGraph &S) override
    {
        const int n = L.n;
        drop_level();
        LSSPG_TRY(csf(
            ctx, n, n, n,
            [&](int *cnt) -> int {
                LSSPG_LAUNCH(ctx, k_amg_strong_count, rows_grid(n), 256, 0, n, dA->p, dA->j, dA->x, pr.strong_threshold, pr.max_row_sum, cnt);
                return 0;
            },
            [&](lsspg_dmat *G) -> int {
                LSSPG_LAUNCH(ctx, k_amg_strong_fill, rows_grid(n), 256, 0, n, dA->p, dA->j, dA->x, pr.strong_threshold, pr.max_row_sum, G->p, G->j);
                return 0;
            },
            &dS));
        return download(ctx, dS, S.p, S.j, nullptr);   // the C/F splitting is the host's
    }
    int interpolation(AmgLevelHost &L, const AmgGraph &, const AmgGraph &, const lsspg_amg_pars &pr) override
    {
        const int n = L.n;
        LSSPG_TRY(upload_ints(ctx, L.cf, &d_cf));
        LSSPG_CUDA(cudaMalloc(&d_cidx, sizeof(int) * ((size_t)n + 1 + 8)));
        LSSPG_LAUNCH(ctx, k_amg_cflag, rows_grid(n), 256, 0, n, d_cf, d_cidx);
        long long nc = 0;
        LSSPG_TRY(dev_exclusive_scan(ctx, d_cidx, n, &nc));
        L.nc = (int)nc;
        LSSPG_TRY(csf(
            ctx, n, n, (int)nc,
            [&](int *cnt) -> int {
                LSSPG_LAUNCH(ctx, k_amg_interp_count, rows_grid(n), 256, 0, n, dA->p, dA->j, dA->x, dS->p, dS->j, d_cf, d_cidx, pr.trunc_threshold, cnt);
                return 0;
            },
            [&](lsspg_dmat *P) -> int {
                LSSPG_LAUNCH(ctx, k_amg_interp_fill, rows_grid(n), 256, 0, n, dA->p, dA->j, dA->x, dS->p, dS->j, d_cf, d_cidx, pr.trunc_threshold, P->p,
                             P->j, P->x);
                return 0;
            },
            &dP));
        return download(ctx, dP, L.Pp, L.Pj, &L.Px);
    }
    int restriction(AmgLevelHost &L, const AmgGraph &T) override
    {
        const int n = L.n, nc = L.nc;
        int *dTp = nullptr, *dTj = nullptr;
        const bool prof = getenv("LSSPG_SETUP_PROF") && atoi(getenv("LSSPG_SETUP_PROF")) > 1;
        auto stamp = [&](const char *what) {
            if (!prof) return;
            cudaStreamSynchronize(ctx->stream);
            timespec ts;
            clock_gettime(CLOCK_MONOTONIC, &ts);
            static double last = 0.0;
            const double t = ts.tv_sec + 1e-9 * ts.tv_nsec;
            fprintf(stderr, "[amg restriction] %-18s %.3f s\n", what, last > 0.0 ? t - last : 0.0);
            last = t;
        };
        stamp("start");
        int rc = upload_ints(ctx, T.p, &dTp);
        if (!rc) rc = upload_ints(ctx, T.j, &dTj);
        stamp("upload of T");
        if (!rc && cudaMalloc(&d_cpoint, sizeof(int) * std::max<size_t>((size_t)nc, 1)) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "amg restriction", __FILE__, __LINE__);
        if (!rc) {
            auto body = [&]() -> int {
                LSSPG_LAUNCH(ctx, k_amg_cpoint, rows_grid(n), 256, 0, n, d_cf, d_cidx, d_cpoint);
                stamp("cpoint");
                LSSPG_TRY(csf(
                    ctx, nc, nc, n,
                    [&](int *cnt) -> int {
                        LSSPG_LAUNCH(ctx, k_amg_restrict_count, rows_grid(nc), 256, 0, nc, d_cpoint, dTp, dTj, dP->p, dP->j, dP->x, cnt);
                        stamp("count kernel");
                        return 0;
                    },
                    [&](lsspg_dmat *R) -> int {
                        LSSPG_LAUNCH(ctx, k_amg_restrict_fill, rows_grid(nc), 256, 0, nc, d_cpoint, dTp, dTj, dP->p, dP->j, dP->x, R->p, R->j, R->x);
                        return 0;
                    },
                    &dR));
                stamp("scan + fill");
                const int rd = download(ctx, dR, L.Rp, L.Rj, &L.Rx);
                stamp("download of R");
                return rd;
            };
            rc = body();
        }
        cudaStreamSynchronize(ctx->stream);
        cudaFree(dTp);
        cudaFree(dTj);
        return rc;
    }
    int galerkin(const AmgLevelHost &, AmgLevelHost &C) override
    {
        lsspg_dmat *dT = nullptr, *dC = nullptr;
        int rc = spgemm_gpu(ctx, dA, dP, &dT);
        if (!rc) rc = spgemm_gpu(ctx, dR, dT, &dC);
        dmat_free(dT);
        if (rc) { dmat_free(dC); return rc; }
        rc = download(ctx, dC, C.Ap, C.Aj, &C.Ax);
        dmat_free(dA);
        dA = dC;   // the next level's operator stays on the device
        return rc;
    }
    const char *name() const override { return "device"; }
};

}  // namespace
}  // namespace lsspg

using namespace lsspg;

extern "C" {

/* The set-up of lsspg_amg_setup_host with its per-row phases on the GPU; the hierarchy object is the same (and is used the
 * same way: lsspg_pc_create_amg). */
int lsspg_amg_setup_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                           lsspg_amg_host **out)
{
    LSSPG_CHECK(ctx, "lsspg_amg_setup_device: NULL context");
    DevicePhases ph(ctx);
    return amg_setup_with(ph, n, hAp, hAj, hAx, pars, out);
}

}  // extern "C"

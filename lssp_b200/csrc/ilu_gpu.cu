// ilu_gpu.cu -- numeric phase of ILU(k) on the GPU (replaces the IKJ loop of
// src/pc-iluk.cxx:347-409).  The symbolic phase (fill pattern, diagonal repair, block restriction)
// stays on the host (ilu_host.cpp: ilu_prepare) and so does the split into L / U.
//
// Row i of the IKJ factorisation needs the FINISHED rows of its strictly lower columns: the same
// dependency graph as the forward sweep with L.  Rows are grouped by that level and every level is
// one launch, one thread per row.  A thread walks its lower entries in ascending order, scales by
// the pivot's inverse, and subtracts a_ik * a_kj from the entries of its row that exist in row k
// (two-pointer merge of the sorted column lists) -- the same operations in the same order as the
// host loop, without FMA: the factors are bit-identical to ilu_host.cpp's and hence to the
// reference's (tests/test_gpu_kernels.py).
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "host_par.h"

struct lsspg_factors;

namespace lsspg {
int ilu_prepare(int n, const int *Ap, const int *Aj, const double *Ax, int level, int bs, IVec &Mp, IVec &Mj, DVec &Mx);
lsspg_factors *ilu_split(int n, IVec &Mp, IVec &Mj, DVec &Mx);

constexpr double kPivotTolG = 1e-10;    // mat_zero_diag_tol,   reference src/pc.cxx:7
constexpr double kPivotValueG = 1e-3;   // mat_zero_diag_value, reference src/pc.cxx:6

__global__ void __launch_bounds__(kBlock) k_ilu_level(const int *__restrict__ rows, int cnt, const int *__restrict__ P,
                                                      const int *__restrict__ C, double *X, double *inv, int bs)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt) return;
    const int i = rows[t];
    const int e = P[i + 1];
    int k = P[i];
    if (i % bs == 0) {
        // first row of a block: its leading entry is the pivot; the repaired value only enters the
        // inverse, the stored entry is left alone (as the host loop)
        const double d = X[k];
        inv[i] = 1. / (fabs(d) < kPivotTolG ? (d > 0 ? kPivotValueG : -kPivotValueG) : d);
        return;
    }
    for (; C[k] < i; k++) {
        const int pr = C[k];
        const double a_ik = X[k] * inv[pr];
        X[k] = a_ik;
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
    double d = kPivotValueG;
    if (C[k] == i) {
        if (fabs(X[k]) < kPivotTolG) X[k] = kPivotValueG;
        d = X[k];
    }
    inv[i] = 1. / d;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" int lsspg_ilu_factor_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx, int level,
                                       int blk_size, lsspg_factors **out)
{
    LSSPG_CHECK(ctx && out && hAp && hAj && hAx && n > 0, "lsspg_ilu_factor_device: bad argument");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    if (level < 0) level = 0;
    const int bs = (blk_size <= 0 || blk_size > n) ? n : blk_size;
    IVec Mp, Mj;
    DVec Mx;
    LSSPG_TRY(ilu_prepare(n, hAp, hAj, hAx, level, bs, Mp, Mj, Mx));
    const size_t nnz = Mj.size();
    // dependency level of every row (its strictly lower columns), rows grouped by level
    std::vector<int> lev(n, 0);
    int nlev = 0;
    for (int i = 0; i < n; i++) {
        int l = 0;
        for (int k = Mp[i]; k < Mp[i + 1] && Mj[k] < i; k++) l = std::max(l, lev[Mj[k]] + 1);
        lev[i] = l;
        nlev = std::max(nlev, l + 1);
    }
    std::vector<int> start(nlev + 1, 0), order(n);
    for (int i = 0; i < n; i++) start[lev[i] + 1]++;
    for (int l = 0; l < nlev; l++) start[l + 1] += start[l];
    {
        std::vector<int> pos(start.begin(), start.end() - 1);
        for (int i = 0; i < n; i++) order[pos[lev[i]]++] = i;
    }
    int *dP = nullptr, *dC = nullptr, *dRows = nullptr;
    double *dX = nullptr, *dInv = nullptr;
    int rc = 0;
    auto fail = [&](const char *what) {
        set_error("lsspg_ilu_factor_device: %s", what);
        rc = 1;
    };
    if (cudaMalloc(&dP, sizeof(int) * ((size_t)n + 1)) != cudaSuccess || cudaMalloc(&dC, sizeof(int) * std::max<size_t>(nnz, 1)) != cudaSuccess ||
        cudaMalloc(&dX, sizeof(double) * std::max<size_t>(nnz, 1)) != cudaSuccess || cudaMalloc(&dInv, sizeof(double) * (size_t)n) != cudaSuccess ||
        cudaMalloc(&dRows, sizeof(int) * (size_t)n) != cudaSuccess)
        fail("out of device memory");
    if (!rc) {
        cudaMemcpyAsync(dP, Mp.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(dC, Mj.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(dX, Mx.data(), sizeof(double) * nnz, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(dRows, order.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
        for (int l = 0; l < nlev && !rc; l++) {
            const int cnt = start[l + 1] - start[l];
            if (cnt == 0) continue;
            k_ilu_level<<<(cnt + kBlock - 1) / kBlock, kBlock, 0, ctx->stream>>>(dRows + start[l], cnt, dP, dC, dX, dInv, bs);
            ctx->launches++;
            if (cudaPeekAtLastError() != cudaSuccess) fail(cudaGetErrorString(cudaGetLastError()));
        }
        if (!rc && cudaMemcpyAsync(Mx.data(), dX, sizeof(double) * nnz, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) fail("download failed");
        if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) fail(cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(dP); cudaFree(dC); cudaFree(dX); cudaFree(dInv); cudaFree(dRows);
    if (rc) return rc;
    *out = ilu_split(n, Mp, Mj, Mx);
    return 0;
}

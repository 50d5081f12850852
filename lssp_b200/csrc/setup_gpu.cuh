// setup_gpu.cuh -- device-resident CSR container of the set-up path: matrices that are
// generated, sorted, repaired, restricted and factorised on the GPU without a host copy of Aj / Ax.
#pragma once
#include "common.cuh"

// CSR (bs == 1) or BCSR (bs > 1: n block rows, nnz blocks, x holds nnz * bs * bs values, blocks column-major as
// include/type-defs.h:26-37) in device memory.  Owned arrays, allocated with the slack the SpMV kernels expect.
struct lsspg_dmat {
    int n = 0, m = 0, bs = 1;
    long long nnz = 0;
    int *p = nullptr;      // [n + 1 + 8]
    int *j = nullptr;      // [nnz + 16]
    double *x = nullptr;   // [nnz * bs * bs + 16]
};

namespace lsspg {

int dmat_alloc(lsspg_ctx *ctx, int n, int m, long long nnz, int bs, bool with_p, lsspg_dmat **out);
void dmat_free(lsspg_dmat *M);
// in-place exclusive scan of d[0 .. n) (ints); d[n] receives the total, which is also returned through *total (one sync)
int dev_exclusive_scan(lsspg_ctx *ctx, int *d, long long n, long long *total);

}  // namespace lsspg

// krylov.cuh -- shared scaffolding of the Krylov drivers.
#pragma once
#include <math.h>
#include <vector>
#include "blas1.cuh"
#include "pc.cuh"
#include "spmv.cuh"

namespace lsspg {

// reference defaults, src/lssp.cxx:5-14
constexpr int kDefRestart = 50, kDefAugK = 3, kDefBgsl = 4, kDefIdrs = 4, kDefMaxit = 1000;
constexpr double kDefAtol = 1e-7, kDefRtol = 1e-7, kDefRb = 1e-7, kBreakdown = 1e-40;

// Work-vector arena of one solve: device vectors of n doubles, zero-initialised
// (the reference mallocs without zeroing; drivers that read before writing are
// compared against the zero-initialising oracle build).
struct Workspace {
    lsspg_ctx *ctx;
    long long n;
    std::vector<double *> ptrs;
    Workspace(lsspg_ctx *c, long long n_) : ctx(c), n(n_) {}
    std::vector<size_t> sizes;
    ~Workspace()
    {
        cudaStreamSynchronize(ctx->stream);
        for (size_t i = 0; i < ptrs.size(); i++) ctx->pool.push_back(std::make_pair(ptrs[i], sizes[i]));
    }
    double *vec()
    {
        const size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
        double *p = nullptr;
        size_t got = 0;
        for (size_t i = 0; i < ctx->pool.size(); i++) {
            if (ctx->pool[i].second >= bytes && ctx->pool[i].second <= 2 * bytes) {
                p = ctx->pool[i].first;
                got = ctx->pool[i].second;
                ctx->pool.erase(ctx->pool.begin() + i);
                break;
            }
        }
        if (!p) {
            if (cudaMalloc(&p, bytes + 64) != cudaSuccess) return nullptr;
            got = bytes;
        }
        cudaMemsetAsync(p, 0, bytes, ctx->stream);
        ptrs.push_back(p);
        sizes.push_back(got);
        return p;
    }
};

struct KrylovArgs {
    lsspg_ctx *ctx;
    const lsspg_csr *A;
    lsspg_pc *pc;
    const double *b;
    double *x;
    int n;      // owned rows (length of every reduction / update)
    int nvec;   // allocated vector length: owned rows + ghost columns of a distributed matrix
    // resolved options
    double tol_abs, tol_rel, tol_rb;
    int maxit, restart, aug_k, bgsl, idrs, verb;
    double *hist;
    int hist_len;
    lsspg_solve_info *info;
};

// tol = max(rtol*||r0||, atol, rbtol*||b||)  (e.g. src/solver-cg.cxx:56-70)
inline double stop_tolerance(const KrylovArgs &k, double res0, double bnorm)
{
    double tol = k.tol_rel * res0;
    const double tol_rb = k.tol_rb * bnorm;
    if (tol < k.tol_abs) tol = k.tol_abs;
    if (tol < tol_rb) tol = tol_rb;
    return tol;
}

inline void record(const KrylovArgs &k, int it, double res)
{
    if (k.hist && it < k.hist_len) {
        k.hist[it] = res;
        if (k.info->hist_used < it + 1) k.info->hist_used = it + 1;
    }
}

int krylov_cg(KrylovArgs &k);
int krylov_bicgstab(KrylovArgs &k);
int krylov_gmres(KrylovArgs &k);
int krylov_idrs(KrylovArgs &k, const lsspg_solver_opts *raw);

}  // namespace lsspg

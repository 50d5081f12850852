// amg.cu -- the SX-AMG-style cycle on sm_100a (replaces sx_solver_amg_solve / sx_solver_amg as
// called from src/pc-sxamg.cxx:42-73 and src/solver-sxamg.cxx:25-99; libsxamg is not in the
// reference tree -- DESIGN.md "AMG" is the specification, parity with libsxamg is UNPINNED).
//
// One V-cycle, all on the device, no host synchronisation:
//   level l < last:  pre_iter Gauss-Seidel sweeps; r = b - A x (spmv.cu, row-sequential);
//                    b_{l+1} = R r; x_{l+1} = 0;  ...;  x += P x_{l+1}; post_iter sweeps
//   last level:      x = A^-1 b with the dense inverse (row-sequential dot products), or
//                    coarse_sweeps natural-order sweeps when it is too large for that
//
// Gauss-Seidel sweep (gs_sweep_kernel): sequential semantics -- rows are visited C points then
// F points (pre) / F then C (post), ascending inside a block, every row using the newest values
// -- reproduced exactly by a dependency schedule: the sweep reads x_old and writes x_new; an
// operand comes from x_new when its row precedes the current one in the sweep order and from
// x_old otherwise; x_new is pre-filled with a sentinel NaN that doubles as the ready flag (as the
// triangular solves, tri.cu).  Warps draw 32-row slices in dependency order from an atomic ticket
// counter, so a slice only ever waits for slices held by resident warps.  Each row subtracts its
// products one by one in column order and divides by the diagonal: bit-identical to the serial
// sweep.  With C/F ordering the fine level of a 7-point operator has ONE dependency level per
// block (the C points are mutually independent and so are the F points): the sweep is a plain
// streaming kernel bounded by HBM, 12 nnz + 4 n (perm) + 8 n (diag) + 8 n (b) + 16 n (x_new fill
// and store) + 8 n (x_old) bytes.
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "amg_host.h"
#include "blas1.cuh"
#include "pc.cuh"
#include "spmv.cuh"

namespace lsspg {

constexpr unsigned long long kGsSentinelBits = 0xFFF8DEADBEEF0002ull;

struct GsDev {
    int n = 0, num_slices = 0, slices_c = 0, depth = 0, mode = 0, levels_c = 0;
    std::vector<int> level_ptr;   // host copy: slices of every dependency level (per-level launches)
    int *perm = nullptr;
    double *diag = nullptr;
    int *slice_ptr = nullptr;
    int *col = nullptr;
    double *val = nullptr;
    unsigned int *counter = nullptr;
    double bytes = 0.0;
};

struct AmgLevelDev {
    int n = 0, nc = 0;
    lsspg_csr *A = nullptr, *P = nullptr, *R = nullptr;
    bool own_A = true;
    GsDev gs;
    double *xa = nullptr, *xb = nullptr, *b = nullptr, *r = nullptr;
};

}  // namespace lsspg

struct lsspg_amg {
    std::vector<lsspg::AmgLevelDev> lv;
    lsspg_amg_pars pars;
    bool coarse_dense = false;
    double *d_inv = nullptr;   // column-major inverse of the last operator
};

namespace lsspg {

struct GsArgs {
    const int *perm;
    const double *diag;
    const int *slice_ptr;
    const int *col;
    const double *val;
    unsigned int *counter;
    int num_slices, slices_c, post, exact;
    const double *xold;
    double *xnew;
    const double *rhs;
    const int *stop;
    int *err;
};

__device__ __forceinline__ double gs_ld_relaxed(const double *p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ bool gs_pending(double v)
{
    return (unsigned long long)__double_as_longlong(v) == kGsSentinelBits;
}

// Slice schedule, one row per lane.  kGsChunk entries of the row are in flight at a time: 8 for
// the shallow (streaming) schedules; 32 for deep ones, where a whole row must be loaded BEFORE the
// wait for its operands starts -- otherwise every further chunk adds two memory latencies to the
// critical path of its dependency level.
template <int MINB, int kGsChunk>
__global__ void __launch_bounds__(kBlock, MINB) gs_sweep_kernel(const GsArgs a)
{
    if (a.stop && *a.stop) return;
    const int lane = threadIdx.x & 31;
    const unsigned int total = (unsigned int)a.num_slices + gridDim.x * (blockDim.x >> 5);
    const int nf = a.num_slices - a.slices_c;
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicInc(a.counter, total - 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned int)a.num_slices) break;
        const int s = a.post ? ((int)t < nf ? a.slices_c + (int)t : (int)t - nf) : (int)t;
        const bool row_c = s < a.slices_c;
        const bool first_c = !a.post;
        const long long slot = (long long)s * 32 + lane;
        const int row = __ldg(a.perm + slot);
        const double dg = __ldg(a.diag + slot);
        const int p0 = __ldg(a.slice_ptr + s);
        const int w = __ldg(a.slice_ptr + s + 1) - p0;
        double r = (row >= 0) ? __ldg(a.rhs + row) : 0.0;
        const long long base = (long long)p0 * 32 + lane;
        for (int k0 = 0; k0 < w; k0 += kGsChunk) {
            int c[kGsChunk];
            bool fresh[kGsChunk];
            double v[kGsChunk], xv[kGsChunk];
#pragma unroll
            for (int j = 0; j < kGsChunk; j++) {
                int enc = -1;
                v[j] = 0.0;
                if (k0 + j < w) {
                    enc = __ldg(a.col + base + (long long)(k0 + j) * 32);
                    v[j] = __ldg(a.val + base + (long long)(k0 + j) * 32);
                }
                c[j] = enc >> 2;   // -1 stays -1
                const bool col_c = enc & 1;
                fresh[j] = (enc >= 0) && ((col_c == row_c) ? ((enc & 2) != 0) : (col_c == first_c));
            }
#pragma unroll
            for (int j = 0; j < kGsChunk; j++)
                xv[j] = (c[j] < 0) ? 0.0 : (fresh[j] ? gs_ld_relaxed(a.xnew + c[j]) : __ldg(a.xold + c[j]));
            bool pending;
            int spins = 0;
            do {
                pending = false;
#pragma unroll
                for (int j = 0; j < kGsChunk; j++) {
                    if (fresh[j] && gs_pending(xv[j])) {
                        xv[j] = gs_ld_relaxed(a.xnew + c[j]);
                        pending |= gs_pending(xv[j]);
                    }
                }
                if (pending && ++spins > 64) {
                    __nanosleep(64);
                    if (spins > (1 << 21)) {   // watchdog, as tri.cu
                        *a.err = 1;
                        pending = false;
                    }
                }
            } while (pending);
#pragma unroll
            for (int j = 0; j < kGsChunk; j++)
                if (c[j] >= 0) r = r - v[j] * xv[j];
        }
        if (row >= 0) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(a.xnew + row), "d"(r / dg) : "memory");
    }
}

// Shallow schedules (a handful of dependency levels: the fine level of a stencil operator under C/F
// ordering has two, multicolour levels have about ten): ONE LAUNCH PER DEPENDENCY LEVEL.  Stream order
// then guarantees that every x_new operand of the level exists -- no sentinel fill, no polling, a lean
// streaming kernel at full occupancy; a warp takes a 32-row slice, a lane its row, products subtracted
// in column order as everywhere else.
__global__ void __launch_bounds__(kBlock) gs_level_kernel(const GsArgs a, int s_begin, int s_end, int row_c, int first_c)
{
    if (a.stop && *a.stop) return;
    const int lane = threadIdx.x & 31;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int s = s_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < s_end; s += warps) {
        const long long slot = (long long)s * 32 + lane;
        const int row = __ldg(a.perm + slot);
        const double dg = __ldg(a.diag + slot);
        const int p0 = __ldg(a.slice_ptr + s);
        const int w = __ldg(a.slice_ptr + s + 1) - p0;
        double r = (row >= 0) ? __ldg(a.rhs + row) : 0.0;
        const long long base = (long long)p0 * 32 + lane;
        for (int k0 = 0; k0 < w; k0 += 4) {
            int enc[4];
            double v[4], xv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                enc[j] = -1;
                v[j] = 0.0;
                if (k0 + j < w) {
                    enc[j] = __ldg(a.col + base + (long long)(k0 + j) * 32);
                    v[j] = __ldg(a.val + base + (long long)(k0 + j) * 32);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = enc[j] >> 2;
                const bool col_c = enc[j] & 1;
                const bool fresh = (col_c == (bool)row_c) ? ((enc[j] & 2) != 0) : (col_c == (bool)first_c);
                xv[j] = (enc[j] < 0) ? 0.0 : (fresh ? __ldcg(a.xnew + c) : __ldg(a.xold + c));
            }
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (enc[j] >= 0) r = r - v[j] * xv[j];
        }
        if (row >= 0) a.xnew[row] = r / dg;
    }
}

// Row schedule (deep schedules of wide rows, GsHost mode 1): one ticket = one row, worked on by a
// warp.  The lanes fetch the row's entries and operands side by side (one memory latency for up to
// 128 entries), form the products, and every lane then subtracts them in column order from shared
// memory -- the same sequence of IEEE operations as the serial sweep.
constexpr int kRowGroup = 128;

template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB) gs_rows_kernel(const GsArgs a)
{
    __shared__ double sprod[kBlock / 32][kRowGroup];
    if (a.stop && *a.stop) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int total = (unsigned int)a.num_slices + gridDim.x * (blockDim.x >> 5);
    const int nf = a.num_slices - a.slices_c;
    double *prod = sprod[wid];
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicInc(a.counter, total - 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned int)a.num_slices) break;
        const int p = a.post ? ((int)t < nf ? a.slices_c + (int)t : (int)t - nf) : (int)t;
        const bool row_c = p < a.slices_c;
        const bool first_c = !a.post;
        const int row = __ldg(a.perm + p);
        const double dg = __ldg(a.diag + p);
        const int beg = __ldg(a.slice_ptr + p), end = __ldg(a.slice_ptr + p + 1);
        const bool tree = !a.exact && end - beg > 64;
        double r = __ldg(a.rhs + row);
        for (int base = beg; base < end; base += kRowGroup) {
            int c[4];
            bool fresh[4];
            double v[4], xv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int k = base + q * 32 + lane;
                int enc = -1;
                v[q] = 0.0;
                if (k < end) {
                    enc = __ldg(a.col + k);
                    v[q] = __ldg(a.val + k);
                }
                c[q] = enc >> 2;
                const bool col_c = enc & 1;
                fresh[q] = (enc >= 0) && ((col_c == row_c) ? ((enc & 2) != 0) : (col_c == first_c));
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                xv[q] = (c[q] < 0) ? 0.0 : (fresh[q] ? gs_ld_relaxed(a.xnew + c[q]) : __ldg(a.xold + c[q]));
            bool pending;
            int spins = 0;
            do {
                pending = false;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (fresh[q] && gs_pending(xv[q])) {
                        xv[q] = gs_ld_relaxed(a.xnew + c[q]);
                        pending |= gs_pending(xv[q]);
                    }
                }
                if (pending && ++spins > 32) {
                    __nanosleep(32);
                    if (spins > (1 << 21)) {
                        *a.err = 1;
                        pending = false;
                    }
                }
            } while (pending);
            if (tree) {
                // rows longer than 64 entries (as the SpMV does with them, spmv.cu): the lanes add their own
                // products, a shuffle tree adds the lanes -- <= 1e-14 relative instead of bit-identical, and
                // the row no longer costs one fp64 add latency per entry.  LSSPG_OPT_SPMV_EXACT turns it off.
                double s = (v[0] * xv[0] + v[1] * xv[1]) + (v[2] * xv[2] + v[3] * xv[3]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                r = r - s;
                continue;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) prod[q * 32 + lane] = v[q] * xv[q];
            __syncwarp();
            const int cnt = min(kRowGroup, end - base);
            int j = 0;
            for (; j + 4 <= cnt; j += 4) {
                const double p0 = prod[j], p1 = prod[j + 1], p2 = prod[j + 2], p3 = prod[j + 3];
                r = r - p0;
                r = r - p1;
                r = r - p2;
                r = r - p3;
            }
            for (; j < cnt; j++) r = r - prod[j];
            __syncwarp();
        }
        if (lane == 0) asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(a.xnew + row), "d"(r / dg) : "memory");
    }
}

// x = inv * b for the last level; inv column-major so that thread i walks a coalesced column
__global__ void __launch_bounds__(128) k_dense_apply(int n, const double *__restrict__ inv, const double *__restrict__ b,
                                                     double *__restrict__ x, const int *stop)
{
    if (stop && *stop) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int j = 0; j < n; j++) s += inv[(size_t)j * n + i] * b[j];
    x[i] = s;
}

static int gs_upload(lsspg_ctx *ctx, const GsHost &G, GsDev &D)
{
    D.n = G.n;
    D.num_slices = G.num_slices;
    D.slices_c = G.slices_c;
    D.depth = G.levels_c + G.levels_f;
    D.mode = G.mode;
    D.levels_c = G.levels_c;
    D.level_ptr = G.level_ptr;
    const size_t slots = G.perm.size();
    LSSPG_CUDA(cudaMalloc(&D.perm, sizeof(int) * std::max<size_t>(slots, 1)));
    LSSPG_CUDA(cudaMalloc(&D.diag, sizeof(double) * std::max<size_t>(slots, 1)));
    LSSPG_CUDA(cudaMalloc(&D.slice_ptr, sizeof(int) * G.slice_ptr.size()));
    LSSPG_CUDA(cudaMalloc(&D.col, sizeof(int) * std::max<size_t>((size_t)G.padded_nnz, 1)));
    LSSPG_CUDA(cudaMalloc(&D.val, sizeof(double) * std::max<size_t>((size_t)G.padded_nnz, 1)));
    LSSPG_CUDA(cudaMalloc(&D.counter, 256));
    LSSPG_CUDA(cudaMemsetAsync(D.counter, 0, 256, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(D.perm, G.perm.data(), sizeof(int) * slots, cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(D.diag, G.diag.data(), sizeof(double) * slots, cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(D.slice_ptr, G.slice_ptr.data(), sizeof(int) * G.slice_ptr.size(),
                               cudaMemcpyHostToDevice, ctx->stream));
    if (G.padded_nnz) {
        LSSPG_CUDA(cudaMemcpyAsync(D.col, G.col.data(), sizeof(int) * (size_t)G.padded_nnz, cudaMemcpyHostToDevice, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(D.val, G.val.data(), sizeof(double) * (size_t)G.padded_nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    D.bytes = 12.0 * (double)G.offdiag_nnz + 44.0 * G.n;
    return 0;
}

static void gs_free(GsDev &D)
{
    cudaFree(D.perm); cudaFree(D.diag); cudaFree(D.slice_ptr); cudaFree(D.col); cudaFree(D.val); cudaFree(D.counter);
    D = GsDev();
}

// one sweep: xnew <- GS(xold)
static int gs_sweep(lsspg_ctx *ctx, const GsDev &D, int post, const double *xold, double *xnew, const double *rhs, bool guarded)
{
    if (D.n == 0) return 0;
    GsArgs a;
    a.perm = D.perm; a.diag = D.diag; a.slice_ptr = D.slice_ptr; a.col = D.col; a.val = D.val;
    a.counter = D.counter; a.num_slices = D.num_slices; a.slices_c = D.slices_c; a.post = post;
    a.xold = xold; a.xnew = xnew; a.rhs = rhs;
    a.stop = guarded ? ctx->d_flags + FLAG_STOP : nullptr;
    a.err = ctx->d_flags + FLAG_TRI_TIMEOUT;
    a.exact = ctx->opt_spmv_exact;
    static int per_level = -1;
    if (per_level < 0) per_level = getenv("LSSPG_GS_PER_LEVEL") ? atoi(getenv("LSSPG_GS_PER_LEVEL")) : 1;
    if (D.mode == 0 && D.depth <= kGsStreamDepth && per_level) {
        // one launch per dependency level, block order as the sweep visits them
        const int nlev = (int)D.level_ptr.size() - 1;
        for (int t = 0; t < nlev; t++) {
            const int nf = nlev - D.levels_c;
            const int l = post ? (t < nf ? D.levels_c + t : t - nf) : t;
            const int s0 = D.level_ptr[l], s1 = D.level_ptr[l + 1];
            if (s1 <= s0) continue;
            const int grid = std::max(1, std::min((s1 - s0 + 7) / 8, ctx->num_sms * 6));   // 40 registers: 6 CTAs / SM
            LSSPG_LAUNCH(ctx, gs_level_kernel, grid, kBlock, 0, a, s0, s1, l < D.levels_c ? 1 : 0, post ? 0 : 1);
        }
        return 0;
    }
    double sentinel;
    const unsigned long long bits = kGsSentinelBits;
    memcpy(&sentinel, &bits, sizeof(double));
    LSSPG_TRY(vec_set(ctx, D.n, xnew, sentinel, guarded));
    static int env_shallow = -1, env_deep = -1;
    if (env_shallow < 0) {
        const char *e;
        env_shallow = (e = getenv("LSSPG_GS_CTAS_PER_SM")) ? atoi(e) : 3;
        env_deep = (e = getenv("LSSPG_GS_DEEP_CTAS_PER_SM")) ? atoi(e) : 1;
        env_shallow = std::min(std::max(env_shallow, 1), 8);
        env_deep = std::min(std::max(env_deep, 1), 8);
    }
    // few dependency levels: a streaming kernel, fill the SMs; many: polling, keep residency low (tri.cu)
    a.exact = ctx->opt_spmv_exact;
    const bool shallow = D.depth <= kGsShallowDepth;
    const bool streaming = D.depth <= kGsStreamDepth;
    if (D.mode == 1) {   // a warp per row: 4 CTAs per SM unless the schedule is deep (polling)
        const int per_sm = shallow ? 4 : env_deep;
        const int grid = std::max(1, std::min((D.num_slices + 7) / 8, ctx->num_sms * per_sm));
        if (shallow) LSSPG_LAUNCH(ctx, gs_rows_kernel<4>, grid, kBlock, 0, a);
        else LSSPG_LAUNCH(ctx, gs_rows_kernel<1>, grid, kBlock, 0, a);
        return 0;
    }
    const int per_sm = streaming ? env_shallow : env_deep;
    const int grid = std::max(1, std::min((D.num_slices + 7) / 8, ctx->num_sms * per_sm));
    if (streaming) LSSPG_LAUNCH(ctx, (gs_sweep_kernel<3, 8>), grid, kBlock, 0, a);
    else LSSPG_LAUNCH(ctx, (gs_sweep_kernel<1, 32>), grid, kBlock, 0, a);
    return 0;
}

static int smooth(lsspg_ctx *ctx, AmgLevelDev &L, int post, int sweeps, double *&cur, double *&other, const double *rhs,
                  bool guarded)
{
    for (int s = 0; s < sweeps; s++) {
        LSSPG_TRY(gs_sweep(ctx, L.gs, post, cur, other, rhs, guarded));
        std::swap(cur, other);
    }
    return 0;
}

// one cycle from the initial guess in dx (src/pc-sxamg.cxx:58-64); dx and drhs have lv[0].n entries
int amg_cycle(lsspg_ctx *ctx, lsspg_amg *M, double *dx, const double *drhs, bool guarded)
{
    LSSPG_CHECK(M && dx && drhs && dx != drhs, "amg_cycle: bad operands");
    const int nl = (int)M->lv.size();
    std::vector<double *> cur(nl), other(nl);
    std::vector<const double *> rhs(nl);
    for (int l = 0; l < nl; l++) {
        cur[l] = l ? M->lv[l].xa : dx;
        other[l] = M->lv[l].xb;
        rhs[l] = l ? M->lv[l].b : drhs;
    }
    const Coef one = coef_imm(1.0), minus = coef_imm(-1.0);
    if (M->pars.zero_guess) LSSPG_TRY(vec_set(ctx, M->lv[0].n, dx, 0.0, guarded));
    for (int l = 0; l + 1 < nl; l++) {
        AmgLevelDev &L = M->lv[l];
        LSSPG_TRY(smooth(ctx, L, 0, M->pars.pre_iter, cur[l], other[l], rhs[l], guarded));
        // r = b*1 + (-1)*(A x)
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBYZ, L.A, minus, cur[l], one, rhs[l], L.r, nullptr, guarded));
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_MXY, L.R, one, L.r, coef_imm(0.0), nullptr, M->lv[l + 1].b, nullptr, guarded));
        LSSPG_TRY(vec_set(ctx, M->lv[l + 1].n, cur[l + 1], 0.0, guarded));
    }
    {
        AmgLevelDev &L = M->lv[nl - 1];
        if (M->coarse_dense) {
            // the dense solve does not read the initial guess; write into the buffer that is current
            LSSPG_LAUNCH(ctx, k_dense_apply, (L.n + 127) / 128, 128, 0, L.n, M->d_inv, rhs[nl - 1], cur[nl - 1],
                         guarded ? ctx->d_flags + FLAG_STOP : (const int *)nullptr);
        }
        else {
            LSSPG_TRY(smooth(ctx, L, 0, M->pars.coarse_sweeps, cur[nl - 1], other[nl - 1], rhs[nl - 1], guarded));
        }
    }
    for (int l = nl - 2; l >= 0; l--) {
        AmgLevelDev &L = M->lv[l];
        // x = (P x_c)*1 + x*1
        LSSPG_TRY(spmv_launch(ctx, LSSPG_MV_AMXPBY, L.P, one, cur[l + 1], one, cur[l], cur[l], nullptr, guarded));
        LSSPG_TRY(smooth(ctx, L, 1, M->pars.post_iter, cur[l], other[l], rhs[l], guarded));
    }
    if (cur[0] != dx) LSSPG_TRY(vec_copy(ctx, M->lv[0].n, dx, cur[0]));
    return 0;
}

void amg_free(lsspg_ctx *ctx, lsspg_amg *M)
{
    if (!M) return;
    for (auto &L : M->lv) {
        if (L.own_A) lsspg_csr_destroy(ctx, L.A);
        lsspg_csr_destroy(ctx, L.P);
        lsspg_csr_destroy(ctx, L.R);
        gs_free(L.gs);
        cudaFree(L.xa); cudaFree(L.xb); cudaFree(L.b); cudaFree(L.r);
    }
    cudaFree(M->d_inv);
    delete M;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

int lsspg_pc_create_amg(lsspg_ctx *ctx, const lsspg_amg_host *H, const lsspg_csr *A0, lsspg_pc **out)
{
    LSSPG_CHECK(ctx && H && out && !H->levels.empty(), "lsspg_pc_create_amg: bad argument");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const int nl = (int)H->levels.size();
    if (A0)
        LSSPG_CHECK(A0->num_rows == H->levels[0].n && A0->num_nnzs == H->levels[0].Ap[H->levels[0].n] && !A0->halo,
                    "lsspg_pc_create_amg: the shared level-0 operator does not match the hierarchy");
    lsspg_amg *M = new lsspg_amg();
    M->pars = H->pars;
    M->coarse_dense = H->coarse_dense;
    M->lv.resize(nl);
    lsspg_pc *pc = new lsspg_pc();
    pc->kind = LSSPG_PC_AMG;
    pc->n = H->levels[0].n;
    pc->amg = M;
    int rc = 0;
    double bytes = 0.0;
    for (int l = 0; l < nl && !rc; l++) {
        const AmgLevelHost &Lh = H->levels[l];
        AmgLevelDev &L = M->lv[l];
        L.n = Lh.n;
        L.nc = Lh.nc;
        const bool last = (l == nl - 1);
        if (l == 0 && A0) {
            L.A = const_cast<lsspg_csr *>(A0);
            L.own_A = false;
        }
        else if (!last || !H->coarse_dense || nl == 1)
            rc = lsspg_csr_upload(ctx, Lh.n, Lh.n, Lh.Ap.data(), Lh.Aj.data(), Lh.Ax.data(), &L.A);
        if (!rc && !last) {
            rc = lsspg_csr_upload(ctx, Lh.n, Lh.nc, Lh.Pp.data(), Lh.Pj.data(), Lh.Px.data(), &L.P);
            if (!rc) rc = lsspg_csr_upload(ctx, Lh.nc, Lh.n, Lh.Rp.data(), Lh.Rj.data(), Lh.Rx.data(), &L.R);
        }
        if (!rc && (!last || !H->coarse_dense)) {
            GsHost G;
            const bool cf_on = H->pars.cf_order && !last;
            rc = gs_build_host(Lh.n, Lh.Ap.data(), Lh.Aj.data(), Lh.Ax.data(), cf_on ? Lh.cf.data() : nullptr,
                               cf_on ? Lh.rank.data() : nullptr, G);
            if (!rc) rc = gs_upload(ctx, G, L.gs);
        }
        const size_t vb = sizeof(double) * (size_t)std::max(Lh.n, 1);
        if (!rc && cudaMalloc(&L.xb, vb) != cudaSuccess) rc = 1;
        if (!rc && l > 0 && (cudaMalloc(&L.xa, vb) != cudaSuccess || cudaMalloc(&L.b, vb) != cudaSuccess)) rc = 1;
        if (!rc && !last && cudaMalloc(&L.r, vb) != cudaSuccess) rc = 1;
        if (rc == 1 && cudaPeekAtLastError() != cudaSuccess) {
            cudaGetLastError();
            set_error("lsspg_pc_create_amg: out of device memory at level %d", l);
        }
        if (!rc) {
            const double nnzA = Lh.Ap[Lh.n];
            if (!last) {
                const double sweeps = H->pars.pre_iter + H->pars.post_iter;
                bytes += sweeps * L.gs.bytes + (12.0 * nnzA + 28.0 * Lh.n)              // residual
                         + (12.0 * Lh.Rp[Lh.nc] + 4.0 * Lh.nc + 8.0 * Lh.n + 8.0 * Lh.nc)   // restriction
                         + 8.0 * Lh.nc                                                   // x_c = 0
                         + (12.0 * Lh.Pp[Lh.n] + 4.0 * Lh.n + 8.0 * Lh.nc + 16.0 * Lh.n);   // prolongation
            }
            else bytes += H->coarse_dense ? 8.0 * Lh.n * (double)Lh.n : H->pars.coarse_sweeps * L.gs.bytes;
        }
    }
    if (!rc && H->coarse_dense) {
        const int n = H->levels[nl - 1].n;
        std::vector<double> cm((size_t)n * n);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) cm[(size_t)j * n + i] = H->coarse_inv[(size_t)i * n + j];
        if (cudaMalloc(&M->d_inv, sizeof(double) * cm.size()) != cudaSuccess) {
            set_error("lsspg_pc_create_amg: out of device memory (coarse inverse)");
            rc = 1;
        }
        else if (cudaMemcpy(M->d_inv, cm.data(), sizeof(double) * cm.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("lsspg_pc_create_amg: upload of the coarse inverse failed");
            rc = 1;
        }
    }
    if (rc) {
        lsspg_pc_destroy(ctx, pc);
        return rc;
    }
    pc->bytes = bytes;
    *out = pc;
    return 0;
}

int lsspg_amg_solve(lsspg_ctx *ctx, lsspg_pc *amg, const double *db, double *dx, double tol, int maxit, int *nits,
                    double *ares)
{
    LSSPG_CHECK(ctx && amg && amg->kind == LSSPG_PC_AMG && amg->amg && db && dx, "lsspg_amg_solve: bad argument");
    LSSPG_CHECK(ctx->comm == nullptr, "lsspg_amg_solve: the stand-alone AMG iteration is single-GPU (block-local when sharded)");
    lsspg_amg *M = amg->amg;
    AmgLevelDev &L0 = M->lv[0];
    LSSPG_CHECK(L0.A, "lsspg_amg_solve: no level-0 operator");
    const int n = L0.n;
    double *r = nullptr;
    LSSPG_CUDA(cudaMalloc(&r, sizeof(double) * (size_t)std::max(n, 1)));
    double bnorm = 0.0, res = 0.0;
    int rc = lsspg_vec_norm(ctx, n, db, &bnorm);
    int it = 0;
    const Coef one = coef_imm(1.0), minus = coef_imm(-1.0);
    if (!rc) rc = spmv_launch(ctx, LSSPG_MV_AMXPBYZ, L0.A, minus, dx, one, db, r, nullptr);
    if (!rc) rc = lsspg_vec_norm(ctx, n, r, &res);
    const double denom = bnorm > 1e-20 ? bnorm : 1e-20;
    while (!rc && it < maxit && res / denom > tol) {
        const int keep = M->pars.zero_guess;   // the iteration always continues from the current x
        M->pars.zero_guess = 0;
        rc = amg_cycle(ctx, M, dx, db, false);
        M->pars.zero_guess = keep;
        if (!rc) rc = spmv_launch(ctx, LSSPG_MV_AMXPBYZ, L0.A, minus, dx, one, db, r, nullptr);
        if (!rc) rc = lsspg_vec_norm(ctx, n, r, &res);
        it++;
        if (!rc && M->pars.verb > 0) log_printf("amg: cycle %3d, residual %.8e, relative %.8e\n", it, res, res / denom);
    }
    cudaFree(r);
    if (rc) return rc;
    if (nits) *nits = it;
    if (ares) *ares = res;
    return 0;
}

int lsspg_amg_solve_host(lsspg_ctx *ctx, lsspg_pc *amg, const double *hb, double *hx, double tol, int maxit, int *nits,
                         double *ares)
{
    LSSPG_CHECK(amg && hb && hx, "lsspg_amg_solve_host: NULL operand");
    const size_t n = amg->n;
    LSSPG_TRY(ensure_stage(ctx, n));
    LSSPG_CUDA(cudaMemcpyAsync(ctx->stage[0], hx, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(ctx->stage[1], hb, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_TRY(lsspg_amg_solve(ctx, amg, ctx->stage[1], ctx->stage[0], tol, maxit, nits, ares));
    LSSPG_CUDA(cudaMemcpyAsync(hx, ctx->stage[0], n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return read_scalars(ctx, 0, 1, true);
}

}  // extern "C"

// common.cuh -- shared internals of liblsspg (context, error plumbing, grid reductions).
//
// Whole library is compiled with -fmad=false: the reference arithmetic is plain
// x86-64 SSE2 double without fused multiply-add, so every
// a*b+c in device code must round twice.  fp64 division and sqrt are IEEE
// round-to-nearest on sm_100a by default.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/lsspg.h"

namespace lsspg {

void set_error(const char *fmt, ...);
void log_printf(const char *fmt, ...);   // driver output: the registered printer (lssp_printf of the facade) or stdout
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define LSSPG_CUDA(call)                                                             \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) return lsspg::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define LSSPG_CHECK(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            lsspg::set_error(__VA_ARGS__);     \
            return 1;                          \
        }                                      \
    } while (0)

#define LSSPG_TRY(call)            \
    do {                           \
        int r__ = (call);          \
        if (r__) return r__;       \
    } while (0)

// launch bookkeeping: every kernel launch of the library goes through this macro
#define LSSPG_LAUNCH(ctx, kernel, grid, block, smem, ...)                          \
    do {                                                                           \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);           \
        (ctx)->launches++;                                                         \
        cudaError_t e__ = cudaPeekAtLastError();                                   \
        if (e__ != cudaSuccess) return lsspg::cuda_fail(e__, #kernel, __FILE__, __LINE__); \
    } while (0)

constexpr int kMaxRedBlocks = 4096;  // upper bound on gridDim.x of any reducing kernel
constexpr int kMaxRedK = 8;          // sums reduced together by one kernel
constexpr int kNumScalars = 512;     // device scalar slab (doubles)
constexpr int kBlock = 256;          // default CTA size of streaming kernels

}  // namespace lsspg

struct lsspg_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    double *d_partials = nullptr;     // [kMaxRedK][kMaxRedBlocks]
    unsigned int *d_ticket = nullptr; // wraps to 0 after every reduction
    double *d_scal = nullptr;         // scalar slab, device
    double *h_scal = nullptr;         // scalar slab, pinned host mirror
    int *d_flags = nullptr;           // device-side status flags (breakdown, ...)
    int *h_flags = nullptr;
    long long launches = 0;
    int opt_spmv_kernel = 0;
    int opt_spmv_exact = 0;
    int opt_check_every = 8;   // CG / BiCGStab: host reads the residuals every 8 iterations (device-side stop flag)
    int opt_graphs = 1;              // CG: replay the steady-state launch train as a CUDA graph
    int opt_reduce_sequential = 0;   // sums in the reference's sequential order: 1 = one thread adds (verification), 2 = the same result in parallel (exact_sum.cu)
    double *d_seq = nullptr;         // [kMaxRedK][seq_len] per-element terms of the sums (sequential mode)
    size_t seq_len = 0;
    void *d_xs = nullptr;            // scratch of the parallel reference-order sum (exact_sum.cu), xs_blocks blocks per sum
    size_t xs_blocks = 0;
    // grow-only device staging for the *_host entry points
    double *stage[3] = {nullptr, nullptr, nullptr};
    size_t stage_len = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    int tri_timeouts = 0;   // sweeps aborted by the watchdog so far (pencil schedules then empty their mailboxes)
    void *comm = nullptr;   // multi-GPU communicator (comm.cu), NULL on a single GPU
    // work-vector pool: the drivers allocate their vectors per solve as the reference does, but
    // cudaMalloc/cudaFree cost milliseconds, so released vectors are kept here for the next solve
    std::vector<std::pair<double *, size_t>> pool;
};

namespace lsspg {

// grid size of a bandwidth-bound grid-stride kernel: a multiple of the SM count
inline int stream_grid(const lsspg_ctx *ctx, long long work_items, int per_block, int ctas_per_sm = 8)
{
    long long need = (work_items + per_block - 1) / per_block;
    long long cap = (long long)ctx->num_sms * ctas_per_sm;
    if (cap > kMaxRedBlocks) cap = kMaxRedBlocks;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

int ensure_stage(lsspg_ctx *ctx, size_t n);
int ensure_seq(lsspg_ctx *ctx, size_t n);

#ifdef __CUDACC__
// ---- deterministic block / grid reductions --------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum K values across the CTA.  Result valid in thread 0.  Fixed order.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double (*sred)[32])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = warp_sum(v[k]);
    __syncthreads();  // protects sred reuse across calls
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) sred[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            double t = (lane < nw) ? sred[k][lane] : 0.0;
            v[k] = warp_sum(t);
        }
    }
}

// Grid-wide sum of K per-thread values; the last CTA to arrive adds the
// per-CTA partials in a fixed order and calls fin(sums) from its thread 0.
// Run-to-run deterministic for a fixed grid.  All threads must call it.
template <int K, class Fin>
__device__ __forceinline__ void grid_sum(double (&v)[K], double *partials, unsigned int *ticket, Fin fin)
{
    __shared__ double sred[K][32];
    __shared__ int s_last;
    block_sum<K>(v, sred);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) partials[k * kMaxRedBlocks + blockIdx.x] = v[k];
        __threadfence();
        unsigned int t = atomicInc(ticket, gridDim.x - 1);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double w[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            double s = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
                s += __ldcg(&partials[k * kMaxRedBlocks + i]);
            w[k] = s;
        }
        block_sum<K>(w, sred);
        if (threadIdx.x == 0) fin(w);
    }
}

// inclusive scan of one value per thread over a CTA of 1024 threads (32 warps); s_w: 32 words of shared memory.
// Returns the inclusive prefix; *total = the sum over the CTA.  Two barriers.
template <class T>
__device__ __forceinline__ T cta_scan_1024(T v, T *s_w, T *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T w = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += w;
    }
    __syncthreads();   // s_w may still be read from the previous call
    if (lane == 31) s_w[wid] = v;
    __syncthreads();
    T w = s_w[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T x = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += x;
    }
    *total = __shfl_sync(0xffffffffu, wi, 31);
    return v + __shfl_sync(0xffffffffu, wi - w, wid);
}

__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
#endif

}  // namespace lsspg

// ctx.cu -- device context, memory plumbing and error reporting of liblsspg.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace lsspg {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Driver messages (per-iteration lines, breakdown notices): the reference prints them with lssp_printf, which flushes and
// mirrors into the file given to lssp_solver_set_log (src/utils.cxx:93-112, src/solver-cg.cxx:108-112).  The library
// above this one registers its lssp_printf here; without a printer the text goes to stdout, flushed.
static void (*g_printer)(const char *) = nullptr;

void log_printf(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (g_printer) g_printer(buf);
    else {
        fputs(buf, stdout);
        fflush(stdout);
    }
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    return 1;
}

int ensure_stage(lsspg_ctx *ctx, size_t n)
{
    if (n <= ctx->stage_len) return 0;
    ctx->stage_len = 0;   // a failed allocation below must not leave a length that vouches for NULL buffers
    for (int i = 0; i < 3; i++) {
        if (ctx->stage[i]) LSSPG_CUDA(cudaFree(ctx->stage[i]));
        ctx->stage[i] = nullptr;
        LSSPG_CUDA(cudaMalloc(&ctx->stage[i], n * sizeof(double)));
    }
    ctx->stage_len = n;
    return 0;
}

int ensure_seq(lsspg_ctx *ctx, size_t n)
{
    if (n <= ctx->seq_len) return 0;
    n = (n + 31) / 32 * 32;   // every sum's slice starts 16-byte aligned (exact_sum.cu reads pairs)
    if (ctx->d_seq) LSSPG_CUDA(cudaFree(ctx->d_seq));
    ctx->d_seq = nullptr;
    LSSPG_CUDA(cudaMalloc(&ctx->d_seq, sizeof(double) * kMaxRedK * n));
    ctx->seq_len = n;
    return 0;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

void lsspg_set_printer(void (*fn)(const char *msg)) { lsspg::g_printer = fn; }


const char *lsspg_last_error(void) { return g_err; }
const char *lsspg_version(void) { return "lsspg 0.1 (sm_100a)"; }

int lsspg_ctx_create(int device, lsspg_ctx **out)
{
    LSSPG_CHECK(out != nullptr, "lsspg_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("lsspg_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 1;
    }
    LSSPG_CHECK(device >= 0 && device < count, "lsspg_ctx_create: device %d out of range [0,%d)", device, count);
    LSSPG_CUDA(cudaSetDevice(device));
    lsspg_ctx *c = new lsspg_ctx();
    c->device = device;
    cudaDeviceProp prop;
    LSSPG_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    LSSPG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    LSSPG_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * kMaxRedK * kMaxRedBlocks));
    LSSPG_CUDA(cudaMalloc(&c->d_ticket, sizeof(unsigned int) * 16));
    LSSPG_CUDA(cudaMemset(c->d_ticket, 0, sizeof(unsigned int) * 16));
    LSSPG_CUDA(cudaMalloc(&c->d_scal, sizeof(double) * kNumScalars));
    LSSPG_CUDA(cudaMemset(c->d_scal, 0, sizeof(double) * kNumScalars));
    LSSPG_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * kNumScalars));
    LSSPG_CUDA(cudaMalloc(&c->d_flags, sizeof(int) * 16));
    LSSPG_CUDA(cudaMemset(c->d_flags, 0, sizeof(int) * 16));
    LSSPG_CUDA(cudaMallocHost(&c->h_flags, sizeof(int) * 16));
    LSSPG_CUDA(cudaEventCreate(&c->ev0));
    LSSPG_CUDA(cudaEventCreate(&c->ev1));
    if (const char *g = getenv("LSSPG_GRAPHS")) c->opt_graphs = atoi(g) != 0;   // 0: no CUDA-graph replay in the drivers
    LSSPG_CUDA(cudaDeviceSynchronize());
    *out = c;
    return 0;
}

int lsspg_ctx_destroy(lsspg_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 3; i++)
        if (c->stage[i]) cudaFree(c->stage[i]);
    for (auto &pr : c->pool) cudaFree(pr.first);
    if (c->d_seq) cudaFree(c->d_seq);
    if (c->d_xs) cudaFree(c->d_xs);
    cudaFree(c->d_partials);
    cudaFree(c->d_ticket);
    cudaFree(c->d_scal);
    cudaFreeHost(c->h_scal);
    cudaFree(c->d_flags);
    cudaFreeHost(c->h_flags);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int lsspg_sync(lsspg_ctx *ctx)
{
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// CUDA-event timer on the context's own stream (torch.cuda.Event would only
// see torch's current stream).  Nestable up to 4 deep by slot.
int lsspg_timer_start(lsspg_ctx *ctx, int slot)
{
    LSSPG_CHECK(slot >= 0 && slot < 4, "lsspg_timer_start: slot out of range");
    if (!ctx->tev[slot][0]) {
        LSSPG_CUDA(cudaEventCreate(&ctx->tev[slot][0]));
        LSSPG_CUDA(cudaEventCreate(&ctx->tev[slot][1]));
    }
    LSSPG_CUDA(cudaEventRecord(ctx->tev[slot][0], ctx->stream));
    return 0;
}

int lsspg_timer_stop(lsspg_ctx *ctx, int slot, double *ms)
{
    LSSPG_CHECK(slot >= 0 && slot < 4 && ctx->tev[slot][0], "lsspg_timer_stop: timer not started");
    LSSPG_CUDA(cudaEventRecord(ctx->tev[slot][1], ctx->stream));
    LSSPG_CUDA(cudaEventSynchronize(ctx->tev[slot][1]));
    float f = 0.f;
    LSSPG_CUDA(cudaEventElapsedTime(&f, ctx->tev[slot][0], ctx->tev[slot][1]));
    if (ms) *ms = f;
    return 0;
}

void *lsspg_ctx_stream(lsspg_ctx *ctx) { return (void *)ctx->stream; }
long long lsspg_ctx_launches(lsspg_ctx *ctx) { return ctx->launches; }

int lsspg_ctx_set_option(lsspg_ctx *ctx, int option, int value)
{
    switch (option) {
        case LSSPG_OPT_SPMV_KERNEL: ctx->opt_spmv_kernel = value; return 0;
        case LSSPG_OPT_SPMV_EXACT: ctx->opt_spmv_exact = value; return 0;
        case LSSPG_OPT_CHECK_EVERY: ctx->opt_check_every = value < 1 ? 1 : value; return 0;
        case LSSPG_OPT_REDUCE_SEQUENTIAL: ctx->opt_reduce_sequential = value; return 0;
        case LSSPG_OPT_GRAPHS: ctx->opt_graphs = value; return 0;
    }
    set_error("lsspg_ctx_set_option: unknown option %d", option);
    return 1;
}

int lsspg_malloc(lsspg_ctx *ctx, size_t bytes, void **dptr)
{
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    *dptr = nullptr;
    if (bytes == 0) return 0;
    // +64 B slack: the streaming kernels may over-read up to one 16-byte vector
    LSSPG_CUDA(cudaMalloc(dptr, bytes + 64));
    return 0;
}

int lsspg_free(lsspg_ctx *ctx, void *dptr)
{
    if (!dptr) return 0;
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    LSSPG_CUDA(cudaFree(dptr));
    return 0;
}

int lsspg_h2d(lsspg_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    if (bytes == 0) return 0;
    LSSPG_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lsspg_d2h(lsspg_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    if (bytes == 0) return 0;
    LSSPG_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int lsspg_memset_zero(lsspg_ctx *ctx, void *dptr, size_t bytes)
{
    if (bytes == 0) return 0;
    LSSPG_CUDA(cudaMemsetAsync(dptr, 0, bytes, ctx->stream));
    return 0;
}

int lsspg_host_alloc(size_t bytes, void **hptr)
{
    LSSPG_CUDA(cudaMallocHost(hptr, bytes ? bytes : 1));
    return 0;
}

int lsspg_host_free(void *hptr)
{
    if (hptr) LSSPG_CUDA(cudaFreeHost(hptr));
    return 0;
}

}  // extern "C"

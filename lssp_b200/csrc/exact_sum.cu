// exact_sum.cu -- kernels K1..K4 of exact_sum.cuh (LSSPG_OPT_REDUCE_SEQUENTIAL = 2: every dot product / norm equals the
// reference's sequential sum bit for bit, at the cost of two more passes over the parked terms instead of an n-step
// chain) and their host replay for the CPU test-suite.
#include "exact_sum.cuh"
#include <algorithm>
#include <vector>
#include "blas1.cuh"
#include "comm.cuh"

namespace lsspg {

// scratch per sum (ctx->d_xs, [kMaxRedK] slices of `stride` blocks): approximate block sums, excursion bounds (xs_dev), predicted
// binades, D and its inclusive wrap-around scan
struct XsScratch {
    double *approx, *dev;
    unsigned long long *D, *scan;
    int *e;
    unsigned int *ticket;
    long long stride;
};

static XsScratch xs_scratch(lsspg_ctx *ctx)
{
    XsScratch sc;
    const size_t nb = ctx->xs_blocks;
    char *p = reinterpret_cast<char *>(ctx->d_xs);
    sc.approx = reinterpret_cast<double *>(p); p += sizeof(double) * kMaxRedK * nb;
    sc.dev = reinterpret_cast<double *>(p); p += sizeof(double) * kMaxRedK * nb;
    sc.D = reinterpret_cast<unsigned long long *>(p); p += sizeof(unsigned long long) * kMaxRedK * nb;
    sc.scan = reinterpret_cast<unsigned long long *>(p); p += sizeof(unsigned long long) * kMaxRedK * nb;
    sc.e = reinterpret_cast<int *>(p); p += sizeof(int) * kMaxRedK * nb;
    sc.ticket = reinterpret_cast<unsigned int *>(p);
    sc.stride = (long long)nb;
    return sc;
}

size_t xs_scratch_bytes(size_t nb) { return (size_t)kMaxRedK * nb * (8 + 8 + 8 + 8 + 4) + 64; }

LSSPG_HD long long xs_min(long long a, long long b) { return a < b ? a : b; }

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_scan(double v, int lane)   // inclusive
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double w = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += w;
    }
    return v;
}

// K1: one warp per block of kXsB terms; lane l holds the kXsB / 32 CONSECUTIVE terms l kXsB/32 ..: its prefix sums are
// local additions and one warp scan of the lane totals does the rest (the terms array is 16-byte aligned: ensure_seq)
__global__ void __launch_bounds__(256) k_xs_blocks(long long n, long long nb, const double *__restrict__ terms, long long seq_n,
                                                   XsScratch sc, const int *stop)
{
    if (stop && *stop) return;
    constexpr int PER = kXsB / 32;
    const int lane = threadIdx.x & 31, k = blockIdx.y;
    const double *t = terms + (size_t)k * seq_n;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nb; b += warps) {
        double v[PER];
        const long long i0 = b * kXsB + lane * PER;
        if ((b + 1) * kXsB <= n) {
#pragma unroll
            for (int j = 0; j < PER; j += 2) {
                const double2 w = *reinterpret_cast<const double2 *>(t + i0 + j);
                v[j] = w.x;
                v[j + 1] = w.y;
            }
        }
        else {
#pragma unroll
            for (int j = 0; j < PER; j++) v[j] = (i0 + j < n) ? t[i0 + j] : 0.0;
        }
        double run = 0.0, a = 0.0;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            run += v[j];
            a += fabs(v[j]);
            v[j] = run;          // the lane's own inclusive prefixes
        }
        const double incl = warp_scan(run, lane), off = incl - run;   // lanes before this one
        double x = 0.0;
#pragma unroll
        for (int j = 0; j < PER; j++) x = fmax(x, fabs(off + v[j]));
        a = warp_sum(a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        const double total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
            sc.approx[k * sc.stride + b] = total;
            sc.dev[k * sc.stride + b] = xs_dev(x, a);
        }
    }
}

// K2: one CTA of 1024 threads per sum: exclusive prefix of the approximate block sums -> predicted binades
// (1024 consecutive blocks per round: coalesced)
__global__ void __launch_bounds__(1024) k_xs_predict(long long nb, XsScratch sc, const int *stop)
{
    if (stop && *stop) return;
    __shared__ double s_w[32];
    const int k = blockIdx.x, tid = threadIdx.x;
    const double *ap = sc.approx + k * sc.stride;
    int *eb = sc.e + k * sc.stride;
    double carry = 0.0;
    constexpr int kBatch = 8;   // chunks loaded together: their L2 latencies overlap
    for (long long b0 = 0; b0 < nb; b0 += 1024 * kBatch) {
        double v[kBatch];
#pragma unroll
        for (int r = 0; r < kBatch; r++) {
            const long long b = b0 + r * 1024 + tid;
            v[r] = (b < nb) ? ap[b] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < kBatch; r++) {
            const long long b = b0 + r * 1024 + tid;
            double total;
            const double inc = cta_scan_1024<double>(v[r], s_w, &total);
            if (b < nb) eb[b] = xs_exponent(carry + (inc - v[r]));
            carry += total;
        }
    }
}

// K3: one warp per block: D_b = sum_i rn_u(t_i) / u for the predicted binade; a tie makes the block unclean
__global__ void __launch_bounds__(256) k_xs_round(long long n, long long nb, const double *__restrict__ terms, long long seq_n,
                                                  XsScratch sc, const int *stop)
{
    if (stop && *stop) return;
    const int lane = threadIdx.x & 31, k = blockIdx.y;
    const double *t = terms + (size_t)k * seq_n;
    int *eb = sc.e + k * sc.stride;
    unsigned long long *D = sc.D + k * sc.stride;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nb; b += warps) {
        const int e = eb[b];
        long long d = 0;
        bool tie = false;
        if (e != kXsUnclean) {
#pragma unroll
            for (int j = 0; j < kXsB / 32; j++) {
                const long long i = b * kXsB + j * 32 + lane;
                const double v = (i < n) ? t[i] : 0.0;
                d += xs_round(v, e, &tie);
            }
        }
        d = warp_sum_ll(d);
        tie = __any_sync(0xffffffffu, tie);
        if (lane == 0) {
            if (tie) eb[b] = kXsUnclean;
            D[b] = (tie || e == kXsUnclean) ? 0ull : (unsigned long long)d;
        }
    }
}

// K4: the walk (one CTA of kXsChunk threads per sum).  The blocks are taken in windows of kXsWin; a window's binades,
// excursion bounds and the inclusive wrap-around scan of its D are staged in shared memory, so that the rounds of a window
// -- one per run of verified blocks, one more per block that has to be taken apart -- cost barriers, not L2 round trips.
constexpr int kXsWin = kXsChunk * kXsPer;
constexpr int kXsPre = 12;   // blocks of a window whose terms are fetched into shared memory ahead of the walk
struct XsWindow {
    unsigned long long scan[kXsWin];
    double dev[kXsWin];
    double pre[kXsPre][kXsB];   // terms of the blocks expected to be taken apart (no binade, or the binade changes after them)
    int e[kXsWin];
    int pre_block[kXsPre];      // their window-relative numbers (any order); unused entries: -1
};

__global__ void __launch_bounds__(kXsChunk) k_xs_walk(long long n, long long nb, int K, const double *__restrict__ terms,
                                                      long long seq_n, XsScratch sc, double *scal, int *flags, int out_slot,
                                                      int defer_fin, FinProg fin, const int *stop)
{
    if (stop && *stop) return;
    extern __shared__ __align__(16) unsigned char xs_smem[];
    XsWindow &W = *reinterpret_cast<XsWindow *>(xs_smem);
    __shared__ unsigned long long s_w[32];
    __shared__ double s_terms[kXsB], s_fdev[kXsB / kXsFine];
    __shared__ long long s_fD[kXsB / kXsFine];
    __shared__ int s_ftie[kXsB / kXsFine];
    __shared__ double s_s;
    __shared__ int s_pos, s_bad, s_npre;
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const double *t = terms + (size_t)k * seq_n;
    const double *dev = sc.dev + k * sc.stride;
    const int *eb = sc.e + k * sc.stride;
    const unsigned long long *D = sc.D + k * sc.stride;
    if (tid == 0) s_s = 0.0;   // src/vector.cxx:127: the sum starts at +0.0
    for (long long w0 = 0; w0 < nb; w0 += kXsWin) {
        const int wlen = (int)xs_min(kXsWin, nb - w0);
        __syncthreads();   // the previous window is done with W
        {   // stage the window; thread tid scans blocks kXsPer tid .. + kXsPer - 1 (only differences inside a run of equal
            // binades are ever used, so the scan may wrap around and restart with every window)
            for (int r = tid; r < wlen; r += kXsChunk) {
                W.scan[r] = D[w0 + r];
                W.dev[r] = dev[w0 + r];
                W.e[r] = eb[w0 + r];
            }
            __syncthreads();
            unsigned long long loc[kXsPer], sum = 0;
#pragma unroll
            for (int q = 0; q < kXsPer; q++) {
                const int r = tid * kXsPer + q;
                loc[q] = (r < wlen) ? W.scan[r] : 0ull;
                sum += loc[q];
            }
            unsigned long long total;
            unsigned long long run = cta_scan_1024<unsigned long long>(sum, s_w, &total) - sum;
#pragma unroll
            for (int q = 0; q < kXsPer; q++) {
                const int r = tid * kXsPer + q;
                run += loc[q];
                if (r < wlen) W.scan[r] = run;
            }
            // blocks that will most likely be taken apart (no binade, or the binade changes after them): the first kXsPre of
            // the window have their terms fetched into shared memory now (warp w fetches candidate w), the rest go to L2
            if (tid < kXsPre) W.pre_block[tid] = -1;
            if (tid == 0) s_npre = 0;
            __syncthreads();
            for (int r = tid; r < wlen; r += kXsChunk) {
                const int e = W.e[r];
                if (e == kXsUnclean || (r + 1 < wlen && W.e[r + 1] != e)) {
                    const int slot = atomicAdd(&s_npre, 1);
                    if (slot < kXsPre) W.pre_block[slot] = r;
                    else {
                        const char *pf = reinterpret_cast<const char *>(t + (w0 + r) * kXsB);
                        const long long lim = (n - (w0 + r) * kXsB) * 8;
                        for (int o = 0; o < kXsB * 8 && o < lim; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + o));
                    }
                }
            }
            __syncthreads();
            if (wid < kXsPre && W.pre_block[wid] >= 0) {
                const long long i0 = (w0 + W.pre_block[wid]) * kXsB;
#pragma unroll
                for (int q = 0; q < kXsB / 32; q++) {
                    const long long i = i0 + q * 32 + lane;
                    W.pre[wid][q * 32 + lane] = (i < n) ? t[i] : 0.0;
                }
            }
        }
        if (tid == 0) s_pos = 0;
        for (;;) {
            __syncthreads();   // s_s, s_pos (and, the first time, W) are visible
            const int pos = s_pos;
            const double s = s_s;
            if (pos >= wlen) break;
            const int e = W.e[pos];
            bool seq = (e == kXsUnclean) || xs_exponent(s) != e;
            if (tid == 0) s_bad = wlen;
            __syncthreads();
            if (!seq) {
                const long long m0 = xs_to_int(s, e);
                const unsigned long long base = pos > 0 ? W.scan[pos - 1] : 0ull;
                int mine = wlen;
#pragma unroll
                for (int q = 0; q < kXsPer; q++) {
                    const int r = pos + q * kXsChunk + tid;
                    if (r < wlen) {
                        bool ok = (W.e[r] == e);
                        if (ok) {
                            const long long m = m0 + (long long)((r > 0 ? W.scan[r - 1] : 0ull) - base);
                            ok = xs_int_in_binade(m) && xs_verify(xs_from_int(m, e), e, W.dev[r]);
                        }
                        if (!ok && r < mine) mine = r;
                    }
                }
                if (mine < wlen) atomicMin(&s_bad, mine);
                __syncthreads();
                const int bad = s_bad;
                if (bad == pos) seq = true;   // (uniform: every thread reads the same s_bad)
                else {
                    if (tid == 0) {
                        s_s = xs_from_int(m0 + (long long)(W.scan[bad - 1] - base), e);
                        s_pos = bad;
                    }
                    continue;
                }
            }
            // The block cannot be advanced as a whole.  Second level: its kXsFine-term pieces, rounded for the binade s is
            // in NOW (8 warps, one piece each); thread 0 then advances piece by piece, and only pieces that fail the same
            // verification -- the one with the binade crossing or the tie -- are added term by term, as the reference does.
            const long long i0 = (w0 + pos) * kXsB;
            const int cnt = (int)xs_min(kXsB, n - i0);
            int pre = -1;
#pragma unroll
            for (int q = 0; q < kXsPre; q++)
                if (W.pre_block[q] == pos) pre = q;
            if (tid < kXsB) s_terms[tid] = (pre >= 0) ? W.pre[pre][tid] : ((tid < cnt) ? t[i0 + tid] : 0.0);
            __syncthreads();
            const int es = xs_exponent(s);
            if (tid < kXsB && es != kXsUnclean) {
                bool tie = false;
                const double v = s_terms[tid];
                long long d = xs_round(v, es, &tie);
                double a = fabs(v), x = fabs(warp_scan(v, lane));
                d = warp_sum_ll(d);
                a = warp_sum(a);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
                tie = __any_sync(0xffffffffu, tie);
                if (lane == 0) { s_fD[wid] = d; s_fdev[wid] = xs_dev(x, a); s_ftie[wid] = tie ? 1 : 0; }
            }
            __syncthreads();
            if (tid == 0) {
                double acc = s;
                for (int j = 0; j * kXsFine < cnt; j++) {
                    if (es != kXsUnclean && !s_ftie[j] && xs_exponent(acc) == es && xs_verify(acc, es, s_fdev[j])) {
                        const long long m = xs_to_int(acc, es) + s_fD[j];
                        if (xs_int_in_binade(m)) { acc = xs_from_int(m, es); continue; }
                    }
                    // term by term (terms beyond the end of the data are +0.0: acc is never -0.0, so adding them changes
                    // nothing); loaded first, then added: the chain is 32 dependent additions, not 32 load-add pairs
                    double tv[kXsFine];
#pragma unroll
                    for (int i = 0; i < kXsFine; i++) tv[i] = s_terms[j * kXsFine + i];
#pragma unroll
                    for (int i = 0; i < kXsFine; i++) acc += tv[i];
                }
                s_s = acc;
                s_pos = pos + 1;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        scal[out_slot + k] = s_s;
        __threadfence();
        const unsigned int tk = atomicInc(sc.ticket, (unsigned int)K - 1);
        if (tk == (unsigned int)K - 1 && !defer_fin) {
            __threadfence();
            fin_run(fin, scal, flags);
        }
    }
}

int ensure_xs(lsspg_ctx *ctx, size_t n)
{
    const size_t nb = (n + kXsB - 1) / kXsB + 1;
    if (nb <= ctx->xs_blocks) return 0;
    if (ctx->d_xs) LSSPG_CUDA(cudaFree(ctx->d_xs));
    ctx->d_xs = nullptr;
    ctx->xs_blocks = 0;
    LSSPG_CUDA(cudaMalloc(&ctx->d_xs, xs_scratch_bytes(nb)));
    LSSPG_CUDA(cudaMemsetAsync(ctx->d_xs, 0, xs_scratch_bytes(nb), ctx->stream));
    ctx->xs_blocks = nb;
    return 0;
}

// K sums of the n parked terms each (ctx->d_seq), in the reference's order, into scal[out_slot ..]; then the FinProg
int exact_seq_sum(lsspg_ctx *ctx, long long n, int K, const RedOut &o)
{
    LSSPG_TRY(ensure_xs(ctx, (size_t)std::max<long long>(n, 1)));
    const XsScratch sc = xs_scratch(ctx);
    const long long nb = (n + kXsB - 1) / kXsB;
    const int *stop = o.guarded ? ctx->d_flags + FLAG_STOP : nullptr;
    const long long seq_n = (long long)ctx->seq_len;
    if (nb > 0) {
        const dim3 grid((unsigned int)std::max<long long>(1, std::min<long long>((nb + 7) / 8, (long long)ctx->num_sms * 8)), (unsigned int)K);
        LSSPG_LAUNCH(ctx, k_xs_blocks, grid, 256, 0, n, nb, ctx->d_seq, seq_n, sc, stop);
        LSSPG_LAUNCH(ctx, k_xs_predict, K, 1024, 0, nb, sc, stop);
        LSSPG_LAUNCH(ctx, k_xs_round, grid, 256, 0, n, nb, ctx->d_seq, seq_n, sc, stop);
    }
    static bool attr_done = false;
    if (!attr_done) {
        LSSPG_CUDA(cudaFuncSetAttribute(k_xs_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(XsWindow)));
        attr_done = true;
    }
    LSSPG_LAUNCH(ctx, k_xs_walk, K, kXsChunk, sizeof(XsWindow), n, nb, K, ctx->d_seq, seq_n, sc, ctx->d_scal, ctx->d_flags, o.out_slot,
                 distributed(ctx) ? 1 : 0, o.fin, stop);
    return 0;
}

// ---- host replay: the same four phases with the same shared functions (CPU test-suite) -------------------------------
static double xs_host(long long n, const double *t, long long *stats)
{
    const long long nb = (n + kXsB - 1) / kXsB;
    std::vector<double> approx(nb), dev(nb);
    std::vector<int> eb(nb);
    std::vector<unsigned long long> D(nb), scan(nb);
    for (long long b = 0; b < nb; b++) {   // K1 (pairwise order inside a block; any order will do)
        double s = 0.0, a = 0.0, x = 0.0;
        for (long long i = b * kXsB; i < std::min(n, (b + 1) * kXsB); i++) { s += t[i]; a += fabs(t[i]); x = std::max(x, fabs(s)); }
        approx[b] = s; dev[b] = xs_dev(x, a);
    }
    {   // K2
        double run = 0.0;
        for (long long b = 0; b < nb; b++) { eb[b] = xs_exponent(run); run += approx[b]; }
    }
    for (long long b = 0; b < nb; b++) {   // K3
        const int e = eb[b];
        long long d = 0;
        bool tie = false;
        if (e != kXsUnclean)
            for (long long i = b * kXsB; i < std::min(n, (b + 1) * kXsB); i++) d += xs_round(t[i], e, &tie);
        if (tie) eb[b] = kXsUnclean;
        D[b] = (tie || e == kXsUnclean) ? 0ull : (unsigned long long)d;
    }
    {   // the scan restarts with every window (wrap-around arithmetic: only differences inside a run are used)
        unsigned long long run = 0;
        for (long long b = 0; b < nb; b++) {
            if (b % kXsWin == 0) run = 0;
            run += D[b];
            scan[b] = run;
        }
    }
    auto scan_before = [](long long) { return 0ull; };   // the scan value "before" the first block of a window
    double s = 0.0;   // K4: windows of kXsWin blocks, as k_xs_walk
    long long rounds = 0, seq_blocks = 0, seq_pieces = 0;
    for (long long w0 = 0; w0 < nb; w0 += kXsWin) {
        const long long wend = std::min<long long>(nb, w0 + kXsWin);
        long long pos = w0;
        while (pos < wend) {
            rounds++;
            const int e = eb[pos];
            bool seq = (e == kXsUnclean) || xs_exponent(s) != e;
            if (!seq) {
                const long long m0 = xs_to_int(s, e);
                const unsigned long long base = pos > w0 ? scan[pos - 1] : scan_before(w0);
                long long bad = wend;
                for (long long b = pos; b < wend; b++) {
                    bool ok = (eb[b] == e);
                    if (ok) {
                        const long long m = m0 + (long long)((b > w0 ? scan[b - 1] : scan_before(w0)) - base);
                        ok = xs_int_in_binade(m) && xs_verify(xs_from_int(m, e), e, dev[b]);
                    }
                    if (!ok) { bad = b; break; }
                }
                if (bad == pos) seq = true;
                else {
                    s = xs_from_int(m0 + (long long)(scan[bad - 1] - base), e);
                    pos = bad;
                    continue;
                }
            }
            {   // second level, as in k_xs_walk
                const long long i0 = pos * kXsB;
                const int cnt = (int)std::min<long long>(kXsB, n - i0);
                const int es = xs_exponent(s);
                double acc = s;
                for (int j = 0; j * kXsFine < cnt; j++) {
                    const int i1 = std::min((j + 1) * kXsFine, cnt);
                    if (es != kXsUnclean && xs_exponent(acc) == es) {
                        bool tie = false;
                        long long d = 0;
                        double a = 0.0, x = 0.0, pre = 0.0;
                        for (int i = j * kXsFine; i < i1; i++) {
                            d += xs_round(t[i0 + i], es, &tie);
                            a += fabs(t[i0 + i]);
                            pre += t[i0 + i];
                            x = std::max(x, fabs(pre));
                        }
                        if (!tie && xs_verify(acc, es, xs_dev(x, a))) {
                            const long long m = xs_to_int(acc, es) + d;
                            if (xs_int_in_binade(m)) { acc = xs_from_int(m, es); continue; }
                        }
                    }
                    for (int i = j * kXsFine; i < i1; i++) acc += t[i0 + i];
                    seq_pieces++;
                }
                s = acc;
            }
            seq_blocks++;
            pos++;
        }
    }
    if (stats) { stats[0] = nb; stats[1] = rounds; stats[2] = seq_blocks; stats[3] = seq_pieces; }
    return s;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

// CPU self-check for the test-suite (never called by a product path): the reference-order sum of t[0..n) by the host
// replay of the kernels above.  stats[4]: blocks, rounds of the walk, blocks not advanced as a whole, pieces (32 terms) added term by term.
int lsspg_debug_exact_seq_sum_host(long long n, const double *t, double *out, long long *stats)
{
    if (n < 0 || !out) return 1;
    *out = xs_host(n, t, stats);
    return 0;
}

}  // extern "C"

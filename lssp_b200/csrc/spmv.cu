// spmv.cu -- CSR SpMV / residual kernels for sm_100a (replaces src/mvops.cxx).
//
// Layout in HBM: plain CSR (Ap int32[n+1], Aj int32[nnz], Ax fp64[nnz]), padded
// so that 16-byte vector loads may start at the previous multiple of 4 elements.
// At upload the rows are cut into contiguous ROW TILES:
//   STREAM tile : <= 256 rows and <= 2048 nnz, every row short (<= 64 nnz).
//                 The CTA copies the tile's contiguous val/col segment into
//                 shared memory with coalesced 128-bit loads, then one thread
//                 per row adds its products SEQUENTIALLY in storage order:
//                 bit-identical to the reference loop (src/mvops.cxx:130-132).
//   BALANCED    : irregular tiles (longest row > 2 x average + 4, from the tile's
//   tile          row-length histogram): the products x[col]*val of the whole tile
//                 are formed first, spread evenly over the threads (8 independent
//                 gathers in flight each), and written back over val; then one
//                 thread per row adds its products in storage order from shared
//                 memory.  Same roundings in the same order: still bit-identical.
//   MIXED tile  : BALANCED, but rows of 65 .. 2048 entries are added by a warp
//                 (lanes stride the products, shuffle tree; <= 1e-14).  Tiles are
//                 never cut at a long row, so they stay ~2048 nnz.
//   BLOCK tile  : rows longer than a tile (> 2048 nnz): the whole CTA strides the
//                 row from global memory, fixed-order block reduction.
//   SERIAL tile : rows longer than a tile in LSSPG_OPT_SPMV_EXACT mode, one thread
//                 per row straight from global memory (bit-exact, slow; opt-in);
//                 in that mode every other row is summed sequentially (STREAM).
// Algorithmic bytes per launch: 12 nnz + 4 (n+1) + 16 n (+ 8 n when y is read).
// x is gathered through L1/L2 (read-only path); val/col/Ap/y/z cross HBM once.
#include <algorithm>
#include "blas1.cuh"
#include "comm.cuh"
#include "spmv.cuh"
#include "setup_gpu.cuh"
#include "host_par.h"

namespace lsspg {

__device__ __forceinline__ double2 ld_stream_f2(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ int4 ld_stream_i4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

template <int KIND>
__device__ __forceinline__ double epilogue(double sum, double alpha, double beta, const double *y, int r)
{
    // operand order exactly as the reference
    if (KIND == LSSPG_MV_MXY) return sum;
    if (KIND == LSSPG_MV_AMXY) return sum * alpha;                     // src/mvops.cxx:99
    if (KIND == LSSPG_MV_AMXPBY) return sum * alpha + y[r] * beta;     // src/mvops.cxx:23
    return y[r] * beta + alpha * sum;                                  // src/mvops.cxx:61
}

struct SpmvArgs {
    const int *Ap;
    const int *Aj;
    const double *Ax;
    const int *tile_row;
    const unsigned char *tile_kind;
    int num_tiles;
    int cap;  // smem capacity in nnz (multiple of 4)
    const double *x;
    const double *y;
    double *z;
    Coef alpha, beta;
    const double *w0, *w1;
    double *scal;
    int *flags;
    double *partials;
    unsigned int *ticket;
    int out_slot;
    const int *stop;
    double *seq;          // sequential-order verification mode (see blas1.cu)
    long long seq_n;
    int defer_fin;        // multi-GPU: `fin` runs after the cross-rank all-reduce (comm.cu)
    FinProg fin;
};

template <int NDOT>
__device__ __forceinline__ void dots_add(const SpmvArgs &a, double (&acc)[NDOT > 0 ? NDOT : 1], int r, double out)
{
    if (NDOT >= 1) {
        const double t = out * (a.w0 ? a.w0[r] : out);
        if (a.seq) a.seq[r] = t;
        else acc[0] += t;
    }
    if (NDOT >= 2) {
        const double t = out * (a.w1 ? a.w1[r] : out);
        if (a.seq) a.seq[a.seq_n + r] = t;
        else acc[NDOT >= 2 ? 1 : 0] += t;
    }
}

// IRR = false: the matrix has STREAM (and SERIAL) tiles only -- the lean instantiation, 32 registers,
// 8 resident CTAs per SM; IRR = true adds the BALANCED / MIXED / BLOCK paths of irregular matrices.
template <int KIND, int NDOT, bool IRR>
__global__ void __launch_bounds__(kBlock) spmv_tiles_kernel(const SpmvArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sval = reinterpret_cast<double *>(smem_raw);
    int *scol = reinterpret_cast<int *>(sval + a.cap + 8);
    int *sap = scol + a.cap + 8;

    if (a.stop && *a.stop) return;
    const int tid = threadIdx.x;
    const double alpha = coef_get(a.alpha, a.scal);
    const double beta = coef_get(a.beta, a.scal);
    const double *__restrict__ x = a.x;
    double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int k = 0; k < (NDOT > 0 ? NDOT : 1); k++) acc[k] = 0.0;

    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        const int r0 = a.tile_row[tile];
        const int nr = a.tile_row[tile + 1] - r0;
        const int kind = a.tile_kind[tile];
        if (kind == TILE_STREAM || (IRR && (kind == TILE_MIXED || kind == TILE_BALANCED))) {
            const bool mixed = (kind == TILE_MIXED);
            for (int i = tid; i <= nr; i += kBlock) sap[i] = a.Ap[r0 + i];
            __syncthreads();
            const int e0 = sap[0], e1 = sap[nr];
            const int a0 = e0 & ~3;
            const int nvec = (e1 - a0 + 3) >> 2;  // groups of 4 elements
            const double2 *gv = reinterpret_cast<const double2 *>(a.Ax + a0);
            const int4 *gc = reinterpret_cast<const int4 *>(a.Aj + a0);
            double2 *sv = reinterpret_cast<double2 *>(sval);
            int4 *sc = reinterpret_cast<int4 *>(scol);
#pragma unroll 4
            for (int i = tid; i < 2 * nvec; i += kBlock) sv[i] = ld_stream_f2(gv + i);
#pragma unroll 2
            for (int i = tid; i < nvec; i += kBlock) sc[i] = ld_stream_i4(gc + i);
            __syncthreads();
            if (!IRR || kind == TILE_STREAM) {
                // regular rows: one thread per row gathers and adds in storage order
                for (int r = tid; r < nr; r += kBlock) {
                    int k = sap[r] - a0;
                    const int k1 = sap[r + 1] - a0;
                    double sum = 0.0;
                    for (; k + 4 <= k1; k += 4) {
                        const double x0 = __ldg(x + scol[k]), x1 = __ldg(x + scol[k + 1]);
                        const double x2 = __ldg(x + scol[k + 2]), x3 = __ldg(x + scol[k + 3]);
                        const double p0 = x0 * sval[k], p1 = x1 * sval[k + 1];
                        const double p2 = x2 * sval[k + 2], p3 = x3 * sval[k + 3];
                        sum += p0; sum += p1; sum += p2; sum += p3;
                    }
                    for (; k < k1; k++) sum += __ldg(x + scol[k]) * sval[k];
                    const double out = epilogue<KIND>(sum, alpha, beta, a.y, r0 + r);
                    a.z[r0 + r] = out;
                    dots_add<NDOT>(a, acc, r0 + r, out);
                }
            }
            else {
                // irregular rows: first every product x[col]*val of the tile, spread evenly over the
                // threads whatever the row lengths (8 independent gathers in flight per thread), written
                // back over val; then one thread per row adds ITS products in storage order from shared
                // memory -- the same two roundings per entry, in the same order, as src/mvops.cxx:130-132
                const int kb = e0 - a0, ke = e1 - a0;
                int k = kb + tid;
                for (; k + 7 * kBlock < ke; k += 8 * kBlock) {
                    double xv[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) xv[j] = __ldg(x + scol[k + j * kBlock]);
#pragma unroll
                    for (int j = 0; j < 8; j++) sval[k + j * kBlock] = xv[j] * sval[k + j * kBlock];
                }
                for (; k < ke; k += kBlock) sval[k] = __ldg(x + scol[k]) * sval[k];
                __syncthreads();
                for (int r = tid; r < nr; r += kBlock) {
                    int q = sap[r] - a0;
                    const int q1 = sap[r + 1] - a0;
                    if (mixed && q1 - q > kLongRow) continue;   // left to the warps below
                    double sum = 0.0;
                    for (; q + 4 <= q1; q += 4) {
                        const double p0 = sval[q], p1 = sval[q + 1], p2 = sval[q + 2], p3 = sval[q + 3];
                        sum += p0; sum += p1; sum += p2; sum += p3;
                    }
                    for (; q < q1; q++) sum += sval[q];
                    const double out = epilogue<KIND>(sum, alpha, beta, a.y, r0 + r);
                    a.z[r0 + r] = out;
                    dots_add<NDOT>(a, acc, r0 + r, out);
                }
                if (mixed) {   // rows of 65 .. 2048 entries: a warp per row, lanes stride, shuffle tree (<= 1e-14)
                    const int w = tid >> 5, lane = tid & 31;
                    for (int r = w; r < nr; r += kBlock / 32) {
                        int q = sap[r] - a0 + lane;
                        const int q1 = sap[r + 1] - a0;
                        if (q1 - (q - lane) <= kLongRow) continue;
                        double s0 = 0.0, s1 = 0.0;
                        for (; q + 32 < q1; q += 64) {
                            s0 += sval[q];
                            s1 += sval[q + 32];
                        }
                        if (q < q1) s0 += sval[q];
                        const double s = warp_sum(s0 + s1);
                        if (lane == 0) {
                            const double out = epilogue<KIND>(s, alpha, beta, a.y, r0 + r);
                            a.z[r0 + r] = out;
                            dots_add<NDOT>(a, acc, r0 + r, out);
                        }
                    }
                }
            }
            __syncthreads();
        }
        else if (IRR && kind == TILE_BLOCK) {
            for (int rr = 0; rr < nr; rr++) {
                const int r = r0 + rr;
                const int e1 = a.Ap[r + 1];
                int k = a.Ap[r] + tid;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                for (; k + 3 * kBlock < e1; k += 4 * kBlock) {
                    const int c0 = a.Aj[k], c1 = a.Aj[k + kBlock], c2 = a.Aj[k + 2 * kBlock], c3 = a.Aj[k + 3 * kBlock];
                    s0 += __ldg(x + c0) * a.Ax[k]; s1 += __ldg(x + c1) * a.Ax[k + kBlock];
                    s2 += __ldg(x + c2) * a.Ax[k + 2 * kBlock]; s3 += __ldg(x + c3) * a.Ax[k + 3 * kBlock];
                }
                for (; k < e1; k += kBlock) s0 += __ldg(x + a.Aj[k]) * a.Ax[k];
                double s = warp_sum((s0 + s1) + (s2 + s3));
                __syncthreads();   // sval is free here: no stream tile is in flight in this CTA
                if ((tid & 31) == 0) sval[tid >> 5] = s;
                __syncthreads();
                if (tid < 32) {
                    s = warp_sum(tid < kBlock / 32 ? sval[tid] : 0.0);
                    if (tid == 0) {
                        const double out = epilogue<KIND>(s, alpha, beta, a.y, r);
                        a.z[r] = out;
                        dots_add<NDOT>(a, acc, r, out);
                    }
                }
            }
            __syncthreads();
        }
        else {  // TILE_SERIAL
            for (int rr = tid; rr < nr; rr += kBlock) {
                const int r = r0 + rr;
                const int e1 = a.Ap[r + 1];
                double sum = 0.0;
                for (int k = a.Ap[r]; k < e1; k++) sum += __ldg(x + a.Aj[k]) * a.Ax[k];
                const double out = epilogue<KIND>(sum, alpha, beta, a.y, r);
                a.z[r] = out;
                dots_add<NDOT>(a, acc, r, out);
            }
        }
    }
    if (NDOT > 0 && a.seq == nullptr) {
        double *scal = a.scal;
        int *flags = a.flags;
        const int slot = a.out_slot;
        const FinProg &fin = a.fin;
        const int defer = a.defer_fin;
        grid_sum<(NDOT > 0 ? NDOT : 1)>(acc, a.partials, a.ticket, [&](double(&s)[NDOT > 0 ? NDOT : 1]) {
#pragma unroll
            for (int k = 0; k < NDOT; k++) scal[slot + k] = s[k];
            if (!defer) fin_run(fin, scal, flags);
        });
    }
}

// ---- bulk-copy pipeline variant (regular matrices: STREAM tiles only) ---------------------------
// Same tiles, same arithmetic (one thread per row, products added in storage order: bit-identical),
// but the tile's val / col / Ap segments arrive by cp.async.bulk (the TMA engine) into one of TWO
// shared-memory stages, signalled on an mbarrier: while the CTA computes tile k, tile k+1 is already
// in flight, no registers are tied up staging it and no thread waits on a load it issued itself.
// e0 of every tile comes from a small host-built array so that the copy of tile k+1 can be issued
// without having seen its Ap.
__device__ __forceinline__ unsigned int spmv_smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }

template <int KIND, int NDOT>
__global__ void __launch_bounds__(kBlock) spmv_pipe_kernel(const SpmvArgs a, const int *__restrict__ tile_e0)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (a.stop && *a.stop) return;
    const int tid = threadIdx.x;
    const size_t val_bytes = (size_t)(a.cap + 8) * sizeof(double);
    const size_t col_bytes = (size_t)(a.cap + 8) * sizeof(int);
    const size_t ap_bytes = (size_t)(kTileRows + 8) * sizeof(int);
    const size_t stage_bytes = (val_bytes + col_bytes + ap_bytes + 127) & ~(size_t)127;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw);   // [2]
    unsigned char *stage0 = smem_raw + 128;
    const double alpha = coef_get(a.alpha, a.scal);
    const double beta = coef_get(a.beta, a.scal);
    const double *__restrict__ x = a.x;
    double acc[NDOT > 0 ? NDOT : 1];
#pragma unroll
    for (int k = 0; k < (NDOT > 0 ? NDOT : 1); k++) acc[k] = 0.0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(spmv_smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(spmv_smem_u32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int tile, int b) {   // thread 0 only
        const int r0 = a.tile_row[tile], nr = a.tile_row[tile + 1] - r0;
        const int e0 = tile_e0[tile], e1 = tile_e0[tile + 1];
        const int a0 = e0 & ~3, cnt = (e1 - a0 + 3) & ~3;
        const int p0 = r0 & ~3, pcnt = (r0 + nr + 1 - p0 + 3) & ~3;
        unsigned char *st = stage0 + (size_t)b * stage_bytes;
        const unsigned int bs = spmv_smem_u32(bar + b);
        const unsigned int bytes = (unsigned int)cnt * 12u + (unsigned int)pcnt * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bs), "r"(bytes) : "memory");
        if (cnt > 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(spmv_smem_u32(st)), "l"(a.Ax + a0), "r"((unsigned int)cnt * 8u), "r"(bs) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(spmv_smem_u32(st + val_bytes)), "l"(a.Aj + a0), "r"((unsigned int)cnt * 4u), "r"(bs) : "memory");
        }
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(spmv_smem_u32(st + val_bytes + col_bytes)), "l"(a.Ap + p0), "r"((unsigned int)pcnt * 4u), "r"(bs) : "memory");
    };
    unsigned int phase[2] = {0, 0};
    if (tid == 0 && (int)blockIdx.x < a.num_tiles) issue(blockIdx.x, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, it++) {
        const int b = it & 1;
        if (tid == 0 && tile + (int)gridDim.x < a.num_tiles) issue(tile + gridDim.x, b ^ 1);
        {
            const unsigned int bs = spmv_smem_u32(bar + b);
            unsigned int ok = 0;
            while (!ok) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(bs), "r"(phase[b]) : "memory");
            }
            phase[b] ^= 1;
        }
        const unsigned char *st = stage0 + (size_t)b * stage_bytes;
        const double *sval = reinterpret_cast<const double *>(st);
        const int *scol = reinterpret_cast<const int *>(st + val_bytes);
        const int r0 = a.tile_row[tile], nr = a.tile_row[tile + 1] - r0;
        const int *sap = reinterpret_cast<const int *>(st + val_bytes + col_bytes) + (r0 & 3);
        const int a0 = sap[0] & ~3;
        for (int r = tid; r < nr; r += kBlock) {
            int k = sap[r] - a0;
            const int k1 = sap[r + 1] - a0;
            double sum = 0.0;
            for (; k + 4 <= k1; k += 4) {
                const double x0 = __ldg(x + scol[k]), x1 = __ldg(x + scol[k + 1]);
                const double x2 = __ldg(x + scol[k + 2]), x3 = __ldg(x + scol[k + 3]);
                const double p0 = x0 * sval[k], p1 = x1 * sval[k + 1];
                const double p2 = x2 * sval[k + 2], p3 = x3 * sval[k + 3];
                sum += p0; sum += p1; sum += p2; sum += p3;
            }
            for (; k < k1; k++) sum += __ldg(x + scol[k]) * sval[k];
            const double out = epilogue<KIND>(sum, alpha, beta, a.y, r0 + r);
            a.z[r0 + r] = out;
            dots_add<NDOT>(a, acc, r0 + r, out);
        }
        __syncthreads();   // stage b is free again: it is refilled by the issue of the next iteration
    }
    if (NDOT > 0 && a.seq == nullptr) {
        double *scal = a.scal;
        int *flags = a.flags;
        const int slot = a.out_slot;
        const FinProg &fin = a.fin;
        const int defer = a.defer_fin;
        grid_sum<(NDOT > 0 ? NDOT : 1)>(acc, a.partials, a.ticket, [&](double(&s)[NDOT > 0 ? NDOT : 1]) {
#pragma unroll
            for (int k = 0; k < NDOT; k++) scal[slot + k] = s[k];
            if (!defer) fin_run(fin, scal, flags);
        });
    }
}

// zero matrix: z = epilogue(0)
template <int KIND>
__global__ void __launch_bounds__(kBlock) spmv_zero_kernel(int n, Coef ca, Coef cb, const double *scal,
                                                            const double *y, double *z)
{
    const double alpha = coef_get(ca, scal), beta = coef_get(cb, scal);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // reference zero-matrix branches: mxy/amxy write 0 (src/mvops.cxx:110-114,145-149),
        // amxpby / amxpbyz scale y by beta (src/mvops.cxx:33-38, :72-76)
        if (KIND == LSSPG_MV_MXY || KIND == LSSPG_MV_AMXY) z[i] = 0.0;
        else z[i] = y[i] * beta;
    }
    (void)alpha;
}

// Persistent grid: exactly the CTAs that are resident at once (a grid sized for more than the
// register / shared-memory limit allows leaves a tail wave at a fraction of the occupancy).
template <int KIND, int NDOT, bool IRR>
static int launch_one(lsspg_ctx *ctx, const lsspg_csr *A, const SpmvArgs &args, size_t smem)
{
    int &per_sm = A->occupancy[KIND][NDOT];   // per matrix: the limit depends on its tile size (smem)
    if (per_sm == 0) {
        cudaFuncSetAttribute(spmv_tiles_kernel<KIND, NDOT, IRR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int nb = 0;
        LSSPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_tiles_kernel<KIND, NDOT, IRR>, kBlock, smem));
        per_sm = nb > 0 ? nb : 1;
    }
    int grid = std::min(A->num_tiles, ctx->num_sms * per_sm);
    if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
    LSSPG_LAUNCH(ctx, (spmv_tiles_kernel<KIND, NDOT, IRR>), grid, kBlock, smem, args);
    return 0;
}

template <int KIND, int NDOT>
static int launch_pipe(lsspg_ctx *ctx, const lsspg_csr *A, const SpmvArgs &args)
{
    const size_t stage = ((size_t)(args.cap + 8) * 12 + (size_t)(kTileRows + 8) * 4 + 127) & ~(size_t)127;
    const size_t smem = 128 + 2 * stage;
    int &per_sm = A->occupancy_pipe[KIND][NDOT];
    if (per_sm == 0) {
        cudaFuncSetAttribute(spmv_pipe_kernel<KIND, NDOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int nb = 0;
        LSSPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmv_pipe_kernel<KIND, NDOT>, kBlock, smem));
        per_sm = nb > 0 ? nb : 1;
    }
    int grid = std::min(A->num_tiles, ctx->num_sms * per_sm);
    if (grid > kMaxRedBlocks) grid = kMaxRedBlocks;
    LSSPG_LAUNCH(ctx, (spmv_pipe_kernel<KIND, NDOT>), grid, kBlock, smem, args, A->d_tile_e0);
    return 0;
}

template <int KIND>
static int launch_kind(lsspg_ctx *ctx, const lsspg_csr *A, const SpmvArgs &args, int ndot, size_t smem)
{
    // bulk-copy pipeline (default; LSSPG_OPT_SPMV_KERNEL = 1 selects the register-staged kernel): regular
    // matrices whose tiles are all STREAM tiles.  256^3 7-point: 6.76 TB/s against 5.74 TB/s.
    if (ctx->opt_spmv_kernel != 1 && !A->irregular && A->num_stream_tiles == A->num_tiles && A->d_tile_e0) {
        switch (ndot) {
            case 0: return launch_pipe<KIND, 0>(ctx, A, args);
            case 1: return launch_pipe<KIND, 1>(ctx, A, args);
            default: return launch_pipe<KIND, 2>(ctx, A, args);
        }
    }
    if (A->irregular) {
        switch (ndot) {
            case 0: return launch_one<KIND, 0, true>(ctx, A, args, smem);
            case 1: return launch_one<KIND, 1, true>(ctx, A, args, smem);
            default: return launch_one<KIND, 2, true>(ctx, A, args, smem);
        }
    }
    switch (ndot) {
        case 0: return launch_one<KIND, 0, false>(ctx, A, args, smem);
        case 1: return launch_one<KIND, 1, false>(ctx, A, args, smem);
        default: return launch_one<KIND, 2, false>(ctx, A, args, smem);
    }
}

int spmv_launch(lsspg_ctx *ctx, int kind, const lsspg_csr *A, Coef alpha, const double *dx, Coef beta,
                const double *dy, double *dz, const SpmvDots *dots, bool guarded)
{
    LSSPG_CHECK(A && dx && dz, "spmv: NULL operand");
    LSSPG_CHECK(kind >= 0 && kind <= 3, "spmv: bad kind %d", kind);
    LSSPG_CHECK(kind < 2 || dy != nullptr, "spmv: y required for kind %d", kind);
    LSSPG_CHECK(dx != dz, "spmv: x and z must not alias");
    const int n = A->num_rows;
    if (n == 0) return 0;
    // row shard of a distributed matrix: refresh the ghost tail of x (NCCL send/recv over NVLink)
    if (A->halo) LSSPG_TRY(halo_exchange(ctx, A->halo, const_cast<double *>(dx)));
    if (A->zero) {
        LSSPG_CHECK(!dots || dots->ndot == 0, "spmv: fused dots are not supported on the zero matrix");
        const int grid = stream_grid(ctx, n, kBlock);
        switch (kind) {
            case 0: LSSPG_LAUNCH(ctx, spmv_zero_kernel<0>, grid, kBlock, 0, n, alpha, beta, ctx->d_scal, dy, dz); break;
            case 1: LSSPG_LAUNCH(ctx, spmv_zero_kernel<1>, grid, kBlock, 0, n, alpha, beta, ctx->d_scal, dy, dz); break;
            case 2: LSSPG_LAUNCH(ctx, spmv_zero_kernel<2>, grid, kBlock, 0, n, alpha, beta, ctx->d_scal, dy, dz); break;
            default: LSSPG_LAUNCH(ctx, spmv_zero_kernel<3>, grid, kBlock, 0, n, alpha, beta, ctx->d_scal, dy, dz); break;
        }
        return 0;
    }
    SpmvArgs args;
    args.Ap = A->dAp; args.Aj = A->dAj; args.Ax = A->dAx;
    args.tile_row = A->d_tile_row; args.tile_kind = A->d_tile_kind;
    args.num_tiles = A->num_tiles;
    args.cap = (A->max_tile_nnz + 3) & ~3;
    args.x = dx; args.y = dy; args.z = dz;
    args.alpha = alpha; args.beta = beta;
    args.w0 = dots ? dots->w[0] : nullptr;
    args.w1 = dots ? dots->w[1] : nullptr;
    args.scal = ctx->d_scal; args.flags = ctx->d_flags;
    args.partials = ctx->d_partials; args.ticket = ctx->d_ticket;
    args.out_slot = dots ? dots->out_slot : 0;
    args.stop = guarded ? ctx->d_flags : nullptr;  // FLAG_STOP == 0
    args.seq = nullptr; args.seq_n = 0;
    args.defer_fin = distributed(ctx) ? 1 : 0;
    if (dots) args.fin = dots->fin;
    const int ndot = dots ? dots->ndot : 0;
    LSSPG_CHECK(ndot >= 0 && ndot <= 2, "spmv: ndot %d out of range", ndot);
    if (ndot > 0 && ctx->opt_reduce_sequential) {
        LSSPG_TRY(seq_prepare(ctx, n));
        args.seq = ctx->d_seq; args.seq_n = (long long)ctx->seq_len;
    }
    const size_t smem = (size_t)(args.cap + 8) * (sizeof(double) + sizeof(int)) + (kTileRows + 1) * sizeof(int);
    int rc;
    switch (kind) {
        case 0: rc = launch_kind<0>(ctx, A, args, ndot, smem); break;
        case 1: rc = launch_kind<1>(ctx, A, args, ndot, smem); break;
        case 2: rc = launch_kind<2>(ctx, A, args, ndot, smem); break;
        default: rc = launch_kind<3>(ctx, A, args, ndot, smem); break;
    }
    if (rc || ndot == 0 || !(ctx->opt_reduce_sequential || distributed(ctx))) return rc;
    RedOut o;
    o.out_slot = dots->out_slot; o.fin = dots->fin; o.guarded = guarded;
    return seq_finish(ctx, n, ndot, o);
}

// Host-side row-tile partition (O(n)); see the header comment for the rules.
static void build_tiles(int n, const int *Ap, bool exact, std::vector<int> &rows, std::vector<unsigned char> &kinds,
                        int &max_nnz, int &nstream)
{
    rows.clear();
    kinds.clear();
    max_nnz = 0;
    nstream = 0;
    auto len = [&](int i) { return Ap[i + 1] - Ap[i]; };
    auto is_huge = [&](int i) { return len(i) > kTileNnzCap - 4; };   // does not fit a tile (4 = alignment slack)
    int i = 0;
    while (i < n) {
        int j = i;
        if (is_huge(i)) {
            const int lim = exact ? kTileRows : kWarpRowsPerTile;
            while (j < n && j - i < lim && is_huge(j)) j++;
            kinds.push_back(exact ? TILE_SERIAL : TILE_BLOCK);
        }
        else {
            int cnt = 0, longest = 0;
            bool has_long = false;
            while (j < n && j - i < kTileRows && !is_huge(j) && cnt + len(j) <= kTileNnzCap - 4) {
                cnt += len(j);
                longest = std::max(longest, len(j));
                has_long |= len(j) > kLongRow;
                j++;
            }
            // row-length histogram of the tile: rows much longer than the average would leave most
            // threads of a thread-per-row pass idle -> products first, balanced over the threads
            const bool irregular = longest > 2 * (cnt / std::max(j - i, 1)) + 4;
            // the tile is loaded from the previous multiple of 4 elements
            const int span = (Ap[j] - (Ap[i] & ~3) + 3) & ~3;
            max_nnz = std::max(max_nnz, span);
            // exact mode: every row of a tile is summed sequentially by one thread, however long
            const bool mixed = has_long && !exact;
            kinds.push_back(mixed ? TILE_MIXED : (irregular || has_long) ? TILE_BALANCED : TILE_STREAM);
            if (!mixed) nstream++;
        }
        rows.push_back(i);
        i = j;
    }
    rows.push_back(n);
}

__global__ void __launch_bounds__(256) k_check_cols(long long nnz, const int *__restrict__ j, int m, int *bad)
{
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x)
        if (j[e] < 0 || j[e] >= m) *bad = 1;
}

}  // namespace lsspg

using namespace lsspg;

extern "C" {

int lsspg_csr_upload(lsspg_ctx *ctx, int num_rows, int num_cols, const int *hAp, const int *hAj, const double *hAx,
                     lsspg_csr **out)
{
    LSSPG_CHECK(ctx && out, "lsspg_csr_upload: NULL argument");
    LSSPG_CHECK(num_rows >= 0 && num_cols >= 0, "lsspg_csr_upload: negative dimension");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    lsspg_csr *A = new lsspg_csr();
    A->num_rows = num_rows;
    A->num_cols = num_cols;
    if (hAp == nullptr) {
        A->zero = true;
        *out = A;
        return 0;
    }
    const int nnz = hAp[num_rows];
    LSSPG_CHECK(nnz >= 0 && hAp[0] == 0, "lsspg_csr_upload: malformed Ap");
    LSSPG_CHECK(nnz == 0 || (hAj && hAx), "lsspg_csr_upload: NULL Aj/Ax");
    A->num_nnzs = nnz;
    {   // a column index outside [0, num_cols) would become an out-of-bounds gather on the device
        const int np = host_threads();
        std::vector<char> bad(np, 0);
        parallel_ranges(nnz, [&](long long e0, long long e1, int p) {
            for (long long e = e0; e < e1; e++)
                if (hAj[e] < 0 || hAj[e] >= num_cols) { bad[p] = 1; return; }
        }, np);
        for (char b : bad) LSSPG_CHECK(!b, "lsspg_csr_upload: column index out of range");
    }
    std::vector<int> rows;
    std::vector<unsigned char> kinds;
    build_tiles(num_rows, hAp, ctx->opt_spmv_exact != 0, rows, kinds, A->max_tile_nnz, A->num_stream_tiles);
    A->num_tiles = (int)kinds.size();
    LSSPG_CUDA(cudaMalloc(&A->dAp, sizeof(int) * ((size_t)num_rows + 1 + 8)));   // + slack: 16-byte bulk copies
    LSSPG_CUDA(cudaMemsetAsync(A->dAp + num_rows + 1, 0, sizeof(int) * 8, ctx->stream));
    LSSPG_CUDA(cudaMalloc(&A->dAj, sizeof(int) * ((size_t)nnz + 16)));
    LSSPG_CUDA(cudaMalloc(&A->dAx, sizeof(double) * ((size_t)nnz + 16)));
    LSSPG_CUDA(cudaMemsetAsync(A->dAj + nnz, 0, sizeof(int) * 16, ctx->stream));
    LSSPG_CUDA(cudaMemsetAsync(A->dAx + nnz, 0, sizeof(double) * 16, ctx->stream));
    LSSPG_CUDA(cudaMalloc(&A->d_tile_row, sizeof(int) * rows.size()));
    LSSPG_CUDA(cudaMalloc(&A->d_tile_kind, std::max<size_t>(kinds.size(), 1)));
    LSSPG_CUDA(cudaMemcpyAsync(A->dAp, hAp, sizeof(int) * ((size_t)num_rows + 1), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz) {
        LSSPG_CUDA(cudaMemcpyAsync(A->dAj, hAj, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(A->dAx, hAx, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    }
    LSSPG_CUDA(cudaMemcpyAsync(A->d_tile_row, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!kinds.empty())
        LSSPG_CUDA(cudaMemcpyAsync(A->d_tile_kind, kinds.data(), kinds.size(), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> e0(rows.size());
    for (size_t t = 0; t < rows.size(); t++) e0[t] = hAp[rows[t]];
    LSSPG_CUDA(cudaMalloc(&A->d_tile_e0, sizeof(int) * e0.size()));
    LSSPG_CUDA(cudaMemcpyAsync(A->d_tile_e0, e0.data(), sizeof(int) * e0.size(), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (unsigned char k : kinds) A->irregular |= (k == TILE_MIXED || k == TILE_BALANCED || k == TILE_BLOCK);
    *out = A;
    return 0;
}

/* The SpMV matrix of a device-resident CSR (setup_gpu.cuh): only the row pointer visits the host (the row-tile schedule is
 * built there, 4 (n + 1) bytes); Aj / Ax stay where they are -- adopted when take != 0 (M is emptied), copied device to
 * device otherwise.  Column indices are range-checked by a kernel. */
int lsspg_csr_from_dmat(lsspg_ctx *ctx, lsspg_dmat *M, int take, lsspg_csr **out)
{
    LSSPG_CHECK(ctx && M && out && M->bs == 1 && M->p, "lsspg_csr_from_dmat: needs a device CSR matrix");
    LSSPG_CUDA(cudaSetDevice(ctx->device));
    const int n = M->n;
    const long long nnz = M->nnz;
    std::vector<int> hp((size_t)n + 1);
    int *flag = ctx->d_flags + FLAG_SETUP;
    LSSPG_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
    if (nnz > 0) LSSPG_LAUNCH(ctx, k_check_cols, stream_grid(ctx, nnz, 256), 256, 0, nnz, M->j, M->m, flag);
    int bad = 0;
    LSSPG_CUDA(cudaMemcpyAsync(hp.data(), M->p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    LSSPG_CHECK(!bad, "lsspg_csr_from_dmat: column index out of range");
    LSSPG_CHECK(hp[0] == 0 && hp[n] == nnz, "lsspg_csr_from_dmat: malformed row pointer");
    lsspg_csr *A = new lsspg_csr();
    A->num_rows = n; A->num_cols = M->m; A->num_nnzs = (int)nnz;
    std::vector<int> rows;
    std::vector<unsigned char> kinds;
    build_tiles(n, hp.data(), ctx->opt_spmv_exact != 0, rows, kinds, A->max_tile_nnz, A->num_stream_tiles);
    A->num_tiles = (int)kinds.size();
    if (take) {
        A->dAp = M->p; A->dAj = M->j; A->dAx = M->x;   // allocated with the slack the kernels expect (dmat_alloc)
        M->p = nullptr; M->j = nullptr; M->x = nullptr; M->nnz = 0; M->n = 0;
    }
    else {
        LSSPG_CUDA(cudaMalloc(&A->dAp, sizeof(int) * ((size_t)n + 1 + 8)));
        LSSPG_CUDA(cudaMalloc(&A->dAj, sizeof(int) * ((size_t)nnz + 16)));
        LSSPG_CUDA(cudaMalloc(&A->dAx, sizeof(double) * ((size_t)nnz + 16)));
        LSSPG_CUDA(cudaMemcpyAsync(A->dAp, M->p, sizeof(int) * ((size_t)n + 1 + 8), cudaMemcpyDeviceToDevice, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(A->dAj, M->j, sizeof(int) * ((size_t)nnz + 16), cudaMemcpyDeviceToDevice, ctx->stream));
        LSSPG_CUDA(cudaMemcpyAsync(A->dAx, M->x, sizeof(double) * ((size_t)nnz + 16), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    LSSPG_CUDA(cudaMalloc(&A->d_tile_row, sizeof(int) * rows.size()));
    LSSPG_CUDA(cudaMalloc(&A->d_tile_kind, std::max<size_t>(kinds.size(), 1)));
    LSSPG_CUDA(cudaMemcpyAsync(A->d_tile_row, rows.data(), sizeof(int) * rows.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!kinds.empty())
        LSSPG_CUDA(cudaMemcpyAsync(A->d_tile_kind, kinds.data(), kinds.size(), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> e0(rows.size());
    for (size_t t = 0; t < rows.size(); t++) e0[t] = hp[rows[t]];
    LSSPG_CUDA(cudaMalloc(&A->d_tile_e0, sizeof(int) * e0.size()));
    LSSPG_CUDA(cudaMemcpyAsync(A->d_tile_e0, e0.data(), sizeof(int) * e0.size(), cudaMemcpyHostToDevice, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    for (unsigned char k : kinds) A->irregular |= (k == TILE_MIXED || k == TILE_BALANCED || k == TILE_BLOCK);
    *out = A;
    return 0;
}

int lsspg_csr_destroy(lsspg_ctx *ctx, lsspg_csr *A)
{
    if (!A) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(A->dAp);
    cudaFree(A->dAj);
    cudaFree(A->dAx);
    cudaFree(A->d_tile_row);
    cudaFree(A->d_tile_kind);
    cudaFree(A->d_tile_e0);
    delete A;
    return 0;
}

int lsspg_csr_dims(const lsspg_csr *A, int *num_rows, int *num_cols, int *num_nnzs)
{
    if (num_rows) *num_rows = A->num_rows;
    if (num_cols) *num_cols = A->num_cols;
    if (num_nnzs) *num_nnzs = A->num_nnzs;
    return 0;
}

int lsspg_csr_schedule_info(const lsspg_csr *A, int *num_tiles, int *num_stream_tiles, int *max_tile_nnz)
{
    if (num_tiles) *num_tiles = A->num_tiles;
    if (num_stream_tiles) *num_stream_tiles = A->num_stream_tiles;
    if (max_tile_nnz) *max_tile_nnz = A->max_tile_nnz;
    return 0;
}

double lsspg_csr_spmv_bytes(const lsspg_csr *A)
{
    return 12.0 * A->num_nnzs + 4.0 * (A->num_rows + 1.0) + 16.0 * A->num_rows;
}

int lsspg_mv(lsspg_ctx *ctx, int kind, const lsspg_csr *A, double alpha, const double *dx, double beta,
             const double *dy, double *dz)
{
    return spmv_launch(ctx, kind, A, coef_imm(alpha), dx, coef_imm(beta), dy, dz, nullptr);
}

int lsspg_mv_host(lsspg_ctx *ctx, int kind, const lsspg_csr *A, double alpha, const double *hx, double beta,
                  const double *hy, double *hz)
{
    LSSPG_CHECK(A && hx && hz, "lsspg_mv_host: NULL operand");
    const size_t nr = A->num_rows, nc = A->num_cols;
    LSSPG_TRY(ensure_stage(ctx, std::max(nr, nc)));
    double *dx = ctx->stage[0], *dy = ctx->stage[1], *dz = ctx->stage[2];
    LSSPG_CUDA(cudaMemcpyAsync(dx, hx, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (kind >= 2) {
        LSSPG_CHECK(hy != nullptr, "lsspg_mv_host: y required");
        LSSPG_CUDA(cudaMemcpyAsync(dy, hy, nr * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    LSSPG_TRY(lsspg_mv(ctx, kind, A, alpha, dx, beta, dy, dz));
    LSSPG_CUDA(cudaMemcpyAsync(hz, dz, nr * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LSSPG_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"

/* lsspg.h -- C ABI of the B200-native LSSP solve-loop library (liblsspg.so).
 *
 * This is the drop-in boundary for the hot path of huiscliu/lssp: CSR SpMV and
 * residual, BLAS-1, preconditioner application and the Krylov drivers.  Plain
 * pointers and sizes only; no C++ or torch types.  The C++ facade in
 * include/lssp/ (signature-identical to the reference's lssp.h / mvops.h /
 * vector.h / solver-*.h / pc-*.h) and the Python host mirror (lssp_b200/) both
 * sit on top of exactly these entry points.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     lsspg_last_error() gives the message (CUDA error string or argument
 *     check).  The C++ facade maps a failure to lssp_error(1, ...) which
 *     prints and exit()s, as the reference does (src/utils.cxx:114-135).
 *   - "h" pointers are host memory, "d" pointers are device memory obtained
 *     from lsspg_malloc() on the same context.
 *   - fp64 values, int32 indices, as in the reference (include/type-defs.h:15-24).
 *   - There is NO CPU fallback: without a CUDA device lsspg_ctx_create fails.
 *
 * Each entry point cites the reference interface (file:line under
 * /root/reference) that it replaces.
 */
#ifndef LSSPG_H
#define LSSPG_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsspg_ctx lsspg_ctx;   /* one device, one stream, reduction scratch        */
typedef struct lsspg_csr lsspg_csr;   /* device-resident CSR matrix + SpMV row-tile schedule */
typedef struct lsspg_tri lsspg_tri;   /* device-resident triangular factor, level-ordered  */
typedef struct lsspg_pc  lsspg_pc;    /* preconditioner application object                 */
typedef struct lsspg_dmat lsspg_dmat; /* device-resident CSR / BCSR matrix of the set-up path (no SpMV schedule) */

const char *lsspg_last_error(void);
const char *lsspg_version(void);

/* ---- context / memory ---------------------------------------------------- */
int lsspg_ctx_create(int device, lsspg_ctx **out);
int lsspg_ctx_destroy(lsspg_ctx *ctx);
int lsspg_sync(lsspg_ctx *ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on */
void *lsspg_ctx_stream(lsspg_ctx *ctx);
/* CUDA-event timer on that stream (slot 0..3): milliseconds between start and stop */
/* Where driver messages go (per-iteration residual lines, breakdown notices): the facade registers the reference's
 * lssp_printf (src/utils.cxx:93-112), so that lssp_solver_set_log files receive them.  NULL: stdout, flushed. */
void lsspg_set_printer(void (*fn)(const char *msg));
int lsspg_timer_start(lsspg_ctx *ctx, int slot);
int lsspg_timer_stop(lsspg_ctx *ctx, int slot, double *ms);
/* number of kernels this context has launched so far (bench.py: gpu_launches) */
long long lsspg_ctx_launches(lsspg_ctx *ctx);
/* options: see LSSPG_OPT_* */
int lsspg_ctx_set_option(lsspg_ctx *ctx, int option, int value);

#define LSSPG_OPT_SPMV_KERNEL   1   /* 0 auto, 1 stream (LDG staging), 2 stream (bulk-copy pipeline), 3 vector only */
#define LSSPG_OPT_SPMV_EXACT    2   /* 1: never use the shuffle-reduced long-row path, in the SpMV and in the AMG
                                       smoother (rows of more than 64 entries): bit-exact always */
#define LSSPG_OPT_CHECK_EVERY   3   /* CG / BiCGStab: the host reads the residuals back every k iterations (default 8;
                                       a device-side stop flag freezes the state at convergence, so the iteration
                                       count, history and solution do not depend on k) */
#define LSSPG_OPT_REDUCE_SEQUENTIAL 4 /* every dot/norm equals the reference's sequential sum (src/vector.cxx:129)
                                       bit for bit -> whole solves become bit-identical to the CPU reference.
                                       1: one thread does the adds (verification only, ~4 ns per term);
                                       2: the same result computed in parallel (exact_sum.cu: binade-wise exact
                                          integer sums, term-by-term only across binade crossings and ties) */
#define LSSPG_OPT_GRAPHS 5            /* 1 (default): CG replays the launch train of a batch of iterations as a CUDA graph */

int lsspg_malloc(lsspg_ctx *ctx, size_t bytes, void **dptr);
int lsspg_free(lsspg_ctx *ctx, void *dptr);
int lsspg_h2d(lsspg_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int lsspg_d2h(lsspg_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int lsspg_memset_zero(lsspg_ctx *ctx, void *dptr, size_t bytes);
/* pinned host staging memory (for the host-buffer entry points and bench.py) */
int lsspg_host_alloc(size_t bytes, void **hptr);
int lsspg_host_free(void *hptr);

/* ---- CSR matrix (replaces lssp_mat_csr, include/type-defs.h:15-24) -------- */
/* Upload a host CSR matrix (copied; the caller keeps its arrays) and build the
 * row-tile schedule used by the SpMV kernels.  hAp == NULL gives the
 * reference's "zero matrix" (src/mvops.cxx:33-38). */
int lsspg_csr_upload(lsspg_ctx *ctx, int num_rows, int num_cols, const int *hAp,
                     const int *hAj, const double *hAx, lsspg_csr **out);
int lsspg_csr_destroy(lsspg_ctx *ctx, lsspg_csr *A);
int lsspg_csr_dims(const lsspg_csr *A, int *num_rows, int *num_cols, int *num_nnzs);
/* schedule introspection for tests: number of row tiles and how many of them
 * take the row-sequential (bit-exact) path */
int lsspg_csr_schedule_info(const lsspg_csr *A, int *num_tiles, int *num_stream_tiles,
                            int *max_tile_nnz);
/* algorithmic bytes of one y = A x: 12 nnz + 4 (n+1) + 16 n */
double lsspg_csr_spmv_bytes(const lsspg_csr *A);

/* ---- device-side ingest & matrix utilities (matops_gpu.cu) ---------------------------------
 * What lssp_solver_assemble and the preconditioner set-up do to the caller's matrix, on matrices that live in device
 * memory: results are byte-identical to the host utilities of liblssp.so (include/lssp/matrix-utils.h), which are
 * pinned against the reference. */
int lsspg_dmat_upload(lsspg_ctx *ctx, int num_rows, int num_cols, const int *hAp, const int *hAj, const double *hAx,
                      lsspg_dmat **out);
int lsspg_dmat_download(lsspg_ctx *ctx, const lsspg_dmat *M, int *hAp, int *hAj, double *hAx);   /* NULL: skip that array */
int lsspg_dmat_dims(const lsspg_dmat *M, int *num_rows, int *num_cols, long long *num_nnzs, int *blk_size);
int lsspg_dmat_destroy(lsspg_ctx *ctx, lsspg_dmat *M);
int lsspg_dmat_copy(lsspg_ctx *ctx, const lsspg_dmat *A, lsspg_dmat **out);            /* deep copy, src/lssp.cxx:169-171 */
int lsspg_dmat_is_sorted(lsspg_ctx *ctx, const lsspg_dmat *A, int *sorted);           /* lssp_mat_csr_is_sorted, src/matrix-utils.cxx:249-279 */
int lsspg_dmat_sort_columns(lsspg_ctx *ctx, lsspg_dmat *A);                           /* lssp_mat_sort_column, :387-481 (stable) */
int lsspg_dmat_adjust_zero_diag(lsspg_ctx *ctx, const lsspg_dmat *A, double tol, lsspg_dmat **out);   /* :483-587 */
int lsspg_dmat_get_block_diag(lsspg_ctx *ctx, const lsspg_dmat *A, int blk_size, lsspg_dmat **out);   /* :589-698 */
int lsspg_dmat_to_bcsr(lsspg_ctx *ctx, const lsspg_dmat *A, int blk_size, lsspg_dmat **out);          /* lssp_mat_csr_to_bcsr, :62-162 */
/* rows [r0, r1) of the 7-point operator on an nx x ny x nz grid in natural order (nz == 1: the 5-point operator of
 * example/exam.cxx:4-59), columns ascending, minus col_shift; stencil[7] = the values at -nx ny, -nx, -1, 0, +1, +nx,
 * +nx ny (byte-identical to lssp_b200/generators.py) */
int lsspg_dmat_gen_stencil(lsspg_ctx *ctx, int nx, int ny, int nz, long long r0, long long r1, long long col_shift,
                           const double *stencil, lsspg_dmat **out);
/* the SpMV matrix of a device-resident CSR: only the row pointer visits the host (row-tile schedule); take != 0 adopts
 * M's arrays (M is left empty), else they are copied device to device */
int lsspg_csr_from_dmat(lsspg_ctx *ctx, lsspg_dmat *M, int take, lsspg_csr **out);

/* ---- mvops (replaces include/mvops.h:8-19, src/mvops.cxx) ------------------ */
#define LSSPG_MV_MXY      0   /* z = A x                  lssp_mv_mxy      src/mvops.cxx:118-150 */
#define LSSPG_MV_AMXY     1   /* z = (A x) * a            lssp_mv_amxy     src/mvops.cxx:81-115  */
#define LSSPG_MV_AMXPBY   2   /* z = (A x) * a + y * b    lssp_mv_amxpby   src/mvops.cxx:5-39  (reference: in place, z == y) */
#define LSSPG_MV_AMXPBYZ  3   /* z = y * b + a * (A x)    lssp_mv_amxpbyz  src/mvops.cxx:42-78   */
/* device-resident operands */
int lsspg_mv(lsspg_ctx *ctx, int kind, const lsspg_csr *A, double alpha, const double *dx,
             double beta, const double *dy, double *dz);
/* host operands (the reference-facing call: lssp_vec lives in host memory);
 * copies x (and y) up, runs the kernel, copies z back */
int lsspg_mv_host(lsspg_ctx *ctx, int kind, const lsspg_csr *A, double alpha, const double *hx,
                  double beta, const double *hy, double *hz);

/* ---- vector (replaces include/vector.h:7-40, src/vector.cxx) --------------- */
int lsspg_vec_set(lsspg_ctx *ctx, int n, double *dx, double val);                       /* :31-38  */
int lsspg_vec_copy(lsspg_ctx *ctx, int n, double *ddst, const double *dsrc);            /* :74-84  */
int lsspg_vec_axy(lsspg_ctx *ctx, int n, double a, const double *dx, double *dy);       /* :86-96   y = x*a       */
int lsspg_vec_axpby(lsspg_ctx *ctx, int n, double a, const double *dx, double b, double *dy);  /* :98-108  y = y*b + x*a */
int lsspg_vec_axpbyz(lsspg_ctx *ctx, int n, double a, const double *dx, double b,
                     const double *dy, double *dz);                                     /* :110-121 z = y*b + x*a */
int lsspg_vec_scale(lsspg_ctx *ctx, int n, double *dx, double a);                       /* :141-146 */
int lsspg_vec_dot(lsspg_ctx *ctx, int n, const double *dx, const double *dy, double *h_out);   /* :123-133 */
int lsspg_vec_norm(lsspg_ctx *ctx, int n, const double *dx, double *h_out);             /* :135-138 */
/* k dot products x_i . y in one pass over y (k <= 8); results to h_out[0..k) */
int lsspg_vec_multidot(lsspg_ctx *ctx, int n, int k, const double *const *dxs, const double *dy,
                       double *h_out);

/* ---- sparse triangular solves (replaces include/solver-tri.h:8-12) -------- */
/* Analyse a host triangular factor stored as the reference stores it
 * (lower: diagonal LAST in each row, src/solver-tri.cxx:4-24; upper: diagonal
 * FIRST, off-diagonals applied in DESCENDING storage order, :26-46), compute
 * the dependency levels, and upload it in level order (sliced-ELL, 32-row
 * slices).  The block-ILU factors use the same layout (src/pc-biluk.cxx:37-59). */
#define LSSPG_TRI_LOWER 0
#define LSSPG_TRI_UPPER 1
int lsspg_tri_analyse(lsspg_ctx *ctx, int which, int n, const int *hTp,
                      const int *hTj, const double *hTx, lsspg_tri **out);
int lsspg_tri_destroy(lsspg_ctx *ctx, lsspg_tri *T);
int lsspg_tri_info(const lsspg_tri *T, int *num_levels, int *num_slices, long long *padded_nnz);
/* x = T^-1 rhs (device operands; x and rhs must not alias) */
/* Synchronises the stream and returns non-zero (lsspg_last_error) when a sweep since the last check was aborted by
 * the watchdog; the flag is cleared.  The *_host entry points and the Krylov drivers check by themselves. */
int lsspg_check_flags(lsspg_ctx *ctx);
int lsspg_tri_solve(lsspg_ctx *ctx, const lsspg_tri *T, double *dx, const double *drhs);
/* host-side schedule for tests: level of every row (length n) */
int lsspg_tri_levels_host(int which, int n, const int *hTp, const int *hTj,
                          int *h_level, int *num_levels);

/* Layout self-check for the CPU test-suite: builds the level-ordered sliced-ELL
 * image on the host and walks it slice by slice in ticket order.  NOT a
 * fallback: no driver, preconditioner or facade function ever calls it. */
int lsspg_debug_tri_walk_layout_host(int which, int n, const int *hTp, const int *hTj,
                                     const double *hTx, double *hx, const double *hrhs,
                                     int *num_slices, long long *padded_nnz);

int lsspg_debug_tri_walk_tiled_host(int which, int n, const int *hTp, const int *hTj,
                                    const double *hTx, double *hx, const double *hrhs,
                                    int *applicable, int *info /* [8]: boxes, box levels, max rows
                                    per box, row levels, grid nx ny nz, box edge */);
/* Fingerprint (64-bit hash) and size of the device image lsspg_tri_analyse would upload for this factor,
 * built on the host without any CUDA call; kind: 0 slice schedule, 1 box blobs, 2 ELL box blobs.
 * For the CPU test-suite (pins the set-up code's output) and set-up profiling; seconds[2] = schedule, packing. */
int lsspg_debug_tri_pack_host(int which, int n, const int *hTp, const int *hTj, const double *hTx,
                              int *kind, unsigned long long *fingerprint, long long *bytes,
                              double *seconds);
/* CPU emulation of the ELL box kernels from the PACKED blobs (the bytes the device reads), boxes advancing one
 * chunk per round: x must equal the serial sweep; info[4] = rounds (longest chain of hand-offs), chunks per box,
 * boxes, levels of the box graph.  Test-suite only (also verifies the experimental LSSPG_TRI_CHUNKS schedule). */
int lsspg_debug_tri_walk_packed_host(int which, int n, const int *hTp, const int *hTj, const double *hTx,
                                     double *hx, const double *hrhs, int *applicable, int *info);
/* CPU replay of the PENCIL schedule (tri_pencil.cu) from the image the device reads: value stream, line
 * descriptors, ghost lanes, mailboxes.  x must equal the serial sweep of src/solver-tri.cxx:4-46 bit for bit.
 * info[8]: pencils, threads per pencil, most ghost lines, largest operand distance, slots per row, values per
 * row, most steps, skew.  Test-suite only. */
int lsspg_debug_tri_walk_pencil_host(int which, int n, const int *hTp, const int *hTj, const double *hTx,
                                     double *hx, const double *hrhs, int *applicable, int *info);
/* LSSPG_TRI_PROF=1: per-pencil timers of the last sweep (8 words per pencil, ticket order: start ns, end ns, cycles
 * total / waiting for ghost lanes / in the step barrier, steps, CTA, SM); returns the number of pencils copied */
int lsspg_debug_tri_pencil_prof(lsspg_ctx *ctx, const lsspg_tri *T, unsigned long long *out, int max_pencils);
/* CPU replay of the parallel reference-order summation (exact_sum.cu, LSSPG_OPT_REDUCE_SEQUENTIAL = 2): *out must equal
 * `s = 0; for (i = 0; i < n; i++) s += t[i];` bit for bit.  stats[4]: blocks, rounds of the walk, blocks not advanced as a
 * whole, 32-term pieces added term by term.  Test-suite only. */
int lsspg_debug_exact_seq_sum_host(long long n, const double *t, double *out, long long *stats);
/* schedule of a device-resident factor: tiled = 2 pencil schedule, 1 box schedule, 0 slice schedule */
int lsspg_tri_schedule(const lsspg_tri *T, int *tiled, int *num_tiles, int *num_tile_levels,
                       int *max_tile_rows);

/* ---- host-side incomplete factorisations (setup; replaces src/pc-iluk.cxx,
 *      src/pc-ilut.cxx incl. lssp_mat_adjust_zero_diag / get_block_diag) ----- */
typedef struct lsspg_factors lsspg_factors;   /* host L and U in the reference's layout */
#define LSSPG_ILUK 0
#define LSSPG_ILUT 1
/* blk_size <= 0 or >= n: one global factorisation (what the reference always
 * does, src/pc-iluk.cxx:574); smaller: block-Jacobi with uniform blocks
 * (the reference's blocked driver, src/pc-iluk.cxx:411-552). */
int lsspg_ilu_factor(int kind, int n, const int *hAp, const int *hAj, const double *hAx,
                     int level, int p, double tol, int blk_size, lsspg_factors **out);
/* ILU(k) set-up ON THE GPU (ilu_gpu.cu): ingest (column sort, missing diagonals), the symbolic
 * level-of-fill phase (src/pc-iluk.cxx:22-135, :279-345, level-raising rule :101), the block restriction (:441-446), the
 * numeric IKJ phase (:347-409) and the L / U split (:501-532) run on the device -- one persistent kernel per phase, a
 * thread per row, rows wait for the finished rows of their lower columns; same operations in the same order as the host
 * loops -> bit-identical factors.  Only the finished factors come back to the host. */
int lsspg_ilu_factor_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx,
                            int level, int blk_size, lsspg_factors **out);
/* the same from a matrix that is already in device memory (see "device-side ingest" below) */
int lsspg_ilu_factor_dmat(lsspg_ctx *ctx, const lsspg_dmat *A, int level, int blk_size, lsspg_factors **out);
/* ILUT(p, tol) set-up on the GPU (src/pc-ilut.cxx:51-286, :429-456): the dual-threshold row recurrence incl. the
 * reference's quick-select (:7-49, the stored order of the kept entries is part of the result), one thread per row in a
 * persistent kernel as above; p <= 0 and tol < 0 select the reference's defaults (:436-442).  Bit-identical factors. */
int lsspg_ilut_factor_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx, int p, double tol,
                             int blk_size, lsspg_factors **out);
int lsspg_ilut_factor_dmat(lsspg_ctx *ctx, const lsspg_dmat *A, int p, double tol, int blk_size, lsspg_factors **out);
/* CPU replay of the device factorisations (the row functions of ilu_rows.cuh that the kernels run, rows in ascending
 * order): kind 0 ILU(k), 1 ILUT; *applicable = 0 when the input is not sorted / lacks diagonals (the device path repairs
 * that first).  The factors must equal lsspg_ilu_factor's bit for bit.  Test-suite only. */
int lsspg_debug_ilu_gpu_replay_host(int kind, int n, const int *hAp, const int *hAj, const double *hAx, int level, int p,
                                    double tol, int blk_size, int *applicable, lsspg_factors **out);
int lsspg_factors_sizes(const lsspg_factors *F, int *n, int *nnzL, int *nnzU);
int lsspg_factors_get(const lsspg_factors *F, int *Lp, int *Lj, double *Lx, int *Up, int *Uj,
                      double *Ux);
int lsspg_factors_destroy(lsspg_factors *F);

/* Block ILU(k) set-up (reference lssp_pc_biluk_assemble / _assemble_mat, src/pc-biluk.cxx:377-431; there only
 * `#if USE_BLAS && USE_LAPACK`): blocks of n / num_blks rows (s.num_blks, src/pc-biluk.cxx:427), CSR -> BCSR
 * (src/matrix-utils.cxx:62-162), level-of-fill symbolic phase on the block graph, block IKJ with dense block
 * products and inverted pivot blocks.  Result: L (strict lower blocks, unit diagonal last), D (block diagonal of
 * the inverted pivot blocks), U (inverse pivot times the upper blocks, unit diagonal first) -- the operands of
 * lsspg_pc_create_bilu.  The dense kernels are the netlib reference DGEMM / DGETF2 / DGETRI (unblocked). */
typedef struct lsspg_bfactors lsspg_bfactors;
int lsspg_bilu_factor(int n, const int *hAp, const int *hAj, const double *hAx, int num_blks,
                      int level, lsspg_bfactors **out);
int lsspg_bfactors_sizes(const lsspg_bfactors *F, int *n, int *nnzL, int *nnzD, int *nnzU);
int lsspg_bfactors_get(const lsspg_bfactors *F, int *Lp, int *Lj, double *Lx, int *Dp, int *Dj,
                       double *Dx, int *Up, int *Uj, double *Ux);
int lsspg_bfactors_destroy(lsspg_bfactors *F);

/* Number of host threads the set-up code (factorisations, schedules, packing) runs on: LSSPG_HOST_THREADS, default
 * the cores of the process' affinity mask divided by LOCAL_WORLD_SIZE, at most 32.  Results do not depend on it. */
int lsspg_host_threads(void);

/* ---- preconditioner application (replaces LSSP_PC.solve,
 *      include/type-defs.h:104,144) ------------------------------------------ */
#define LSSPG_PC_NON   0   /* x = rhs                          src/pc.cxx:67-70            */
#define LSSPG_PC_ILU   1   /* x = U^-1 L^-1 rhs                src/solver-tri.cxx:48-60    */
#define LSSPG_PC_BILU  2   /* x = U^-1 D L^-1 rhs              src/pc-biluk.cxx:22-60      */
#define LSSPG_PC_AMG   3   /* one SX-AMG-style V-cycle from x  src/pc-sxamg.cxx:42-73      */
#define LSSPG_PC_USER  4   /* host callback (LSSP_PC_USER)     src/pc.cxx:219-227          */
int lsspg_pc_create_non(lsspg_ctx *ctx, int n, lsspg_pc **out);
/* takes host L/U in the reference layout (as produced by lsspg_ilu_factor) */
int lsspg_pc_create_ilu(lsspg_ctx *ctx, int n, const int *Lp, const int *Lj, const double *Lx,
                        const int *Up, const int *Uj, const double *Ux, lsspg_pc **out);
int lsspg_pc_create_bilu(lsspg_ctx *ctx, int n, const int *Lp, const int *Lj, const double *Lx,
                         const int *Dp, const int *Dj, const double *Dx, const int *Up,
                         const int *Uj, const double *Ux, lsspg_pc **out);
/* user-supplied host preconditioner: fn(user, x, rhs, n) with HOST vectors; x holds the incoming
 * contents on entry.  Costs a device<->host round trip per application. */
int lsspg_pc_create_user(lsspg_ctx *ctx, int n, void (*fn)(void *user, double *hx, const double *hrhs, int n),
                         void *user, lsspg_pc **out);
int lsspg_pc_destroy(lsspg_ctx *ctx, lsspg_pc *pc);
int lsspg_pc_kind(const lsspg_pc *pc);
int lsspg_pc_info(const lsspg_pc *pc, int *levels_L, int *levels_U, long long *padded_L,
                  long long *padded_U);
/* algorithmic bytes of one application */
double lsspg_pc_bytes(const lsspg_pc *pc);
int lsspg_pc_apply(lsspg_ctx *ctx, lsspg_pc *pc, double *dx, const double *drhs);
int lsspg_pc_apply_host(lsspg_ctx *ctx, lsspg_pc *pc, double *hx, const double *hrhs);

/* ---- SX-AMG-style classical AMG (replaces the libsxamg calls of src/pc-sxamg.cxx and
 *      src/solver-sxamg.cxx; libsxamg itself is not in the reference tree -- the algorithm is
 *      specified in DESIGN.md "AMG", parity with libsxamg is UNPINNED) ---------------------- */
typedef struct lsspg_amg_pars {          /* the role of SX_AMG_PARS (sx_amg_pars_init) */
    int    max_levels;        /* 30  */
    int    coarse_dof;        /* 100: stop coarsening at or below this many unknowns */
    double strong_threshold;  /* 0.3 */
    double max_row_sum;       /* 0.9 */
    double trunc_threshold;   /* 0.2: interpolation truncation */
    int    pre_iter;          /* 2 Gauss-Seidel sweeps before restriction */
    int    post_iter;         /* 2 after prolongation */
    int    cf_order;          /* 1: sweeps visit C points then F points (pre) / F then C (post), ascending index
                                 inside a block; 0: natural; 2: as 1, but colour by colour inside a block
                                 (multicolour Gauss-Seidel: short dependency chains, made for the GPU) */
    int    zero_guess;        /* 0: the cycle starts from the incoming x, as the reference's adapter does
                                 (src/pc-sxamg.cxx:58-64); 1: from x = 0 (a fixed linear operator) */
    int    coarse_dense_max;  /* 4096: coarsest level solved with its dense inverse up to this size */
    int    coarse_sweeps;     /* 40 natural-order sweeps on a coarsest level larger than that */
    double tol;               /* 1e-8  stand-alone solver: stop when ||b - A x|| / ||b|| <= tol */
    int    maxit;             /* 100   stand-alone solver: most cycles */
    int    verb;
} lsspg_amg_pars;
int lsspg_amg_pars_default(lsspg_amg_pars *p);
typedef struct lsspg_amg_host lsspg_amg_host;   /* host image of the hierarchy */
/* Setup (host; GPU setup is SURVEY.md 8f row 2).  Columns of A must be sorted. */
int lsspg_amg_setup_host(int n, const int *hAp, const int *hAj, const double *hAx,
                         const lsspg_amg_pars *pars, lsspg_amg_host **out);
/* The same set-up with its per-row phases -- strong couplings, interpolation, restriction, Galerkin products -- on the GPU
 * (amg_gpu.cu); the Ruge-Stueben C/F splitting stays on the host.  Same hierarchy, array by array. */
int lsspg_amg_setup_device(lsspg_ctx *ctx, int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                           lsspg_amg_host **out);
/* CPU replay of the device set-up (amg_gpu.cu): the per-row phases (strong couplings, interpolation, restriction, Galerkin
 * products) through the row functions of amg_rows.cuh, row after row; the hierarchy must equal lsspg_amg_setup_host's array
 * by array.  Test-suite only. */
int lsspg_debug_amg_setup_replay_host(int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                                      lsspg_amg_host **out);
int lsspg_amg_host_levels(const lsspg_amg_host *H, int *num_levels, int *coarse_dense);
/* level l: unknowns, C points, nnz of A_l, of P_l (n_l x n_{l+1}) and of R_l = P_l^T; the last
 * level has nc = nnzP = nnzR = 0 */
int lsspg_amg_host_level_sizes(const lsspg_amg_host *H, int l, int *n, int *nc, int *nnzA,
                               int *nnzP, int *nnzR);
/* any pointer may be NULL; cf[i] = 1 for C points, 0 for F points */
int lsspg_amg_host_level_get(const lsspg_amg_host *H, int l, int *Ap, int *Aj, double *Ax,
                             int *Pp, int *Pj, double *Px, int *Rp, int *Rj, double *Rx, int *cf);
/* visiting rank of every point of level l inside its C / F block (cf_order 0/1: the index;
 * cf_order 2: multicolour order, see amg_host.cpp) */
int lsspg_amg_host_level_rank(const lsspg_amg_host *H, int l, int *rank);
/* row-major n x n inverse of the coarsest operator (only when coarse_dense != 0) */
int lsspg_amg_host_coarse_inverse(const lsspg_amg_host *H, double *inv);
int lsspg_amg_host_pars(const lsspg_amg_host *H, lsspg_amg_pars *pars);
int lsspg_amg_host_destroy(lsspg_amg_host *H);
/* CPU self-check of the smoother layout (as lsspg_debug_tri_walk_layout_host): walks the
 * level-ordered image of level l in ticket order, one sweep x_new <- GS(x_old).  post: bit 0 =
 * post-smoothing order; bits 1..2 = 0: the schedule the device would use, 1: slices, 2: rows. */
int lsspg_debug_amg_walk_gs_host(const lsspg_amg_host *H, int l, int post, const double *hb,
                                 const double *hx_old, double *hx_new, int *info /* [4]: slices,
                                 C-block levels, F-block levels, padded entries / 32 */);
/* device hierarchy; A0 (optional) is an already uploaded copy of the level-0 operator to share */
int lsspg_pc_create_amg(lsspg_ctx *ctx, const lsspg_amg_host *H, const lsspg_csr *A0, lsspg_pc **out);
/* stand-alone AMG iteration (lssp_solver_sxamg, src/solver-sxamg.cxx:25-99): cycles from x until
 * ||b - A x|| / ||b|| <= tol or maxit cycles; returns cycles in *nits, ||b - A x|| in *ares */
int lsspg_amg_solve(lsspg_ctx *ctx, lsspg_pc *amg, const double *db, double *dx, double tol, int maxit,
                    int *nits, double *ares);
int lsspg_amg_solve_host(lsspg_ctx *ctx, lsspg_pc *amg, const double *hb, double *hx, double tol,
                         int maxit, int *nits, double *ares);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink / NVSwitch -------------------------
 * The reference is serial; the path shards by contiguous row blocks of ceil(n/P) rows, the
 * block boundaries of lssp_mat_get_block_diag (src/matrix-utils.cxx:615,626-628), so that the
 * per-GPU ILU equals the reference's blocked ILU (src/pc-iluk.cxx:411-552) = block-Jacobi.
 * A rank holds its rows with columns renumbered [owned ; ghost]; every vector handed to
 * lsspg_mv / the drivers as SpMV input has num_cols = owned + ghost entries. */
typedef struct lsspg_halo lsspg_halo;
int lsspg_comm_unique_id(void *out128);                       /* rank 0: 128-byte NCCL id to broadcast */
int lsspg_comm_init(lsspg_ctx *ctx, int rank, int nranks, const void *id128);
/* One-shot peer-to-peer all-reduce for the dot products (k_p2p_allreduce_fin, comm.cu) instead of ncclAllReduce:
 * (1) every rank allocates its mailbox and returns its 64-byte CUDA IPC handle; (2) after the launcher has all-gathered
 * the handles (rank order) and passed a barrier, every rank maps its peers' mailboxes.  Optional: without it the sums
 * travel through NCCL. */
int lsspg_comm_p2p_local(lsspg_ctx *ctx, void *handle64);
int lsspg_comm_p2p_connect(lsspg_ctx *ctx, const void *handles);
int lsspg_comm_destroy(lsspg_ctx *ctx);
int lsspg_comm_size(lsspg_ctx *ctx, int *rank, int *nranks);
int lsspg_allreduce_sum(lsspg_ctx *ctx, double *d_buf, int count);   /* in place */
/* peers[p] receives x[send_idx[send_off[p] .. send_off[p+1])] (owned row indices) and sends the
 * recv_counts[p] entries that fill this rank's ghost segment, peers in the given order */
int lsspg_halo_create(lsspg_ctx *ctx, int n_owned, int npeers, const int *peers, const int *send_counts,
                      const int *h_send_idx, const int *recv_counts, lsspg_halo **out);
int lsspg_halo_destroy(lsspg_ctx *ctx, lsspg_halo *H);
int lsspg_halo_sizes(const lsspg_halo *H, int *n_owned, int *n_ghost, int *n_send);
int lsspg_halo_exchange(lsspg_ctx *ctx, const lsspg_halo *H, double *dx);
/* attach to a matrix uploaded with num_rows = owned, num_cols = owned + ghost: every SpMV
 * on it refreshes the ghost tail of x first */
int lsspg_csr_set_halo(lsspg_csr *A, lsspg_halo *H);

/* ---- Krylov drivers (replace int lssp_solver_<m>(LSSP_SOLVER&, LSSP_PC&),
 *      src/solver-*.cxx; numbering = LSSP_SOLVER_TYPE with every USE_* = 0,
 *      include/type-defs.h:156-174) ------------------------------------------ */
#define LSSPG_GMRES      0
#define LSSPG_LGMRES     1
#define LSSPG_RGMRES     2
#define LSSPG_RLGMRES    3
#define LSSPG_BICGSTAB   4
#define LSSPG_BICGSTABL  5
#define LSSPG_BICGSAFE   6
#define LSSPG_CG         7
#define LSSPG_CGS        8
#define LSSPG_GPBICG     9
#define LSSPG_CR        10
#define LSSPG_CRS       11
#define LSSPG_BICRSTAB  12
#define LSSPG_BICRSAFE  13
#define LSSPG_GPBICR    14
#define LSSPG_QMRCGSTAB 15
#define LSSPG_TFQMR     16
#define LSSPG_ORTHOMIN  17
#define LSSPG_IDRS      18

typedef struct lsspg_solver_opts {
    double tol_rel, tol_abs, tol_rb;   /* LSSP_SOLVER.tol_rel/abs/rb, include/type-defs.h:237-239 */
    int maxit, restart, aug_k, bgsl, idrs;
    int verb;                          /* >=1: per-iteration lines in the reference's format */
    int hist_len;                      /* capacity of hist (0: none) */
    double *hist;                      /* host: ||r|| after each iteration, full precision */
} lsspg_solver_opts;

typedef struct lsspg_solve_info {
    int nits;              /* return value of the reference driver                 */
    double residual;       /* LSSP_SOLVER.residual                                  */
    int hist_used;
    double solve_ms;       /* device time of the solve loop (CUDA events)           */
    long long launches;    /* kernels launched by this solve                        */
    int breakdown;         /* 1 when a breakdown branch of the reference was taken  */
} lsspg_solve_info;

int lsspg_solver_opts_default(lsspg_solver_opts *o);   /* src/lssp.cxx:5-14 defaults */
int lsspg_solver_supported(int solver);
/* device-resident b and x (x: initial guess in, solution out) */
int lsspg_krylov_solve(lsspg_ctx *ctx, int solver, const lsspg_csr *A, lsspg_pc *pc,
                       const double *db, double *dx, const lsspg_solver_opts *opts,
                       lsspg_solve_info *info);
/* host b and x, as the reference's lssp_solver_solve sees them (src/lssp.cxx:250):
 * uploads b and x, solves, downloads x */
int lsspg_krylov_solve_host(lsspg_ctx *ctx, int solver, const lsspg_csr *A, lsspg_pc *pc,
                            const double *hb, double *hx, const lsspg_solver_opts *opts,
                            lsspg_solve_info *info);

#ifdef __cplusplus
}
#endif
#endif /* LSSPG_H */

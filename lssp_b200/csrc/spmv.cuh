// spmv.cuh -- device CSR container and the fused SpMV launcher shared by
// mvops (spmv.cu) and the Krylov drivers.
#pragma once
#include "common.cuh"
#include "scalars.cuh"

struct lsspg_halo;

struct lsspg_csr {
    int num_rows = 0, num_cols = 0, num_nnzs = 0;
    bool zero = false;          // reference "Ap == NULL" zero-matrix branch
    int *dAp = nullptr;         // [num_rows + 1]
    int *dAj = nullptr;         // [num_nnzs] (+ slack)
    double *dAx = nullptr;      // [num_nnzs] (+ slack)
    // row-tile schedule (built on the host at upload)
    int num_tiles = 0;
    int num_stream_tiles = 0;   // tiles whose rows are all summed sequentially (bit-exact)
    int max_tile_nnz = 0;       // smem sizing of the stream kernel
    bool irregular = false;     // has BALANCED / MIXED / BLOCK tiles: needs the full kernel instantiation
    mutable int occupancy[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // resident CTAs / SM per (kind, ndot)
    int *d_tile_row = nullptr;  // [num_tiles + 1] first row of every tile
    unsigned char *d_tile_kind = nullptr;  // [num_tiles]
    int *d_tile_e0 = nullptr;              // [num_tiles + 1] first nnz of every tile (bulk-copy pipeline kernel)
    mutable int occupancy_pipe[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    lsspg_halo *halo = nullptr;            // row shard of a distributed matrix: ghost columns follow the owned ones
};

namespace lsspg {

constexpr int kTileRows = 256;     // rows per stream tile (one per thread)
constexpr int kTileNnzCap = 2048;  // nnz per stream tile
constexpr int kLongRow = 64;       // rows longer than this take the warp-per-row path
constexpr int kWarpRowsPerTile = 8;

enum TileKind : unsigned char { TILE_STREAM = 0, TILE_MIXED = 1, TILE_SERIAL = 2, TILE_BLOCK = 3, TILE_BALANCED = 4 };

// Fused SpMV: z = epilogue(A x) and up to two dot products of the result in the
// same pass:  sums[k] = sum_r z_r * (w_k ? w_k[r] : z_r).  The sums land in
// ctx->d_scal[out_slot + k]; `fin` then runs on the device (last CTA).
struct SpmvDots {
    int ndot = 0;
    const double *w[2] = {nullptr, nullptr};
    int out_slot = 0;
    FinProg fin;
};

int spmv_launch(lsspg_ctx *ctx, int kind, const lsspg_csr *A, Coef alpha, const double *dx, Coef beta,
                const double *dy, double *dz, const SpmvDots *dots, bool guarded = false);

}  // namespace lsspg

// tri.cuh -- level-scheduled sparse triangular solves (K5/K6 of SURVEY.md 2.2).
#pragma once
#include <vector>
#include "common.cuh"
#include "host_par.h"

namespace lsspg {

// Host image of one triangular factor in level order, sliced-ELL with 32-row
// slices (one slice = one warp-ticket).  Slot s*32+lane holds row perm[slot]
// (-1: empty).  Slice s owns columns [slice_ptr[s], slice_ptr[s+1]) of width-32
// panels: entry k of the row in `lane` sits at (slice_ptr[s] + k) * 32 + lane,
// in APPLICATION order (lower: ascending storage order; upper: descending
// storage order, as src/solver-tri.cxx:17-19 / :39-41).  col < 0 marks padding.
struct TriHost {
    int n = 0, which = 0;
    int num_levels = 0, num_slices = 0;
    long long padded_nnz = 0, offdiag_nnz = 0;
    std::vector<int> level;      // [n]
    // filled by the host threads (host_par.h: no serial zero-fill on resize)
    IVec perm;       // [num_slices*32]
    DVec diag;       // [num_slices*32] divisor of the row
    IVec slice_ptr;  // [num_slices+1]
    IVec slice_need; // [num_slices] progress hint: #slices that precede level(slice) - 1
    IVec col;        // [padded_nnz]
    DVec val;        // [padded_nnz]
};

// returns 0 on success; levels only when want_layout == false
int tri_build_host(int which, int n, const int *Tp, const int *Tj, const double *Tx, bool want_layout, TriHost &H);

}  // namespace lsspg

struct lsspg_tri {
    int n = 0, which = 0;
    int num_levels = 0, num_slices = 0;
    long long padded_nnz = 0, offdiag_nnz = 0;
    int *d_perm = nullptr;
    double *d_diag = nullptr;
    int *d_slice_ptr = nullptr;
    int *d_slice_need = nullptr;
    int *d_col = nullptr;
    double *d_val = nullptr;
    unsigned int *d_counter = nullptr;
    // tile schedule (tri_tiled.cu), used instead of the slice schedule when the factor comes
    // from a structured grid; all arrays in tile-major row order
    bool tiled = false;
    int num_tiles = 0, max_tile_rows = 0, num_tile_levels = 0, tile_dims[3] = {0, 0, 0}, grid_dims[3] = {0, 0, 0};
    unsigned char *t_blob = nullptr;   // packed boxes (see tri_tiled.cu)
    void *t_desc = nullptr;            // BoxDesc[num_tiles], ticket order
    int blob_cap = 0;                  // largest blob (bytes)
    int max_ext = 0;                   // most operands a box reads from other boxes
    bool box_flags = false;            // acyclic box graph: wait on per-box completion flags
    unsigned int *t_flags = nullptr;   // [num_tiles] epoch of the last sweep that finished the box
    unsigned int epoch = 0;
    // pencil schedule (tri_pencil.cu): lattice factors; num_tiles = pencils, tile_dims = pencil cross-section
    bool pencil = false;
    void *p_hdr = nullptr, *p_thr = nullptr, *p_ghost = nullptr, *p_outs = nullptr;
    double *p_vals = nullptr, *p_mail = nullptr;
    long long p_mail_len = 0;
    int p_T = 0, p_RS = 0, p_W = 0, p_diag = 0, p_dir = 1, p_pv = 0, p_pw = 0, p_holes = 0, p_own_last = 0;
    int p_seen_timeouts = 0;
    unsigned long long *p_prof = nullptr;
};

namespace lsspg {

// Host image of the tile schedule.
struct TiledHost {
    int n = 0, which = 0, num_tiles = 0, max_tile_rows = 0, num_tile_levels = 0, num_levels = 0;
    int tile_dims[3] = {0, 0, 0}, grid_dims[3] = {0, 0, 0};
    long long offdiag_nnz = 0;
    IVec perm, ptr, col;               // [n], [n+1], [offdiag_nnz]: filled by the host threads
    std::vector<int> tile_ptr, lev_off, lev_ptr;
    bool acyclic = false;              // box graph has no cycles: boxes may wait for whole predecessor boxes
    std::vector<int> pred_ptr, pred;   // per box (ticket order): tickets of the boxes it reads from
    DVec diag, val;
};
// returns 0 when a tile schedule was built, 2 when the factor is not a structured-grid factor
// (caller falls back to the slice schedule), 1 on error
int tri_tiled_build_host(int which, int n, const int *Tp, const int *Tj, const double *Tx, TiledHost &H);
// The device image of a tile schedule: one 16-byte aligned blob per box plus BoxDesc[num_tiles] (raw bytes here;
// the struct is private to tri_tiled.cu).  Packed on the host, no CUDA call involved.
struct PackedBoxes {
    std::vector<unsigned char, default_init_allocator<unsigned char>> blob;   // zeroed box by box by the packing threads
    std::vector<unsigned char> desc_bytes;
    size_t cap = 0;       // largest blob
    int max_ext = 0;      // most operands a box reads from other boxes
    bool flags = false;   // ELL blobs for the completion-flag kernel (acyclic box graphs)
};
int tri_tiled_pack_host(const TiledHost &H, PackedBoxes &P);
int tri_tiled_upload(lsspg_ctx *ctx, const TiledHost &H, lsspg_tri *T);
void tri_tiled_free(lsspg_tri *T);
int tri_tiled_solve(lsspg_ctx *ctx, const lsspg_tri *T, double *dx, const double *drhs, bool guarded);
// lattice test shared by the box and the pencil schedules: grid dimensions and the distinct |column - row| offsets
bool detect_lattice(int n, const int *Tp, const int *Tj, int dims[3], std::vector<long long> &offsets);

// ---- pencil schedule (tri_pencil.cu) ----
constexpr int kPenMaxW = 8;          // off-diagonal entries per row
constexpr int kPenMaxOut = 4;        // mailboxes one line feeds
constexpr int kPenRing = 16;         // hyperplanes kept in shared memory
constexpr int kPenMaxDk = 7;         // largest distance (in steps) between a row and an operand
constexpr int kPenMaxGhost = 96;     // ghost lines per pencil
constexpr int kPenMaxThreads = 256;  // lines per pencil
constexpr int kPenZeroLane = 0xffff; // operand descriptor of a neighbour that does not exist: the ring lane holding +0.0
// ring lanes per hyperplane: own lines, ghost lines, the +0.0 lane (last)
// (odd, so that a column of the ring -- one lane over consecutive steps -- spreads over the banks)
inline int pencil_ring_stride(int T, int max_ghost) { return T + ((max_ghost + 1 + 15) / 16) * 16 + 1; }
struct PencilHdr {      // one per pencil, ticket order (48 bytes)
    long long val_off;  // first value of the pencil in the value stream: vals[val_off + ((k * NV + w) * T + t)]
    int nsteps, nghost, thr_off, ghost_off;
    int kgend;          // last virtual step + 1 the ghost prefetcher handles
    int nout, out_off;  // mailboxes the pencil feeds: PencilOut[out_off, out_off + nout)
    int pad[3];
};
struct PencilOut {      // one per (pencil, fed mailbox): steps [kstart, kend) of ring lane `lane` go to mail[mail0 + k]
    int lane, mail0, kstart, kend;
};
struct PencilThread {   // one per (pencil, thread) = per line (64 bytes)
    int kstart, kend;   // the line is active in steps [kstart, kend)
    int row0;           // row of step k: row0 + dir * k
    int pad;
    int op[kPenMaxW];   // per slot: (distance in steps << 16) | ring lane  (lanes >= T: ghost lines)
    int out[kPenMaxOut];// mailbox position of step 0 (INT_MIN: none): the value of step k also goes to mail[out + k]
};
struct PencilGhost {    // one per (pencil, ghost line): virtual step kg in [kg0, kg1) is mail[mail0 + kg]
    int mail0, kg0, kg1, pad;
};
struct PencilHost {
    int n = 0, which = 0, W = 0, nv = 0, T = 0, pv = 0, pw = 0, dir = 1;
    bool holes = false;   // rows lack entries whose neighbour exists: the value stream marks missing entries
    int num_pencils = 0, max_ghost = 0, max_dk = 0, max_steps = 0, num_levels = 0;
    int grid_dims[3] = {0, 0, 0}, skew[3] = {0, 0, 0};
    long long mail_len = 0, offdiag_nnz = 0;
    std::vector<PencilHdr> hdr;
    std::vector<PencilThread> thr;
    std::vector<PencilGhost> ghost;
    std::vector<PencilOut> outs;
    bool own_last = false;   // the last slot of every row is the row's predecessor on its own line (kept in a register)
    DVec vals;
};
// 0: built; 2: not a lattice factor (caller falls back); 1: error
int tri_pencil_build_host(int which, int n, const int *Tp, const int *Tj, const double *Tx, int num_sms, PencilHost &H);
int tri_pencil_walk_host(const PencilHost &H, double *x, const double *rhs, int *info);
int tri_pencil_upload(lsspg_ctx *ctx, const PencilHost &H, lsspg_tri *T);
void tri_pencil_free(lsspg_tri *T);
int tri_pencil_solve(lsspg_ctx *ctx, const lsspg_tri *T, double *dx, const double *drhs, bool guarded);
// x = T^-1 rhs; `guarded`: skip when ctx->d_flags[FLAG_STOP] is set
int tri_solve(lsspg_ctx *ctx, const lsspg_tri *T, double *dx, const double *drhs, bool guarded);
}  // namespace lsspg

// amg_host.h -- host image of the SX-AMG-style hierarchy and of the Gauss-Seidel smoother layout.
#pragma once
#include <vector>
#include "host_par.h"
#include "../../include/lsspg.h"

namespace lsspg {

struct AmgLevelHost {
    int n = 0, nc = 0;
    std::vector<int> Ap, Aj;            // level operator, columns ascending
    std::vector<double> Ax;
    std::vector<int> Pp, Pj;            // prolongation, n x nc (empty on the last level)
    std::vector<double> Px;
    std::vector<int> Rp, Rj;            // restriction = P^T, nc x n
    std::vector<double> Rx;
    std::vector<int> cf;                // 1 = C point, 0 = F point (all 1 on the last level)
    std::vector<int> rank;              // visiting rank of a point inside its block (cf_order 0/1: its index;
                                        // cf_order 2: colour by colour, see amg_host.cpp)
};

// One Gauss-Seidel sweep as a dependency schedule.  Rows are split into the C block and the F
// block; inside a block they are grouped by dependency level (row i waits for the rows j < i of
// its own block that it references) and packed into 32-row slices, C slices first.  Entry k of
// the row in `lane` of slice s sits at (slice_ptr[s] + k) * 32 + lane, off-diagonals only, in
// ascending column order; col = (column << 2) | (visited earlier in the same block ? 2 : 0) | cf[column],
// -1 = padding.  `rank` (NULL: the row index) is the visiting order inside a block.
// mode 1 (deep schedules of wide rows): one ticket per ROW instead of per slice -- perm[p] / diag[p]
// per ordered row p, entries of row p contiguous at [slice_ptr[p], slice_ptr[p+1]), slices_c = #C rows.
constexpr int kGsShallowDepth = 64;   // up to this many dependency levels the SMs are filled (little polling)
constexpr int kGsStreamDepth = 16;    // ... and up to this many a slice sweep keeps only 8 entries per row in flight
struct GsHost {
    int n = 0, num_slices = 0, slices_c = 0, levels_c = 0, levels_f = 0, mode = 0;
    long long padded_nnz = 0, offdiag_nnz = 0;
    std::vector<int> perm;        // [num_slices*32] row of the slot, -1 = empty
    std::vector<double> diag;     // [num_slices*32]
    std::vector<int> slice_ptr;   // [num_slices+1]
    std::vector<int> level_ptr;   // mode 0: first slice of every dependency level, C levels then F levels (+ end)
    lsspg::IVec col;     // filled by the host threads (host_par.h)
    lsspg::DVec val;
};

// cf == NULL: every row in the C block (natural-order sweep)
// mode: 0 slices, 1 rows, -1 chosen from depth and row width
int gs_build_host(int n, const int *Ap, const int *Aj, const double *Ax, const int *cf, const int *rank, GsHost &G,
                  int mode = -1);

}  // namespace lsspg

struct lsspg_amg_host {
    lsspg_amg_pars pars;
    std::vector<lsspg::AmgLevelHost> levels;
    bool coarse_dense = false;
    std::vector<double> coarse_inv;   // row-major inverse of the last operator
};

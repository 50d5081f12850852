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

// strength graph (no values): row i lists the points that strongly influence i, in the row's column order
struct AmgGraph {
    std::vector<int> p, j;
};

// The phases of one coarsening step that are independent per row.  Three providers share the set-up loop
// (amg_setup_with, amg_host.cpp): the host code of amg_host.cpp (lsspg_amg_setup_host), the row functions of
// amg_rows.cuh replayed on the CPU (lsspg_debug_amg_setup_replay_host) and the same row functions one row per thread on
// the device (amg_gpu.cu, lsspg_amg_setup_device).  The serial phases -- transposed strength graph, Ruge-Stueben C/F
// splitting, visiting ranks, dense inverse of the last level -- are the loop's own.
struct AmgPhases {
    virtual ~AmgPhases() {}
    virtual int begin(const AmgLevelHost &L0) { (void)L0; return 0; }                       // level 0 is ready on the host
    virtual int strength(const AmgLevelHost &L, const lsspg_amg_pars &pr, AmgGraph &S) = 0;
    // L.cf is set; fills L.nc, L.Pp / Pj / Px
    virtual int interpolation(AmgLevelHost &L, const AmgGraph &S, const AmgGraph &T, const lsspg_amg_pars &pr) = 0;
    virtual int restriction(AmgLevelHost &L, const AmgGraph &T) = 0;                        // fills L.Rp / Rj / Rx
    virtual int galerkin(const AmgLevelHost &L, AmgLevelHost &C) = 0;                       // C.Ap / Aj / Ax = R A P
    virtual const char *name() const = 0;
};
int amg_setup_with(AmgPhases &ph, int n, const int *hAp, const int *hAj, const double *hAx, const lsspg_amg_pars *pars,
                   lsspg_amg_host **out);

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

// Prints size and member offsets of the structs that cross the LSSP C++ API by value or by reference.  The test-suite
// compiles this file twice -- against /root/reference/include (config.h generated with the same USE_* switches as
// include/lssp/config.h) and against include/lssp -- and compares the output: an object file built against the
// reference headers can then be linked against liblssp.so (reference include/type-defs.h:15-62, 107-151, 225-304).
#include <stddef.h>
#include <stdio.h>
#include "config.h"
#include "type-defs.h"

#define OFF(T, m) printf("  %-22s %zu\n", #T "." #m, offsetof(T, m))

int main()
{
    printf("sizeof lssp_mat_csr %zu lssp_mat_coo %zu lssp_mat_bcsr %zu lssp_vec %zu LSSP_PC %zu LSSP_SOLVER %zu\n", sizeof(lssp_mat_csr),
           sizeof(lssp_mat_coo), sizeof(lssp_mat_bcsr), sizeof(lssp_vec), sizeof(LSSP_PC), sizeof(LSSP_SOLVER));
    printf("enum LSSP_PC_NON %d ILUK %d ILUT %d BILUK %d SXAMG %d USER %d; LSSP_SOLVER_GMRES %d CG %d IDRS %d SXAMG %d\n", (int)LSSP_PC_NON,
           (int)LSSP_PC_ILUK, (int)LSSP_PC_ILUT, (int)LSSP_PC_BILUK, (int)LSSP_PC_SXAMG, (int)LSSP_PC_USER, (int)LSSP_SOLVER_GMRES,
           (int)LSSP_SOLVER_CG, (int)LSSP_SOLVER_IDRS, (int)LSSP_SOLVER_SXAMG);
    OFF(lssp_mat_csr, num_rows); OFF(lssp_mat_csr, num_cols); OFF(lssp_mat_csr, num_nnzs); OFF(lssp_mat_csr, Ap); OFF(lssp_mat_csr, Aj);
    OFF(lssp_mat_csr, Ax);
    OFF(lssp_vec, n); OFF(lssp_vec, d);
    OFF(LSSP_PC, iluk_level); OFF(LSSP_PC, ilut_p); OFF(LSSP_PC, ilut_tol); OFF(LSSP_PC, A); OFF(LSSP_PC, L); OFF(LSSP_PC, D); OFF(LSSP_PC, U);
    OFF(LSSP_PC, sxamg); OFF(LSSP_PC, data); OFF(LSSP_PC, cache); OFF(LSSP_PC, type); OFF(LSSP_PC, assemble); OFF(LSSP_PC, solve);
    OFF(LSSP_PC, destroy); OFF(LSSP_PC, log); OFF(LSSP_PC, verb); OFF(LSSP_PC, assembled);
    OFF(LSSP_SOLVER, tol_rel); OFF(LSSP_SOLVER, tol_abs); OFF(LSSP_SOLVER, tol_rb); OFF(LSSP_SOLVER, maxit); OFF(LSSP_SOLVER, restart);
    OFF(LSSP_SOLVER, aug_k); OFF(LSSP_SOLVER, bgsl); OFF(LSSP_SOLVER, idrs); OFF(LSSP_SOLVER, A); OFF(LSSP_SOLVER, Ab); OFF(LSSP_SOLVER, num_blks);
    OFF(LSSP_SOLVER, blk_size); OFF(LSSP_SOLVER, type); OFF(LSSP_SOLVER, rhs); OFF(LSSP_SOLVER, x); OFF(LSSP_SOLVER, residual);
    OFF(LSSP_SOLVER, nits); OFF(LSSP_SOLVER, sxamg); OFF(LSSP_SOLVER, verb); OFF(LSSP_SOLVER, log); OFF(LSSP_SOLVER, assembled);
    return 0;
}

/* solver-tri.h -- triangular sweeps of the ILU preconditioners (reference include/solver-tri.h:8-12). */
#ifndef LSSP_SOLVER_TRI_H
#define LSSP_SOLVER_TRI_H

#include "matrix-utils.h"
#include "type-defs.h"

void lssp_pc_ilu_solve_lower_matrix(lssp_mat_csr L, double *x, double *rhs);
void lssp_pc_ilu_solve_upper_matrix(lssp_mat_csr L, double *x, double *rhs);
void lssp_pc_ilu_solve_lu_matrix(lssp_mat_csr L, lssp_mat_csr U, double *x, double *rhs, double *cache);
void lssp_pc_ilu_solve(LSSP_PC *pc, lssp_vec x, lssp_vec rhs);

#endif

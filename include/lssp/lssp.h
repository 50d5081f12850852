/* lssp.h -- the LSSP solver facade (reference include/lssp.h:44-94), B200 build.
 * Source-compatible with the reference: a program written against huiscliu/lssp recompiles
 * against include/lssp/ and links liblssp.so + liblsspg.so unchanged. */
#ifndef LSSP_LSSP_H
#define LSSP_LSSP_H

#include "pc.h"
#include "solver-bicgsafe.h"
#include "solver-bicgstab.h"
#include "solver-bicgstabl.h"
#include "solver-bicrsafe.h"
#include "solver-bicrstab.h"
#include "solver-cg.h"
#include "solver-cgs.h"
#include "solver-cr.h"
#include "solver-crs.h"
#include "solver-gmres.h"
#include "solver-gpbicg.h"
#include "solver-gpbicr.h"
#include "solver-idrs.h"
#include "solver-lgmres.h"
#include "solver-orthomin.h"
#include "solver-qmrcgstab.h"
#include "solver-sxamg.h"
#include "solver-tfqmr.h"

void lssp_solver_create(LSSP_SOLVER &s, LSSP_SOLVER_TYPE s_type, LSSP_PC &pc, LSSP_PC_TYPE p_type);
void lssp_solver_assemble(LSSP_SOLVER &s, lssp_mat_csr &Ax, lssp_vec x, lssp_vec b, LSSP_PC &pc);
void lssp_solver_destroy(LSSP_SOLVER &s, LSSP_PC &pc);
int lssp_solver_solve(LSSP_SOLVER &solver, LSSP_PC &pc);
void lssp_solver_reset_rhs(LSSP_SOLVER &s, lssp_vec rhs);
void lssp_solver_reset_unknown(LSSP_SOLVER &s, lssp_vec x);
void lssp_solver_reset_type(LSSP_SOLVER &s, LSSP_SOLVER_TYPE type);
void lssp_solver_set_rtol(LSSP_SOLVER &s, double tol);
void lssp_solver_set_atol(LSSP_SOLVER &s, double tol);
void lssp_solver_set_rbtol(LSSP_SOLVER &s, double tol);
void lssp_solver_set_maxit(LSSP_SOLVER &s, int maxit);
void lssp_solver_set_restart(LSSP_SOLVER &s, int m);
void lssp_solver_set_augk(LSSP_SOLVER &s, int k);
void lssp_solver_set_bgsl(LSSP_SOLVER &s, int k);
void lssp_solver_set_idrs(LSSP_SOLVER &s, int k);
void lssp_solver_reset_verbosity(LSSP_SOLVER &s, int v);
double lssp_solver_get_residual(LSSP_SOLVER s);
int lssp_solver_get_nits(LSSP_SOLVER s);
void lssp_solver_set_log(LSSP_SOLVER &s, FILE *io);

#endif

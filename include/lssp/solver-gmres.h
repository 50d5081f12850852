/* solver-gmres.h -- reference include/solver-gmres.h */
#ifndef LSSP_SOLVER_GMRES_H
#define LSSP_SOLVER_GMRES_H

#include "mvops.h"

int lssp_solver_gmres(LSSP_SOLVER &solver, LSSP_PC &pc);
int lssp_solver_gmres_r(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

/* solver-crs.h -- reference include/solver-crs.h */
#ifndef LSSP_SOLVER_CRS_H
#define LSSP_SOLVER_CRS_H

#include "mvops.h"

int lssp_solver_crs(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

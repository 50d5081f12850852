/* solver-bicgstabl.h -- reference include/solver-bicgstabl.h */
#ifndef LSSP_SOLVER_BICGSTABL_H
#define LSSP_SOLVER_BICGSTABL_H

#include "mvops.h"

int lssp_solver_bicgstabl(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

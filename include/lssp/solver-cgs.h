/* solver-cgs.h -- reference include/solver-cgs.h */
#ifndef LSSP_SOLVER_CGS_H
#define LSSP_SOLVER_CGS_H

#include "mvops.h"

int lssp_solver_cgs(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

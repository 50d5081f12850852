/* solver-cg.h -- reference include/solver-cg.h */
#ifndef LSSP_SOLVER_CG_H
#define LSSP_SOLVER_CG_H

#include "mvops.h"

int lssp_solver_cg(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

/* solver-idrs.h -- reference include/solver-idrs.h */
#ifndef LSSP_SOLVER_IDRS_H
#define LSSP_SOLVER_IDRS_H

#include "mvops.h"

int lssp_solver_idrs(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

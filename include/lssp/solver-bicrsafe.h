/* solver-bicrsafe.h -- reference include/solver-bicrsafe.h */
#ifndef LSSP_SOLVER_BICRSAFE_H
#define LSSP_SOLVER_BICRSAFE_H

#include "mvops.h"

int lssp_solver_bicrsafe(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

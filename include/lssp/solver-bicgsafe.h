/* solver-bicgsafe.h -- reference include/solver-bicgsafe.h */
#ifndef LSSP_SOLVER_BICGSAFE_H
#define LSSP_SOLVER_BICGSAFE_H

#include "mvops.h"

int lssp_solver_bicgsafe(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

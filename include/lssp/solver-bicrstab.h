/* solver-bicrstab.h -- reference include/solver-bicrstab.h */
#ifndef LSSP_SOLVER_BICRSTAB_H
#define LSSP_SOLVER_BICRSTAB_H

#include "mvops.h"

int lssp_solver_bicrstab(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

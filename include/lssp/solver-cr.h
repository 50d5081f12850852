/* solver-cr.h -- reference include/solver-cr.h */
#ifndef LSSP_SOLVER_CR_H
#define LSSP_SOLVER_CR_H

#include "mvops.h"

int lssp_solver_cr(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

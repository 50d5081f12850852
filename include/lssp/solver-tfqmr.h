/* solver-tfqmr.h -- reference include/solver-tfqmr.h */
#ifndef LSSP_SOLVER_TFQMR_H
#define LSSP_SOLVER_TFQMR_H

#include "mvops.h"

int lssp_solver_tfqmr(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

/* solver-gpbicr.h -- reference include/solver-gpbicr.h */
#ifndef LSSP_SOLVER_GPBICR_H
#define LSSP_SOLVER_GPBICR_H

#include "mvops.h"

int lssp_solver_gpbicr(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

/* solver-gpbicg.h -- reference include/solver-gpbicg.h */
#ifndef LSSP_SOLVER_GPBICG_H
#define LSSP_SOLVER_GPBICG_H

#include "mvops.h"

int lssp_solver_gpbicg(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

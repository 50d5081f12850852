/* solver-orthomin.h -- reference include/solver-orthomin.h */
#ifndef LSSP_SOLVER_ORTHOMIN_H
#define LSSP_SOLVER_ORTHOMIN_H

#include "mvops.h"

int lssp_solver_orthomin(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

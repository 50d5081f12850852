/* solver-bicgstab.h -- reference include/solver-bicgstab.h */
#ifndef LSSP_SOLVER_BICGSTAB_H
#define LSSP_SOLVER_BICGSTAB_H

#include "mvops.h"

int lssp_solver_bicgstab(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

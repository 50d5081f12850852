/* solver-qmrcgstab.h -- reference include/solver-qmrcgstab.h */
#ifndef LSSP_SOLVER_QMRCGSTAB_H
#define LSSP_SOLVER_QMRCGSTAB_H

#include "mvops.h"

int lssp_solver_qmrcgstab(LSSP_SOLVER &solver, LSSP_PC &pc);

#endif

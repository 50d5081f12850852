/* pc-iluk.h -- ILU(k) preconditioner (reference include/pc-iluk.h). */
#ifndef LSSP_PC_ILUK_H
#define LSSP_PC_ILUK_H

#include "matrix-utils.h"
#include "solver-tri.h"
#include "type-defs.h"

void lssp_pc_iluk_assemble(LSSP_PC &pc, LSSP_SOLVER s);
void lssp_pc_iluk_destroy(LSSP_PC *pc);
void lssp_pc_iluk_set_level(LSSP_PC &pc, int level);

#endif

/* pc.h -- preconditioner life cycle (reference include/pc.h:15-17). */
#ifndef LSSP_PC_H
#define LSSP_PC_H

#include "pc-iluk.h"
#include "pc-ilut.h"
#include "pc-biluk.h"
#include "pc-sxamg.h"

void lssp_pc_create(LSSP_PC &pc, LSSP_PC_TYPE type);
void lssp_pc_destroy(LSSP_PC &pc);
void lssp_pc_assemble(LSSP_PC &pc, LSSP_SOLVER s);

#endif

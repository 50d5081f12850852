/* matrix-utils.h -- CSR helpers of the LSSP API that callers and the set-up path use
 * (reference include/matrix-utils.h).  Host-side; the BCSR/COO converters and the transpose
 * belong to the third-party adapters and are not part of this build. */
#ifndef LSSP_MATRIX_UTILS_H
#define LSSP_MATRIX_UTILS_H

#include "type-defs.h"
#include "utils.h"

void lssp_mat_init(lssp_mat_csr &A);
void lssp_mat_init(lssp_mat_coo &A);
void lssp_mat_init(lssp_mat_bcsr &A);
void lssp_mat_destroy(lssp_mat_csr &A);
void lssp_mat_destroy(lssp_mat_coo &A);
void lssp_mat_destroy(lssp_mat_bcsr &A);
lssp_mat_csr lssp_mat_create(int nrows, int ncols, int *Ap, int *Aj, double *Ax);
bool lssp_mat_csr_is_sorted(const lssp_mat_csr A);
void lssp_mat_sort_column(lssp_mat_csr &A);

#endif

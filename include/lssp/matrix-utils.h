/* matrix-utils.h -- matrix containers and utilities of the LSSP API (reference include/matrix-utils.h): the
 * data formats either side of the solve loop.  Host-side, written fresh; results equal the reference's bit for bit
 * (tests/cxx/mat_utils_abi_check.cpp calls both libraries through the same binary interface). */
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
/* format converters (reference include/matrix-utils.h:22-31) */
lssp_mat_bcsr lssp_mat_csr_to_bcsr(const lssp_mat_csr A, int bs);
lssp_mat_csr lssp_mat_bcsr_to_csr(const lssp_mat_bcsr A);
lssp_mat_coo lssp_mat_csr_to_coo(const lssp_mat_csr csr);
lssp_mat_csr lssp_mat_coo_to_csr(const lssp_mat_coo A);

bool lssp_mat_csr_is_sorted(const lssp_mat_csr A);
bool lssp_mat_bcsr_is_sorted(const lssp_mat_bcsr A);
void lssp_mat_sort_column(lssp_mat_csr &A);

/* insert (i, tol) where a row stores no diagonal; block-Jacobi restriction; transpose (:43-49) */
lssp_mat_csr lssp_mat_adjust_zero_diag(const lssp_mat_csr A, double tol);
lssp_mat_csr lssp_mat_get_block_diag(const lssp_mat_csr A, int blk_size);
lssp_mat_csr lssp_mat_transpose(const lssp_mat_csr A);

#endif

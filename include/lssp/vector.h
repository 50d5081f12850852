/* vector.h -- BLAS-1 interface of the LSSP API (reference include/vector.h:7-40).  Vectors are
 * host objects; every arithmetic entry point runs the sm_100a kernel behind the C ABI. */
#ifndef LSSP_VECTOR_H
#define LSSP_VECTOR_H

#include "matrix-utils.h"

lssp_vec lssp_vec_create(int n);
void lssp_vec_destroy(lssp_vec &v);
void lssp_vec_set_value(lssp_vec x, double val);
void lssp_vec_set_value_by_array(lssp_vec x, double *val);
void lssp_vec_set_value_by_index(lssp_vec x, int i, double val);
void lssp_vec_get_value(double *val, lssp_vec x);
double lssp_vec_get_value_by_index(lssp_vec x, int i);
void lssp_vec_copy(lssp_vec des, const lssp_vec src);
void lssp_vec_axy(double alpha, const lssp_vec x, lssp_vec y);                         /* y = alpha x */
void lssp_vec_axpby(double alpha, const lssp_vec x, double beta, lssp_vec y);          /* y = beta y + alpha x */
void lssp_vec_axpbyz(double alpha, const lssp_vec x, double beta, lssp_vec y, lssp_vec z);
double lssp_vec_dot(const lssp_vec x, const lssp_vec y);
double lssp_vec_norm(const lssp_vec x);
void lssp_vec_scale(lssp_vec x, double a);

#endif

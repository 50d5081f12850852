/* type-defs.h -- containers and handles of the LSSP API (same names, fields and enumerator order
 * as the reference's include/type-defs.h with every USE_* = 0), written fresh for the B200 build.
 * Vectors and matrices stay HOST objects, exactly as callers of the reference see them.  The structs are
 * binary-identical to the reference's for the same USE_* switches (tests/cxx/struct_layout_check.cpp): device
 * images hang off pc.data (built-in preconditioners) or a side table inside liblssp.so (lssp_facade.cpp). */
#ifndef LSSP_TYPES_H
#define LSSP_TYPES_H

#include <assert.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "config.h"

typedef struct lssp_mat_csr_ {
    int num_rows, num_cols, num_nnzs;
    int *Ap, *Aj;
    double *Ax;
} lssp_mat_csr;

typedef struct lssp_mat_coo_ {
    int num_rows, num_cols, num_nnzs;
    int *Ai, *Aj;
    double *Ax;
} lssp_mat_coo;

typedef struct lssp_mat_entry_ {
    int i, j;
    double x;
} lssp_mat_entry;

typedef struct lssp_mat_bcsr_ {
    int num_rows, num_cols, num_nnzs, blk_size;
    int *Ap, *Aj;
    double *Ax;
} lssp_mat_bcsr;

typedef struct lssp_vec_ {
    int n;
    double *d;
} lssp_vec;

typedef enum LSSP_PC_TYPE_ {
    LSSP_PC_NON,
    LSSP_PC_ILUK,
    LSSP_PC_ILUT,
#if USE_BLAS
#if USE_LAPACK
    LSSP_PC_BILUK,              /* block-wise ILUK (reference include/type-defs.h:70-74) */
#endif
#endif
#if USE_SXAMG
    LSSP_PC_SXAMG,              /* AMG, SX-AMG style (reference include/type-defs.h:92-94) */
#endif
    LSSP_PC_USER,
} LSSP_PC_TYPE;

struct LSSP_PC_;
struct LSSP_SOLVER_;

typedef void (*LSSP_PC_ASSEMBLE)(struct LSSP_PC_ &pc, struct LSSP_SOLVER_ s);
typedef void (*LSSP_PC_SOLVE)(struct LSSP_PC_ *s, lssp_vec x, lssp_vec rhs);
typedef void (*LSSP_PC_DESTROY)(struct LSSP_PC_ *s);

typedef struct LSSP_PC_ {
    int iluk_level;
    int ilut_p;
    double ilut_tol;

    lssp_mat_csr A, L, D, U;     /* host copies of the factors (L: diagonal last, U: diagonal first) */

#if USE_SXAMG
    struct SX_DATA_ *sxamg;      /* reference include/type-defs.h:135-137 */
#endif

    void *data;
    double *cache;

    LSSP_PC_TYPE type;
    LSSP_PC_ASSEMBLE assemble;
    LSSP_PC_SOLVE solve;
    LSSP_PC_DESTROY destroy;

    FILE *log;
    int verb;
    bool assembled;
} LSSP_PC;

typedef enum LSSP_SOLVER_TYPE_ {
    LSSP_SOLVER_GMRES,
    LSSP_SOLVER_LGMRES,
    LSSP_SOLVER_RGMRES,
    LSSP_SOLVER_RLGMRES,
    LSSP_SOLVER_BICGSTAB,
    LSSP_SOLVER_BICGSTABL,
    LSSP_SOLVER_BICGSAFE,
    LSSP_SOLVER_CG,
    LSSP_SOLVER_CGS,
    LSSP_SOLVER_GPBICG,
    LSSP_SOLVER_CR,
    LSSP_SOLVER_CRS,
    LSSP_SOLVER_BICRSTAB,
    LSSP_SOLVER_BICRSAFE,
    LSSP_SOLVER_GPBICR,
    LSSP_SOLVER_QMRCGSTAB,
    LSSP_SOLVER_TFQMR,
    LSSP_SOLVER_ORTHOMIN,
    LSSP_SOLVER_IDRS,
#if USE_SXAMG
    LSSP_SOLVER_SXAMG,           /* stand-alone AMG (reference include/type-defs.h:219-221) */
#endif
} LSSP_SOLVER_TYPE;

typedef struct LSSP_SOLVER_ {
    double tol_rel, tol_abs, tol_rb;
    int maxit, restart, aug_k, bgsl, idrs;

#if USE_SXAMG
    struct SXAMG_DATA_ *sxamg;   /* reference include/type-defs.h:280-282 (before the matrix, as there) */
#endif

    lssp_mat_csr A;              /* host deep copy, columns sorted (as the reference keeps it) */
    lssp_mat_bcsr Ab;
    int num_blks;
    int *blk_size;

    LSSP_SOLVER_TYPE type;
    lssp_vec rhs, x;             /* ALIASES of the caller's vectors */

    double residual;
    int nits;

    int verb;
    FILE *log;
    bool assembled;
} LSSP_SOLVER;

#endif

/* sxamg.h -- stand-in for the header of libsxamg (https://github.com/huiscliu/sxamg), which the
 * reference includes when configured with SX-AMG (include/pc-sxamg.h:10, include/solver-sxamg.h:10)
 * but which is not part of the reference tree.  Only what the LSSP adapters touch is declared:
 * the parameter block and its initialiser.  Field names follow the library as recollected
 * (SURVEY.md App. C); the AMG itself is the B200 build's own (DESIGN.md "AMG"; parity with
 * libsxamg unpinned).  Fields this build honours are marked (*). */
#ifndef LSSP_SXAMG_STANDIN_H
#define LSSP_SXAMG_STANDIN_H

typedef int    SX_INT;
typedef double SX_FLT;

typedef struct SX_AMG_PARS_ {
    SX_INT verb;               /* (*) */
    SX_INT cycle_itr;          /* 1 = V-cycle (the only one implemented) */
    SX_FLT tol;                /* (*) stand-alone solver: relative residual */
    SX_FLT ctol;               /* coarsest-level tolerance: the last level is solved directly here */
    SX_INT maxit;              /* (*) stand-alone solver: cycles; the preconditioner uses 1 */
    SX_INT cs_type;            /* coarsening: classical Ruge-Stueben */
    SX_INT interp_type;        /* interpolation: direct, truncated */
    SX_INT max_levels;         /* (*) 30  */
    SX_INT max_coarsest_dof;   /* (*) 100 */
    SX_FLT strong_threshold;   /* (*) 0.3 */
    SX_FLT max_row_sum;        /* (*) 0.9 */
    SX_FLT trunc_threshold;    /* (*) 0.2 */
    SX_INT smoother;           /* Gauss-Seidel */
    SX_FLT relaxation;         /* 1.0 */
    SX_INT cf_order;           /* (*) 1: C/F-ordered sweeps */
    SX_INT pre_iter;           /* (*) 2 */
    SX_INT post_iter;          /* (*) 2 */
    SX_INT zero_guess;         /* (*) extension: 1 = every preconditioner application starts from x = 0;
                                  0 = from the vector the driver hands in, as src/pc-sxamg.cxx:58-64 */
} SX_AMG_PARS;

void sx_amg_pars_init(SX_AMG_PARS *pars);

#endif

// e2e_shim.cpp -- extern "C" handle around the LSSP C++ API of liblssp.so, for harnesses that cannot call C++ directly
// (bench.py's end-to-end leg, through ctypes).  It does exactly what a C++ caller of the reference does
// (example/exam.cxx:61-127): lssp_solver_create / setters / lssp_solver_assemble once, then lssp_solver_solve per
// right-hand side with the caller's HOST vectors -- the library copies b and x0 to the device and x back inside
// lssp_solver_solve.  Built into lssp_b200/liblssp_e2e.so; not part of the drop-in library itself.
#include <stdlib.h>
#include <string.h>

#include "lssp.h"

extern "C" {

typedef struct lssp_e2e_ {
    LSSP_SOLVER s;
    LSSP_PC pc;
    lssp_vec x, b;     // the caller's vectors (aliased by the solver, src/lssp.cxx:175-176)
    int n;
} lssp_e2e;

// solver_type / pc_type: LSSP_SOLVER_TYPE / LSSP_PC_TYPE enumerators; x and b: caller-owned host arrays of n doubles that
// stay valid for the lifetime of the handle
void *lssp_e2e_create(int solver_type, int pc_type, int n, int *Ap, int *Aj, double *Ax, double *x, double *b, int iluk_level,
                      int maxit, int restart, double rtol)
{
    lssp_e2e *h = (lssp_e2e *)calloc(1, sizeof(lssp_e2e));
    lssp_mat_csr A;
    A.num_rows = A.num_cols = n;
    A.num_nnzs = Ap[n];
    A.Ap = Ap; A.Aj = Aj; A.Ax = Ax;          // deep-copied by lssp_solver_assemble (src/lssp.cxx:169-171)
    h->n = n;
    h->x.n = n; h->x.d = x;
    h->b.n = n; h->b.d = b;
    lssp_solver_create(h->s, (LSSP_SOLVER_TYPE)solver_type, h->pc, (LSSP_PC_TYPE)pc_type);
    lssp_solver_set_maxit(h->s, maxit);
    lssp_solver_set_restart(h->s, restart);
    if (rtol > 0) lssp_solver_set_rtol(h->s, rtol);
    lssp_pc_iluk_set_level(h->pc, iluk_level);
    lssp_solver_assemble(h->s, A, h->x, h->b, h->pc);
    return h;
}

// one lssp_solver_solve(LSSP_SOLVER &, LSSP_PC &) from whatever the caller put into x; returns the iteration count
int lssp_e2e_solve(void *hh, double *residual)
{
    lssp_e2e *h = (lssp_e2e *)hh;
    const int nits = lssp_solver_solve(h->s, h->pc);
    if (residual) *residual = h->s.residual;
    return nits;
}

void lssp_e2e_destroy(void *hh)
{
    lssp_e2e *h = (lssp_e2e *)hh;
    lssp_solver_destroy(h->s, h->pc);
    free(h);
}

}  // extern "C"

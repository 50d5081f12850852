// Two behaviours of the reference's C++ API that the facade must keep (run on the GPU box by tests/test_facade.py):
//  (1) a LSSP_PC_USER preconditioner may call library functions from inside pc.solve while a solve is running
//      (include/type-defs.h:104, src/pc.cxx:219-227): here it calls lssp_mv_mxy and lssp_pc_ilu_solve on a second,
//      independently assembled ILU preconditioner -- both stage through the context and must not disturb the
//      x and b of the running Krylov solve.  The solve must behave exactly like the one with the built-in ILUK.
//  (2) per-iteration lines go through lssp_printf, so a file given to lssp_solver_set_log receives them
//      (src/utils.cxx:93-112, src/solver-cg.cxx:108-112).
#include <stdio.h>
#include <string.h>

#include "lssp.h"

static lssp_mat_csr poisson2d(int N)
{
    lssp_mat_csr A;
    A.num_rows = A.num_cols = N * N;
    A.num_nnzs = 5 * N * N - 4 * N;
    A.Ap = lssp_malloc<int>(A.num_rows + 1);
    A.Aj = lssp_malloc<int>(A.num_nnzs);
    A.Ax = lssp_malloc<double>(A.num_nnzs);
    int k = 0;
    A.Ap[0] = 0;
    for (int row = 0; row < N * N; row++) {
        const int gy = row / N, gx = row % N;
        const int cand[5] = {row - N, row - 1, row, row + 1, row + N};
        const bool ok[5] = {gy > 0, gx > 0, true, gx < N - 1, gy < N - 1};
        for (int q = 0; q < 5; q++) {
            if (!ok[q]) continue;
            A.Aj[k] = cand[q];
            A.Ax[k] = (q == 2) ? 4. : -1.;
            k++;
        }
        A.Ap[row + 1] = k;
    }
    return A;
}

static LSSP_PC g_inner;            // a real ILUK(1), assembled by the user preconditioner
static lssp_mat_csr g_A;
static lssp_vec g_tmp;
static int g_calls = 0;

static void user_assemble(LSSP_PC &pc, LSSP_SOLVER s)
{
    lssp_pc_create(g_inner, LSSP_PC_ILUK);
    lssp_pc_assemble(g_inner, s);
    g_tmp = lssp_vec_create(s.A.num_rows);
    pc.cache = NULL;
}
static void user_solve(LSSP_PC *pc, lssp_vec x, lssp_vec rhs)
{
    (void)pc;
    lssp_mv_mxy(g_A, rhs, g_tmp);              // a library call that stages through the context
    g_inner.solve(&g_inner, x, rhs);           // and the inner preconditioner's own application
    g_calls++;
}
static void user_destroy(LSSP_PC *pc)
{
    (void)pc;
    lssp_pc_destroy(g_inner);
    lssp_vec_destroy(g_tmp);
}

static int run(lssp_mat_csr A, LSSP_PC_TYPE pt, double *residual, double *xnorm, FILE *log)
{
    const int n = A.num_rows;
    lssp_vec x = lssp_vec_create(n), b = lssp_vec_create(n);
    lssp_vec_set_value(x, 0.);
    lssp_vec_set_value(b, 1.);
    LSSP_SOLVER solver;
    LSSP_PC pc;
    lssp_solver_create(solver, LSSP_SOLVER_CG, pc, pt);
    lssp_solver_set_maxit(solver, 3000);
    if (log) lssp_solver_set_log(solver, log);
    if (pt == LSSP_PC_USER) {
        pc.assemble = user_assemble;
        pc.solve = user_solve;
        pc.destroy = user_destroy;
    }
    lssp_solver_assemble(solver, A, x, b, pc);
    const int nits = lssp_solver_solve(solver, pc);
    *residual = solver.residual;
    *xnorm = lssp_vec_norm(x);
    lssp_solver_destroy(solver, pc);
    lssp_vec_destroy(x);
    lssp_vec_destroy(b);
    return nits;
}

int main(int argc, char **argv)
{
    const char *logname = argc > 1 ? argv[1] : "user_pc_and_log.log";
    lssp_mat_csr A = poisson2d(100);
    g_A = A;
    double r0, r1, n0, n1;
    lssp_verbosity = 0;
    const int k0 = run(A, LSSP_PC_ILUK, &r0, &n0, NULL);
    const int k1 = run(A, LSSP_PC_USER, &r1, &n1, NULL);
    printf("iluk: %d %.17g %.17g\nuser: %d %.17g %.17g (pc.solve called %d times)\n", k0, r0, n0, k1, r1, n1, g_calls);
    const bool same = (k0 == k1 && r0 == r1 && n0 == n1 && g_calls >= k1);
    printf("reentrant user preconditioner: %s\n", same ? "OK" : "MISMATCH");
    lssp_verbosity = 2;
    FILE *log = fopen(logname, "w");
    double r2, n2;
    const int k2 = run(A, LSSP_PC_ILUK, &r2, &n2, log);
    fclose(log);
    int lines = 0;
    char buf[512];
    log = fopen(logname, "r");
    while (fgets(buf, sizeof(buf), log))
        if (strstr(buf, "cg: itr:")) lines++;
    fclose(log);
    printf("log file: %d iteration lines for %d iterations: %s\n", lines, k2, (lines >= k2 && k2 == k0) ? "OK" : "MISMATCH");
    lssp_mat_destroy(A);
    return 0;
}

// exam.cxx -- smoke program for the LSSP C++ API on the B200 build, in the spirit of the
// reference's example/exam.cxx: 2-D 5-point Poisson problem on an N x N grid, b = 1, x0 = 0,
// GMRES(60) + ILUK(level 1), then an independent verification residual ||b - A x||.
// Usage: exam [N] [solver: gmres|cg|bicgstab|idrs] [pc: non|iluk|ilut]
// Expected for the defaults (N = 100), from the unmodified reference:
//   49 iterations, residual 8.18058783e-06, ||x|| = 4.25082937e+04.
#include <string>

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

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 100;
    const std::string sname = argc > 2 ? argv[2] : "gmres", pname = argc > 3 ? argv[3] : "iluk";
    LSSP_SOLVER_TYPE st = LSSP_SOLVER_GMRES;
    if (sname == "cg") st = LSSP_SOLVER_CG;
    if (sname == "bicgstab") st = LSSP_SOLVER_BICGSTAB;
    if (sname == "idrs") st = LSSP_SOLVER_IDRS;
    if (sname == "sxamg") st = LSSP_SOLVER_SXAMG;      // stand-alone AMG iteration
    LSSP_PC_TYPE pt = LSSP_PC_ILUK;
    if (pname == "non") pt = LSSP_PC_NON;
    if (pname == "ilut") pt = LSSP_PC_ILUT;
    if (pname == "sxamg") pt = LSSP_PC_SXAMG;          // one V-cycle per application
    if (pname == "biluk") pt = LSSP_PC_BILUK;          // block ILU(k); 5th argument: block size (default 2)

    lssp_mat_csr A = poisson2d(N);
    const int n = A.num_rows;
    lssp_vec x = lssp_vec_create(n), b = lssp_vec_create(n), r = lssp_vec_create(n);
    lssp_vec_set_value(x, 0.);
    lssp_vec_set_value(b, 1.);

    LSSP_SOLVER solver;
    LSSP_PC pc;
    lssp_solver_create(solver, st, pc, pt);
    lssp_solver_set_restart(solver, 60);
    lssp_solver_set_maxit(solver, 3000);
    lssp_solver_reset_verbosity(solver, 0);
    if (pt == LSSP_PC_SXAMG && argc > 4) {             // 5th argument: zero_guess (see include/lssp/sxamg.h)
        SX_AMG_PARS pars;
        sx_amg_pars_init(&pars);
        pars.maxit = 1;
        pars.zero_guess = atoi(argv[4]);
        lssp_pc_sxamg_set_pars(pc, &pars);
    }
    if (pt == LSSP_PC_BILUK) solver.num_blks = n / (argc > 4 ? atoi(argv[4]) : 2);   // block size = n / num_blks
    lssp_solver_assemble(solver, A, x, b, pc);
    const int nits = lssp_solver_solve(solver, pc);

    lssp_mv_amxpbyz(-1, A, x, 1, b, r);      // verification residual with the public SpMV
    lssp_printf("exam: n: %d, solver: %s, pc: %s\n", n, sname.c_str(), pname.c_str());
    lssp_printf("exam: iterations: %d, solver residual: %.8e\n", nits, lssp_solver_get_residual(solver));
    lssp_printf("exam: solution L2 norm: %.8e residual: %.8e\n", lssp_vec_norm(x), lssp_vec_norm(r));

    lssp_solver_destroy(solver, pc);
    lssp_mat_destroy(A);
    lssp_vec_destroy(x);
    lssp_vec_destroy(b);
    lssp_vec_destroy(r);
    return 0;
}

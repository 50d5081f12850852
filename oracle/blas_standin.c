/* TEST INFRASTRUCTURE (oracle) -- never linked into the product.
 *
 * The reference's block ILU(k) (src/pc-biluk.cxx) is only compiled `#if USE_BLAS && USE_LAPACK` and calls the
 * Fortran symbols dgemm_, dgetrf_, dgetri_ (src/pc-biluk.cxx:10-16).  No BLAS/LAPACK is installed here, so the
 * oracle build `_ref/liblssp_refb.so` links the UNMODIFIED pc-biluk.cxx against the three routines below: plain-C
 * restatements of the PUBLISHED netlib reference algorithms (reference BLAS DGEMM 'N','N'; LAPACK 3.x DGETF2
 * = unblocked DGETRF with partial pivoting; DGETRI's unblocked path = DTRTI2 + the column sweep with DGEMV).
 * An optimised BLAS orders the dense sums differently: against such a build the block factors agree to rounding,
 * not bit for bit -- the pin this file gives is "the reference's block algorithm over reference-BLAS arithmetic".
 * Column-major, Fortran calling convention (every argument by pointer).
 */
#include <float.h>
#include <math.h>

/* C := alpha A B + beta C, all n-by-n style general sizes, no transposes (the only use, pc-biluk.cxx:100-101) */
void dgemm_(char *transa, char *transb, int *m_, int *n_, int *k_, double *alpha_, double *a, int *lda_, double *b,
            int *ldb_, double *beta_, double *c, int *ldc_)
{
    const int m = *m_, n = *n_, k = *k_, lda = *lda_, ldb = *ldb_, ldc = *ldc_;
    const double alpha = *alpha_, beta = *beta_;
    (void)transa; (void)transb;
    for (int j = 0; j < n; j++) {
        if (beta == 0.0) for (int i = 0; i < m; i++) c[i + j * ldc] = 0.0;
        else if (beta != 1.0) for (int i = 0; i < m; i++) c[i + j * ldc] = beta * c[i + j * ldc];
        for (int l = 0; l < k; l++) {
            const double temp = alpha * b[l + j * ldb];
            for (int i = 0; i < m; i++) c[i + j * ldc] = c[i + j * ldc] + temp * a[i + l * lda];
        }
    }
}

/* DGETF2: A = P L U, unit lower L, ipiv 1-based */
void dgetrf_(int *m_, int *n_, double *a, int *lda_, int *ipiv, int *info)
{
    const int m = *m_, n = *n_, lda = *lda_;
    const int mn = m < n ? m : n;
    const double sfmin = DBL_MIN;
    *info = 0;
    for (int j = 0; j < mn; j++) {
        int jp = j;                                   /* IDAMAX: first entry of largest magnitude */
        double big = fabs(a[j + j * lda]);
        for (int i = j + 1; i < m; i++)
            if (fabs(a[i + j * lda]) > big) { big = fabs(a[i + j * lda]); jp = i; }
        ipiv[j] = jp + 1;
        if (a[jp + j * lda] != 0.0) {
            if (jp != j)
                for (int c = 0; c < n; c++) { const double t = a[j + c * lda]; a[j + c * lda] = a[jp + c * lda]; a[jp + c * lda] = t; }
            if (j < m - 1) {
                if (fabs(a[j + j * lda]) >= sfmin) {
                    const double r = 1.0 / a[j + j * lda];
                    for (int i = j + 1; i < m; i++) a[i + j * lda] = r * a[i + j * lda];
                }
                else for (int i = j + 1; i < m; i++) a[i + j * lda] = a[i + j * lda] / a[j + j * lda];
            }
        }
        else if (*info == 0) *info = j + 1;
        if (j < mn - 1)                               /* DGER: trailing update */
            for (int c = j + 1; c < n; c++) {
                const double temp = -a[j + c * lda];
                if (a[j + c * lda] != 0.0)
                    for (int i = j + 1; i < m; i++) a[i + c * lda] = a[i + c * lda] + a[i + j * lda] * temp;
            }
    }
}

/* DGETRI, unblocked: inv(U) by DTRTI2, then inv(A) L = inv(U) column by column, then the column interchanges */
void dgetri_(int *n_, double *a, int *lda_, int *ipiv, double *work, int *lwork, int *info)
{
    const int n = *n_, lda = *lda_;
    (void)lwork;
    *info = 0;
    for (int j = 0; j < n; j++)
        if (a[j + j * lda] == 0.0) { *info = j + 1; return; }   /* DTRTRI's singularity check */
    for (int j = 0; j < n; j++) {                     /* DTRTI2, upper, non-unit */
        a[j + j * lda] = 1.0 / a[j + j * lda];
        const double ajj = -a[j + j * lda];
        /* DTRMV upper, no transpose, non-unit: x := U(0:j,0:j) x with x = column j above the diagonal */
        for (int c = 0; c < j; c++) {
            if (a[c + j * lda] != 0.0) {
                const double temp = a[c + j * lda];
                for (int i = 0; i < c; i++) a[i + j * lda] = a[i + j * lda] + temp * a[i + c * lda];
                a[c + j * lda] = a[c + j * lda] * a[c + c * lda];
            }
        }
        for (int i = 0; i < j; i++) a[i + j * lda] = ajj * a[i + j * lda];   /* DSCAL */
    }
    for (int j = n - 1; j >= 0; j--) {
        for (int i = j + 1; i < n; i++) { work[i] = a[i + j * lda]; a[i + j * lda] = 0.0; }
        if (j < n - 1)                                /* DGEMV: A(:,j) -= A(:,j+1:n) work(j+1:n) */
            for (int c = j + 1; c < n; c++) {
                const double temp = -work[c];
                for (int i = 0; i < n; i++) a[i + j * lda] = a[i + j * lda] + temp * a[i + c * lda];
            }
    }
    for (int j = n - 2; j >= 0; j--) {
        const int jp = ipiv[j] - 1;
        if (jp != j)
            for (int i = 0; i < n; i++) { const double t = a[i + j * lda]; a[i + j * lda] = a[i + jp * lda]; a[i + jp * lda] = t; }
    }
}

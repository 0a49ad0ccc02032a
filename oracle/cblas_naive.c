/* oracle/cblas_naive.c -- TEST INFRASTRUCTURE ONLY.
 * Hermetic cblas_dgemm for building the reference sources in oracle/_ref/ (see oracle/cblas.h).
 * Row-major, NoTrans x NoTrans only -- the one call pattern the reference uses
 * (main_MIDASPOM.c:363).  Output may alias neither input (the reference never does that for C).
 * i-k-j loop order so the inner loop streams rows; blocked over k for the 256^3 products of the
 * dieoff/loss modules (main_MIDASPOM_dieoff.c:313).
 */
#include <stdlib.h>
#include "cblas.h"

void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc)
{
    if (order != CblasRowMajor || ta != CblasNoTrans || tb != CblasNoTrans) abort();
    for (int i = 0; i < m; i++) {
        double *ci = c + (size_t)i * ldc;
        if (beta == 0.0) for (int j = 0; j < n; j++) ci[j] = 0.0;
        else             for (int j = 0; j < n; j++) ci[j] *= beta;
        for (int p = 0; p < k; p++) {
            const double aip = alpha * a[(size_t)i * lda + p];
            if (aip == 0.0) continue;
            const double *bp = b + (size_t)p * ldb;
            for (int j = 0; j < n; j++) ci[j] += aip * bp[j];
        }
    }
}

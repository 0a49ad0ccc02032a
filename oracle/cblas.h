/* oracle/cblas.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Minimal stand-in for the CBLAS header the reference includes
 * (/root/reference/sources/main_MIDASPOM.c:6, `#include <cblas.h>`).  The image ships no cblas.h
 * and no libcblas; the only BLAS symbol the reference uses is cblas_dgemm (RowMajor, NoTrans x
 * NoTrans, alpha=1, beta=0 -- e.g. main_MIDASPOM.c:363,379; main_MIDASPOM_dieoff.c:35,43,313).
 * The implementation is oracle/cblas_naive.c.  Third-party dependency it replaces: ATLAS/CBLAS,
 * un-vendored and un-pinned by the reference (makefile:3, `-latlas -lcblas`).
 */
#ifndef ORACLE_CBLAS_SHIM_H
#define ORACLE_CBLAS_SHIM_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc);
#endif

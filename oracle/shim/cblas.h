/* Minimal stand-in for <cblas.h> mapping onto scipy's bundled OpenBLAS (see shim/lapacke.h). */
#pragma once
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
#ifdef __cplusplus
extern "C" {
#endif
void scipy_cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int M, int N, int K,
                       double alpha, const double *A, int lda, const double *B, int ldb, double beta, double *C,
                       int ldc);
#ifdef __cplusplus
}
#endif
#define cblas_dgemm scipy_cblas_dgemm

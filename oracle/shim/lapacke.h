/* Minimal stand-in for <lapacke.h> so the UNMODIFIED reference sources compile in this image (which ships LAPACK
 * only inside scipy's OpenBLAS, with scipy_-prefixed symbols).  Test infrastructure, not product code. */
#pragma once
#include <math.h> /* the reference's benchmark.cpp uses fabs() without including <cmath> */
#include <stdlib.h>
#define LAPACK_COL_MAJOR 102
#define LAPACK_ROW_MAJOR 101
typedef int lapack_int;
#ifdef __cplusplus
extern "C" {
#endif
int scipy_LAPACKE_dgetrf(int layout, int m, int n, double *a, int lda, int *ipiv);
int scipy_LAPACKE_dgetrs(int layout, char trans, int n, int nrhs, const double *a, int lda, const int *ipiv, double *b,
                         int ldb);
#ifdef __cplusplus
}
#endif
#define LAPACKE_dgetrf scipy_LAPACKE_dgetrf
#define LAPACKE_dgetrs scipy_LAPACKE_dgetrs

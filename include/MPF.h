// Drop-in declaration of the reference's only host entry point (reference: MPF.h:3, implemented in MPF.cu:66-256).
// Same name, same C++ linkage (link symbol _Z3MPFPdiiPi), same argument meaning:
//   h_A   [in/out] N x N matrix in HOST memory, column-major, leading dimension N; overwritten by the LAPACK-style
//                  L\U factors (unit-lower L below the diagonal, U on and above it)
//   N     [in]     matrix order
//   r     [in]     panel width of the mixed-precision pre-pivoting factorization (benchmark.cpp:220 passes 32)
//   IPIV  [out]    1-based global pivot rows, sequential-swap (dgetrf) convention.  As in the reference, the entry of
//                  a trailing 1 x 1 panel is left untouched, so callers pre-fill IPIV[i] = i + 1 (benchmark.cpp:215).
// Errors are not reported through this symbol (the reference's is void too): with no CUDA device a message goes to
// stderr and h_A is left unchanged.  The C-ABI twin  int mplu_MPF(double*, int, int, int*)  returns a status.
#pragma once

void MPF(double *h_A, int N, int r, int *IPIV);

#ifdef __cplusplus
extern "C" {
#endif
int mplu_MPF(double *h_A, int N, int r, int *IPIV);
#ifdef __cplusplus
}
#endif

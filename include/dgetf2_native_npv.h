// Drop-in device entry point (reference: dgetf2_native_npv.h:8, dgetf2_native_npv.cu:11-36): fp64 LU WITHOUT
// pivoting of a pre-pivoted m x n column-major panel, in place.  C++ linkage, symbol _Z17dgetf2_native_npviiPdi.
// Cooperative launch required; any geometry works (the reference needs gridDim*blockDim >= m).
#pragma once

__global__ void dgetf2_native_npv(int m, int n, double *panel, int ld);

// Drop-in device entry point (reference: hgetf2_kernel.h:10, hgetf2_kernel.cu:15-120): fp16 LU with partial
// pivoting of a column-major rows x cols panel, used by MPF to DISCOVER pivot rows (the fp16 factors themselves are
// discarded).  C++ linkage, link symbol _Z13HGETF2_kernelP6__halfiiiPi.
//   panel       [in/out] rows x cols halfs, column-major, leading dimension ld
//   ipiv_panel  [out]    cols entries, 1-based, panel-local: row swapped with row j at step j (first maximum of |a|)
// Must be launched with cudaLaunchCooperativeKernel (it synchronises the grid); unlike the reference any grid/block
// geometry works (grid-stride loops), including the reference caller's ceil(rows/256) x 256 (MPF.cu:126-133).
#pragma once
#include "fp16_utils.h"

__global__ void HGETF2_kernel(fp16 *panel, int ld, int rows, int cols, int *ipiv_panel);

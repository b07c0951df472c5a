// fp16 helper contract of the reference (reference: fp16_utils.h:5-27), kept so code that includes
// "hgetf2_kernel.h" from its own .cu files keeps compiling: the `fp16` alias, swap_fp16, double_to_fp16,
// fp16_to_double.  Numerics of double_to_fp16 are part of the pivot-discovery parity: the value goes through
// float, saturates at +-65504 (no infinities), magnitudes under 6.10352e-5 become zero (no subnormals), then
// round-to-nearest-even to half.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

using fp16 = __half;

__device__ inline void swap_fp16(fp16 &x, fp16 &y) {
    const fp16 keep = x;
    x = y;
    y = keep;
}

__host__ __device__ inline fp16 double_to_fp16(double value) {
    float f = static_cast<float>(value);
    const float kMax = 65504.0f;          // largest finite half
    const float kTiny = 6.10352e-05f;     // smallest normal half (as spelled in the reference)
    f = (f > kMax) ? kMax : ((f < -kMax) ? -kMax : f);
    if (f < kTiny && f > -kTiny) f = 0.0f;
    return __float2half_rn(f);
}

__host__ __device__ inline double fp16_to_double(fp16 value) { return static_cast<double>(__half2float(value)); }

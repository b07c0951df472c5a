// On-device synthetic inputs: the reference's value distribution (values k/10, k = 0..99, matrix_generator.cpp:66)
// from a counter-based hash so any size can be generated in place, optionally made strictly column diagonally
// dominant (diag = column off-diagonal abs sum + 1).  Bit-identical to oracle/mplu_oracle.py:counter_matrix.
#include "../../include/mplu.h"
#include <cuda_runtime.h>
#include <cstdint>

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t elem10(uint64_t seed, uint64_t i, uint64_t j) {  // value in tenths, 0..99
    return (uint32_t)(splitmix64((seed << 40) | (i << 20) | j) % 100ull);
}

// one block per column: write the column, reduce its off-diagonal sum, patch the diagonal
__global__ void generate_kernel(double* A, long long lda, int n, uint64_t seed, int dominant) {
    const int j = blockIdx.x;
    unsigned long long s = 0;  // off-diagonal sum in exact integer tenths (so host and device agree bit for bit)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t v = elem10(seed, i, j);
        A[i + (long long)j * lda] = (double)v / 10.0;
        if (i != j) s += v;
    }
    if (!dominant) return;
    __shared__ unsigned long long sm[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
        A[j + (long long)j * lda] = (double)t / 10.0 + 1.0;
    }
}

// b = A * 1 (row sums), thread per row
__global__ void rowsum_kernel(const double* __restrict__ A, long long lda, int n, double* __restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += A[i + (long long)j * lda];
    b[i] = s;
}

}  // namespace

extern "C" int mplu_generate(int n, unsigned long long seed, int dominant, double* dA, long long lda, double* db,
                             void* stream) {
    if (n <= 0 || !dA || lda < n) return MPLU_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    generate_kernel<<<n, 256, 0, st>>>(dA, lda, n, seed, dominant);
    if (db) rowsum_kernel<<<(n + 127) / 128, 128, 0, st>>>(dA, lda, n, db);
    return (int)cudaGetLastError();
}

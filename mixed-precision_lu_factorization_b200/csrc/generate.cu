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

// Block-cyclic local part of the same (dominant) matrix: process (p, q) of a P x Q grid owns the nb x nb tiles (I, J)
// with I % P == p, J % Q == q, stored in ScaLAPACK local order.  One block per LOCAL column; the diagonal needs the
// column's sum over ALL global rows, which every owner recomputes from the hash (no communication).
__global__ void generate_local_kernel(double* A, long long lda, int n, uint64_t seed, int nb, int P, int p, int Q, int q) {
    const int lj = blockIdx.x;
    const int gj = ((lj / nb) * Q + q) * nb + lj % nb;
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t v = elem10(seed, i, gj);
        if (i != gj) s += v;
        const int ti = i / nb;
        if (ti % P == p && i != gj) A[(long long)(ti / P) * nb + i % nb + (long long)lj * lda] = (double)v / 10.0;
    }
    __shared__ unsigned long long sm[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ti = gj / nb;
        if (ti % P == p) {
            unsigned long long t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
            A[(long long)(ti / P) * nb + gj % nb + (long long)lj * lda] = (double)t / 10.0 + 1.0;
        }
    }
}

// b = A * 1 for the dominant matrix, full length, from the hash alone: b_i = (sum_{j != i} v(i,j) + sum_{k != i} v(k,i)) / 10 + 1
__global__ void rhs_from_hash_kernel(double* b, int n, uint64_t seed) {
    const int i = blockIdx.x;
    unsigned long long s = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x)
        if (j != i) s += elem10(seed, i, j) + elem10(seed, j, i);
    __shared__ unsigned long long sm[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
        b[i] = (double)t / 10.0 + 1.0;
    }
}

}  // namespace

extern "C" int mplu_generate_local(int n, unsigned long long seed, int nb, int P, int Q, int p, int q, double* dA_loc,
                                   long long lda, double* db_full, void* stream) {
    if (n <= 0 || nb <= 0 || n % nb || P <= 0 || Q <= 0 || p < 0 || p >= P || q < 0 || q >= Q) return MPLU_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = n / nb;
    const int nt = q < T ? (T - q + Q - 1) / Q : 0;
    if (dA_loc && nt > 0) generate_local_kernel<<<nt * nb, 256, 0, st>>>(dA_loc, lda, n, seed, nb, P, p, Q, q);
    if (db_full) rhs_from_hash_kernel<<<n, 256, 0, st>>>(db_full, n, seed);
    return (int)cudaGetLastError();
}

extern "C" int mplu_generate(int n, unsigned long long seed, int dominant, double* dA, long long lda, double* db,
                             void* stream) {
    if (n <= 0 || !dA || lda < n) return MPLU_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    generate_kernel<<<n, 256, 0, st>>>(dA, lda, n, seed, dominant);
    if (db) rowsum_kernel<<<(n + 127) / 128, 128, 0, st>>>(dA, lda, n, db);
    return (int)cudaGetLastError();
}

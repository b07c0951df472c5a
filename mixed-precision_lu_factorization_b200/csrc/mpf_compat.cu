// Drop-in implementations of the reference's public symbols on this library's own kernels:
//   void MPF(double*, int, int, int*)                        reference MPF.h:3 / MPF.cu:66-256
//   __global__ HGETF2_kernel(fp16*, int, int, int, int*)     reference hgetf2_kernel.cu:15-120
//   __global__ dgetf2_native_npv(int, int, double*, int)     reference dgetf2_native_npv.cu:11-36
// Semantics are the reference's ("mixed-precision pre-pivoting"): per panel of width r the pivot rows are found by
// an fp16 partial-pivot LU of the fp16-cast panel, the swaps are applied to the whole fp64 matrix, the pre-pivoted
// panel is factored in fp64 without pivoting, then U12 = L11^-1 A12 and A22 -= L21 U12 in fp64.  What changes is
// the execution: no per-column cudaMemcpy gathers (MPF.cu:108-115,168-175,193-200: 3r blocking copies per panel),
// no host round trip of the pivots (MPF.cu:146,158), 3 instead of 5 grid barriers per fp16 column, panels factored
// in place (ld = N), fp64 TRSM/GEMM written here instead of cuBLAS.
#include "../../include/MPF.h"
#include "../../include/dgetf2_native_npv.h"
#include "../../include/hgetf2_kernel.h"
#include "../../include/mplu.h"

#include <algorithm>
#include <cstdio>
#include <iostream>

int mplu_coop_blocks_limit(const void* kernel, int threads);  // dropin_kernels.cu (the two kernels live there)

namespace {

// panel16[c*rows + i] = double_to_fp16(A[(k+c)*lda + k+i])     (MPF.cu:106-121 in one kernel)
__global__ void gather_cast_kernel(const double* __restrict__ A, long long lda, int k, int rows, int cols,
                                   fp16* __restrict__ panel16) {
    const long long total = (long long)rows * cols;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e / rows), i = (int)(e - (long long)c * rows);
        panel16[e] = double_to_fp16(A[(long long)(k + c) * lda + k + i]);
    }
}

// LASWP with dlaswp semantics over all N columns (MPF.cu:42-59); also converts the panel-local pivots to global
// 1-based ones and stores them (MPF.cu:150-155) without leaving the device.
__global__ void laswp_kernel(double* A, long long lda, int ncols, int k, int cols, const int* __restrict__ ipiv_panel,
                             int* __restrict__ ipiv_global) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col == 0)
        for (int j = 0; j < cols; ++j) ipiv_global[k + j] = ipiv_panel[j] + k;
    if (col >= ncols) return;
    double* a = A + (long long)col * lda;
    for (int j = 0; j < cols; ++j) {
        const int cur = k + j, piv = ipiv_panel[j] - 1 + k;
        if (piv != cur) {
            const double t = a[cur];
            a[cur] = a[piv];
            a[piv] = t;
        }
    }
}

// U12 = L11^-1 * A12, L11 unit lower pc x pc at A[k,k], A12 = pc x ncols at A[k,k+pc]  (cublasDtrsm, MPF.cu:215-225)
// one thread per column of A12; L11 staged in shared memory (pc <= 64) else read through L1.
__global__ void trsm_unit_lower_kernel(double* A, long long lda, int k, int pc, int ncols) {
    extern __shared__ double sL[];  // pc*pc or nothing
    const bool staged = pc <= 64;
    if (staged) {
        for (int e = threadIdx.x; e < pc * pc; e += blockDim.x) sL[e] = A[(long long)(k + e / pc) * lda + k + e % pc];
        __syncthreads();
    }
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double* x = A + (long long)(k + pc + c) * lda + k;
    for (int i = 1; i < pc; ++i) {
        double s = x[i];
        for (int t = 0; t < i; ++t) {
            const double l = staged ? sL[t * pc + i] : A[(long long)(k + t) * lda + k + i];
            s -= l * x[t];
        }
        x[i] = s;
    }
}

// A22 -= L21 * U12 in fp64 (cublasDgemm, MPF.cu:230-239); rank-pc update, HBM-bound like the reference's.
// 64x64 tile of C per block, 256 threads, 4x4 outputs per thread, K staged through shared memory 32 at a time.
__global__ void __launch_bounds__(256)
dgemm_rank_update_kernel(double* A, long long lda, int k, int pc, int nt) {
    __shared__ double sLt[32][64 + 1];  // [kk][row]
    __shared__ double sU[32][64 + 1];   // [kk][col]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const double* L21 = A + (long long)k * lda + (k + pc);            // (row, kk) -> L21[row + kk*lda]
    const double* U12 = A + (long long)(k + pc) * lda + k;            // (kk, col) -> U12[kk + col*lda]
    double* C = A + (long long)(k + pc) * lda + (k + pc);
    double acc[4][4] = {};
    for (int kk0 = 0; kk0 < pc; kk0 += 32) {
        const int kn = min(32, pc - kk0);
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * 64; e += 256) {
            const int row = e & 63, kk = e >> 6;
            sLt[kk][row] = (kk < kn && r0 + row < nt) ? L21[(r0 + row) + (long long)(kk0 + kk) * lda] : 0.0;
        }
        for (int e = threadIdx.x; e < 32 * 64; e += 256) {
            const int kk = e & 31, col = e >> 5;
            sU[kk][col] = (kk < kn && c0 + col < nt) ? U12[(kk0 + kk) + (long long)(c0 + col) * lda] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sLt[kk][tx + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sU[kk][ty + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = c0 + ty + 16 * j;
        if (col >= nt) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + tx + 16 * i;
            if (row < nt) C[row + (long long)col * lda] -= acc[i][j];
        }
    }
}

int mpf_impl(double* h_A, int N, int r, int* IPIV) {
    if (!h_A || !IPIV || N <= 0 || r <= 0) return MPLU_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MPLU_E_NODEVICE;
    // unlike the reference (cudaSetDevice(0), MPF.cu:77) the caller's current device is kept
    const size_t nn = (size_t)N * (size_t)N;  // the reference computes N*N in int (overflow for N >= 46341)
    double* d_A = nullptr;
    fp16* d_panel16 = nullptr;
    int *d_ipiv_panel = nullptr, *d_ipiv = nullptr;
    cudaError_t e;
#define MPF_CK(x) do { e = (x); if (e != cudaSuccess) goto fail; } while (0)
    MPF_CK(cudaMalloc(&d_A, nn * sizeof(double)));
    MPF_CK(cudaMalloc(&d_panel16, (size_t)N * r * sizeof(fp16)));
    MPF_CK(cudaMalloc(&d_ipiv_panel, r * sizeof(int)));
    MPF_CK(cudaMalloc(&d_ipiv, (size_t)N * sizeof(int)));
    MPF_CK(cudaMemcpy(d_A, h_A, nn * sizeof(double), cudaMemcpyHostToDevice));
    MPF_CK(cudaMemcpy(d_ipiv, IPIV, (size_t)N * sizeof(int), cudaMemcpyHostToDevice));  // untouched entries survive
    {
        const int threads = 256;
        const int max_h = mplu_coop_blocks_limit((const void*)HGETF2_kernel, threads);
        const int max_d = mplu_coop_blocks_limit((const void*)dgetf2_native_npv, threads);
        for (int k = 0; k < N; k += r) {
            int pc = std::min(r, N - k);
            int pr = N - k;
            if (pr <= 1) continue;  // MPF.cu:104
            const long long total = (long long)pr * pc;
            gather_cast_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256>>>(d_A, N, k, pr, pc, d_panel16);
            {
                int blocks = std::min((pr + threads - 1) / threads, max_h);
                void* args[] = {&d_panel16, &pr, &pr, &pc, &d_ipiv_panel};
                MPF_CK(cudaLaunchCooperativeKernel((void*)HGETF2_kernel, dim3(blocks), dim3(threads), args, 0, 0));
            }
            laswp_kernel<<<(N + 255) / 256, 256>>>(d_A, N, N, k, pc, d_ipiv_panel, d_ipiv);
            {
                int blocks = std::min((pr + threads - 1) / threads, max_d);
                double* panel = d_A + (size_t)k * N + k;
                int ld = N;
                void* args[] = {&pr, &pc, &panel, &ld};
                MPF_CK(cudaLaunchCooperativeKernel((void*)dgetf2_native_npv, dim3(blocks), dim3(threads), args, 0, 0));
            }
            const int nt = N - k - pc;
            if (nt > 0) {
                const size_t sh = pc <= 64 ? (size_t)pc * pc * sizeof(double) : 0;
                trsm_unit_lower_kernel<<<(nt + 127) / 128, 128, sh>>>(d_A, N, k, pc, nt);
                dim3 grid((nt + 63) / 64, (nt + 63) / 64);
                dgemm_rank_update_kernel<<<grid, 256>>>(d_A, N, k, pc, nt);
            }
        }
    }
    MPF_CK(cudaGetLastError());
    MPF_CK(cudaMemcpy(h_A, d_A, nn * sizeof(double), cudaMemcpyDeviceToHost));
    MPF_CK(cudaMemcpy(IPIV, d_ipiv, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_A); cudaFree(d_panel16); cudaFree(d_ipiv_panel); cudaFree(d_ipiv);
    return 0;
fail:
    cudaFree(d_A); cudaFree(d_panel16); cudaFree(d_ipiv_panel); cudaFree(d_ipiv);
    return (int)e;
#undef MPF_CK
}

}  // namespace

extern "C" int mplu_MPF(double* h_A, int N, int r, int* IPIV) { return mpf_impl(h_A, N, r, IPIV); }

void MPF(double* h_A, int N, int r, int* IPIV) {
    const int rc = mpf_impl(h_A, N, r, IPIV);
    if (rc == MPLU_E_NODEVICE) std::cerr << "No CUDA devices available." << std::endl;  // as MPF.cu:73
    else if (rc != 0) std::cerr << "MPF: error " << rc << (rc > 0 ? " (" : "") << (rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "")
                                << (rc > 0 ? ")" : "") << std::endl;
}

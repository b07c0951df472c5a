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

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <iostream>

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------------------------------------
// fp16 panel LU with partial pivoting.  Arg-max of |a| over rows j..: among EQUAL maxima the reference keeps the lower
// slot at every merge of its 256-slot shared-memory tree (strict '>', hgetf2_kernel.cu:48-56) and the lower block in
// its scan over the blocks (:72-79), so the winner is the tied row with the smallest
//   order(row) = ((row - j) / 256) << 8 | bitreverse8((row - j) % 256)
// (not the first row: found by the live-reference parity test on a tie-rich input).  Here that is one 64-bit atomicMax
// per block on the key   (bits of |a| as fp16) << 32 | (0xFFFFFFFF - order(row)).
// Like the reference's g_block_max_* scratch (hgetf2_kernel.cu:6-7) the key slots are __device__ globals, so two
// concurrent launches on one device must not overlap (same restriction as the reference).
__device__ unsigned long long g_hgetf2_key[2];

__global__ void HGETF2_kernel(fp16 *panel, int ld, int rows, int cols, int *ipiv_panel) {
    cg::grid_group grid = cg::this_grid();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    __shared__ unsigned long long s_key[32];

    if (gtid == 0) { g_hgetf2_key[0] = 0ull; g_hgetf2_key[1] = 0ull; }
    grid.sync();

    for (int j = 0; j < cols; ++j) {
        unsigned long long* slot = &g_hgetf2_key[j & 1];
        // ---- 1. pivot search over rows j .. rows-1 of column j
        unsigned long long best = 0ull;
        for (long long r = j + gtid; r < rows; r += gsz) {
            const unsigned short bits = __half_as_ushort(__habs(panel[(long long)j * ld + r]));
            const unsigned rel = (unsigned)(r - j);
            const unsigned order = (rel & ~255u) | (__brev(rel & 255u) >> 24);
            const unsigned long long key = ((unsigned long long)bits << 32) | (unsigned long long)(0xFFFFFFFFu - order);
            best = key > best ? key : best;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            best = (threadIdx.x < ((blockDim.x + 31) >> 5)) ? s_key[threadIdx.x] : 0ull;
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if (threadIdx.x == 0 && best != 0ull) atomicMax(slot, best);
        }
        grid.sync();
        const unsigned long long win = *reinterpret_cast<volatile unsigned long long*>(slot);
        // all-zero (or empty) column: the reference's scan keeps its initial index j (hgetf2_kernel.cu:35,69)
        const unsigned word = 0xFFFFFFFFu - (unsigned)(win & 0xFFFFFFFFull);  // order(row) of the winner
        const int piv = ((win >> 32) == 0ull) ? j : j + (int)((word & ~255u) | (__brev(word & 255u) >> 24));
        if (gtid == 0) {
            ipiv_panel[j] = piv + 1;
            g_hgetf2_key[(j + 1) & 1] = 0ull;  // the other slot is idle during this column: clear it for column j+1
        }
        // ---- 2. swap rows j and piv across the panel's columns
        if (piv != j) {
            for (long long c = gtid; c < cols; c += gsz) swap_fp16(panel[c * ld + j], panel[c * ld + piv]);
        }
        grid.sync();
        // ---- 3. multipliers and rank-1 update, all in half arithmetic like the reference (hgetf2_kernel.cu:104-115)
        const fp16 pivot_val = panel[(long long)j * ld + j];
        for (long long r = j + 1 + gtid; r < rows; r += gsz) {
            const fp16 mult = panel[(long long)j * ld + r] / pivot_val;
            panel[(long long)j * ld + r] = mult;
            for (int k = j + 1; k < cols; ++k) panel[(long long)k * ld + r] -= mult * panel[(long long)k * ld + j];
        }
        grid.sync();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// fp64 no-pivot panel LU, in place.  Same arithmetic as the reference (quotient, then a -= m*b contracted to DFMA);
// the pivot row of each step is staged in shared memory and rows are grid-strided.
__global__ void dgetf2_native_npv(int m, int n, double *panel, int ld) {
    cg::grid_group grid = cg::this_grid();
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    __shared__ double s_row[256];  // pivot row, 256 columns at a time
    for (int j = 0; j < n; ++j) {
        const double pivot_val = panel[(long long)j * ld + j];
        for (int k0 = j + 1; k0 < n || k0 == j + 1; k0 += 256) {
            const int kn = min(256, n - k0);
            __syncthreads();
            for (int t = threadIdx.x; t < kn; t += blockDim.x) s_row[t] = panel[(long long)(k0 + t) * ld + j];
            __syncthreads();
            for (long long r = j + 1 + gtid; r < m; r += gsz) {
                double mult;
                if (k0 == j + 1) {
                    mult = panel[(long long)j * ld + r] / pivot_val;
                    panel[(long long)j * ld + r] = mult;
                } else {
                    mult = panel[(long long)j * ld + r];
                }
                for (int t = 0; t < kn; ++t) panel[(long long)(k0 + t) * ld + r] -= mult * s_row[t];
            }
            if (kn <= 0) break;
        }
        grid.sync();
    }
}

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

int coop_blocks_limit(const void* kernel, int threads) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    return std::max(1, sms * per_sm);
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
        const int max_h = coop_blocks_limit((const void*)HGETF2_kernel, threads);
        const int max_d = coop_blocks_limit((const void*)dgetf2_native_npv, threads);
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

// Launch the two drop-in kernels on device-resident panels exactly the way the reference's caller does
// (cooperative, ceil(rows/256) x 256: MPF.cu:126-133,178-185), capped at the co-residency limit.
extern "C" int mplu_hgetf2(void* d_panel, int ld, int rows, int cols, int* d_ipiv, void* stream) {
    if (!d_panel || !d_ipiv || rows <= 0 || cols <= 0 || ld < rows) return MPLU_E_ARG;
    const int threads = 256;
    int blocks = std::min((rows + threads - 1) / threads, coop_blocks_limit((const void*)HGETF2_kernel, threads));
    fp16* panel = (fp16*)d_panel;
    void* args[] = {&panel, &ld, &rows, &cols, &d_ipiv};
    return (int)cudaLaunchCooperativeKernel((void*)HGETF2_kernel, dim3(blocks), dim3(threads), args, 0, (cudaStream_t)stream);
}

extern "C" int mplu_dgetf2_npv(int m, int n, double* d_panel, int ld, void* stream) {
    if (!d_panel || m <= 0 || n <= 0 || ld < m) return MPLU_E_ARG;
    const int threads = 256;
    int blocks = std::min((m + threads - 1) / threads, coop_blocks_limit((const void*)dgetf2_native_npv, threads));
    void* args[] = {&m, &n, &d_panel, &ld};
    return (int)cudaLaunchCooperativeKernel((void*)dgetf2_native_npv, dim3(blocks), dim3(threads), args, 0, (cudaStream_t)stream);
}

void MPF(double* h_A, int N, int r, int* IPIV) {
    const int rc = mpf_impl(h_A, N, r, IPIV);
    if (rc == MPLU_E_NODEVICE) std::cerr << "No CUDA devices available." << std::endl;  // as MPF.cu:73
    else if (rc != 0) std::cerr << "MPF: error " << rc << (rc > 0 ? " (" : "") << (rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "")
                                << (rc > 0 ? ")" : "") << std::endl;
}

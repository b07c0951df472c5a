// The reference's two kernels as drop-ins (same C++ symbols, argument lists and launch geometry contract):
//   __global__ HGETF2_kernel(fp16*, int, int, int, int*)     reference hgetf2_kernel.cu:15-120
//   __global__ dgetf2_native_npv(int, int, double*, int)     reference dgetf2_native_npv.cu:11-36
// A translation unit of its own so that it can also be linked, as an object / static archive, into a FOREIGN program --
// e.g. the reference's unmodified MPF.cu, which launches both with cudaLaunchCooperativeKernel (MPF.cu:126-133,178-185).
// (A kernel's host stub must be registered with the CUDA runtime instance that launches it: libmplu.so links the
// runtime statically, so a foreign TU links this object instead of taking the stub out of the shared library;
// INTEGRATION.md section 2, tests/test_gpu_boundary.py.)
#include "../../include/dgetf2_native_npv.h"
#include "../../include/hgetf2_kernel.h"
#include "../../include/mplu.h"

#include <cooperative_groups.h>
#include <algorithm>

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

int mplu_coop_blocks_limit(const void* kernel, int threads) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0);
    return std::max(1, sms * per_sm);
}

// Launch the two drop-in kernels on device-resident panels exactly the way the reference's caller does
// (cooperative, ceil(rows/256) x 256: MPF.cu:126-133,178-185), capped at the co-residency limit.
extern "C" int mplu_hgetf2(void* d_panel, int ld, int rows, int cols, int* d_ipiv, void* stream) {
    if (!d_panel || !d_ipiv || rows <= 0 || cols <= 0 || ld < rows) return MPLU_E_ARG;
    const int threads = 256;
    int blocks = std::min((rows + threads - 1) / threads, mplu_coop_blocks_limit((const void*)HGETF2_kernel, threads));
    fp16* panel = (fp16*)d_panel;
    void* args[] = {&panel, &ld, &rows, &cols, &d_ipiv};
    return (int)cudaLaunchCooperativeKernel((void*)HGETF2_kernel, dim3(blocks), dim3(threads), args, 0, (cudaStream_t)stream);
}

extern "C" int mplu_dgetf2_npv(int m, int n, double* d_panel, int ld, void* stream) {
    if (!d_panel || m <= 0 || n <= 0 || ld < m) return MPLU_E_ARG;
    const int threads = 256;
    int blocks = std::min((m + threads - 1) / threads, mplu_coop_blocks_limit((const void*)dgetf2_native_npv, threads));
    void* args[] = {&m, &n, &d_panel, &ld};
    return (int)cudaLaunchCooperativeKernel((void*)dgetf2_native_npv, dim3(blocks), dim3(threads), args, 0, (cudaStream_t)stream);
}


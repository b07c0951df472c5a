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
// Round 2: the panel width the reference's driver uses (r <= 32) takes a fast path with the SAME arithmetic per element --
//   * pivot discovery: every thread keeps its fp16 row in registers (the fp64 -> fp16 cast fused in), rows are never
//     physically swapped (only the pivot sequence leaves the kernel, MPF.cu:146-155 discards the fp16 factors): ONE
//     grid barrier per column instead of three; the reference's tie order is kept by tracking each row's position;
//   * fp64 panel: after the pre-pivoting the rows are independent -- the r x r top block is factored by one CTA, then
//     every row solves x U11 = a on its own: no grid barriers (same quotient / DFMA sequence per element);
//   * A22 -= L21 U12 on the fp64 tensor cores (mma.sync m8n8k4 f64), 128x64 tiles, addend loaded before the products;
//   * host <-> device copies through pinned staging buffers filled by several host threads (pageable cudaMemcpy moved the
//     2 x 2 GiB of an n = 16384 call at 12 GB/s: 350 of the call's 1126 ms)
// -- and the g_hgetf2_key device globals are no longer touched by MPF() (per-call workspace).
#include "../../include/MPF.h"
#include "../../include/dgetf2_native_npv.h"
#include "../../include/hgetf2_kernel.h"
#include "../../include/mplu.h"

#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <thread>
#include <vector>

namespace cg = cooperative_groups;

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
__global__ void trsm_unit_lower_kernel(double* A, long long lda, int k, int pc, int col0, int ncols) {  // columns [col0, col0 + ncols)
    extern __shared__ double sL[];  // pc*pc or nothing
    const bool staged = pc <= 64;
    if (staged) {
        for (int e = threadIdx.x; e < pc * pc; e += blockDim.x) sL[e] = A[(long long)(k + e / pc) * lda + k + e % pc];
        __syncthreads();
    }
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    double* x = A + (long long)(col0 + c) * lda + k;
    for (int i = 1; i < pc; ++i) {
        double s = x[i];
        for (int t = 0; t < i; ++t) {
            const double l = staged ? sL[t * pc + i] : A[(long long)(k + t) * lda + k + i];
            s -= l * x[t];
        }
        x[i] = s;
    }
}

// ---- fast path, pivot discovery (reference: double_to_fp16_block + HGETF2_kernel, MPF.cu:106-133, hgetf2_kernel.cu:15-120).
// Each thread owns up to HP_RPT rows of the panel, cast from the fp64 matrix and kept in registers for all `cols` steps.
// Rows stay where they are: a row's POSITION (what the reference's physical swaps would have made of it) is tracked by its
// owner, pos2phys[] maps a position back to the row that sits there.  Step j: arg-max key of the rows at positions >= j
// (same 64-bit key as HGETF2_kernel above: |a| bits, then the reference's slot order of position - j) -> one atomicMax per
// block -> grid barrier -> everybody reads the winner's row from the global mirror (written by its owner before the
// barrier), the row at position j takes the winner's position, the other live rows form their multiplier and update in
// half arithmetic (quotient, product, difference: hgetf2_kernel.cu:104-115) and refresh their mirror rows.
constexpr int HP_RPT = 2, HP_COLS = 32, HP_THREADS = 256;
__global__ void __launch_bounds__(HP_THREADS)
hpivot_kernel(const double* __restrict__ A, long long lda, int k, int rows, int cols, int* __restrict__ ipiv_panel,
              fp16* mirror, int* pos2phys, unsigned long long* keys /* 3 slots, zero */) {
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    __shared__ unsigned long long s_key[HP_THREADS / 32];
    __shared__ fp16 s_u[HP_COLS];
    fp16 row[HP_RPT][HP_COLS];
    int pos[HP_RPT];
#pragma unroll
    for (int q = 0; q < HP_RPT; ++q) {
        const int r = gtid + q * gsz;
        pos[q] = r < rows ? r : -1;
        if (r < rows) {
            pos2phys[r] = r;
#pragma unroll
            for (int c = 0; c < HP_COLS; ++c) {
                row[q][c] = c < cols ? double_to_fp16(A[(long long)(k + c) * lda + k + r]) : __float2half(0.f);
                mirror[(long long)r * HP_COLS + c] = row[q][c];
            }
        }
    }
#pragma unroll 1
    for (int j = 0; j < cols; ++j) {
        unsigned long long best = 0ull;
#pragma unroll
        for (int q = 0; q < HP_RPT; ++q) {
            if (pos[q] < j) continue;
            fp16 v = row[q][0];
#pragma unroll
            for (int c = 1; c < HP_COLS; ++c) v = (c == j) ? row[q][c] : v;
            const unsigned rel = (unsigned)(pos[q] - j);
            const unsigned order = (rel & ~255u) | (__brev(rel & 255u) >> 24);
            // |a| (16 bits) | 0xFFFFFF - order (24) | the row itself (24): the winner's mirror row is found without a look-up
            const unsigned long long key = ((unsigned long long)__half_as_ushort(__habs(v)) << 48) |
                                           ((unsigned long long)(0xFFFFFFu - order) << 24) | (unsigned long long)(gtid + q * gsz);
            best = key > best ? key : best;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            best = threadIdx.x < HP_THREADS / 32 ? s_key[threadIdx.x] : 0ull;
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other > best ? other : best;
            }
            if (threadIdx.x == 0 && best != 0ull) atomicMax(&keys[j % 3], best);
        }
        grid.sync();
        const unsigned long long win = *reinterpret_cast<volatile unsigned long long*>(&keys[j % 3]);
        const unsigned word = 0xFFFFFFu - (unsigned)((win >> 24) & 0xFFFFFFull);
        // all-zero (or empty) column: the reference keeps its initial index j (hgetf2_kernel.cu:35,69)
        const bool zero_col = (win >> 48) == 0ull;
        const int piv = zero_col ? j : j + (int)((word & ~255u) | (__brev(word & 255u) >> 24));
        // the row at position j (only needed for a zero column) was recorded before an earlier barrier
        const int pphys = zero_col ? *reinterpret_cast<volatile int*>(&pos2phys[j]) : (int)(win & 0xFFFFFFull);
        if (gtid == 0) {
            ipiv_panel[j] = piv + 1;
            keys[(j + 2) % 3] = 0ull;  // idle until column j+2, whose atomics come after the next barrier
        }
        if (threadIdx.x < HP_COLS) s_u[threadIdx.x] = mirror[(long long)pphys * HP_COLS + threadIdx.x];
        __syncthreads();
        const fp16 pivot_val = s_u[j];
#pragma unroll
        for (int q = 0; q < HP_RPT; ++q) {
            if (pos[q] < j) continue;
            const int phys = gtid + q * gsz;
            if (phys == pphys) {               // this row is the pivot row of step j: it moves to position j and retires
                if (pos[q] == j) { pos[q] = -1; continue; }
                pos[q] = -1;
                continue;
            }
            if (pos[q] == j) {                 // the row that sat at position j takes the winner's position
                pos[q] = piv;
                pos2phys[piv] = phys;
            }
            fp16 mult = __float2half(0.f);
#pragma unroll
            for (int c = 0; c < HP_COLS; ++c)
                if (c == j) { mult = row[q][c] / pivot_val; row[q][c] = mult; }
#pragma unroll
            for (int c = 1; c < HP_COLS; ++c)
                if (c > j && c < cols) {
                    row[q][c] -= mult * s_u[c];
                    mirror[(long long)phys * HP_COLS + c] = row[q][c];
                }
        }
        __syncthreads();  // s_u is rewritten in the next step
    }
}

// The same pivot discovery inside ONE thread-block cluster (8 or 16 CTAs x 512 threads x 2 rows = up to 16384 rows): the
// grid version above spends ~10 us per column on its grid barrier and three L2 round trips; here the per-CTA arg-max keys
// and the pivot row travel through distributed shared memory (st.shared::cluster into every CTA of the cluster) between
// two barrier.cluster per column, and nothing touches global memory after the initial cast.  The winner recognises itself
// by its key (keys are unique: they contain the row's position), so neither a mirror of the rows nor a position table is
// needed.  Same keys, same tie order, same half arithmetic per element as hpivot_kernel / HGETF2_kernel.
constexpr int HC_THREADS = 512, HC_RPT = 2, HC_MAXCS = 16;
__device__ __forceinline__ unsigned hc_mapa(const void* p, unsigned cta) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(cta));
    return r;
}
__global__ void __launch_bounds__(HC_THREADS, 1)
hpivot_cluster_kernel(const double* __restrict__ A, long long lda, int k, int rows, int cols, int* __restrict__ ipiv_panel) {
    __shared__ unsigned long long s_warp[HC_THREADS / 32];
    __shared__ __align__(16) unsigned long long s_keys[2][HC_MAXCS];  // [column parity][CTA]: every CTA's best key
    __shared__ __align__(16) fp16 s_u[2][HP_COLS];                    // [column parity]: the pivot row
    unsigned crank, csize;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
    const int gtid = (int)crank * HC_THREADS + threadIdx.x, gsz = (int)csize * HC_THREADS;
    fp16 row[HC_RPT][HP_COLS];
    int pos[HC_RPT];
#pragma unroll
    for (int q = 0; q < HC_RPT; ++q) {
        const int r = gtid + q * gsz;
        pos[q] = r < rows ? r : -1;
#pragma unroll
        for (int c = 0; c < HP_COLS; ++c)
            row[q][c] = (r < rows && c < cols) ? double_to_fp16(A[(long long)(k + c) * lda + k + r]) : __float2half(0.f);
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");  // all CTAs are running
#pragma unroll 1
    for (int j = 0; j < cols; ++j) {
        const int par = j & 1;
        unsigned long long mykey[HC_RPT], best = 0ull;
#pragma unroll
        for (int q = 0; q < HC_RPT; ++q) {
            mykey[q] = 0ull;
            if (pos[q] < j) continue;
            fp16 v = row[q][0];
#pragma unroll
            for (int c = 1; c < HP_COLS; ++c) v = (c == j) ? row[q][c] : v;
            const unsigned rel = (unsigned)(pos[q] - j);
            const unsigned order = (rel & ~255u) | (__brev(rel & 255u) >> 24);
            mykey[q] = ((unsigned long long)__half_as_ushort(__habs(v)) << 48) | ((unsigned long long)(0xFFFFFFu - order) << 24) | 1ull;
            best = mykey[q] > best ? mykey[q] : best;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < csize) {  // thread t delivers this CTA's best key to CTA t
            unsigned long long b = 0ull;
#pragma unroll
            for (int w = 0; w < HC_THREADS / 32; ++w) b = s_warp[w] > b ? s_warp[w] : b;
            asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(hc_mapa(&s_keys[par][crank], threadIdx.x)), "l"(b) : "memory");
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        unsigned long long win = 0ull;
        for (unsigned i = 0; i < csize; ++i) win = s_keys[par][i] > win ? s_keys[par][i] : win;
        const bool zero_col = (win >> 48) == 0ull;  // all-zero (or empty) column: the reference keeps index j (hgetf2_kernel.cu:35,69)
        const unsigned word = 0xFFFFFFu - (unsigned)((win >> 24) & 0xFFFFFFull);
        const int piv = zero_col ? j : j + (int)((word & ~255u) | (__brev(word & 255u) >> 24));
        if (gtid == 0) ipiv_panel[j] = piv + 1;
        // the pivot row's owner sends it to every CTA
#pragma unroll
        for (int q = 0; q < HC_RPT; ++q) {
            if (pos[q] != piv) continue;
            unsigned w32[HP_COLS / 2];
#pragma unroll
            for (int c = 0; c < HP_COLS; c += 2) w32[c >> 1] = (unsigned)__half_as_ushort(row[q][c]) | ((unsigned)__half_as_ushort(row[q][c + 1]) << 16);
            for (unsigned i = 0; i < csize; ++i) {
                const unsigned dst = hc_mapa(&s_u[par][0], i);
#pragma unroll
                for (int v4 = 0; v4 < HP_COLS / 8; ++v4)
                    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16 * v4), "r"(w32[4 * v4]), "r"(w32[4 * v4 + 1]),
                                 "r"(w32[4 * v4 + 2]), "r"(w32[4 * v4 + 3]) : "memory");
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        const fp16 pivot_val = s_u[par][j];
#pragma unroll
        for (int q = 0; q < HC_RPT; ++q) {
            if (pos[q] < j) continue;
            if (pos[q] == piv) { pos[q] = -1; continue; }  // the pivot row moves to position j and retires
            if (pos[q] == j) pos[q] = piv;                  // the row that sat at position j takes the winner's position
            fp16 mult = __float2half(0.f);
#pragma unroll
            for (int c = 0; c < HP_COLS; ++c)
                if (c == j) { mult = row[q][c] / pivot_val; row[q][c] = mult; }
#pragma unroll
            for (int c = 1; c < HP_COLS; ++c)
                if (c > j && c < cols) row[q][c] -= mult * s_u[par][c];
        }
    }
    // no CTA may exit while another one can still store into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- fast path, fp64 panel without pivoting (reference: dgetf2_native_npv.cu:11-36 on the pre-pivoted panel).
// The pc x pc top block: one CTA, column by column in shared memory (quotient, then a -= m * b as DFMA, like the kernel).
__global__ void dpanel_top_kernel(double* A, long long lda, int k, int pc) {
    __shared__ double s[HP_COLS][HP_COLS + 1];  // s[r][c]
    const int t = threadIdx.x;
    for (int e = t; e < pc * pc; e += blockDim.x) s[e % pc][e / pc] = A[(long long)(k + e / pc) * lda + k + e % pc];
    __syncthreads();
    for (int j = 0; j < pc; ++j) {
        const double pivot_val = s[j][j];
        __syncthreads();
        if (t > j && t < pc) s[t][j] = s[t][j] / pivot_val;
        __syncthreads();
        for (int e = t; e < pc * pc; e += blockDim.x) {
            const int r = e % pc, c = e / pc;
            if (r > j && c > j) s[r][c] -= s[r][j] * s[j][c];
        }
        __syncthreads();
    }
    for (int e = t; e < pc * pc; e += blockDim.x) A[(long long)(k + e / pc) * lda + k + e % pc] = s[e % pc][e / pc];
}
// The rows below it: row i solves x U11 = a(i, :) on its own -- the very operations the column-by-column elimination
// applies to that row, in the same order.
__global__ void __launch_bounds__(128)
dpanel_rows_kernel(double* A, long long lda, int k, int pc, int nrows /* rows below the top block */) {
    __shared__ double sU[HP_COLS][HP_COLS + 1];  // sU[j][c] = U11(j, c)
    for (int e = threadIdx.x; e < pc * pc; e += blockDim.x) sU[e % pc][e / pc] = A[(long long)(k + e / pc) * lda + k + e % pc];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    double* a = A + (long long)k * lda + k + pc + i;
    double x[HP_COLS];
#pragma unroll
    for (int c = 0; c < HP_COLS; ++c) x[c] = c < pc ? a[(long long)c * lda] : 0.0;
#pragma unroll
    for (int j = 0; j < HP_COLS; ++j) {
        if (j < pc) {
            const double mult = x[j] / sU[j][j];
            x[j] = mult;
#pragma unroll
            for (int c = j + 1; c < HP_COLS; ++c)
                if (c < pc) x[c] -= mult * sU[j][c];
        }
    }
#pragma unroll
    for (int c = 0; c < HP_COLS; ++c)
        if (c < pc) a[(long long)c * lda] = x[c];
}

// ---- A22 -= L21 * U12 on the fp64 tensor cores (cublasDgemm of MPF.cu:230-239).  128 x 64 tile of C per block, 8 warps
// of 32 x 32 (4 x 4 fragments of mma.sync.m8n8k4.f64); the addend is loaded into the accumulators first so that its
// HBM latency hides behind the operand staging; K in chunks of 32 through shared memory (leading dimensions = 4 mod 16:
// the fragment loads of a half-warp hit 16 different bank pairs).  Rank-32 updates are HBM-bound: 16 bytes of C traffic
// per 64 flops.
#ifndef MPLU_DM_BM
#define MPLU_DM_BM 64
#endif
constexpr int DM_BM = MPLU_DM_BM, DM_BN = 64, DM_KC = 32, DM_LDA = DM_BM + 4, DM_LDB = DM_BN + 4, DM_THREADS = 2 * DM_BM;
constexpr int DM_SMEM = (DM_KC * DM_LDA + DM_KC * DM_LDB) * (int)sizeof(double);
// a block's phases (addend loads, products, stores) are serial, so the overlap has to come from independent blocks:
// 64 x 64 tiles / 128 threads / 4 blocks per SM (MPLU_DM_BM=128: 128 x 64 / 256 / 2, measured slower)
__global__ void __launch_bounds__(DM_THREADS, 512 / DM_THREADS)
dmma_rank_update_kernel(double* A, long long lda, int row0, int col0, int kk_first, int pc, int mrows, int ncols) {
    // C = A[row0 .. row0+mrows, col0 .. col0+ncols) -= A[row0.., kk_first .. kk_first+pc) * A[kk_first .. kk_first+pc, col0..)
    extern __shared__ double dm_smem[];
    double* As = dm_smem;                  // As[kk][m]
    double* Bs = dm_smem + DM_KC * DM_LDA;  // Bs[kk][n]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm = warp % (DM_BM / 32), wn = warp / (DM_BM / 32);
    const int r0 = blockIdx.x * DM_BM, c0 = blockIdx.y * DM_BN;
    const double* L21 = A + (long long)kk_first * lda + row0;
    const double* U12 = A + (long long)col0 * lda + kk_first;
    double* C = A + (long long)col0 * lda + row0;
    double acc[4][4][2];
    // this thread's 4 x 8 addend elements: rows rb + 8 mf, columns cb + 8 nf + e
    const int rb = r0 + wm * 32 + g, cb = c0 + wn * 32 + 2 * t;
    double* Ct = C + rb + (long long)cb * lda;
    const bool interior = r0 + DM_BM <= mrows && c0 + DM_BN <= ncols;
    if (interior) {
#pragma unroll
        for (int nf = 0; nf < 4; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mf = 0; mf < 4; ++mf) acc[mf][nf][e] = Ct[mf * 8 + (long long)(nf * 8 + e) * lda];
    } else {
#pragma unroll
        for (int nf = 0; nf < 4; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mf = 0; mf < 4; ++mf)
                    acc[mf][nf][e] = (rb + mf * 8 < mrows && cb + nf * 8 + e < ncols) ? Ct[mf * 8 + (long long)(nf * 8 + e) * lda] : 0.0;
    }
    for (int kk0 = 0; kk0 < pc; kk0 += DM_KC) {
        const int kn = min(DM_KC, pc - kk0);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < DM_KC * DM_BM / DM_THREADS; ++i) {
            const int e = tid + i * DM_THREADS, m = e % DM_BM, kk = e / DM_BM;
            As[kk * DM_LDA + m] = (kk < kn && r0 + m < mrows) ? -L21[(r0 + m) + (long long)(kk0 + kk) * lda] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < DM_KC * DM_BN / DM_THREADS; ++i) {
            const int e = tid + i * DM_THREADS, kk = e & (DM_KC - 1), n = e >> 5;
            Bs[kk * DM_LDB + n] = (kk < kn && c0 + n < ncols) ? U12[(kk0 + kk) + (long long)(c0 + n) * lda] : 0.0;
        }
        __syncthreads();
#pragma unroll 2
        for (int ks = 0; ks < DM_KC / 4; ++ks) {
            double a[4], b[4];
#pragma unroll
            for (int mf = 0; mf < 4; ++mf) a[mf] = As[(ks * 4 + t) * DM_LDA + wm * 32 + mf * 8 + g];
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) b[nf] = Bs[(ks * 4 + t) * DM_LDB + wn * 32 + nf * 8 + g];
#pragma unroll
            for (int mf = 0; mf < 4; ++mf)
#pragma unroll
                for (int nf = 0; nf < 4; ++nf)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                                 : "+d"(acc[mf][nf][0]), "+d"(acc[mf][nf][1])
                                 : "d"(a[mf]), "d"(b[nf]));
        }
    }
    if (interior) {
#pragma unroll
        for (int nf = 0; nf < 4; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mf = 0; mf < 4; ++mf) Ct[mf * 8 + (long long)(nf * 8 + e) * lda] = acc[mf][nf][e];
    } else {
#pragma unroll
        for (int nf = 0; nf < 4; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e)
#pragma unroll
                for (int mf = 0; mf < 4; ++mf)
                    if (rb + mf * 8 < mrows && cb + nf * 8 + e < ncols) Ct[mf * 8 + (long long)(nf * 8 + e) * lda] = acc[mf][nf][e];
    }
}

// ---- host <-> device copies of a pageable buffer through pinned staging: T host threads, each with two 8 MiB pinned
// slots and its own stream, copy interleaved chunks (host memcpy of one chunk overlaps the DMA of the previous one)
struct StagePool {
    static constexpr int kThreads = 8;
    static constexpr size_t kChunk = (size_t)8 << 20;
    std::mutex mu;
    void* slot[kThreads][2] = {};
    cudaStream_t st[kThreads] = {};
    cudaEvent_t ev[kThreads][2] = {};
    int device = -1;
    bool ok = false;
    bool init() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (ok && dev == device) return true;
        if (ok) return false;  // one device per process for the staged path; others use plain cudaMemcpy
        for (int t = 0; t < kThreads; ++t) {
            if (cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking) != cudaSuccess) return false;
            for (int i = 0; i < 2; ++i) {
                if (cudaHostAlloc(&slot[t][i], kChunk, cudaHostAllocDefault) != cudaSuccess) return false;
                if (cudaEventCreateWithFlags(&ev[t][i], cudaEventDisableTiming) != cudaSuccess) return false;
            }
        }
        device = dev;
        ok = true;
        return true;
    }
};
StagePool g_stage;

cudaError_t staged_copy(void* dst, const void* src, size_t bytes, bool h2d) {
    std::unique_lock<std::mutex> lock(g_stage.mu, std::try_to_lock);
    if (bytes < 4 * StagePool::kChunk || !lock.owns_lock() || !g_stage.init())
        return cudaMemcpy(dst, src, bytes, h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost);
    const size_t nchunks = (bytes + StagePool::kChunk - 1) / StagePool::kChunk;
    const int dev = g_stage.device;
    std::vector<cudaError_t> err(StagePool::kThreads, cudaSuccess);
    std::vector<std::thread> th;
    for (int t = 0; t < StagePool::kThreads; ++t)
        th.emplace_back([&, t]() {
            cudaError_t e = cudaSetDevice(dev);
            auto len_of = [&](size_t c) { return std::min(StagePool::kChunk, bytes - c * StagePool::kChunk); };
            int n = 0;
            size_t prev = (size_t)-1;
            for (size_t c = t; e == cudaSuccess && c < nchunks; c += StagePool::kThreads, ++n) {
                const int s = n & 1;
                const size_t off = c * StagePool::kChunk, len = len_of(c);
                if (h2d) {
                    if (n >= 2) e = cudaEventSynchronize(g_stage.ev[t][s]);  // the slot's previous DMA has read it
                    memcpy(g_stage.slot[t][s], (const char*)src + off, len);
                    if (e == cudaSuccess) e = cudaMemcpyAsync((char*)dst + off, g_stage.slot[t][s], len, cudaMemcpyHostToDevice, g_stage.st[t]);
                    if (e == cudaSuccess) e = cudaEventRecord(g_stage.ev[t][s], g_stage.st[t]);
                } else {
                    e = cudaMemcpyAsync(g_stage.slot[t][s], (const char*)src + off, len, cudaMemcpyDeviceToHost, g_stage.st[t]);
                    if (e == cudaSuccess) e = cudaEventRecord(g_stage.ev[t][s], g_stage.st[t]);
                    if (prev != (size_t)-1 && e == cudaSuccess) {  // while this chunk is in flight, unload the previous one
                        e = cudaEventSynchronize(g_stage.ev[t][s ^ 1]);
                        memcpy((char*)dst + prev * StagePool::kChunk, g_stage.slot[t][s ^ 1], len_of(prev));
                    }
                    prev = c;
                }
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_stage.st[t]);
            if (!h2d && prev != (size_t)-1 && e == cudaSuccess) memcpy((char*)dst + prev * StagePool::kChunk, g_stage.slot[t][(n - 1) & 1], len_of(prev));
            err[t] = e;
        });
    for (auto& x : th) x.join();
    for (cudaError_t e : err)
        if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace

namespace mplu_detail {
// Factor the device-resident column-major fp64 matrix in place with the reference's semantics (MPF.cu:100-241); the
// 1-based global pivots go to d_ipiv.  Everything is enqueued on the legacy default stream.  Also the factorization of
// the solver's full-precision fallback (fp64_fallback.cu).
cudaError_t mpf_device(double* d_A, int N, int r, int* d_ipiv) {
    cudaError_t e = cudaSuccess;
#define MPF_CK(x) do { e = (x); if (e != cudaSuccess) goto done; } while (0)
    fp16* d_panel16 = nullptr;
    fp16* d_mirror = nullptr;
    int *d_ipiv_panel = nullptr, *d_pos2phys = nullptr;
    unsigned long long* d_keys = nullptr;
    // per device, and cheap: set on every call rather than cached per process
    cudaFuncSetAttribute(dmma_rank_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM);
    const int threads = 256;
    const bool fast = r <= HP_COLS;
    // super-panel width of the delayed updates: the largest multiple of r up to 256 (MPLU_MPF_SP=0: none)
    int sp = r;
    {
        const char* env = getenv("MPLU_MPF_SP");
        const int want = env ? atoi(env) : 256;
        if (want > r) sp = (want / r) * r;
    }
    const int max_hp = mplu_coop_blocks_limit((const void*)hpivot_kernel, HP_THREADS);
    bool cluster_ok = getenv("MPLU_MPF_NO_CLUSTER") == nullptr &&
                      cudaFuncSetAttribute(hpivot_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (fast) {
        MPF_CK(cudaMalloc(&d_mirror, (size_t)N * HP_COLS * sizeof(fp16)));
        MPF_CK(cudaMalloc(&d_pos2phys, (size_t)N * sizeof(int)));
        MPF_CK(cudaMalloc(&d_keys, 3 * sizeof(unsigned long long)));
    }
    MPF_CK(cudaMalloc(&d_ipiv_panel, (size_t)std::max(r, 1) * sizeof(int)));
    {
        const int max_h = mplu_coop_blocks_limit((const void*)HGETF2_kernel, threads);
        const int max_d = mplu_coop_blocks_limit((const void*)dgetf2_native_npv, threads);
        for (int k = 0; k < N; k += r) {
            int pc = std::min(r, N - k);
            int pr = N - k;
            if (pr <= 1) continue;  // MPF.cu:104
            // ---- pivot discovery in fp16
            int hp_blocks = std::min((pr + HP_THREADS - 1) / HP_THREADS, max_hp);
            bool done_pivots = false;
            if (fast && cluster_ok && pr <= HC_MAXCS * HC_THREADS * HC_RPT) {
                // one cluster: 8 CTAs (portable) up to 8192 rows, 16 (opt-in) up to 16384
                const int cs = pr <= 8 * HC_THREADS * HC_RPT ? 8 : 16;
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(cs);
                cfg.blockDim = dim3(HC_THREADS);
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                const double* Ac = d_A;
                long long lda = N;
                const cudaError_t le = cudaLaunchKernelEx(&cfg, hpivot_cluster_kernel, Ac, lda, k, pr, pc, d_ipiv_panel);
                if (le == cudaSuccess) done_pivots = true;
                else { cudaGetLastError(); cluster_ok = false; }  // e.g. no room for a 16-CTA cluster: the grid version below
            }
            if (done_pivots) {
            } else if (fast && (long long)hp_blocks * HP_THREADS * HP_RPT >= pr) {
                // few rows per thread while the grid stays small: a grid barrier costs more with more blocks
                hp_blocks = std::min(hp_blocks, std::max(1, (pr + 2 * HP_THREADS - 1) / (2 * HP_THREADS)));
                while ((long long)hp_blocks * HP_THREADS * HP_RPT < pr) ++hp_blocks;
                MPF_CK(cudaMemsetAsync(d_keys, 0, 3 * sizeof(unsigned long long), 0));
                long long lda = N;
                const double* Ac = d_A;
                void* args[] = {&Ac, &lda, &k, &pr, &pc, &d_ipiv_panel, &d_mirror, &d_pos2phys, &d_keys};
                MPF_CK(cudaLaunchCooperativeKernel((void*)hpivot_kernel, dim3(hp_blocks), dim3(HP_THREADS), args, 0, 0));
            } else {
                if (!d_panel16) MPF_CK(cudaMalloc(&d_panel16, (size_t)N * r * sizeof(fp16)));
                const long long total = (long long)pr * pc;
                gather_cast_kernel<<<(int)std::min<long long>((total + 255) / 256, 4096), 256>>>(d_A, N, k, pr, pc, d_panel16);
                int blocks = std::min((pr + threads - 1) / threads, max_h);
                void* args[] = {&d_panel16, &pr, &pr, &pc, &d_ipiv_panel};
                MPF_CK(cudaLaunchCooperativeKernel((void*)HGETF2_kernel, dim3(blocks), dim3(threads), args, 0, 0));
            }
            laswp_kernel<<<(N + 255) / 256, 256>>>(d_A, N, N, k, pc, d_ipiv_panel, d_ipiv);
            // ---- fp64 panel without pivoting
            if (fast) {
                dpanel_top_kernel<<<1, 256>>>(d_A, N, k, pc);
                if (pr > pc) dpanel_rows_kernel<<<(pr - pc + 127) / 128, 128>>>(d_A, N, k, pc, pr - pc);
            } else {
                int blocks = std::min((pr + threads - 1) / threads, max_d);
                double* panel = d_A + (size_t)k * N + k;
                int ld = N;
                void* args[] = {&pr, &pc, &panel, &ld};
                MPF_CK(cudaLaunchCooperativeKernel((void*)dgetf2_native_npv, dim3(blocks), dim3(threads), args, 0, 0));
            }
            const int nt = N - k - pc;
            if (nt > 0) {
                const size_t sh = pc <= 64 ? (size_t)pc * pc * sizeof(double) : 0;
                auto update = [&](int row0, int col0, int kfirst, int kw, int mrows, int ncols) {
                    if (mrows <= 0 || ncols <= 0) return;
                    dim3 grid((mrows + DM_BM - 1) / DM_BM, (ncols + DM_BN - 1) / DM_BN);
                    dmma_rank_update_kernel<<<grid, DM_THREADS, DM_SMEM>>>(d_A, N, row0, col0, kfirst, kw, mrows, ncols);
                };
                if (sp <= r) {  // the reference's order: U12 and the whole trailing block after every panel
                    trsm_unit_lower_kernel<<<(nt + 127) / 128, 128, sh>>>(d_A, N, k, pc, k + pc, nt);
                    update(k + pc, k + pc, k, pc, nt, nt);
                } else {
                    // Super-panel [K0, K1) (classical delayed update): after a panel only the columns left of K1 are
                    // brought up to date (U rows + all rows below: what the next panels' pivot searches read).  The columns
                    // right of K1 are touched by the row interchanges only, until the super-panel's last panel: then, panel
                    // by panel, their U rows are solved and the rows above K1 updated (now that every interchange of the
                    // super-panel has been applied to them), and the block below / right of the super-panel receives ONE
                    // rank-(K1 - K0) update.  Same interchanges and panel arithmetic as the reference; the big block's 8
                    // rank-32 sums become one rank-256 sum: HBM traffic of the updates / 8, the kernel turns compute-bound.
                    const int K0 = (k / sp) * sp, K1 = std::min(K0 + sp, N), e = k + pc;
                    if (K1 > e) {
                        trsm_unit_lower_kernel<<<(K1 - e + 127) / 128, 128, sh>>>(d_A, N, k, pc, e, K1 - e);
                        update(e, e, k, pc, N - e, K1 - e);
                    }
                    if (e >= K1 && N > K1) {
                        for (int kk = K0; kk < K1; kk += r) {
                            const int pw = std::min(r, K1 - kk), ee = kk + pw;
                            const size_t shh = pw <= 64 ? (size_t)pw * pw * sizeof(double) : 0;
                            trsm_unit_lower_kernel<<<(N - K1 + 127) / 128, 128, shh>>>(d_A, N, kk, pw, K1, N - K1);
                            update(ee, K1, kk, pw, K1 - ee, N - K1);
                        }
                        update(K1, K1, K0, K1 - K0, N - K1, N - K1);
                    }
                }
            }
        }
    }
    MPF_CK(cudaGetLastError());
    MPF_CK(cudaStreamSynchronize(0));
done:
    cudaFree(d_panel16); cudaFree(d_mirror); cudaFree(d_pos2phys); cudaFree(d_keys); cudaFree(d_ipiv_panel);
    return e;
#undef MPF_CK
}
}  // namespace mplu_detail

namespace {
using mplu_detail::mpf_device;

int mpf_impl(double* h_A, int N, int r, int* IPIV) {
    if (!h_A || !IPIV || N <= 0 || r <= 0) return MPLU_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MPLU_E_NODEVICE;
    // unlike the reference (cudaSetDevice(0), MPF.cu:77) the caller's current device is kept
    const size_t nn = (size_t)N * (size_t)N;  // the reference computes N*N in int (overflow for N >= 46341)
    double* d_A = nullptr;
    int* d_ipiv = nullptr;
    cudaError_t e;
#define MPF_CK(x) do { e = (x); if (e != cudaSuccess) goto fail; } while (0)
    MPF_CK(cudaMalloc(&d_A, nn * sizeof(double)));
    MPF_CK(cudaMalloc(&d_ipiv, (size_t)N * sizeof(int)));
    {
        const bool timing = getenv("MPLU_MPF_TIMING") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        MPF_CK(staged_copy(d_A, h_A, nn * sizeof(double), true));
        MPF_CK(cudaMemcpy(d_ipiv, IPIV, (size_t)N * sizeof(int), cudaMemcpyHostToDevice));  // untouched entries survive
        const auto t1 = std::chrono::steady_clock::now();
        MPF_CK(mpf_device(d_A, N, r, d_ipiv));
        const auto t2 = std::chrono::steady_clock::now();
        MPF_CK(staged_copy(h_A, d_A, nn * sizeof(double), false));
        MPF_CK(cudaMemcpy(IPIV, d_ipiv, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost));
        const auto t3 = std::chrono::steady_clock::now();
        if (timing) {
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "MPF n=%d r=%d: h2d %.1f ms, device %.1f ms, d2h %.1f ms\n", N, r, ms(t0, t1), ms(t1, t2), ms(t2, t3));
        }
    }
    cudaFree(d_A); cudaFree(d_ipiv);
    return 0;
fail:
    cudaFree(d_A); cudaFree(d_ipiv);
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

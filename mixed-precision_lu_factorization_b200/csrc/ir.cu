// Iterative-refinement kernels (fp64 residual, fp32 triangular solves with the stored factors).  The reference has
// no solve / refinement path; these follow BASELINE.json's north_star and LAPACK dsgesv's
// structure: r = b - A*x in fp64, L U d = r in the factorization precision, x += d in fp64.
// All of them are HBM-bound: the residual streams 8*n^2 bytes, the solve pair 4*n^2 bytes.
#include "kernels.h"

namespace mplu {

namespace {

constexpr int RES_THREADS = 128;  // each thread owns 2 consecutive rows

// partial[chunk][row] = sum_{c in chunk} A(row,c) * x(c)
__global__ void __launch_bounds__(RES_THREADS)
residual_partial_kernel(const double* __restrict__ A, long long lda, int n, const double* __restrict__ x,
                        double* __restrict__ partial, int cols_per) {
    extern __shared__ double xs[];
    const int c0 = blockIdx.y * cols_per;
    const int c1 = min(n, c0 + cols_per);
    for (int c = c0 + threadIdx.x; c < c1; c += RES_THREADS) xs[c - c0] = x[c];
    __syncthreads();
    const int r0 = 2 * (blockIdx.x * RES_THREADS + threadIdx.x);
    if (r0 >= n) return;
    const int nc = c1 - c0;
    const double* Ap = A + r0 + (long long)c0 * lda;
    double a0 = 0.0, a1 = 0.0;
    const bool vec = (r0 + 1 < n) && ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (vec) {
        int c = 0;
        for (; c + 8 <= nc; c += 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const double2*>(Ap + (long long)(c + u) * lda));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double xv = xs[c + u];
                a0 = fma(v[u].x, xv, a0);
                a1 = fma(v[u].y, xv, a1);
            }
        }
        for (; c < nc; ++c) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(Ap + (long long)c * lda));
            a0 = fma(v.x, xs[c], a0);
            a1 = fma(v.y, xs[c], a1);
        }
    } else {
        for (int c = 0; c < nc; ++c) {
            a0 = fma(__ldg(Ap + (long long)c * lda), xs[c], a0);
            if (r0 + 1 < n) a1 = fma(__ldg(Ap + 1 + (long long)c * lda), xs[c], a1);
        }
    }
    partial[(long long)blockIdx.y * n + r0] = a0;
    if (r0 + 1 < n) partial[(long long)blockIdx.y * n + r0 + 1] = a1;
}

__device__ __forceinline__ void atomic_max_double_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// r = b - sum_chunks partial; norms[0] = max|r|, norms[1] = max|x|   (norms zeroed by the launcher)
__global__ void residual_finish_kernel(const double* __restrict__ partial, int n, int nchunk,
                                       const double* __restrict__ b, const double* __restrict__ x,
                                       double* __restrict__ r, double* norms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double rv = 0.0, xv = 0.0;
    if (i < n) {
        double s = 0.0;
        for (int c = 0; c < nchunk; ++c) s += partial[(long long)c * n + i];
        rv = b[i] - s;
        r[i] = rv;
        rv = fabs(rv);
        xv = fabs(x[i]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        rv = fmax(rv, __shfl_xor_sync(0xffffffffu, rv, o));
        xv = fmax(xv, __shfl_xor_sync(0xffffffffu, xv, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomic_max_double_nonneg(&norms[0], rv);
        atomic_max_double_nonneg(&norms[1], xv);
    }
}

__global__ void to_float_kernel(const double* __restrict__ src, float* __restrict__ dst, int n, int npad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) dst[i] = (i < n) ? static_cast<float>(src[i]) : 0.f;
}

__global__ void finish_solve_kernel(const float* __restrict__ sol, int n, double* __restrict__ d_out,
                                    double* __restrict__ x_accum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double d = static_cast<double>(sol[i]);
        if (d_out) d_out[i] = d;
        if (x_accum) x_accum[i] += d;
    }
}

// One block-column step of a blocked triangular solve (column sweep).
//   lower:  s = inv(L_jj) * y[j-block];  sol[j-block] = s;  y[r] -= L(r, j-block) * s   for rows r below the block
//   upper:  s = inv(U_jj) * y[j-block];  sol[j-block] = s;  y[r] -= U(r, j-block) * s   for rows r above the block
// Every CTA recomputes s (64 KiB of the fp32 inverse, L2 resident) so one launch does the whole step.
constexpr int TS_THREADS = 256;
constexpr int DBS = kDiagBlock;

__global__ void __launch_bounds__(TS_THREADS)
trsv_step_kernel(const float* __restrict__ W, long long ldw, const float* __restrict__ inv, int j0, int row_begin,
                 int row_end, float* __restrict__ y, float* __restrict__ sol) {
    __shared__ float ys[DBS];
    __shared__ float part[2][DBS];
    __shared__ float s[DBS];
    const int tid = threadIdx.x;
    if (tid < DBS) ys[tid] = y[j0 + tid];
    __syncthreads();
    {
        const int r = tid & (DBS - 1), half = tid >> 7;  // 2 threads per row of the inverse
        const float* ip = inv + r + (long long)(half * 64) * DBS;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < 64; ++c) acc = fmaf(__ldg(ip + (long long)c * DBS), ys[half * 64 + c], acc);
        part[half][r] = acc;
    }
    __syncthreads();
    if (tid < DBS) {
        const float v = part[0][tid] + part[1][tid];
        s[tid] = v;
        if (blockIdx.x == 0) sol[j0 + tid] = v;
    }
    __syncthreads();
    const int r = row_begin + blockIdx.x * TS_THREADS + tid;
    if (r < row_end) {
        const float* wp = W + r + (long long)j0 * ldw;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 8
        for (int c = 0; c < DBS; c += 2) {
            acc0 = fmaf(__ldg(wp + (long long)c * ldw), s[c], acc0);
            acc1 = fmaf(__ldg(wp + (long long)(c + 1) * ldw), s[c + 1], acc1);
        }
        y[r] -= (acc0 + acc1);
    }
}

}  // namespace

int launch_residual(const double* A, long long lda, int n, const double* x, const double* b, double* r,
                    double* partial, int nchunk, double* norms, cudaStream_t st) {
    const int cols_per = (n + nchunk - 1) / nchunk;
    dim3 grid((n + 2 * RES_THREADS - 1) / (2 * RES_THREADS), nchunk);
    residual_partial_kernel<<<grid, RES_THREADS, cols_per * sizeof(double), st>>>(A, lda, n, x, partial, cols_per);
    cudaMemsetAsync(norms, 0, 2 * sizeof(double), st);
    residual_finish_kernel<<<(n + 255) / 256, 256, 0, st>>>(partial, n, nchunk, b, x, r, norms);
    return (int)cudaGetLastError();
}

int launch_lu_solve(const float* W, long long ldw, int n, int npad, const float* Linv32, const float* Uinv32,
                    const double* rhs, float* y, double* d_out, double* x_accum, cudaStream_t st) {
    float* sol = y + npad;  // caller provides 2*npad floats
    to_float_kernel<<<(npad + 255) / 256, 256, 0, st>>>(rhs, y, n, npad);
    const int nblk = npad / DBS;
    for (int j = 0; j < nblk; ++j) {  // L y' = y
        const int j0 = j * DBS;
        const int rb = j0 + DBS, re = npad;
        const int g = (re - rb + TS_THREADS - 1) / TS_THREADS;
        trsv_step_kernel<<<g > 0 ? g : 1, TS_THREADS, 0, st>>>(W, ldw, Linv32 + (long long)j * DBS * DBS, j0, rb, re, y,
                                                               sol);
    }
    // sol now holds y' ; move it back into y for the backward sweep
    cudaMemcpyAsync(y, sol, npad * sizeof(float), cudaMemcpyDeviceToDevice, st);
    for (int j = nblk - 1; j >= 0; --j) {  // U x = y'
        const int j0 = j * DBS;
        const int rb = 0, re = j0;
        const int g = (re - rb + TS_THREADS - 1) / TS_THREADS;
        trsv_step_kernel<<<g > 0 ? g : 1, TS_THREADS, 0, st>>>(W, ldw, Uinv32 + (long long)j * DBS * DBS, j0, rb, re, y,
                                                               sol);
    }
    finish_solve_kernel<<<(n + 255) / 256, 256, 0, st>>>(sol, n, d_out, x_accum);
    return (int)cudaGetLastError();
}

}  // namespace mplu

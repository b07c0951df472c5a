// Host orchestration of the blocked right-looking no-pivot LU + fp64 iterative refinement, and the C ABI of
// include/mplu.h.  This is the B200-native counterpart of the panel loop in /root/reference/MPF.cu:100-241:
//   reference per panel (r = 32):  gather -> fp16 pivot search -> LASWP -> fp64 panel LU -> Dtrsm -> rank-32 Dgemm
//   here per outer block (nb = 2048): the tall panel is factored RECURSIVELY (halving down to 128-wide leaves:
//   diag_lu + L21 = A21*inv(U11) on the tensor cores), the block row by a recursive TRSM whose leaves multiply with
//   the explicit 128x128 inverses, then one rank-nb tcgen05 trailing update; operands are 16-bit shadows,
//   accumulation and the working matrix are fp32.
// Look-ahead: the trailing update of step k is split into the next panel's columns (done first) and the rest; the
// next panel is factored on a second stream restricted to `side_sms` SMs while the rest of the update runs on the
// others.  The whole schedule is captured once into a CUDA graph per (n, options) and replayed.
// No host<->device round trips inside the loop (the reference does one per panel, MPF.cu:146,158).
#include "../../include/mplu.h"
#include "gemm_tc.h"
#include "kernels.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace mplu;

struct mplu_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int n = 0, npad = 0, cap_npad = 0;
    bool factored = false;
    mplu_options opts{};
    // working storage
    float* W = nullptr;       // npad x npad fp32, column-major (ld = npad): becomes L\U
    uint16_t* Wh = nullptr;   // npad x npad 16-bit shadow (scaled), same indexing
    uint16_t* Linv16 = nullptr;  // 128 x npad : block j at columns [128j, 128j+128)
    uint16_t* Uinv16 = nullptr;
    float* Linv32 = nullptr;
    float* Uinv32 = nullptr;
    float* inv_scales = nullptr;  // 4 per diagonal block
    float* scales = nullptr;      // SC_COUNT
    float* amax = nullptr;
    double* anorm = nullptr;      // [0] ||A||inf, [1] ||b||inf
    double* rowsum_part = nullptr;
    int* status = nullptr;
    unsigned* ready = nullptr;    // step counter of the cooperative triangular-solve kernel
    // refinement
    double* r = nullptr;
    double* partial = nullptr;
    double* norms = nullptr;  // [0] ||r||, [1] ||x||
    float* y = nullptr;       // 2*npad
    int nchunk = 64;
    // staging for the host variant
    double* dA_stage = nullptr; size_t dA_cap = 0;
    double* db_stage = nullptr; double* dx_stage = nullptr; size_t dv_cap = 0;
    // tensor maps
    CUtensorMap tmWh_A, tmWh_B1, tmWh_B2, tmLinv_A, tmUinv_B1, tmUinv_B2;
    int gemm_launches = 0, kernel_launches = 0;
    // look-ahead / graph
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    struct GraphKey { int n, npad, nb, precision, gemm_variant, max_sms, lookahead, side_sms, a_exp, l_exp; const void* W; } gkey{};
    int g_gemm_launches = 0, g_kernel_launches = 0, g_trail_count = 0;
    double g_trail_flops = 0, g_trail_bytes = 0;
    bool capturing = false;
    int num_sms = 0;
    // per-launch timing of the trailing updates (events are cheap: <= npad/nb pairs per factorization)
    static constexpr int kMaxTrail = 1024;
    cudaEvent_t trail_ev[2 * kMaxTrail] = {};
    int trail_count = 0;
    double trail_flops = 0, trail_bytes = 0;
};

namespace {

#define CK(expr)                                  \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return (int)_e;    \
    } while (0)
#define CKI(expr)                  \
    do {                           \
        int _e = (expr);           \
        if (_e != 0) return _e;    \
    } while (0)

void free_work(mplu_context* c) {
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    cudaFree(c->W); cudaFree(c->Wh); cudaFree(c->Linv16); cudaFree(c->Uinv16); cudaFree(c->Linv32); cudaFree(c->Uinv32);
    cudaFree(c->inv_scales); cudaFree(c->rowsum_part); cudaFree(c->r); cudaFree(c->partial); cudaFree(c->y);
    c->W = nullptr; c->Wh = nullptr; c->Linv16 = c->Uinv16 = nullptr; c->Linv32 = c->Uinv32 = nullptr;
    c->inv_scales = nullptr; c->rowsum_part = nullptr; c->r = c->partial = nullptr; c->y = nullptr;
    c->cap_npad = 0;
}

int ensure_work(mplu_context* c, int n) {
    const int npad = ((n + kDiagBlock - 1) / kDiagBlock) * kDiagBlock;
    c->n = n;
    c->npad = npad;
    if (npad > c->cap_npad) {
        free_work(c);
        const size_t np = (size_t)npad;
        CK(cudaMalloc(&c->W, np * np * sizeof(float)));
        CK(cudaMalloc(&c->Wh, np * np * sizeof(uint16_t)));
        CK(cudaMalloc(&c->Linv16, np * kDiagBlock * sizeof(uint16_t)));
        CK(cudaMalloc(&c->Uinv16, np * kDiagBlock * sizeof(uint16_t)));
        CK(cudaMalloc(&c->Linv32, np * kDiagBlock * sizeof(float)));
        CK(cudaMalloc(&c->Uinv32, np * kDiagBlock * sizeof(float)));
        CK(cudaMalloc(&c->inv_scales, 4 * (np / kDiagBlock) * sizeof(float)));
        CK(cudaMalloc(&c->rowsum_part, (size_t)c->nchunk * np * sizeof(double)));
        CK(cudaMalloc(&c->r, np * sizeof(double)));
        CK(cudaMalloc(&c->partial, (size_t)c->nchunk * np * sizeof(double)));
        CK(cudaMalloc(&c->y, 2 * np * sizeof(float)));
        c->cap_npad = npad;
    }
    // tensor maps over the parents (dims = npad so that out-of-range boxes are zero filled)
    const uint64_t np = (uint64_t)npad;
    if (make_tmap_16bit(&c->tmWh_A, c->Wh, np, np, np, 64, 64)) return MPLU_E_TMAP;
    if (make_tmap_16bit(&c->tmWh_B1, c->Wh, np, np, np, 64, 256)) return MPLU_E_TMAP;
    if (make_tmap_16bit(&c->tmWh_B2, c->Wh, np, np, np, 64, 128)) return MPLU_E_TMAP;
    if (make_tmap_16bit(&c->tmLinv_A, c->Linv16, kDiagBlock, np, kDiagBlock, 64, 64)) return MPLU_E_TMAP;
    if (make_tmap_16bit(&c->tmUinv_B1, c->Uinv16, kDiagBlock, np, kDiagBlock, 64, 256)) return MPLU_E_TMAP;
    if (make_tmap_16bit(&c->tmUinv_B2, c->Uinv16, kDiagBlock, np, kDiagBlock, 64, 128)) return MPLU_E_TMAP;
    return 0;
}

void resolve_options(mplu_context* c, int n) {
    if (c->opts.nb <= 0) c->opts.nb = n >= 12288 ? 2048 : (n >= 4096 ? 1024 : 512);
}

struct GemmCall {
    // A operand: 0 = Wh block, 1 = Linv16 block ; B operand: 0 = Wh block, 1 = Uinv16 block
    int a_kind, a_r0, a_c0;
    int b_kind, b_r0, b_c0;
    int M, N, K;
    int out_r0, out_c0;  // block origin in W / Wh
    bool accumulate;     // out = W + alpha*acc (else alpha*acc)
    const float* alpha_p1;
    const float* alpha_p2;
    const float* hscale_p;
    int h_rows, h_cols;
};

// Where a piece of the schedule runs: stream + SM budget (0 = all SMs).
struct Lane {
    cudaStream_t st;
    int sms;
};

int run_gemm(mplu_context* c, const Lane& ln, const GemmCall& g) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    int variant;
    if (c->opts.gemm_variant == MPLU_GEMM_CG1) variant = GEMM_CG1_AMN;
    else variant = (g.M > 128) ? GEMM_CG2_AMN : GEMM_CG1_AMN;
    const bool cg2 = (variant == GEMM_CG2_AMN);
    const CUtensorMap* tmA = g.a_kind == 0 ? &c->tmWh_A : &c->tmLinv_A;
    const CUtensorMap* tmB = g.b_kind == 0 ? (cg2 ? &c->tmWh_B2 : &c->tmWh_B1) : (cg2 ? &c->tmUinv_B2 : &c->tmUinv_B1);
    GemmParams p{};
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.a_r0 = g.a_r0; p.a_c0 = g.a_c0; p.b_r0 = g.b_r0; p.b_c0 = g.b_c0;
    const long long ld = c->npad;
    p.C = c->W + g.out_r0 + (long long)g.out_c0 * ld;
    p.ldc = ld;
    p.Cin = g.accumulate ? p.C : nullptr;
    p.ldcin = ld;
    p.H = c->Wh + g.out_r0 + (long long)g.out_c0 * ld;
    p.ldh = ld;
    p.h_rows = g.h_rows; p.h_cols = g.h_cols;
    p.alpha = 1.f; p.alpha_p1 = g.alpha_p1; p.alpha_p2 = g.alpha_p2;
    p.hscale = 1.f; p.hscale_p = g.hscale_p;
    p.bf16 = c->opts.precision == MPLU_BF16;
    p.status = c->status;
    c->gemm_launches++;
    c->kernel_launches++;
    int sms = ln.sms > 0 ? ln.sms : c->num_sms;
    if (c->opts.max_sms > 0 && c->opts.max_sms < sms) sms = c->opts.max_sms;
    return launch_gemm_tc(variant, tmA, tmB, p, sms, ln.st);
}

inline int split_width(int w) { return kDiagBlock * ((w / kDiagBlock + 1) / 2); }

// A[r0:r1, c0:c1) -= L[r0:r1, k0:k1) * U[k0:k1, c0:c1)   (Schur update; the result is A-type, shadow where asked)
int schur_update(mplu_context* c, const Lane& ln, int r0, int r1, int c0, int c1, int k0, int k1, int h_rows,
                 int h_cols) {
    GemmCall g{0, r0, k0, 0, k0, c0, r1 - r0, c1 - c0, k1 - k0, r0, c0, true,
               c->scales + SC_NEG_LA_INV, nullptr, c->scales + SC_A, h_rows, h_cols};
    return run_gemm(c, ln, g);
}

// U[c0:c0+w, n0:n1) = inv(L11[c0:c0+w)) * A[c0:c0+w, n0:n1), in place; recursive, 128-wide leaves multiply with the
// explicit inverse of the diagonal block (replaces cublasDtrsm, MPF.cu:215-225).
int trsm_rec(mplu_context* c, const Lane& ln, int c0, int w, int n0, int n1) {
    if (n1 <= n0) return 0;
    if (w <= kDiagBlock) {
        const int blk = c0 / kDiagBlock;
        GemmCall t{1, 0, blk * kDiagBlock, 0, c0, n0, kDiagBlock, n1 - n0, kDiagBlock, c0, n0, false,
                   c->inv_scales + 4 * blk + 1, c->scales + SC_A_INV, c->scales + SC_A, kDiagBlock, n1 - n0};
        return run_gemm(c, ln, t);
    }
    const int h = split_width(w);
    CKI(trsm_rec(c, ln, c0, h, n0, n1));
    CKI(schur_update(c, ln, c0 + h, c0 + w, n0, n1, c0, c0 + h, w - h, n1 - n0));
    return trsm_rec(c, ln, c0 + h, w - h, n0, n1);
}

// LU of the tall panel: columns [c0, c0+w), rows [c0, npad); recursive halving down to one diagonal block.
int panel_rec(mplu_context* c, const Lane& ln, int c0, int w) {
    const int npad = c->npad;
    if (w <= kDiagBlock) {
        const int blk = c0 / kDiagBlock;
        CKI(launch_diag_lu(c->W, npad, c0, c->Linv16, c->Uinv16, c->Linv32, c->Uinv32, c->inv_scales, blk,
                           c->opts.precision == MPLU_BF16, c->status, ln.st));
        c->kernel_launches++;
        const int below = c0 + kDiagBlock;
        if (npad > below) {  // L21 = A21 * inv(U11), in place
            GemmCall g{0, below, c0, 1, 0, blk * kDiagBlock, npad - below, kDiagBlock, kDiagBlock, below, c0, false,
                       c->scales + SC_A_INV, c->inv_scales + 4 * blk + 3, c->scales + SC_L, npad - below, kDiagBlock};
            CKI(run_gemm(c, ln, g));
        }
        return 0;
    }
    const int h = split_width(w);
    CKI(panel_rec(c, ln, c0, h));
    CKI(trsm_rec(c, ln, c0, h, c0 + h, c0 + w));
    CKI(schur_update(c, ln, c0 + h, npad, c0 + h, c0 + w, c0, c0 + h, npad - c0 - h, w - h));
    return panel_rec(c, ln, c0 + h, w - h);
}

int record_event(mplu_context* c, cudaEvent_t ev, cudaStream_t st) {
    return (int)cudaEventRecordWithFlags(ev, st, c->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

// Everything after the first touch: scales, shadows of the first block column / block row, the panel loop.
int enqueue_factorization(mplu_context* c) {
    const int npad = c->npad;
    const long long ld = npad;
    const int bf16 = c->opts.precision == MPLU_BF16;
    int NB = c->opts.nb;
    if (NB < kDiagBlock) NB = kDiagBlock;
    NB = (NB / kDiagBlock) * kDiagBlock;
    if (NB > npad) NB = npad;
    cudaStream_t st = c->stream;
    int side_sms = c->opts.side_sms > 0 ? c->opts.side_sms : 24;
    side_sms -= side_sms % 2;
    const bool lookahead = c->opts.lookahead != 0 && side_sms >= 2 && side_sms <= c->num_sms - 16;
    const Lane all{st, 0};
    const Lane main_part{st, c->num_sms - side_sms};
    const Lane side{c->side, side_sms};

    CKI(launch_scales(c->amax, c->scales, c->opts.a_exp, c->opts.l_exp, bf16, st));
    CKI(launch_shadow_cast(c->W, ld, c->Wh, ld, npad, NB, c->scales + SC_A, bf16, c->status, st));
    if (npad > NB)
        CKI(launch_shadow_cast(c->W + (long long)NB * ld, ld, c->Wh + (long long)NB * ld, ld, NB, npad - NB,
                               c->scales + SC_A, bf16, c->status, st));
    c->kernel_launches += 3;

    CKI(panel_rec(c, all, 0, NB));
    for (int k = 0; k < npad; k += NB) {
        const int nbk = (NB < npad - k) ? NB : (npad - k);
        const int kend = k + nbk;
        if (kend >= npad) break;
        const int nbn = (NB < npad - kend) ? NB : (npad - kend);
        const int cnext = kend + nbn;  // end of the next panel's columns
        // next panel's columns first: block row of U, then their Schur update (full shadow: the panel consumes it)
        CKI(trsm_rec(c, all, k, nbk, kend, cnext));
        CKI(schur_update(c, all, kend, npad, kend, cnext, k, kend, npad - kend, nbn));
        const bool rest = cnext < npad;
        const bool fork = lookahead && rest;
        if (fork) {
            CK(cudaEventRecord(c->ev_fork, st));
            CK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
            CKI(panel_rec(c, side, kend, nbn));
            CK(cudaEventRecord(c->ev_join, c->side));
        }
        if (rest) {
            const Lane& ln = fork ? main_part : all;
            CKI(trsm_rec(c, ln, k, nbk, cnext, npad));
            const int Mt = npad - kend, Nt = npad - cnext;
            const bool timed = c->trail_count < mplu_context::kMaxTrail;
            if (timed) {
                cudaEvent_t& e0 = c->trail_ev[2 * c->trail_count];
                if (!e0) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&c->trail_ev[2 * c->trail_count + 1])); }
                CKI(record_event(c, e0, st));
            }
            // the rest of the trailing matrix: shadow only for the next panel's block row (its TRSM input)
            CKI(schur_update(c, ln, kend, npad, cnext, npad, k, kend, nbn, 0));
            if (timed) {
                CKI(record_event(c, c->trail_ev[2 * c->trail_count + 1], st));
                c->trail_count++;
                c->trail_flops += 2.0 * Mt * (double)Nt * nbk;
                c->trail_bytes += 8.0 * Mt * (double)Nt;
            }
        }
        if (fork) CK(cudaStreamWaitEvent(st, c->ev_join, 0));
        else CKI(panel_rec(c, all, kend, nbn));
    }
    return 0;
}

int factor_impl(mplu_context* c, int n, const double* dA, long long lda) {
    CKI(ensure_work(c, n));
    const int npad = c->npad;
    cudaStream_t st = c->stream;
    CK(cudaMemsetAsync(c->status, 0, sizeof(int), st));
    CKI(launch_first_touch(dA, lda, n, c->W, npad, npad, c->amax, c->rowsum_part, c->nchunk, c->anorm, st));

    const bool use_graph = c->opts.use_graph != 0;
    mplu_context::GraphKey key{n, npad, c->opts.nb, c->opts.precision, c->opts.gemm_variant, c->opts.max_sms,
                               c->opts.lookahead, c->opts.side_sms, c->opts.a_exp, c->opts.l_exp, c->W};
    const bool hit = use_graph && c->graph_exec && memcmp(&key, &c->gkey, sizeof(key)) == 0;
    if (!hit) {
        c->gemm_launches = 0;
        c->kernel_launches = 2;  // first touch + anorm
        c->trail_count = 0;
        c->trail_flops = c->trail_bytes = 0;
        if (use_graph) {
            if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            c->capturing = true;
            int rc = enqueue_factorization(c);
            c->capturing = false;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess) return (int)e;
            e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { c->graph_exec = nullptr; return (int)e; }
            memset(&c->gkey, 0, sizeof(c->gkey));
            c->gkey = key;
            c->g_gemm_launches = c->gemm_launches; c->g_kernel_launches = c->kernel_launches;
            c->g_trail_count = c->trail_count; c->g_trail_flops = c->trail_flops; c->g_trail_bytes = c->trail_bytes;
        } else {
            CKI(enqueue_factorization(c));
        }
    }
    if (use_graph) {
        c->gemm_launches = c->g_gemm_launches; c->kernel_launches = c->g_kernel_launches;
        c->trail_count = c->g_trail_count; c->trail_flops = c->g_trail_flops; c->trail_bytes = c->g_trail_bytes;
        CK(cudaGraphLaunch(c->graph_exec, st));
    }
    c->factored = true;
    return 0;
}

__global__ void absmax_kernel(const double* v, int n, double* out) {
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmax(m, fabs(v[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0)
        atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
}

__global__ void widen_kernel(const float* __restrict__ W, long long ldw, int n, double* __restrict__ out,
                             long long ldo) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (r < n) out[r + (long long)c * ldo] = (double)W[r + (long long)c * ldw];
}

int solve_impl(mplu_context* c, const double* dA, long long lda, const double* db, double* dx, mplu_stats* stats) {
    if (!c->factored) return MPLU_E_NOTFACTORED;
    const int n = c->n, npad = c->npad;
    const long long ld = npad;
    cudaStream_t st = c->stream;
    const int solve_launches = 1;

    CK(cudaMemsetAsync(c->anorm + 1, 0, sizeof(double), st));
    absmax_kernel<<<64, 256, 0, st>>>(db, n, c->anorm + 1);
    // first solve: x = (LU)^-1 b
    CKI(launch_lu_solve(c->W, ld, n, npad, c->Linv32, c->Uinv32, db, c->y, dx, nullptr, c->ready, st));
    c->kernel_launches += solve_launches + 1;

    double h_norms[2] = {0, 0}, h_an[2] = {0, 0};
    const double eps = 2.220446049250313e-16 / 2.0;  // LAPACK dlamch('E')
    int iters = 0, converged = 0;
    double first_be = -1.0;
    const int max_iters = c->opts.max_iters > 0 ? c->opts.max_iters : 30;
    for (;;) {
        CKI(launch_residual(dA, lda, n, dx, db, c->r, c->partial, c->nchunk, c->norms, st));
        c->kernel_launches += 2;
        CK(cudaMemcpyAsync(h_norms, c->norms, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (first_be < 0) CK(cudaMemcpyAsync(h_an, c->anorm, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const double be = h_norms[0] / (h_an[0] * h_norms[1] + h_an[1]);
        if (first_be < 0) first_be = be;
        double thresh;
        if (c->opts.tol > 0) thresh = c->opts.tol * h_an[0] * h_norms[1];
        else thresh = h_norms[1] * h_an[0] * eps * std::sqrt((double)n);
        if (!(h_norms[0] == h_norms[0])) break;  // NaN: give up
        if (h_norms[0] <= thresh) { converged = 1; break; }
        if (iters >= max_iters) break;
        CKI(launch_lu_solve(c->W, ld, n, npad, c->Linv32, c->Uinv32, c->r, c->y, nullptr, dx, c->ready, st));
        c->kernel_launches += solve_launches;
        ++iters;
    }
    int h_status = 0;
    CK(cudaMemcpy(&h_status, c->status, sizeof(int), cudaMemcpyDeviceToHost));
    if (stats) {
        stats->n = n;
        stats->iters = iters;
        stats->converged = converged;
        stats->status_bits = h_status;
        stats->anorm_inf = h_an[0];
        stats->bnorm_inf = h_an[1];
        stats->xnorm_inf = h_norms[1];
        stats->rnorm_inf = h_norms[0];
        stats->backward_error = h_norms[0] / (h_an[0] * h_norms[1] + h_an[1]);
        stats->first_backward_error = first_be;
        stats->gemm_launches = c->gemm_launches;
        stats->kernel_launches = c->kernel_launches;
        stats->trailing_launches = c->trail_count;
        stats->trailing_flops = c->trail_flops;
        stats->trailing_bytes = c->trail_bytes;
        float tms = 0.f;
        for (int i = 0; i < c->trail_count; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, c->trail_ev[2 * i], c->trail_ev[2 * i + 1]) == cudaSuccess) tms += ms;
        }
        stats->trailing_ms = tms;
    }
    if (converged) return 0;
    if (h_status & 1) return MPLU_E_OVERFLOW;
    if (h_status & 2) return MPLU_E_ZEROPIVOT;
    return MPLU_E_NOCONV;
}

}  // namespace

extern "C" {

void mplu_default_options(mplu_options* o) {
    if (!o) return;
    o->precision = MPLU_FP16;
    o->nb = 0;  // auto: 2048 for n >= 12288, 1024 for n >= 4096, else 512
    o->max_iters = 30;
    o->tol = 0.0;
    o->gemm_variant = MPLU_GEMM_AUTO;
    o->max_sms = 0;
    o->a_exp = 11;
    o->l_exp = 11;
    o->lookahead = 1;
    o->side_sms = 24;
    o->use_graph = 1;
}

int mplu_create(mplu_context** out, int device) {
    if (!out) return MPLU_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MPLU_E_NODEVICE;
    if (device < 0 || device >= ndev) return MPLU_E_ARG;
    CK(cudaSetDevice(device));
    mplu_context* c = new (std::nothrow) mplu_context();
    if (!c) return MPLU_E_ARG;
    c->device = device;
    mplu_default_options(&c->opts);
    CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
    if (gemm_tc_init() != 0) return MPLU_E_TMAP;
    CKI(panel_init());
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (auto& e : c->ev) CK(cudaEventCreate(&e));
    CK(cudaMalloc(&c->scales, SC_COUNT * sizeof(float)));
    CK(cudaMalloc(&c->amax, sizeof(float)));
    CK(cudaMalloc(&c->anorm, 2 * sizeof(double)));
    CK(cudaMalloc(&c->norms, 2 * sizeof(double)));
    CK(cudaMalloc(&c->status, sizeof(int)));
    CK(cudaMalloc(&c->ready, sizeof(unsigned)));
    *out = c;
    return 0;
}

void mplu_destroy(mplu_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_work(c);
    cudaFree(c->scales); cudaFree(c->amax); cudaFree(c->anorm); cudaFree(c->norms); cudaFree(c->status); cudaFree(c->ready);
    cudaFree(c->dA_stage); cudaFree(c->db_stage); cudaFree(c->dx_stage);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& e : c->trail_ev) if (e) cudaEventDestroy(e);
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

void* mplu_stream(mplu_context* c) { return c ? (void*)c->stream : nullptr; }

int mplu_factor_device(mplu_context* c, int n, const double* dA, long long lda, const mplu_options* opts) {
    if (!c || !dA || n <= 0 || lda < n) return MPLU_E_ARG;
    CK(cudaSetDevice(c->device));
    if (opts) c->opts = *opts;
    resolve_options(c, n);
    c->factored = false;
    return factor_impl(c, n, dA, lda);
}

int mplu_solve_device(mplu_context* c, const double* dA, long long lda, const double* db, double* dx,
                      mplu_stats* stats) {
    if (!c || !dA || !db || !dx) return MPLU_E_ARG;
    CK(cudaSetDevice(c->device));
    if (stats) memset(stats, 0, sizeof(*stats));
    return solve_impl(c, dA, lda, db, dx, stats);
}

int mplu_gesv_device(mplu_context* c, int n, const double* dA, long long lda, const double* db, double* dx,
                     const mplu_options* opts, mplu_stats* stats) {
    if (!c || !dA || !db || !dx || n <= 0 || lda < n) return MPLU_E_ARG;
    CK(cudaSetDevice(c->device));
    if (opts) c->opts = *opts;
    resolve_options(c, n);
    if (stats) memset(stats, 0, sizeof(*stats));
    c->factored = false;
    CK(cudaEventRecord(c->ev[0], c->stream));
    int rc = factor_impl(c, n, dA, lda);
    if (rc) return rc;
    CK(cudaEventRecord(c->ev[1], c->stream));
    rc = solve_impl(c, dA, lda, db, dx, stats);
    CK(cudaEventRecord(c->ev[2], c->stream));
    CK(cudaEventSynchronize(c->ev[2]));
    if (stats) {
        cudaEventElapsedTime(&stats->factor_ms, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&stats->solve_ms, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&stats->total_ms, c->ev[0], c->ev[2]);
    }
    return rc;
}

int mplu_gesv_host(mplu_context* c, int n, const double* hA, long long lda, const double* hb, double* hx,
                   const mplu_options* opts, mplu_stats* stats) {
    if (!c || !hA || !hb || !hx || n <= 0 || lda < n) return MPLU_E_ARG;
    CK(cudaSetDevice(c->device));
    const size_t need = (size_t)n * n;
    if (need > c->dA_cap) {
        cudaFree(c->dA_stage);
        c->dA_stage = nullptr; c->dA_cap = 0;
        CK(cudaMalloc(&c->dA_stage, need * sizeof(double)));
        c->dA_cap = need;
    }
    if ((size_t)n > c->dv_cap) {
        cudaFree(c->db_stage); cudaFree(c->dx_stage);
        c->db_stage = c->dx_stage = nullptr; c->dv_cap = 0;
        CK(cudaMalloc(&c->db_stage, n * sizeof(double)));
        CK(cudaMalloc(&c->dx_stage, n * sizeof(double)));
        c->dv_cap = n;
    }
    cudaStream_t st = c->stream;
    CK(cudaEventRecord(c->ev[3], st));
    CK(cudaMemcpy2DAsync(c->dA_stage, (size_t)n * sizeof(double), hA, (size_t)lda * sizeof(double),
                         (size_t)n * sizeof(double), n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->db_stage, hb, n * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = mplu_gesv_device(c, n, c->dA_stage, n, c->db_stage, c->dx_stage, opts, stats);
    float h2d = 0.f;
    cudaEventElapsedTime(&h2d, c->ev[3], c->ev[0]);
    if (rc != 0 && rc != MPLU_E_NOCONV) return rc;
    CK(cudaEventRecord(c->ev[0], st));
    CK(cudaMemcpyAsync(hx, c->dx_stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->ev[1], st));
    CK(cudaEventSynchronize(c->ev[1]));
    if (stats) {
        float d2h = 0.f;
        cudaEventElapsedTime(&d2h, c->ev[0], c->ev[1]);
        stats->h2d_ms = h2d;
        stats->d2h_ms = d2h;
        stats->total_ms += h2d + d2h;
    }
    return rc;
}

int mplu_get_factors(mplu_context* c, double* LU, long long ldlu, int on_device) {
    if (!c || !LU || !c->factored || ldlu < c->n) return MPLU_E_ARG;
    CK(cudaSetDevice(c->device));
    const int n = c->n;
    double* dst = LU;
    double* tmp = nullptr;
    if (!on_device) {
        CK(cudaMalloc(&tmp, (size_t)n * n * sizeof(double)));
        dst = tmp;
    }
    dim3 grid((n + 255) / 256, n);
    widen_kernel<<<grid, 256, 0, c->stream>>>(c->W, c->npad, n, dst, on_device ? ldlu : n);
    CK(cudaGetLastError());
    if (!on_device) {
        CK(cudaMemcpy2DAsync(LU, (size_t)ldlu * sizeof(double), tmp, (size_t)n * sizeof(double),
                             (size_t)n * sizeof(double), n, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    if (tmp) cudaFree(tmp);
    return 0;
}

// ------------------------------------------------------------------------------------------------ kernel hooks
int mplu_gemm16(int variant, int bf16, int M, int N, int K, float alpha, const void* dA, long long lda,
                const void* dB, long long ldb, float beta, float* dC, long long ldc, void* dH, long long ldh,
                float hscale, int max_sms, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || !dA || !dB) return MPLU_E_ARG;
    if (beta != 0.f && beta != 1.f) return MPLU_E_ARG;
    if (gemm_tc_init() != 0) return MPLU_E_TMAP;
    uint32_t abr, abc, bbr, bbc;
    gemm_box_shapes(variant, &abr, &abc, &bbr, &bbc);
    const bool amn = (variant == GEMM_CG1_AMN || variant == GEMM_CG2_AMN);
    CUtensorMap tA, tB;
    if (amn) { if (make_tmap_16bit(&tA, dA, M, K, lda, abr, abc)) return MPLU_E_TMAP; }
    else     { if (make_tmap_16bit(&tA, dA, K, M, lda, abr, abc)) return MPLU_E_TMAP; }
    if (make_tmap_16bit(&tB, dB, K, N, ldb, bbr, bbc)) return MPLU_E_TMAP;
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.C = dC; p.ldc = ldc;
    p.Cin = (beta == 1.f) ? dC : nullptr; p.ldcin = ldc;
    p.H = dH; p.ldh = ldh; p.h_rows = M; p.h_cols = N;
    p.alpha = alpha; p.hscale = hscale; p.bf16 = bf16;
    return launch_gemm_tc(variant, &tA, &tB, p, max_sms, (cudaStream_t)stream);
}

int mplu_diag_lu128(float* dW, long long ldw, float* dLinv, float* dUinv, void* stream) {
    if (!dW || !dLinv || !dUinv) return MPLU_E_ARG;
    CKI(panel_init());
    uint16_t* tmp16 = nullptr;
    float* sc = nullptr;
    CK(cudaMalloc(&tmp16, 2 * 128 * 128 * sizeof(uint16_t)));
    CK(cudaMalloc(&sc, 4 * sizeof(float)));
    int rc = launch_diag_lu(dW, ldw, 0, tmp16, tmp16 + 128 * 128, dLinv, dUinv, sc, 0, 0, nullptr, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(tmp16);
    cudaFree(sc);
    return rc;
}

// development aid: same as mplu_diag_lu128 with clock64() stamps of the kernel's phases written to d_clocks[0..31]
int mplu_diag_lu128_timed(float* dW, long long ldw, float* dLinv, float* dUinv, long long* d_clocks, void* stream) {
    if (!dW || !dLinv || !dUinv) return MPLU_E_ARG;
    CKI(panel_init());
    uint16_t* tmp16 = nullptr;
    float* sc = nullptr;
    CK(cudaMalloc(&tmp16, 2 * 128 * 128 * sizeof(uint16_t)));
    CK(cudaMalloc(&sc, 4 * sizeof(float)));
    int rc = launch_diag_lu(dW, ldw, 0, tmp16, tmp16 + 128 * 128, dLinv, dUinv, sc, 0, 0, nullptr, (cudaStream_t)stream,
                            d_clocks);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(tmp16);
    cudaFree(sc);
    return rc;
}

int mplu_residual(int n, const double* dA, long long lda, const double* dx, const double* db, double* dr,
                  double* dnorms, void* stream) {
    if (n <= 0 || !dA || !dx || !db || !dr || !dnorms) return MPLU_E_ARG;
    double* partial = nullptr;
    const int nchunk = 64;
    CK(cudaMalloc(&partial, (size_t)nchunk * n * sizeof(double)));
    int rc = launch_residual(dA, lda, n, dx, db, dr, partial, nchunk, dnorms, (cudaStream_t)stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(partial);
    return rc;
}

}  // extern "C"

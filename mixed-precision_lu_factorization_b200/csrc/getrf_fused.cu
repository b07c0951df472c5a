// Fused GETRF of one diagonal block (up to a whole nb x nb tile) in ONE persistent launch.
//
// The recursive, inverse-carrying GETRF of lu.cu (Sched::getrf) is a strictly sequential chain: per 2048-tile 16 leaf
// launches (diag_lu) and 45 small grouped GEMM launches, each of which costs a kernel boundary (launch gap, barrier /
// TMEM set-up, cold pipeline, drain) next to a few hundred nanoseconds of tensor-core work: 0.6 ms of the tile's
// 1.75 ms were launch overhead (profiles/r01j_leaf_timeline_n32768.txt).  Here the host records that chain once as a
// step program (FusedStep / FusedProblem) and a small persistent grid interprets it:
//   * GEMM step: the 128x128 output tiles of up to four independent products are dealt round-robin to the CTAs; each
//     CTA runs the same warp-specialised pipeline as gemm_tc.cu (one TMA producer thread, one tcgen05.mma issuer,
//     8 epilogue warps reading the fp32 accumulators out of TMEM, double-buffered) with operands fetched from the
//     L2-resident 16-bit arrays;
//   * LEAF step: the first cluster (CTAs 0, 1) runs the 128x128 leaf (leaf.cuh) out of shared memory;
//   * between steps: a grid barrier on a global counter (release / acquire at gpu scope) plus the cross-proxy fences
//     that order the generic-proxy stores of one step before the async-proxy (TMA) loads of the next.
// Replaces, for the diagonal blocks, the dgetf2_native_npv + cublasDtrsm + cublasDgemm chain of the reference's panel
// loop (/root/reference/MPF.cu:166-239, dgetf2_native_npv.cu:18-35).  Same products in the same order on the same
// 16-bit operands as the unfused path: the factors are bit-identical (tests/test_gpu_solver.py).
#include "getrf_fused.h"

#include "gemm_tc.h"
#include "leaf_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mplu {

namespace {

constexpr int FBM = 128, FBN = 128, FBK = 64, FUK = 16;
constexpr int FSTAGES = 3;
constexpr int F_A_BYTES = FBM * FBK * 2;  // 16 KiB
constexpr int F_B_BYTES = FBN * FBK * 2;  // 16 KiB
constexpr int F_RING_BYTES = FSTAGES * (F_A_BYTES + F_B_BYTES);
constexpr int F_STC_BYTES = FBM * FBN * 4, F_STH_BYTES = FBM * FBN * 2;  // result tile staged for the TMA stores
constexpr int F_GEMM_BYTES = F_RING_BYTES + F_STC_BYTES + F_STH_BYTES;
constexpr int F_MAIN_RAW = leaf::DL_SMEM_BYTES > F_GEMM_BYTES ? leaf::DL_SMEM_BYTES : F_GEMM_BYTES;  // leaf arrays alias ring + staging
constexpr int F_MAIN_BYTES = ((F_MAIN_RAW + 1023) / 1024) * 1024;
constexpr int F_PROG_CAP = 16 * 1024;     // the step program is staged in shared memory when it fits
constexpr int F_BAR_BYTES = 256;
constexpr int F_SMEM_BYTES = F_MAIN_BYTES + F_BAR_BYTES + F_PROG_CAP + 1024;  // + alignment slack
constexpr int F_THREADS = leaf::DL_THREADS;  // 16 warps: 0-7 epilogue, 8 TMA producer, 9 MMA issuer; all 16 in a leaf
constexpr int F_EPI_WARPS = 8;
constexpr int F_TMEM_COLS = 256;             // two 128-column fp32 accumulators
static_assert(F_MAIN_BYTES % 1024 == 0 && F_MAIN_BYTES >= leaf::DL_SMEM_BYTES && F_MAIN_BYTES >= F_GEMM_BYTES, "smem layout");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// All CTAs of the grid: every global store (and every shared-memory store of a leaf) made before the barrier is visible
// to every generic-proxy AND async-proxy (TMA) access made after it.
// The barrier in two halves: what a CTA writes between its arrive and its wait is covered by its NEXT arrive.
__device__ __forceinline__ void grid_step_arrive(unsigned* bar, long long* dbg_slot) {
    fence_proxy_async_all();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (dbg_slot) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); *dbg_slot = t_; }
        // release at gpu scope (cumulative over the CTA's writes ordered before it by the bar.sync above)
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    }
}
__device__ __forceinline__ void grid_step_wait(unsigned* bar, unsigned target) {
    if (threadIdx.x == 0) {
        while (ld_acquire_gpu(bar) < target) { }
        // The fence is what invalidates this SM's L1: the other threads' plain loads after the barrier (addends, leaf
        // input) must not hit lines cached before another CTA -- or this CTA's own TMA store, which bypasses L1 --
        // rewrote them.  (Without it a general matrix lost whole Schur updates: first backward error 4e-2 vs 2e-4.)
        __threadfence();
    }
    __syncthreads();
}

struct TileRef {
    const FusedProblem* p;
    int mt, nt, kb0, kb1;
};

__device__ __forceinline__ TileRef tile_ref(const FusedStep& st, const FusedProblem* probs, int t) {
    int i = 0;
    while (i + 1 < st.num_problems && t >= st.tile_end[i]) ++i;
    const FusedProblem* p = probs + st.first_problem + i;
    const int lt = t - (i ? st.tile_end[i - 1] : 0);
    const int nm = p->M / FBM;
    TileRef r;
    r.p = p;
    r.mt = lt % nm;
    r.nt = lt / nm;
    int k0 = 0, k1 = p->K;
    const int mlo = r.mt * FBM, mhi = mlo + FBM, nlo = r.nt * FBN, nhi = nlo + FBN;
    if (p->tri == TRI_A_LOWER) k1 = min(p->K, mhi);
    else if (p->tri == TRI_A_UPPER) k0 = min(mlo, p->K - FBK);
    else if (p->tri == TRI_B_UPPER) k1 = min(p->K, nhi);
    else if (p->tri == TRI_B_LOWER) k0 = min(nlo, p->K - FBK);
    r.kb0 = k0 / FBK;
    r.kb1 = (k1 + FBK - 1) / FBK;
    return r;
}

__global__ void __launch_bounds__(F_THREADS, 1)
getrf_fused_kernel(const __grid_constant__ FusedMaps maps, const FusedArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + FSTAGES * F_A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + F_MAIN_BYTES);
    uint64_t* empty = full + FSTAGES;
    uint64_t* tfull = empty + FSTAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* lbar = tempty + 2;  // the leaf's tensor-core products (leaf.cuh, kTc)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lbar + 1);
    uint8_t* sprog = smem + F_MAIN_BYTES + F_BAR_BYTES;

    const int tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    // GEMM tiles are dealt starting at CTA 2: CTAs 0 and 1 run the leaves (their instruction cache holds leaf code and
    // they are the last to reach the barrier after a leaf), the others wait at the barrier with the GEMM code hot
    const int first_cta = G > 2 ? 2 : 0;
    const int slot0 = (cta - first_cta + G) % G;
    const bool prof = a.dbg_clk != nullptr && cta == first_cta;  // the CTA that takes tile 0 of every GEMM step

    // ---- the step program: staged in shared memory when it fits (it is read at the head of every step, on the chain)
    const int prog_bytes = a.num_steps * (int)sizeof(FusedStep) + a.num_problems * (int)sizeof(FusedProblem);
    const bool prog_in_smem = prog_bytes <= F_PROG_CAP;
    if (prog_in_smem) {
        const uint4* src = reinterpret_cast<const uint4*>(a.program);
        uint4* dst = reinterpret_cast<uint4*>(sprog);
        for (int i = tid; i < prog_bytes / 16; i += F_THREADS) dst[i] = __ldg(src + i);
    }
    const uint8_t* prog = prog_in_smem ? sprog : reinterpret_cast<const uint8_t*>(a.program);
    const FusedStep* steps = reinterpret_cast<const FusedStep*>(prog);
    const FusedProblem* probs = reinterpret_cast<const FusedProblem*>(prog + (size_t)a.num_steps * sizeof(FusedStep));

    if (warp == F_EPI_WARPS && lane == 0) {
        for (int i = 0; i < FM_COUNT; ++i) {
            ptx::prefetch_tmap(&maps.a[i]);
            ptx::prefetch_tmap(&maps.b[i]);
            ptx::prefetch_tmap(&maps.h[i]);
        }
        ptx::prefetch_tmap(&maps.c);
        for (int i = 0; i < FSTAGES; ++i) {
            ptx::mbar_init(&full[i], 1);   // the producer's arrive.expect_tx
            ptx::mbar_init(&empty[i], 1);  // one tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);             // one tcgen05.commit
            ptx::mbar_init(&tempty[i], F_EPI_WARPS);  // one arrival per epilogue warp
        }
        ptx::mbar_init(lbar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == F_EPI_WARPS + 1) ptx::tmem_alloc<1>(tmem_slot, F_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    uint32_t stage = 0, phase = 0;  // smem ring position (producer and MMA issuer each advance their own copy)
    uint32_t acc_iter = 0;          // accumulator tiles this CTA has gone through (MMA issuer and epilogue warps)
    float mx = 0.f;                 // largest scaled fp16 magnitude written (overflow detection)
    const uint32_t idesc = make_idesc_f16(FBM, FBN, a.bf16 != 0, true, false);
    leaf::LeafTc ltc{tmem_base, lbar, 0u};

    // development aid (mplu_debug_fused_profile): CTA 0 stamps clock64 at the head of every step and at the end,
    // followed by %globaltimer (ns) at both ends of the launch for calibration
    auto stamp = [&](int slot, bool wall) {
        if (prof && tid == 0) {
            long long t_;
            if (wall) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_) :: "memory");
            else asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory");
            a.dbg_clk[slot] = t_;
        }
    };
    stamp(a.num_steps + 1, true);
#pragma unroll 1
    for (int s = 0; s < a.num_steps; ++s) {
        const FusedStep st = steps[s];
        stamp(s, false);
        if (st.kind == FS_LEAF) {
            if (cta < 2) {
                const long long off16 = (long long)(st.k0 - st.T) + (long long)st.k0 * a.ld16;
                leaf::diag_lu_body_tc(reinterpret_cast<float*>(smem), cta, a.W, a.ldw, st.k0,
                                   reinterpret_cast<uint16_t*>(a.Linv16) + off16, reinterpret_cast<uint16_t*>(a.Uinv16) + off16,
                                   a.ld16, a.Linv32, a.Uinv32, a.inv_scales + 4 * (st.T / leaf::DB), st.first_in_tile, st.blk,
                                   a.bf16, a.status, (a.dbg_clk && cta == 0 && st.blk % 16 == 8) ? a.dbg_clk + 300 : nullptr, st.valid, ltc);
            }
        } else {
            const int ntiles = st.tile_end[st.num_problems - 1];
            if (warp == F_EPI_WARPS) {
                // ------------------------------------------------------------ TMA producer
                if (lane == 0) {
                    if (prof && s < 61) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); a.dbg_clk[384 + 4 * s] = t_; }
                    fence_proxy_async_all();  // the previous step's generic-proxy stores (made visible by the barrier) -> TMA
                    if (prof && s < 61) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); a.dbg_clk[384 + 4 * s + 1] = t_; }
                    for (int t = slot0; t < ntiles; t += G) {
                        const TileRef r = tile_ref(st, probs, t);
                        const FusedProblem& p = *r.p;
                        const CUtensorMap* tmA = &maps.a[p.a_map];
                        const CUtensorMap* tmB = &maps.b[p.b_map];
                        const int m0 = r.mt * FBM, n0 = r.nt * FBN;
                        for (int kb = r.kb0; kb < r.kb1; ++kb) {
                            ptx::mbar_wait(&empty[stage], phase ^ 1);
                            ptx::mbar_arrive_expect_tx(&full[stage], F_A_BYTES + F_B_BYTES);
                            uint8_t* a_dst = sA + stage * F_A_BYTES;
                            uint8_t* b_dst = sB + stage * F_B_BYTES;
                            const int k0 = kb * FBK;
                            // A is column-major (M contiguous): two 64(m) x 64(k) boxes; B is K-major: one 64(k) x 128(n) box
                            ptx::tma_load_2d(a_dst, tmA, &full[stage], p.a_r0 + m0, p.a_c0 + k0);
                            ptx::tma_load_2d(a_dst + 8192, tmA, &full[stage], p.a_r0 + m0 + 64, p.a_c0 + k0);
                            ptx::tma_load_2d(b_dst, tmB, &full[stage], p.b_r0 + k0, p.b_c0 + n0);
                            if (++stage == FSTAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
                __syncwarp();
            } else if (warp == F_EPI_WARPS + 1) {
                // ------------------------------------------------------------ MMA issuer
                for (int t = slot0; t < ntiles; t += G, ++acc_iter) {
                    const TileRef r = tile_ref(st, probs, t);
                    const uint32_t as = acc_iter & 1, aphase = (acc_iter >> 1) & 1;
                    ptx::mbar_wait(&tempty[as], aphase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * FBN;
                    for (int kb = r.kb0; kb < r.kb1; ++kb) {
                        ptx::mbar_wait(&full[stage], phase);
                        ptx::tc_fence_after();
                        if (prof && lane == 0 && s < 61 && t == 0 && kb == r.kb0) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); a.dbg_clk[384 + 4 * s + 2] = t_; }
                        if (lane == 0) {
                            const uint32_t a_base = ptx::smem_u32(sA + stage * F_A_BYTES);
                            const uint32_t b_base = ptx::smem_u32(sB + stage * F_B_BYTES);
#pragma unroll
                            for (int k = 0; k < FBK / FUK; ++k) {
                                // A: [k][64 m] rows of 128 B, 8-row swizzle atoms: next atom along K +1024 B, next 64-row slab
                                // along M +8192 B, 16 k-rows per MMA = 2048 B; B: [n][64 k] rows of 128 B, 16 k = 32 B in the row
                                const uint64_t adesc = ptx::make_smem_desc_sw128(a_base + k * 2048, 8192, 1024);
                                const uint64_t bdesc = ptx::make_smem_desc_sw128(b_base + k * 32, 0, 1024);
                                ptx::umma_f16<1>(d_tmem, adesc, bdesc, idesc, (kb != r.kb0 || k != 0) ? 1u : 0u);
                            }
                            ptx::umma_commit<1>(&empty[stage]);
                            if (kb == r.kb1 - 1) ptx::umma_commit<1>(&tfull[as]);
                            if (prof && s < 61 && t == 0 && kb == r.kb1 - 1) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); a.dbg_clk[384 + 4 * s + 3] = t_; }
                        }
                        __syncwarp();
                        if (++stage == FSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            } else if (warp < F_EPI_WARPS) {
                // ------------------------------------------------------------ epilogue: warp w reads TMEM lanes 32*(w%4)..,
                // columns [64*(w/4), +64) of the tile as two chunks of 32 (one row per thread)
                const uint32_t q = warp & 3, half = warp >> 2;
                // the result tile is staged in shared memory as [column][row] (rows contiguous, like the matrix) and leaves
                // as two TMA tile stores: a thread's 64 fp32 + 64 16-bit scalar global stores per tile kept the epilogue
                // at ~6k cycles, LSU-issue bound, on the critical path of every step
                float* stC = reinterpret_cast<float*>(smem + F_RING_BYTES);
                uint16_t* stH = reinterpret_cast<uint16_t*>(smem + F_RING_BYTES + F_STC_BYTES);
                bool stored = false;
                for (int t = slot0; t < ntiles; t += G, ++acc_iter) {
                    const TileRef r = tile_ref(st, probs, t);
                    const FusedProblem& p = *r.p;
                    const int trow = q * 32 + lane;          // row inside the tile
                    const int tcol0 = half * 64;             // first of this warp's 64 columns inside the tile
                    float alpha = p.alpha;
                    if (p.alpha_p1) alpha *= *p.alpha_p1;
                    if (p.alpha_p2) alpha *= *p.alpha_p2;
                    const float hs = p.hscale_p ? *p.hscale_p : 1.f;
                    float cin[2][32];
                    if (p.accumulate) {  // addend loads in flight while the MMAs of this tile run
                        const float* src = a.W + (p.c_r0 + r.mt * FBM + trow) + (long long)(p.c_c0 + r.nt * FBN + tcol0) * a.ldw;
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int j = 0; j < 32; ++j) cin[c][j] = src[(long long)(c * 32 + j) * a.ldw];
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int j = 0; j < 32; ++j) cin[c][j] = 0.f;
                    }
                    const uint32_t as = acc_iter & 1, aphase = (acc_iter >> 1) & 1;
                    ptx::mbar_wait(&tfull[as], aphase);
                    ptx::tc_fence_after();
                    if (prof && tid == 0 && t == 0 && s < 61) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); a.dbg_clk[64 + 3 * s] = t_; }
                    if (stored) {  // the staging area is free once the previous tile's stores have read it
                        if (tid == 0) ptx::bulk_wait_group_read0();
                        ptx::named_bar_sync(1, F_EPI_WARPS * 32);
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t v[32];
                        ptx::tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + as * FBN + half * 64 + c * 32, v);
                        ptx::tmem_ld_wait();
                        if (c == 1) {  // accumulator fully read: hand the TMEM stage back
                            ptx::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(&tempty[as]);
                        }
                        const int sidx = (tcol0 + c * 32) * FBM + trow;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float o = fmaf(alpha, __uint_as_float(v[j]), cin[c][j]);
                            stC[sidx + j * FBM] = o;
                            const float hv = o * hs;
                            if (a.bf16) stH[sidx + j * FBM] = __bfloat16_as_ushort(__float2bfloat16_rn(hv));
                            else { stH[sidx + j * FBM] = __half_as_ushort(__float2half_rn(hv)); mx = fmaxf(mx, fabsf(hv)); }
                        }
                    }
                    ptx::fence_proxy_async();  // the staged tile -> async proxy
                    ptx::named_bar_sync(1, F_EPI_WARPS * 32);
                    if (tid == 0) {
                        if (p.c_r0 >= 0) ptx::tma_store_2d(&maps.c, stC, p.c_r0 + r.mt * FBM, p.c_c0 + r.nt * FBN);
                        if (p.h_map >= 0) ptx::tma_store_2d(&maps.h[p.h_map], stH, p.h_r0 + r.mt * FBM, p.h_c0 + r.nt * FBN);
                        ptx::bulk_commit_group();
                    }
                    stored = true;
                    if (prof && tid == 0 && s < 61) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); a.dbg_clk[64 + 3 * s + 1] = t_; }
                }
                if (stored && tid == 0) ptx::bulk_wait_group0();  // this step's tile stores are performed before the barrier
            }
        }
        grid_step_arrive(a.barrier, (prof && s < 61) ? a.dbg_clk + 64 + 3 * s + 2 : nullptr);
        grid_step_wait(a.barrier, (unsigned)G * (unsigned)(s + 1));
    }
    stamp(a.num_steps, false);
    stamp(a.num_steps + 2, true);
    if (warp < F_EPI_WARPS) {
        const float hmax = a.bf16 ? 3.0e38f : 65504.f;
        if (a.status && __any_sync(0xffffffffu, mx > hmax) && lane == 0) atomicOr(a.status, 1);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == F_EPI_WARPS + 1) ptx::tmem_dealloc<1>(tmem_base, F_TMEM_COLS);
}

}  // namespace

int getrf_fused_init() {
    return (int)cudaFuncSetAttribute(getrf_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES);
}

int launch_getrf_fused(const FusedMaps& maps, const FusedArgs& args, int num_ctas, cudaStream_t st) {
    if (num_ctas < 2 || (num_ctas & 1) || args.num_steps <= 0 || !args.program || !args.barrier) return (int)cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_ctas);
    cfg.blockDim = dim3(F_THREADS);
    cfg.dynamicSmemBytes = F_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // the leaf's two CTAs are one cluster
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, getrf_fused_kernel, maps, args);
}

}  // namespace mplu

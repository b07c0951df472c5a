// Dataflow GETRF of one diagonal block (up to a whole nb x nb tile) in ONE persistent launch without grid barriers.
//
// getrf_fused.cu runs the inverse-carrying recursion as a step program: every product group between two leaves is a
// step, and a step costs four L2 round trips in series (barrier poll, operand fetch, result store, barrier arrive) --
// 15-25k cycles under the bulk lane's memory traffic next to a few hundred cycles of tensor-core work, ~2.8 such steps
// per 128-block on the critical path (profiles/r02_fused_getrf_steps.txt).  Here the block is factored RIGHT-looking at
// 128-block granularity instead, which needs no merged inverse between two leaves:
//     leaf(i):  L\U of D_i + inv(L_i), inv(U_i)                                    (CTAs 0/1, leaf_tc.cuh)
//     P-tasks:  L(j,i) = A(j,i) inv(U_i),  U(i,k) = inv(L_i) A(i,k)                 j, k > i
//     S-tasks:  A(j,k) -= L(j,i) U(i,k)                                             rank-128 updates, j, k > i
//     merges:   the block-recursive merges of the 128-block inverses into the inverse of the whole block (only the
//               panel solves OUTSIDE this launch need it): off the leaf-to-leaf path
// Every task is one 128x128 output tile.  The host writes all tasks of the block into ONE list in priority order (the
// three tasks between leaf i and leaf i+1 first, then what step i+1's panel needs, then the rest); helper CTAs take the
// next list entry with an atomic add on a queue head, wait until the entry's dependency counters have reached their
// targets, run the product through the same TMA -> tcgen05 -> TMEM pipeline as getrf_fused.cu and bump the entry's
// counters once the result tile's TMA stores have completed.  A helper that holds an entry whose inputs are not there yet
// simply waits: the list is a topological order, so the earliest unfinished entry can always run.  Between two leaves
// there are now three dependent tile products handed over through counters instead of ~2.8 barrier-separated steps, and
// the leaf CTAs never run anything but leaves.
// Replaces, for the diagonal blocks, the dgetf2_native_npv + cublasDtrsm + cublasDgemm chain of the reference's panel
// loop (/root/reference/MPF.cu:166-239, dgetf2_native_npv.cu:18-35).  The products are the same tcgen05 products on the
// same 16-bit operands as the other GETRF paths, summed in rank-128 pieces: factors agree to rounding level.
#include "getrf_fused.h"

#include "gemm_tc.h"
#include "leaf_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mplu {

namespace {

constexpr int WBM = 128, WBN = 128, WBK = 64, WUK = 16;
constexpr int WSTAGES = 3;
constexpr int W_A_BYTES = WBM * WBK * 2;  // 16 KiB
constexpr int W_B_BYTES = WBN * WBK * 2;  // 16 KiB
constexpr int W_RING_BYTES = WSTAGES * (W_A_BYTES + W_B_BYTES);
constexpr int W_STC_BYTES = WBM * WBN * 4, W_STH_BYTES = WBM * WBN * 2;  // result tile staged for the TMA stores
constexpr int W_GEMM_BYTES = W_RING_BYTES + W_STC_BYTES + W_STH_BYTES;
constexpr int W_MAIN_RAW = leaf::DL_SMEM_BYTES > W_GEMM_BYTES ? leaf::DL_SMEM_BYTES : W_GEMM_BYTES;
constexpr int W_MAIN_BYTES = ((W_MAIN_RAW + 1023) / 1024) * 1024;
constexpr int W_PROB_CAP = 16 * 1024;  // the problem descriptors are staged in shared memory when they fit
constexpr int W_BAR_BYTES = 512;
constexpr int W_SMEM_BYTES = W_MAIN_BYTES + W_BAR_BYTES + W_PROB_CAP + 1024;  // + alignment slack
constexpr int W_THREADS = leaf::DL_THREADS;  // 16 warps; helpers: 0-7 epilogue, 8 producer, 9 MMA issuer, 10 store + signal
constexpr int W_EPI_WARPS = 8;
constexpr int W_TMEM_COLS = 256;  // two 128-column fp32 accumulators
constexpr int W_TQ = 4;           // depth of the CTA's task FIFO (producer -> MMA issuer / epilogue / store thread)
constexpr uint16_t W_NONE = 0xFFFFu;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ long long globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

struct TaskRegs {
    uint32_t problem, mt, nt, kb0, kb1;
    uint16_t sig[3];
};
__device__ __forceinline__ TaskRegs read_slot(const FlowTask* slot) {
    const volatile FlowTask* s = slot;
    TaskRegs r;
    r.problem = s->problem; r.mt = s->mt; r.nt = s->nt; r.kb0 = s->kb0; r.kb1 = s->kb1;
    r.sig[0] = s->sig_ctr[0]; r.sig[1] = s->sig_ctr[1]; r.sig[2] = s->sig_ctr[2];
    return r;
}

__global__ void __launch_bounds__(W_THREADS, 1)
getrf_flow_kernel(const __grid_constant__ FusedMaps maps, const FlowArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* sA = smem;
    uint8_t* sB = smem + WSTAGES * W_A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + W_MAIN_BYTES);
    uint64_t* empty = full + WSTAGES;
    uint64_t* tfull = empty + WSTAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* lbar = tempty + 2;       // the leaf's tensor-core products
    uint64_t* tqf = lbar + 1;          // task FIFO slot filled (producer)
    uint64_t* tqe = tqf + W_TQ;        // ... released by the MMA issuer, the 8 epilogue warps and the store thread
    uint64_t* st_ready = tqe + W_TQ;   // result tile staged (one arrival per epilogue warp)
    uint64_t* st_free = st_ready + 1;  // its TMA stores have read the staging area
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_free + 1);
    FlowTask* slots = reinterpret_cast<FlowTask*>(smem + W_MAIN_BYTES + 256);
    uint8_t* sprob = smem + W_MAIN_BYTES + W_BAR_BYTES;

    const int tid = threadIdx.x;
    const uint32_t warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x;

    const int prob_bytes = a.num_problems * (int)sizeof(FusedProblem);
    const bool prob_in_smem = prob_bytes <= W_PROB_CAP;
    if (prob_in_smem && cta >= 2) {
        const uint4* src = reinterpret_cast<const uint4*>(a.problems);
        uint4* dst = reinterpret_cast<uint4*>(sprob);
        for (int i = tid; i < prob_bytes / 16; i += W_THREADS) dst[i] = __ldg(src + i);
    }
    const FusedProblem* probs = prob_in_smem ? reinterpret_cast<const FusedProblem*>(sprob) : a.problems;

    if (warp == W_EPI_WARPS && lane == 0) {
        for (int i = 0; i < FM_COUNT; ++i) {
            ptx::prefetch_tmap(&maps.a[i]);
            ptx::prefetch_tmap(&maps.b[i]);
            ptx::prefetch_tmap(&maps.h[i]);
        }
        ptx::prefetch_tmap(&maps.c);
        for (int i = 0; i < WSTAGES; ++i) {
            ptx::mbar_init(&full[i], 1);   // the producer's arrive.expect_tx
            ptx::mbar_init(&empty[i], 1);  // one tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);             // one tcgen05.commit
            ptx::mbar_init(&tempty[i], W_EPI_WARPS);  // one arrival per epilogue warp
        }
        ptx::mbar_init(lbar, 1);
        for (int i = 0; i < W_TQ; ++i) {
            ptx::mbar_init(&tqf[i], 1);
            ptx::mbar_init(&tqe[i], W_EPI_WARPS + 2);
        }
        ptx::mbar_init(st_ready, W_EPI_WARPS);
        ptx::mbar_init(st_free, 1);
        ptx::fence_mbar_init();
    }
    if (warp == W_EPI_WARPS + 1) ptx::tmem_alloc<1>(tmem_slot, W_TMEM_COLS);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    float mx = 0.f;  // largest scaled fp16 magnitude written (overflow detection)

    if (cta < 2) {
        // ------------------------------------------------------------------------------------------ the leaf chain
        leaf::LeafTc ltc{tmem_base, lbar, 0u};
#pragma unroll 1
        for (int i = 0; i < a.num_leaves; ++i) {
            const FlowLeaf lf = a.leaves[i];
            if (tid == 0) {
                if (lf.wait_ctr != W_NONE)
                    while (ld_acquire_gpu(a.counters + lf.wait_ctr) < lf.wait_val) { }
                __threadfence();
                if (a.dbg && cta == 0) a.dbg[2 * i] = globaltimer();
            }
            __syncthreads();
            const long long off16 = (long long)(lf.k0 - lf.T) + (long long)lf.k0 * a.ld16;
            leaf::diag_lu_body_tc(reinterpret_cast<float*>(smem), cta, a.W, a.ldw, lf.k0,
                                  reinterpret_cast<uint16_t*>(a.Linv16) + off16, reinterpret_cast<uint16_t*>(a.Uinv16) + off16, a.ld16,
                                  a.Linv32, a.Uinv32, a.inv_scales + 4 * (lf.T / leaf::DB), lf.first_in_tile, lf.blk, a.bf16, a.status,
                                  nullptr, lf.valid, ltc);
            // L\U, the 16-bit inverses and the tile scales (generic-proxy stores of all threads) -> visible to the helpers'
            // TMA loads and plain loads once they have seen the counter
            fence_proxy_async_all();
            __syncthreads();
            if (tid == 0) {
                red_release_gpu(a.counters + lf.sig_ctr);
                if (a.dbg && cta == 0) a.dbg[2 * i + 1] = globaltimer();
            }
        }
    } else if (warp == W_EPI_WARPS) {
        // ------------------------------------------------------------------------------------------ producer
        long long* tdbg = a.dbg ? a.dbg + 2 * a.num_leaves : nullptr;
        uint32_t stage = 0, phase = 0;
        // two lists: the main one (everything the leaf chain and the updates need) and the inverse merges, which only the
        // launches after this one read.  A few helpers serve the merges first so that they neither delay the main list
        // nor pile up behind it; a helper whose first list is exhausted moves on to the other one.
        int qsel = cta >= (int)gridDim.x - a.merge_ctas ? 1 : 0;
        bool switched = false;
#pragma unroll 1
        for (int n = 0;; ++n) {
            unsigned idx = 0;
            bool valid;
            for (;;) {
                unsigned got = 0;
                if (lane == 0) got = atomicAdd(a.counters + qsel, 1u);
                got = __shfl_sync(0xffffffffu, got, 0);
                const unsigned limit = qsel ? (unsigned)(a.num_tasks - a.num_main) : (unsigned)a.num_main;
                if (got < limit) { idx = qsel ? got + (unsigned)a.num_main : got; valid = true; break; }
                if (switched) { valid = false; break; }
                switched = true;
                qsel ^= 1;
            }
            uint4 q0 = make_uint4(0xFFFFu, 0, 0, 0), q1 = make_uint4(0, 0, 0, 0);
            if (valid) {
                const uint4* src = reinterpret_cast<const uint4*>(a.tasks + idx);
                q0 = __ldg(src);
                q1 = __ldg(src + 1);
                if (tdbg && lane == 0) { tdbg[4 * idx] = globaltimer(); tdbg[4 * idx + 3] = cta; }
                // 16-bit fields of the record: 0 problem, 1 mt|nt, 2 kb0|kb1, 3..6 wait_ctr, 7..10 wait_val, 11..13 sig_ctr
                const uint32_t w[6] = {q0.y, q0.z, q0.w, q1.x, q1.y, q1.z};  // 32-bit words 1..6 hold fields 2..13
                auto field = [&](uint32_t i) -> uint32_t {                    // i in 3..10
                    uint32_t v = 0;
#pragma unroll
                    for (uint32_t j = 1; j <= 5; ++j)
                        if ((i >> 1) == j) v = w[j - 1];
                    return (i & 1) ? v >> 16 : v & 0xFFFFu;
                };
                if (lane < 4) {  // one dependency per lane
                    const uint32_t ctr = field(3 + lane), val = field(7 + lane);
                    if (ctr != W_NONE)
                        while (ld_acquire_gpu(a.counters + ctr) < val) { }
                }
                __syncwarp();
            }
            if (lane == 0) {
                if (valid) {
                    fence_proxy_async_all();  // other CTAs' results (made visible by the counters) -> this thread's TMA loads
                    if (tdbg) tdbg[4 * idx + 1] = globaltimer();
                }
                const int slot = n % W_TQ;
                ptx::mbar_wait(&tqe[slot], ((n / W_TQ) & 1) ^ 1);
                uint4* dst = reinterpret_cast<uint4*>(&slots[slot]);
                dst[0] = q0;
                dst[1] = make_uint4(q1.x, q1.y, q1.z, valid ? idx : 0u);  // the padding carries the list index (debug stamps)
                ptx::mbar_arrive(&tqf[slot]);
                if (valid) {
                    const uint32_t problem = q0.x & 0xFFFFu, mt = (q0.x >> 16) & 0xFFu, nt = q0.x >> 24;
                    const uint32_t kb0 = q0.y & 0xFFu, kb1 = (q0.y >> 8) & 0xFFu;
                    const FusedProblem& p = probs[problem];
                    const CUtensorMap* tmA = &maps.a[p.a_map];
                    const CUtensorMap* tmB = &maps.b[p.b_map];
                    const int m0 = mt * WBM, n0 = nt * WBN;
                    for (uint32_t kb = kb0; kb < kb1; ++kb) {
                        ptx::mbar_wait(&empty[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&full[stage], W_A_BYTES + W_B_BYTES);
                        uint8_t* a_dst = sA + stage * W_A_BYTES;
                        uint8_t* b_dst = sB + stage * W_B_BYTES;
                        const int k0 = kb * WBK;
                        // A is column-major (M contiguous): two 64(m) x 64(k) boxes; B is K-major: one 64(k) x 128(n) box
                        ptx::tma_load_2d(a_dst, tmA, &full[stage], p.a_r0 + m0, p.a_c0 + k0);
                        ptx::tma_load_2d(a_dst + 8192, tmA, &full[stage], p.a_r0 + m0 + 64, p.a_c0 + k0);
                        ptx::tma_load_2d(b_dst, tmB, &full[stage], p.b_r0 + k0, p.b_c0 + n0);
                        if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            __syncwarp();
            if (!valid) break;
        }
    } else if (warp == W_EPI_WARPS + 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_f16(WBM, WBN, a.bf16 != 0, true, false);
        uint32_t stage = 0, phase = 0, acc_iter = 0;
#pragma unroll 1
        for (int n = 0;; ++n, ++acc_iter) {
            const int slot = n % W_TQ;
            ptx::mbar_wait(&tqf[slot], (n / W_TQ) & 1);
            const TaskRegs t = read_slot(&slots[slot]);
            if (t.problem == W_NONE) break;
            const uint32_t as = acc_iter & 1, aphase = (acc_iter >> 1) & 1;
            ptx::mbar_wait(&tempty[as], aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * WBN;
            for (uint32_t kb = t.kb0; kb < t.kb1; ++kb) {
                ptx::mbar_wait(&full[stage], phase);
                ptx::tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_base = ptx::smem_u32(sA + stage * W_A_BYTES);
                    const uint32_t b_base = ptx::smem_u32(sB + stage * W_B_BYTES);
#pragma unroll
                    for (int k = 0; k < WBK / WUK; ++k) {
                        // A: [k][64 m] rows of 128 B, 8-row swizzle atoms: next atom along K +1024 B, next 64-row slab along M
                        // +8192 B, 16 k-rows per MMA = 2048 B; B: [n][64 k] rows of 128 B, 16 k = 32 B in the row
                        const uint64_t adesc = ptx::make_smem_desc_sw128(a_base + k * 2048, 8192, 1024);
                        const uint64_t bdesc = ptx::make_smem_desc_sw128(b_base + k * 32, 0, 1024);
                        ptx::umma_f16<1>(d_tmem, adesc, bdesc, idesc, (kb != t.kb0 || k != 0) ? 1u : 0u);
                    }
                    ptx::umma_commit<1>(&empty[stage]);
                    if (kb == t.kb1 - 1) ptx::umma_commit<1>(&tfull[as]);
                }
                __syncwarp();
                if (++stage == WSTAGES) { stage = 0; phase ^= 1; }
            }
            if (lane == 0) ptx::mbar_arrive(&tqe[slot]);
        }
    } else if (warp < W_EPI_WARPS) {
        // ------------------------------------------------------------------------------------------ epilogue: warp w reads
        // TMEM lanes 32*(w%4).., columns [64*(w/4), +64) of the tile as two chunks of 32 (one row per thread); the result
        // tile is staged in shared memory as [column][row] and leaves as TMA tile stores issued by the store thread
        const uint32_t q = warp & 3, half = warp >> 2;
        float* stC = reinterpret_cast<float*>(smem + W_RING_BYTES);
        uint16_t* stH = reinterpret_cast<uint16_t*>(smem + W_RING_BYTES + W_STC_BYTES);
        uint32_t acc_iter = 0;
#pragma unroll 1
        for (int n = 0;; ++n, ++acc_iter) {
            const int slot = n % W_TQ;
            ptx::mbar_wait(&tqf[slot], (n / W_TQ) & 1);
            const TaskRegs t = read_slot(&slots[slot]);
            if (t.problem == W_NONE) break;
            const FusedProblem& p = probs[t.problem];
            const int trow = q * 32 + lane;  // row inside the tile
            const int tcol0 = half * 64;     // first of this warp's 64 columns inside the tile
            // scales written by a leaf of THIS launch (inverse scales of the tile) must not come out of a stale L1 line
            float alpha = p.alpha;
            if (p.alpha_p1) alpha *= __ldcg(p.alpha_p1);
            if (p.alpha_p2) alpha *= __ldcg(p.alpha_p2);
            const float hs = p.hscale_p ? __ldcg(p.hscale_p) : 1.f;
            float cin[2][32];
            if (p.accumulate) {  // addend loads (L2: another CTA's TMA store may have rewritten the tile) in flight under the MMAs
                const float* src = a.W + (p.c_r0 + t.mt * WBM + trow) + (long long)(p.c_c0 + t.nt * WBN + tcol0) * a.ldw;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) cin[c][j] = __ldcg(src + (long long)(c * 32 + j) * a.ldw);
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) cin[c][j] = 0.f;
            }
            const uint32_t as = acc_iter & 1, aphase = (acc_iter >> 1) & 1;
            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            if (n > 0) ptx::mbar_wait(st_free, (uint32_t)(n - 1) & 1u);  // the previous tile's stores have read the staging area
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                ptx::tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + as * WBN + half * 64 + c * 32, v);
                ptx::tmem_ld_wait();
                if (c == 1) {  // accumulator fully read: hand the TMEM stage back
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tempty[as]);
                }
                const int sidx = (tcol0 + c * 32) * WBM + trow;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float o = fmaf(alpha, __uint_as_float(v[j]), cin[c][j]);
                    stC[sidx + j * WBM] = o;
                    const float hv = o * hs;
                    if (a.bf16) stH[sidx + j * WBM] = __bfloat16_as_ushort(__float2bfloat16_rn(hv));
                    else { stH[sidx + j * WBM] = __half_as_ushort(__float2half_rn(hv)); mx = fmaxf(mx, fabsf(hv)); }
                }
            }
            ptx::fence_proxy_async();  // the staged tile -> async proxy
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(st_ready);
                ptx::mbar_arrive(&tqe[slot]);
            }
        }
    } else if (warp == W_EPI_WARPS + 2 && lane == 0) {
        // ------------------------------------------------------------------------------------------ store + signal thread
        float* stC = reinterpret_cast<float*>(smem + W_RING_BYTES);
        uint16_t* stH = reinterpret_cast<uint16_t*>(smem + W_RING_BYTES + W_STC_BYTES);
        long long* tdbg = a.dbg ? a.dbg + 2 * a.num_leaves : nullptr;
#pragma unroll 1
        for (int n = 0;; ++n) {
            const int slot = n % W_TQ;
            ptx::mbar_wait(&tqf[slot], (n / W_TQ) & 1);
            const TaskRegs t = read_slot(&slots[slot]);
            if (t.problem == W_NONE) break;
            const unsigned idx = reinterpret_cast<const volatile uint32_t*>(&slots[slot])[7];
            const FusedProblem& p = probs[t.problem];
            ptx::mbar_wait(st_ready, (uint32_t)n & 1u);
            if (p.c_r0 >= 0) ptx::tma_store_2d(&maps.c, stC, p.c_r0 + t.mt * WBM, p.c_c0 + t.nt * WBN);
            if (p.h_map >= 0) ptx::tma_store_2d(&maps.h[p.h_map], stH, p.h_r0 + t.mt * WBM, p.h_c0 + t.nt * WBN);
            ptx::bulk_commit_group();
            ptx::bulk_wait_group_read0();
            ptx::mbar_arrive(st_free);
            ptx::bulk_wait_group0();  // the tile is written: tell the tasks (and leaves) that wait for it
            fence_proxy_async_all();
#pragma unroll
            for (int s = 0; s < 3; ++s)
                if (t.sig[s] != W_NONE) red_release_gpu(a.counters + t.sig[s]);
            if (tdbg) tdbg[4 * idx + 2] = globaltimer();
            ptx::mbar_arrive(&tqe[slot]);
        }
    }

    if (cta >= 2 && warp < W_EPI_WARPS) {
        const float hmax = a.bf16 ? 3.0e38f : 65504.f;
        if (a.status && __any_sync(0xffffffffu, mx > hmax) && lane == 0) atomicOr(a.status, 1);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == W_EPI_WARPS + 1) ptx::tmem_dealloc<1>(tmem_base, W_TMEM_COLS);
}

}  // namespace

int getrf_flow_init() {
    return (int)cudaFuncSetAttribute(getrf_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BYTES);
}

int launch_getrf_flow(const FusedMaps& maps, const FlowArgs& args, int num_ctas, cudaStream_t st) {
    if (num_ctas < 4 || (num_ctas & 1) || args.num_leaves <= 0 || !args.counters) return (int)cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(num_ctas);
    cfg.blockDim = dim3(W_THREADS);
    cfg.dynamicSmemBytes = W_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // the leaf's two CTAs are one cluster
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, getrf_flow_kernel, maps, args);
}

}  // namespace mplu

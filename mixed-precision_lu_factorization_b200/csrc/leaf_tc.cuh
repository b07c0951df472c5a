// The 128x128 leaf of the recursive GETRF as it runs inside the fused GETRF kernel (getrf_fused.cu), which owns TMEM:
// same algorithm and outputs as leaf.cuh (no-pivot LU of the block, the column elimination of
// /root/reference/dgetf2_native_npv.cu:18-35, plus explicit inv(L11), inv(U11)), rewritten for two measured facts
// (profiles/r02_fused_getrf_steps.txt):
//   1. The stand-alone leaf is ~8k fully unrolled SASS instructions (130 KB) that execute once per block.  Inside the
//      fused kernel the same two SMs run leaf after leaf, but the code does not fit the instruction cache, and while the
//      bulk lane keeps L2 busy every first use of a code region stalls on instruction fetch: 110-135k cycles per leaf
//      next to 75k on an idle GPU.  Here every phase is a ROLLED loop over groups of 8 columns (P1, the substitutions) or
//      over 16-byte operand chunks (staging, write-back): ~2.5k instructions that stay resident.
//   2. The O(n^3) parts -- the rank-32 Schur updates (P3: 21k cycles) and the block-recursive merges of the inverses
//      (I2: 19k cycles) -- were bound by shared-memory bandwidth on the FMA pipe.  Here they are tcgen05.mma products:
//      fp32 values are split x = p0 + p1 + p2 into three bf16 parts (24 mantissa bits) and a product is the six part
//      products of order <= 2 (p0 p0 + p0 p1 + p1 p0 + p0 p2 + p1 p1 + p2 p0) with fp32 accumulation in TMEM: fp32-class
//      accuracy like the FMA leaf.  (Two parts, 16 bits, were measured first: fine on dominant matrices, but the bf16
//      operand mode at kappa = 1e6 went from 31 to 114 GMRES steps -- ill-conditioned diagonal blocks need the leaf's
//      full precision.)  Operand tiles are staged by all 512 threads in the SWIZZLE_128B layout of gemm_tc.cu's descriptors.
// Two CTAs (one cluster) per block as before: both factor the block, CTA 0 delivers L\U + inv(L11), CTA 1 inv(U11).
#pragma once
#include "leaf.cuh"

namespace mplu {
namespace leaf {

struct LeafTc {
    uint32_t tmem;   // TMEM base address: 128 lanes x 128 fp32 columns are used
    uint64_t* bar;   // mbarrier (count 1) the MMAs are committed to
    uint32_t phase;  // its phase parity, tracked by every thread alike
};

// staging region (third 128 x 129 fp32 array of the leaf's shared memory, 1024-byte aligned): operand slabs of up to
// 128 rows x 128 bytes
// K = 32 operands keep two parts in one 128-byte row [p0(32) | p1(32)] of their first slab (A: +0, B: +32768) and p2 in
// the first half-rows of their second slab (A: +16384, B: +49152); K = 64 operands and MN-major operands (rows of up to
// 64 values) use three slabs of 8 KiB (A: +0, +8192, +16384 -- the MMA reads 128 rows, the upper 64 are don't-care rows;
// B: +32768, +40960, +49152).
constexpr int TC_A0 = 0, TC_A1 = 8192, TC_A2 = 16384, TC_B0 = 32768, TC_B1 = 40960, TC_B2 = 49152;
constexpr int TC_STAGE_BYTES = 65536;
static_assert(TC_STAGE_BYTES <= DB * LDS * 4 && (2 * DB * LDS * 4) % 1024 == 0, "staging area");

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {  // a in the low half
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// one 16-byte chunk (8 consecutive elements) of tile row `row` as three bf16 parts: part i to chunk ch[i] of slab sl[i];
// chunk index XOR row%8 inside the 128-byte row (SWIZZLE_128B)
struct TcSlabs {
    uint8_t* sl[3];
    int ch0[3];  // first chunk of each part inside its slab's rows
};
__device__ __forceinline__ void stage_chunk(const TcSlabs& t, int row, int c, const float (&v)[8]) {
    uint32_t p0[4], p1[4], p2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        p0[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        const float r0 = v[2 * i] - __uint_as_float(p0[i] << 16), r1 = v[2 * i + 1] - __uint_as_float(p0[i] & 0xffff0000u);
        p1[i] = pack_bf16x2(r0, r1);
        p2[i] = pack_bf16x2(r0 - __uint_as_float(p1[i] << 16), r1 - __uint_as_float(p1[i] & 0xffff0000u));
    }
    const uint32_t rbase = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u, x = (uint32_t)(row & 7);
    *reinterpret_cast<uint4*>(t.sl[0] + rbase + ((((uint32_t)(t.ch0[0] + c)) ^ x) << 4)) = make_uint4(p0[0], p0[1], p0[2], p0[3]);
    *reinterpret_cast<uint4*>(t.sl[1] + rbase + ((((uint32_t)(t.ch0[1] + c)) ^ x) << 4)) = make_uint4(p1[0], p1[1], p1[2], p1[3]);
    *reinterpret_cast<uint4*>(t.sl[2] + rbase + ((((uint32_t)(t.ch0[2] + c)) ^ x) << 4)) = make_uint4(p2[0], p2[1], p2[2], p2[3]);
}
// slabs of a K = 32 operand ([p0 | p1] rows + p2 rows) / of a 64-wide or MN-major operand (three slabs) at `base`
__device__ __forceinline__ TcSlabs slabs32(uint8_t* first, uint8_t* second) { return TcSlabs{{first, first, second}, {0, 4, 0}}; }
__device__ __forceinline__ TcSlabs slabs3(uint8_t* base, int off = 0) { return TcSlabs{{base + off, base + 8192 + off, base + 16384 + off}, {0, 0, 0}}; }

// An operand element (r, k) = sign * p[r * rs + k * ks], read as zero where the never-written upper-right 32x32 block of
// a 64x64 triangular matrix would be touched: zmode 1: r < 32 <= k, zmode 2: k < 32 <= r.
struct TcSrc {
    const float* p;
    int rs, ks;
    float sign;
    int zmode;
};
// Stage rows [0, nrows) of the operand as tile rows row0 + r, 8 * nch elements per row, by all threads of the CTA.
//   nch == 4 (K = 32): both parts in one slab, row = [h(32) | l(32)]  (pass ls == hs)
//   nch == 8 (K = 64): parts in two slabs
struct TcDst {
    TcSlabs slabs;
    int nrows, row0;  // operand rows [0, nrows) become tile rows row0 + r
};
__device__ __forceinline__ void stage_item(const TcDst& d, const TcSrc& s, int nch, int idx) {
    const int lsh = nch == 4 ? 2 : 3;
    const int r = idx >> lsh, c = idx & (nch - 1);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = 8 * c + i;
        const bool z = (s.zmode == 1 && r < SB && k >= SB) || (s.zmode == 2 && k < SB && r >= SB);
        v[i] = z ? 0.f : s.sign * s.p[r * s.rs + k * s.ks];
    }
    stage_chunk(d.slabs, d.row0 + r, c, v);
}
// up to two operands in one pass over the CTA's threads (the second starts at a warp boundary of the item space)
__device__ __forceinline__ void stage_tiles(const TcDst d0, const TcSrc s0, const TcDst d1, const TcSrc s1, int nch, int tid) {
    const int lsh = nch == 4 ? 2 : 3;
    const int n0 = d0.nrows << lsh, n1 = d1.nrows << lsh, n0r = (n0 + 31) & ~31;
#pragma unroll 1
    for (int idx = tid; idx < n0r + n1; idx += DL_THREADS) {
        if (idx < n0) stage_item(d0, s0, nch, idx);
        else if (idx >= n0r) stage_item(d1, s1, nch, idx - n0r);
    }
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

// Descriptors (k = 0) of the staging area's A and B bases, built once per leaf: a part / k-step / sub-operand is an add
// to the 16-byte-granular start-address field (descriptor construction by the single issuing thread cost ~1k cycles per
// product, on the critical path).
struct TcDesc {
    uint64_t a, b;  // K-major, no leading-byte offset; MN-major adds kTcMnLbo
};
constexpr uint64_t kTcMnLbo = (uint64_t)(8192 >> 4) << 16;
// D[128 x N] (TMEM, fp32) = A * B over K = 16 * KSTEPS from the six part products of order <= 2, small terms first;
// issued by ONE thread.  kK32: K = 32 operands in the [p0 | p1] + p2 layout, else three 8 KiB slabs.  B K-major (32 bytes per
// k-step) or MN-major (kBmn: three slabs, 16 k-rows = 2048 bytes per k-step); b_off: byte offset of a second operand.
template <int KSTEPS, bool kK32, bool kBmn>
__device__ __forceinline__ void tc_product(uint32_t d_tmem, const TcDesc& t, uint32_t b_off, int N) {
    const uint32_t idesc = make_idesc_f16(128, N, true, false, kBmn);
    uint64_t ap[3], bp[3];
    ap[0] = t.a;
    ap[1] = t.a + (kK32 ? 4 : (TC_A1 >> 4));
    ap[2] = t.a + (TC_A2 >> 4);
    const uint64_t bb = t.b + (b_off >> 4) + (kBmn ? kTcMnLbo : 0);
    bp[0] = bb;
    bp[1] = bb + ((kK32 && !kBmn) ? 4 : ((TC_B1 - TC_B0) >> 4));
    bp[2] = bb + ((TC_B2 - TC_B0) >> 4);
    constexpr uint64_t bstep = kBmn ? 128 : 2;
    constexpr int pa[6] = {2, 1, 0, 1, 0, 0}, pb[6] = {0, 1, 2, 0, 1, 0};
#pragma unroll
    for (int t6 = 0; t6 < 6; ++t6) {
#pragma unroll
        for (int k = 0; k < KSTEPS; ++k) ptx::umma_f16<1>(d_tmem, ap[pa[t6]] + 2 * k, bp[pb[t6]] + bstep * k, idesc, (t6 | k) ? 1u : 0u);
    }
}
// staged operands -> MMAs (one thread) -> completion, executed by every thread of the CTA
#define LEAF_TC_ISSUE_BEGIN()  \
    ptx::fence_proxy_async();  \
    __syncthreads();           \
    if (threadIdx.x == 0) {    \
        ptx::tc_fence_after();
#define LEAF_TC_ISSUE_END(tc)              \
        ptx::umma_commit<1>((tc).bar);     \
    }                                      \
    ptx::mbar_wait((tc).bar, (tc).phase);  \
    (tc).phase ^= 1;                       \
    ptx::tc_fence_after();

// x <- solution of a 32-step triangular substitution, one vector per thread, in groups of 8 steps over a register
// window that slides by 8 (the same ~300 instructions for every group, warp and use):
//   step k:  if (scale) x[k] *= s_rd[k];   x[c] -= x[k] * mt[k][c]  for c > k
// in(i) gives the initial x[i]; out(i, v) receives the final x[i].  Serves the panel solves (P2: rows of L21 with
// mt = U_D and the reciprocal pivots, columns of U12 with mt = L_D^T) and the sub-block inverses (I1: x = e_lane).
template <class In, class Out>
__device__ __forceinline__ void substitute32(const float (*mt)[SB], const float* rd, bool scale, In in, Out out) {
    float w[SB];
#pragma unroll
    for (int i = 0; i < SB; ++i) w[i] = in(i);
#pragma unroll 1
    for (int g = 0; g < SB / 8; ++g) {
        const int kb = 8 * g;  // window element i is x[kb + i]
        float rdg[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) rdg[jj] = scale ? rd[kb + jj] : 1.f;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int k = kb + jj;
            // window columns beyond 31 are dead: they read on into the next row of mt (one padding row at the end) and
            // are never stored -- cheaper than a predicate per group on this dependent chain.  Loads first: they do not
            // depend on the chain.
            float4 cv[SB / 4];
#pragma unroll
            for (int c4 = ((jj + 1) & ~3); c4 < SB; c4 += 4) cv[c4 >> 2] = *reinterpret_cast<const float4*>(&mt[k][kb + c4]);
            w[jj] *= rdg[jj];
            const float xk = w[jj];
#pragma unroll
            for (int c4 = ((jj + 1) & ~3); c4 < SB; c4 += 4) {
                const float4 v = cv[c4 >> 2];
                if (c4 > jj) w[c4] = fmaf(-xk, v.x, w[c4]);
                if (c4 + 1 > jj) w[c4 + 1] = fmaf(-xk, v.y, w[c4 + 1]);
                if (c4 + 2 > jj) w[c4 + 2] = fmaf(-xk, v.z, w[c4 + 2]);
                w[c4 + 3] = fmaf(-xk, v.w, w[c4 + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out(kb + i, w[i]);
#pragma unroll
        for (int i = 0; i < SB - 8; ++i) w[i] = w[i + 8];
    }
}

// `dl_smem`: the fused kernel's main shared-memory region (>= DL_SMEM_BYTES, 1024-byte aligned); `which`: 0 = this CTA
// delivers inv(L11) and the L\U block, 1 = inv(U11).  Called by all DL_THREADS threads of BOTH CTAs of a 2-CTA cluster.
__device__ __forceinline__ void diag_lu_body_tc(float* dl_smem, const int which, float* __restrict__ W, long long ldw, int k0,
                                                void* __restrict__ Linv16, void* __restrict__ Uinv16, long long ld16,
                                                float* __restrict__ Linv32, float* __restrict__ Uinv32, float* tile_scales,
                                                int first_in_tile, int blk, int bf16, int* status, long long* dbg_clk,
                                                int valid, LeafTc& tc) {
    int dbg_i = 0;
#define DBG_CLK() do { if (dbg_clk && threadIdx.x == 0) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); dbg_clk[dbg_i++] = t_; } } while (0)
    DBG_CLK();
    float* S = dl_smem;                 // the block -> L11\U11
    float* Xh = S + DB * LDS;           // this CTA's inverse, lower triangular: inv(L11), or inv(U11)^T = inv(U11^T)
    uint8_t* stage = reinterpret_cast<uint8_t*>(S + 2 * DB * LDS);
    __shared__ float s_rd[SB];
    __shared__ __align__(16) float s_ut[SB + 1][SB];  // s_ut[k][c] = U_D(k,c)   (+ one padding row: see substitute32)
    __shared__ __align__(16) float s_lt[SB + 1][SB];  // s_lt[k][r] = L_D(r,k)
    __shared__ __align__(16) float s_prow[2][2 * SB];  // pivot rows of the current / next step (+ padding for the window's dead tail)
    __shared__ float s_red[DL_THREADS / 32];
    __shared__ int s_zero;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* Wb = W + k0 + (long long)k0 * ldw;
    if (tid == 0) s_zero = 0;
    const TcDesc td{ptx::make_smem_desc_sw128(ptx::smem_u32(stage + TC_A0), 0, 1024), ptx::make_smem_desc_sw128(ptx::smem_u32(stage + TC_B0), 0, 1024)};
    {   // lanes -> consecutive rows: coalesced; 8 loads of a thread in flight per round
        const int r = tid & (DB - 1), cq = tid >> 7;
#pragma unroll 1
        for (int i0 = 0; i0 < DB / 4; i0 += 8) {
            float t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = Wb[r + (long long)(cq + 4 * (i0 + i)) * ldw];
#pragma unroll
            for (int i = 0; i < 8; ++i) S[r + (cq + 4 * (i0 + i)) * LDS] = t[i];
        }
    }
    // the two CTAs are one cluster: nobody writes the block back before both have read it
    ptx::cluster_sync_all();
    DBG_CLK();

#pragma unroll 1
    for (int kb = 0; kb < DB / SB; ++kb) {
        const int o = kb * SB;
        // ---- P1: 32x32 diagonal sub-block inside one warp, lane = row.  The pivot row of step j travels through shared
        // memory (the lane that owns it publishes it, 128-bit stores / broadcast loads); 8 steps are unrolled over a
        // register window of the row that slides by 8 columns per group.
        if (warp == 0) {
            float w[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) w[c] = S[(o + lane) + (o + c) * LDS];
            bool zp = false;
            if (lane == 0) {
#pragma unroll
                for (int c4 = 0; c4 < SB; c4 += 4)
                    *reinterpret_cast<float4*>(&s_prow[0][c4]) = make_float4(w[c4], w[c4 + 1], w[c4 + 2], w[c4 + 3]);
            }
            __syncwarp();
#pragma unroll 1
            for (int g = 0; g < SB / 8; ++g) {
                const int cb = 8 * g;  // window element i is column cb + i of the sub-block
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const int j = cb + jj;
                    const float* prow = s_prow[j & 1];
                    float4 pv[SB / 4];  // the pivot row's window: loads in flight under the reciprocal chain
#pragma unroll
                    for (int c4 = ((jj + 1) & ~3); c4 < SB; c4 += 4) pv[c4 >> 2] = *reinterpret_cast<const float4*>(&prow[cb + c4]);
                    const float piv = prow[j];
                    zp |= (piv == 0.f);
                    const float rp = fast_rcp(piv);
                    if (lane == 0) s_rd[j] = rp;
                    const float l = (lane > j) ? w[jj] * rp : 0.f;
                    w[jj] = (lane > j) ? l : w[jj];
#pragma unroll
                    for (int c4 = ((jj + 1) & ~3); c4 < SB; c4 += 4) {  // columns beyond 31: dead tail, no predicate
                        const float4 v = pv[c4 >> 2];
                        if (c4 > jj) w[c4] = fmaf(-l, v.x, w[c4]);
                        if (c4 + 1 > jj) w[c4 + 1] = fmaf(-l, v.y, w[c4 + 1]);
                        if (c4 + 2 > jj) w[c4 + 2] = fmaf(-l, v.z, w[c4 + 2]);
                        w[c4 + 3] = fmaf(-l, v.w, w[c4 + 3]);
                    }
                    if (j + 1 < SB) {
                        if (lane == j + 1) {  // this lane's row is final: publish it from its 4-aligned group on
                            float* nrow = s_prow[(j + 1) & 1];
#pragma unroll
                            for (int c4 = ((jj + 1) & ~3); c4 < SB; c4 += 4)
                                *reinterpret_cast<float4*>(&nrow[cb + c4]) = make_float4(w[c4], w[c4 + 1], w[c4 + 2], w[c4 + 3]);
                        }
                        __syncwarp();
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) S[(o + lane) + (o + cb + i) * LDS] = w[i];
#pragma unroll
                for (int i = 0; i < SB - 8; ++i) w[i] = w[i + 8];
            }
            if (zp && lane == 0) s_zero = 1;
        }
        __syncthreads();
#pragma unroll 1
        for (int e = tid; e < SB * SB; e += DL_THREADS) {
            const int k = e >> 5, c = e & 31;
            s_ut[k][c] = S[(o + k) + (o + c) * LDS];
            s_lt[k][c] = S[(o + c) + (o + k) * LDS];
        }
        if (tid < SB) { s_ut[SB][tid] = 0.f; s_lt[SB][tid] = 0.f; }  // keep the dead tail finite
        __syncthreads();
        DBG_CLK();
        const int m = DB - o - SB;  // rows below / columns right
        const int mw = m / 32;
        // ---- P2 + I1, one routine: warps 1..mw solve the rows of L21 = A21 inv(U_D), warps mw+1..2mw the columns of
        // U12 = inv(L_D) A12, warp 7 + which the sub-block's inverse (inv(L_D), or inv(U_D^T): columns = e_lane)
        {
            const bool rowsL = warp >= 1 && warp <= mw, colsU = warp > mw && warp <= 2 * mw, inv = warp == 7 + which;
            if (rowsL || colsU || inv) {
                const bool useU = rowsL || (inv && which);   // coefficients U_D (with the reciprocal pivots) or L_D^T
                float* base;
                int stride;
                if (rowsL) { base = S + (o + SB + (warp - 1) * 32 + lane) + o * LDS; stride = LDS; }
                else if (colsU) { base = S + o + (o + SB + (warp - 1 - mw) * 32 + lane) * LDS; stride = 1; }
                else { base = Xh + o + (o + lane) * LDS; stride = 1; }
                substitute32(useU ? s_ut : s_lt, s_rd, useU,
                             [&](int i) { return inv ? (i == lane ? 1.f : 0.f) : base[i * stride]; },
                             [&](int i, float v) { base[i * stride] = v; });
            }
        }
        if (m == 0) break;
        __syncthreads();
        DBG_CLK();
        // ---- P3: S22 -= L21 U12 on the tensor cores.  A(r, k) = -L21(r, o+k) for matrix rows r >= o+32 (tile row = r, so
        // the TMEM lane of a result is its matrix row), B(n, k) = U12(o+k, o+32+n); D = TMEM columns [0, m)
#define SUB_CLK(i) do { if (dbg_clk && threadIdx.x == 0 && kb == 1) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); dbg_clk[40 + (i)] = t_; } } while (0)
        SUB_CLK(0);
        stage_tiles(TcDst{slabs32(stage + TC_A0, stage + TC_A2), m, o + SB}, TcSrc{S + (o + SB) + o * LDS, 1, LDS, -1.f, 0},
                    TcDst{slabs32(stage + TC_B0, stage + TC_B2), m, 0}, TcSrc{S + o + (o + SB) * LDS, LDS, 1, 1.f, 0}, 4, tid);
        SUB_CLK(1);
        ptx::fence_proxy_async();
        __syncthreads();
        SUB_CLK(2);
        if (threadIdx.x == 0) {
            ptx::tc_fence_after();
            tc_product<2, true, false>(tc.tmem, td, 0, m);
            ptx::umma_commit<1>(tc.bar);
        }
        SUB_CLK(3);
        ptx::mbar_wait(tc.bar, tc.phase);
        tc.phase ^= 1;
        ptx::tc_fence_after();
        SUB_CLK(4);
#define SUB3_CLK(i) do { if (dbg_clk && threadIdx.x == 96 && kb == 1) { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); dbg_clk[46 + (i)] = t_; } } while (0)
        SUB3_CLK(0);
        {   // S22 += D: warp w reads TMEM lane quarter w % 4 (matrix rows), 8 of every 32 columns
            const int q = warp & 3, part = warp >> 2;
            if (q > kb) {
                float* dst = S + (q * 32 + lane) + (o + SB + part * 8) * LDS;
#pragma unroll 1
                for (int cc = 0; cc < mw; ++cc) {
                    uint32_t d[8];
                    tmem_ld_32x8(tc.tmem + ((uint32_t)(q * 32) << 16) + cc * 32 + part * 8, d);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[(cc * 32 + j) * LDS] += __uint_as_float(d[j]);
                }
            }
        }
        SUB3_CLK(1);
        ptx::tc_fence_before();
        __syncthreads();
        SUB3_CLK(2);
        SUB_CLK(5);
        DBG_CLK();
    }
    __syncthreads();
    DBG_CLK();
    // The L\U block is final: CTA 0 sends it off now, so that its 64 KiB of stores drain under the merges below instead of
    // in front of the step barrier's release (measured: the drain, not the issue, is what a leaf waits for while the bulk
    // lane keeps L2 busy).
    if (which == 0) {
        const int r = tid & (DB - 1), cq = tid >> 7;
#pragma unroll 4
        for (int i = 0; i < DB / 4; ++i) Wb[r + (long long)(cq + 4 * i) * ldw] = S[r + (cq + 4 * i) * LDS];
    }

    // ---- I2: off-diagonal blocks of this CTA's inverse by block-recursive doubling (32 -> 64 -> 128) on the tensor
    // cores.  M(r, k) = the factor read as lower triangular: L11(r, k), or U11^T(r, k) = U11(k, r).
    {
        const int mrs = which ? LDS : 1, mks = which ? 1 : LDS;  // M(r, k) = S[r * mrs + k * mks]
        const int q = warp & 3, part = warp >> 2;
        const TcSlabs a32 = slabs32(stage + TC_A0, stage + TC_A2), b32 = slabs32(stage + TC_B0, stage + TC_B2);
        const TcSlabs a64 = slabs3(stage + TC_A0), b64 = slabs3(stage + TC_B0);
        // level 1, both 64x64 problems (d = 0, 64) at once.  T_d = M21_d X11_d: matrix rows d+32 .. d+63 (= tile rows), K = N = 32
        stage_tiles(TcDst{a32, SB, SB}, TcSrc{S + SB * mrs, mrs, mks, 1.f, 0},
                    TcDst{a32, SB, 64 + SB}, TcSrc{S + (64 + SB) * mrs + 64 * mks, mrs, mks, 1.f, 0}, 4, tid);
        stage_tiles(TcDst{b32, SB, 0}, TcSrc{Xh, LDS, 1, 1.f, 0},                        // B(n, k) = X11_0(k, n)
                    TcDst{b32, SB, SB}, TcSrc{Xh + 64 + 64 * LDS, LDS, 1, 1.f, 0}, 4, tid);  // rows 32..63: X11_64
        LEAF_TC_ISSUE_BEGIN()
            tc_product<2, true, false>(tc.tmem + 0, td, 0, 32);      // rows 32..63 are T_0
            tc_product<2, true, false>(tc.tmem + 32, td, 4096, 32);  // rows 96..127 are T_64
        LEAF_TC_ISSUE_END(tc)
        // X21_d = -X22_d T_d.  T comes back through TMEM lane quarters 1 and 3, 8 columns per warp = one 16-byte chunk of
        // the MN-major B slabs (row = k); the MMAs that read the slabs have completed (the wait above)
        if (q == 1 || q == 3) {
            const int so = q == 3 ? 4096 : 0;
            uint32_t d[8];
            tmem_ld_32x8(tc.tmem + ((uint32_t)(q * 32) << 16) + (q == 3 ? 32 : 0) + part * 8, d);
            ptx::tmem_ld_wait();
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(d[j]);
            stage_chunk(slabs3(stage + TC_B0, so), lane, part, v);
        }
        stage_tiles(TcDst{a32, SB, SB}, TcSrc{Xh + SB + SB * LDS, 1, LDS, -1.f, 0},                          // A(r, k) = -X22_0
                    TcDst{a32, SB, 64 + SB}, TcSrc{Xh + (64 + SB) + (64 + SB) * LDS, 1, LDS, -1.f, 0}, 4, tid);  //           -X22_64
        ptx::tc_fence_before();
        LEAF_TC_ISSUE_BEGIN()
            tc_product<2, true, true>(tc.tmem + 64, td, 0, 32);
            tc_product<2, true, true>(tc.tmem + 96, td, 4096, 32);
        LEAF_TC_ISSUE_END(tc)
        if (q == 1 || q == 3) {
            const int d0 = q == 3 ? 64 : 0, r = q * 32 + lane;
            uint32_t d[8];
            tmem_ld_32x8(tc.tmem + ((uint32_t)(q * 32) << 16) + (q == 3 ? 96 : 64) + part * 8, d);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) Xh[r + (d0 + part * 8 + j) * LDS] = __uint_as_float(d[j]);
        }
        ptx::tc_fence_before();
        __syncthreads();  // X21 of both 64x64 blocks is in Xh
        // level 2: T = M21 X11 (matrix rows 64..127 as tile rows 0..63, K = N = 64); the upper-right 32x32 blocks of the
        // 64x64 triangular operands were never written and are read as zero
        stage_tiles(TcDst{a64, 64, 0}, TcSrc{S + 64 * mrs, mrs, mks, 1.f, 0},
                    TcDst{b64, 64, 0}, TcSrc{Xh, LDS, 1, 1.f, 2}, 8, tid);  // B(n, k) = X11(k, n): zero for k < 32 <= n
        LEAF_TC_ISSUE_BEGIN()
            tc_product<4, false, false>(tc.tmem + 0, td, 0, 64);
        LEAF_TC_ISSUE_END(tc)
        if (q < 2) {  // T row k = TMEM lane = tile row; 2 x 8 of its 64 columns per warp -> chunks of the MN-major B slabs
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t d[8];
                tmem_ld_32x8(tc.tmem + ((uint32_t)(q * 32) << 16) + hh * 32 + part * 8, d);
                ptx::tmem_ld_wait();
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(d[j]);
                stage_chunk(b64, q * 32 + lane, hh * 4 + part, v);
            }
        }
        stage_tiles(TcDst{a64, 64, 0}, TcSrc{Xh + 64 + 64 * LDS, 1, LDS, -1.f, 1}, TcDst{a64, 0, 0}, TcSrc{Xh, 1, 1, 1.f, 0}, 8, tid);  // A(i, k) = -X22(i, k): zero for i < 32 <= k
        ptx::tc_fence_before();
        LEAF_TC_ISSUE_BEGIN()
            tc_product<4, false, true>(tc.tmem + 64, td, 0, 64);
        LEAF_TC_ISSUE_END(tc)
        if (q < 2) {
            const int i = q * 32 + lane;
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t d[8];
                tmem_ld_32x8(tc.tmem + ((uint32_t)(q * 32) << 16) + 64 + hh * 32 + part * 8, d);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) Xh[(64 + i) + (hh * 32 + part * 8 + j) * LDS] = __uint_as_float(d[j]);
            }
        }
        ptx::tc_fence_before();
        __syncthreads();
    }
    DBG_CLK();

    // ---- scales of the 16-bit inverses: one power-of-two pair per nb-wide diagonal TILE, chosen by the tile's first
    // 128-block from the magnitudes of its inverses with 2^8 of headroom; later blocks reuse it (as leaf.cuh)
    float sI = 1.f;
    if (first_in_tile) {
        float mI = 0.f;
#pragma unroll 1
        for (int idx = tid; idx < DB * DB; idx += DL_THREADS) {
            const int r = idx & (DB - 1), c = idx >> 7;
            if (r >= c && r < valid) mI = fmaxf(mI, fabsf(Xh[r + c * LDS]));
        }
        for (int o = 16; o > 0; o >>= 1) mI = fmaxf(mI, __shfl_xor_sync(FULL, mI, o));
        if (lane == 0) s_red[warp] = mI;
        __syncthreads();
        mI = 0.f;
        for (int i = 0; i < DL_THREADS / 32; ++i) mI = fmaxf(mI, s_red[i]);
        if (!bf16) {
            int e;
            if (mI > 0.f && isfinite(mI)) { frexpf(mI, &e); sI = ldexpf(1.f, 8 - e); }
        }
        if (tid == 0) {
            tile_scales[2 * which] = sI;
            tile_scales[2 * which + 1] = 1.f / sI;
        }
    } else if (!bf16) {
        sI = tile_scales[2 * which];
    }

    // ---- write back: the scaled 16-bit inverse into the band and its fp32 copy for the triangular solves (triangles
    // only: the other halves stay zero).  The L\U block itself left right after the elimination loop (see there).
    uint16_t* I16 = reinterpret_cast<uint16_t*>(which ? Uinv16 : Linv16);
    float* I32 = which ? Uinv32 : Linv32;
    if (I32) I32 += (long long)blk * DB * DB;
    float mx = 0.f;
    constexpr float PADMAX = 32768.f;
    {
        const int r = tid & (DB - 1), cq = tid >> 7;
#pragma unroll 4
        for (int i = 0; i < DB / 4; ++i) {
            const int c = cq + 4 * i;
            if (which == 0) {
                if (r >= c) {
                    const float xl = Xh[r + c * LDS];  // inv(L11)(r,c)
                    float v = xl * sI;
                    if (r < valid) mx = fmaxf(mx, fabsf(v)); else v = fminf(fmaxf(v, -PADMAX), PADMAX);
                    store16(I16, r + (long long)c * ld16, v, bf16);
                    if (I32) I32[r + c * DB] = xl;
                }
            } else if (r <= c) {
                const float zu = Xh[c + r * LDS];      // inv(U11)(r,c) = inv(U11^T)(c,r)
                float v = zu * sI;
                if (c < valid) mx = fmaxf(mx, fabsf(v)); else v = fminf(fmaxf(v, -PADMAX), PADMAX);
                store16(I16, r + (long long)c * ld16, v, bf16);
                if (I32) I32[r + c * DB] = zu;
            }
        }
    }
    if (status) {
        const float hmax = bf16 ? 3.0e38f : 65504.f;
        const bool bad = !(mx <= hmax);  // also true for NaN
        if (__any_sync(FULL, bad) && lane == 0) atomicOr(status, isfinite(mx) ? 1 : 4);
        if (tid == 0 && s_zero && which == 0) atomicOr(status, 2);
    }
    DBG_CLK();
#undef DBG_CLK
}

}  // namespace leaf
}  // namespace mplu

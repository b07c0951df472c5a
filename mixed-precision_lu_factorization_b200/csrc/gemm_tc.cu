// Trailing-update GEMM on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma kind::f16 issued by one thread
//   -> fp32 accumulators in TMEM (2 x 256 columns, double buffered) -> tcgen05.ld epilogue that applies
//   C <- C + alpha*acc in fp32 and emits the scaled 16-bit shadow the next panel / TRSM will consume.
//
// Replaces the rank-r cublasDgemm / cublasDtrsm pair of the reference (/root/reference/MPF.cu:215-239).
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM lanes 32*(w%4).., column half w/4), warp 8 TMA producer,
// warp 9 TMEM allocator + MMA issuer.  Persistent: each CTA (or CTA pair, cta_group::2) walks tiles t, t+G, t+2G, ...
#include "gemm_tc.h"

#include <cstdlib>
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <mutex>

namespace mplu {

namespace {

constexpr int BM = 128;  // rows per CTA (= TMEM lanes)
constexpr int BN = 256;  // UMMA N (accumulator columns per stage)
constexpr int BK = 64;   // K elements per smem stage (= one 128-byte swizzle row)
constexpr int UK = 16;   // K per tcgen05.mma for 16-bit inputs
constexpr int GROUP_M = 8;
constexpr int EPI_WARPS = 8;   // 2 column halves x 4 TMEM lane quarters
constexpr int NTHREADS = (EPI_WARPS + 2) * 32;

template <int kCG>
struct Cfg {
    static constexpr int LOAD_BN = BN / kCG;           // B columns each CTA of the pair loads
    static constexpr int A_BYTES = BM * BK * 2;        // 16 KiB
    static constexpr int B_BYTES = LOAD_BN * BK * 2;   // 32 / 16 KiB
    static constexpr int STAGES = (kCG == 1) ? 4 : 6;  // 192 KiB of operands either way
    static constexpr int BAR_OFF = STAGES * (A_BYTES + B_BYTES);
    static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;  // barriers + tmem ptr + alignment slack
};

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& mt, int& nt) {
    const int group_size = GROUP_M * num_n;
    const int g = t / group_size;
    const int first_m = g * GROUP_M;
    const int gm = min(GROUP_M, num_m - first_m);
    const int r = t - g * group_size;
    mt = first_m + r % gm;
    nt = r / gm;
}

// ---- epilogue helpers: one 32-column chunk of one row per thread, ONE call site each (software-pipelined chunk loop
// below).  The first version of this epilogue was 13.7k SASS instructions and instruction-fetch bound; the second
// kept the steady state compact but still carried 13.5k instructions of variants, so every small panel GEMM ran
// 25-35 us on a cold instruction cache (gpurun_out/launches_r01b_n16384.csv).  kRagged = per-element bounds checks
// (arbitrary M, N: the mplu_gemm16 hook); otherwise M % 32 == 0, N % 32 == 0, h_cols % 32 == 0 and a chunk is
// either entirely inside the matrix or entirely outside (warp-uniform test, no per-element predicates).
// kStream: the fp32 addend / result and the 16-bit shadow are touched once per launch and far larger than L2 (the bulk
// lane's tall rank-nb updates): load / store them with the streaming (evict-first) cache policy so that they do not
// push the operand panels -- and the chain lane's diagonal tile -- out of L2.
template <bool kRagged, bool kStream>
__device__ __forceinline__ void epi_load(float (&dst)[32], const GemmParams& p, int row, int col0, bool ok) {
    const float* src = p.Cin + row + (long long)col0 * p.ldcin;
    if constexpr (kRagged) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = (row < p.M && col0 + j < p.N) ? src[(long long)j * p.ldcin] : 0.f;
    } else {
        if (ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j] = kStream ? __ldcs(src + (long long)j * p.ldcin) : src[(long long)j * p.ldcin];
        }
    }
}

// addend of a tile's first update: the original fp64 matrix, cast on the fly (aligned path only: the host uses it for
// matrices whose order is a multiple of 128)
template <bool kStream>
__device__ __forceinline__ void epi_load64(float (&dst)[32], const GemmParams& p, const ARef& ar, int row, int col0, bool ok) {
    if (ok) {
        const double* src = ar.A + (p.cin64_r0 + row) + (long long)(p.cin64_c0 + col0) * ar.lda;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = static_cast<float>(kStream ? __ldcs(src + (long long)j * ar.lda) : src[(long long)j * ar.lda]);
    }
}

template <bool kRagged, bool kStream>
__device__ __forceinline__ void epi_store(const uint32_t (&v)[32], const float (&cin)[32], const GemmParams& p,
                                          float alpha, float hs, int row, int col0, bool ok, float& mx) {
    if (!kRagged && !ok) return;
    float out[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = fmaf(alpha, __uint_as_float(v[j]), cin[j]);
    if (p.C) {
        float* dst = p.C + row + (long long)col0 * p.ldc;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (!kRagged || (row < p.M && col0 + j < p.N)) {
                if constexpr (kStream) __stcs(dst + (long long)j * p.ldc, out[j]);
                else dst[(long long)j * p.ldc] = out[j];
            }
    }
    if (p.H) {
        // shadow region: every row of the columns < h_cols, and the rows < h_rows of every column
        const bool row_in = row < p.h_rows;
        if (kRagged || row_in || col0 < p.h_cols) {
            if (p.bf16) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.H) + row + (long long)col0 * p.ldh;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (!kRagged || (row < p.M && col0 + j < p.N && (row_in || col0 + j < p.h_cols))) {
                        const __nv_bfloat16 hv = __float2bfloat16_rn(out[j] * hs);
                        if constexpr (kStream) __stcs(reinterpret_cast<unsigned short*>(dst) + (long long)j * p.ldh, __bfloat16_as_ushort(hv));
                        else dst[(long long)j * p.ldh] = hv;
                    }
            } else {
                __half* dst = reinterpret_cast<__half*>(p.H) + row + (long long)col0 * p.ldh;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float hv = out[j] * hs;
                    if (!kRagged || (row < p.M && col0 + j < p.N && (row_in || col0 + j < p.h_cols))) {
                        if constexpr (kStream) __stcs(reinterpret_cast<unsigned short*>(dst) + (long long)j * p.ldh, __half_as_ushort(__float2half_rn(hv)));
                        else dst[(long long)j * p.ldh] = __float2half_rn(hv);
                        mx = fmaxf(mx, fabsf(hv));  // inf propagates; a NaN needs an inf operand, caught when it was made
                    }
                }
            }
        }
    }
}

template <int kCG, bool kAMN, bool kRagged, bool kStream = false>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmGroup G) {
    // Up to kMaxGroup independent problems share one launch ("grouped"): tiles [tile_end[i-1], tile_end[i]) belong to
    // problem i.  The panel code groups the L-side and U-side products of one recursion node (and the node's Schur
    // update with the first products of its inverse merge), which cuts the number of dependent launches on the
    // factorization's critical path from 7 to 3 per node.
    const GemmParams& p0 = G.p[0];
    auto problem_of = [&](int t) {
        int i = 0;
        while (i + 1 < G.count && t >= G.tile_end[i]) ++i;
        return i;
    };
    // k-block range [kb0, kb1) of a tile: everything, or the part where a triangular operand is non-zero
    auto k_range = [&](const GemmParams& p, int mt, int nt, int& kb0, int& kb1) {
        int k0 = 0, k1 = p.K;
        const int mlo = mt * BM * kCG, mhi = mlo + BM * kCG, nlo = nt * BN, nhi = nlo + BN;
        if (p.tri == TRI_A_LOWER) k1 = min(p.K, mhi);
        else if (p.tri == TRI_A_UPPER) k0 = min(mlo, p.K - BK);
        else if (p.tri == TRI_B_UPPER) k1 = min(p.K, nhi);
        else if (p.tri == TRI_B_LOWER) k0 = min(nlo, p.K - BK);
        kb0 = k0 / BK;
        kb1 = (k1 + BK - 1) / BK;
    };
    auto tile_of = [&](int t, int i, int& mt, int& nt) {
        const GemmParams& p = G.p[i];
        const int nm = (p.M + BM * kCG - 1) / (BM * kCG), nn = (p.N + BN - 1) / BN;
        tile_coords(t - (i ? G.tile_end[i - 1] : 0), nm, nn, mt, nt);
    };
    using C = Cfg<kCG>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

    uint8_t* sA = smem;
    uint8_t* sB = smem + C::STAGES * C::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    // CTA-pair kernels may be launched in clusters of several pairs (ranks 2i, 2i+1 = pair i; the pairs of one cluster work
    // on neighbouring tiles at the same time): cta_rank = rank inside the pair, pair_base = cluster rank of its leader
    uint32_t cta_rank = 0, pair_base = 0;
    if constexpr (kCG == 2) {
        const uint32_t cr = ptx::cluster_ctarank();
        cta_rank = cr & 1u;
        pair_base = cr & ~1u;
    }
    ptx::griddep_launch();  // the stream successor may start its prologue now (it blocks in its own griddep_wait)

    if (warp == EPI_WARPS && lane == 0) {
        for (int i = 0; i < G.count; ++i) {
            ptx::prefetch_tmap(&G.tmA[i]);
            ptx::prefetch_tmap(&G.tmB[i]);
        }
        for (int i = 0; i < C::STAGES; ++i) {
            ptx::mbar_init(&full[i], kCG);  // one producer arrival per CTA of the pair (leader's barrier is used)
            ptx::mbar_init(&empty[i], 1);   // one tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&tfull[i], 1);         // one tcgen05.commit
            ptx::mbar_init(&tempty[i], EPI_WARPS * kCG);  // one arrival per epilogue warp per CTA (leader's barrier)
        }
        ptx::fence_mbar_init();
    }
    if constexpr (kCG == 2) ptx::cluster_sync_all();  // both CTAs resident before the paired TMEM allocation
    if (warp == EPI_WARPS + 1) ptx::tmem_alloc<kCG>(tmem_slot, 512);
    ptx::tc_fence_before();
    if constexpr (kCG == 2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    // nothing above touched global memory: with a programmatic launch the prologue overlapped the predecessor
    ptx::griddep_wait();

    const int num_tiles = G.tile_end[G.count - 1];
    const int first_tile = blockIdx.x / kCG;
    const int tile_step = gridDim.x / kCG;

    if (warp == EPI_WARPS) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step) {
                const int pi = problem_of(t);
                const GemmParams& p = G.p[pi];
                const CUtensorMap& tmA = G.tmA[pi];
                const CUtensorMap& tmB = G.tmB[pi];
                int mt, nt;
                tile_of(t, pi, mt, nt);
                const int m0 = mt * BM * kCG + cta_rank * BM;
                const int n0 = nt * BN + cta_rank * C::LOAD_BN;
                int kb0, kb1;
                k_range(p, mt, nt, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    if constexpr (kCG == 1) {
                        ptx::mbar_arrive_expect_tx(&full[stage], C::A_BYTES + C::B_BYTES);
                    } else {
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * (C::A_BYTES + C::B_BYTES));
                        else ptx::mbar_arrive_cluster(&full[stage], pair_base);
                    }
                    uint8_t* a_dst = sA + stage * C::A_BYTES;
                    uint8_t* b_dst = sB + stage * C::B_BYTES;
                    const int k0 = kb * BK;
                    if constexpr (kCG == 1) {
                        if constexpr (kAMN) {
                            ptx::tma_load_2d(a_dst, &tmA, &full[stage], p.a_r0 + m0, p.a_c0 + k0);
                            ptx::tma_load_2d(a_dst + 8192, &tmA, &full[stage], p.a_r0 + m0 + 64, p.a_c0 + k0);
                        } else {
                            ptx::tma_load_2d(a_dst, &tmA, &full[stage], p.a_r0 + k0, p.a_c0 + m0);
                        }
                        ptx::tma_load_2d(b_dst, &tmB, &full[stage], p.b_r0 + k0, p.b_c0 + n0);
                    } else {
                        if constexpr (kAMN) {
                            ptx::tma_load_2d_pair(a_dst, &tmA, &full[stage], p.a_r0 + m0, p.a_c0 + k0);
                            ptx::tma_load_2d_pair(a_dst + 8192, &tmA, &full[stage], p.a_r0 + m0 + 64, p.a_c0 + k0);
                        } else {
                            ptx::tma_load_2d_pair(a_dst, &tmA, &full[stage], p.a_r0 + k0, p.a_c0 + m0);
                        }
                        ptx::tma_load_2d_pair(b_dst, &tmB, &full[stage], p.b_r0 + k0, p.b_c0 + n0);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == EPI_WARPS + 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA of the pair only)
        if (cta_rank == 0) {
            const uint32_t idesc = make_idesc_f16(BM * kCG, BN, p0.bf16 != 0, kAMN, false);
            uint32_t stage = 0, phase = 0, iter = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step, ++iter) {
                int kb0, kb1;
                {
                    const int pi = problem_of(t);
                    int mt, nt;
                    tile_of(t, pi, mt, nt);
                    k_range(G.p[pi], mt, nt, kb0, kb1);
                }
                const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_base = ptx::smem_u32(sA + stage * C::A_BYTES);
                        const uint32_t b_base = ptx::smem_u32(sB + stage * C::B_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k) {
                            uint64_t adesc, bdesc;
                            if constexpr (kAMN) {
                                // [k][64 m] rows of 128 B; 8-row swizzle atoms: next atom along K +1024 B,
                                // next 64-row slab along M +8192 B; 16 k-rows per MMA = 2048 B
                                adesc = ptx::make_smem_desc_sw128(a_base + k * 2048, 8192, 1024);
                            } else {
                                // [m][64 k] rows of 128 B; 8-row atoms along M +1024 B; 16 k = 32 B inside the row
                                adesc = ptx::make_smem_desc_sw128(a_base + k * 32, 0, 1024);
                            }
                            bdesc = ptx::make_smem_desc_sw128(b_base + k * 32, 0, 1024);
                            ptx::umma_f16<kCG>(d_tmem, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
                        }
                        ptx::umma_commit<kCG>(&empty[stage], pair_base);
                        if (kb == kb1 - 1) ptx::umma_commit<kCG>(&tfull[as], pair_base);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue warps 0..7
        // warp w: TMEM lanes 32*(w%4).. (rows), columns [128*(w/4), +128) of the tile in 4 chunks of 32.
        // Software pipeline over the flat chunk sequence of all tiles of this CTA: iteration i issues the addend
        // loads of chunk i (also across tiles, i.e. while the MMAs of that tile are still running) and then
        // finishes chunk i-1, so each warp keeps 32 x 128 B of loads in flight behind its stores.
        const uint32_t q = warp & 3, half = warp >> 2;
        float mx = 0.f;

        float cin_cur[32], cin_nxt[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { cin_cur[j] = 0.f; cin_nxt[j] = 0.f; }
        const int my_tiles = first_tile < num_tiles ? (num_tiles - first_tile + tile_step - 1) / tile_step : 0;
        const int nchunks = my_tiles * 4;
        int t_row0 = 0, t_colbase = 0;              // tile of the chunk being loaded
        int cur_row0 = 0, cur_col0 = 0;             // chunk being finished
        bool cur_ok = false;
        const GemmParams* lp = &p0;                 // problem of the chunk being loaded / finished
        const GemmParams* cp = &p0;
        ARef l_aref{nullptr, 0};
        float l_alpha = 0.f, l_hs = 0.f, c_alpha = 0.f, c_hs = 0.f;
#pragma unroll 1
        for (int i = 0; i <= nchunks; ++i) {
            int nx_col0 = 0;
            bool nx_ok = false;
            if (i < nchunks) {
                if ((i & 3) == 0) {
                    const int t = first_tile + (i >> 2) * tile_step;
                    int mt, nt;
                    const int pi = problem_of(t);
                    lp = &G.p[pi];
                    tile_of(t, pi, mt, nt);
                    t_row0 = mt * BM * kCG + cta_rank * BM + q * 32;
                    t_colbase = nt * BN + half * 128;
                    l_alpha = lp->alpha;
                    if (lp->alpha_p1) l_alpha *= __ldg(lp->alpha_p1);
                    if (lp->alpha_p2) l_alpha *= __ldg(lp->alpha_p2);
                    l_hs = lp->hscale;
                    if (lp->hscale_p) l_hs *= __ldg(lp->hscale_p);
                    if (lp->cin64) l_aref = *lp->cin64;
                }
                nx_col0 = t_colbase + (i & 3) * 32;
                nx_ok = (t_row0 < lp->M) && (nx_col0 < lp->N);
            }
            uint32_t v[32];
            if (i > 0) {
                const uint32_t c = (i - 1) & 3, it = (i - 1) >> 2;
                const uint32_t as = it & 1, aphase = (it >> 1) & 1;
                if (c == 0) {
                    ptx::mbar_wait(&tfull[as], aphase);
                    ptx::tc_fence_after();
                }
                ptx::tmem_ld_32x32(tmem_base + ((q * 32u) << 16) + as * BN + half * 128 + c * 32, v);
            }
            if (i < nchunks) {
                if (lp->Cin != nullptr) {
                    epi_load<kRagged, kStream>(cin_nxt, *lp, t_row0 + lane, nx_col0, nx_ok);
                } else if (!kRagged && lp->cin64 != nullptr) {
                    epi_load64<kStream>(cin_nxt, *lp, l_aref, t_row0 + lane, nx_col0, nx_ok);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) cin_nxt[j] = 0.f;
                }
            }
            if (i > 0) {
                ptx::tmem_ld_wait();
                if (((i - 1) & 3) == 3) {
                    // accumulator fully read: hand the TMEM stage back before the last stores
                    const uint32_t as = ((i - 1) >> 2) & 1;
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (kCG == 1) ptx::mbar_arrive(&tempty[as]);
                        else ptx::mbar_arrive_cluster(&tempty[as], pair_base);
                    }
                }
                epi_store<kRagged, kStream>(v, cin_cur, *cp, c_alpha, c_hs, cur_row0 + lane, cur_col0, cur_ok, mx);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) cin_cur[j] = cin_nxt[j];
            cur_row0 = t_row0; cur_col0 = nx_col0; cur_ok = nx_ok;
            cp = lp; c_alpha = l_alpha; c_hs = l_hs;
        }
        const float hmax = p0.bf16 ? 3.0e38f : 65504.f;
        if (p0.status && __any_sync(0xffffffffu, mx > hmax) && lane == 0) atomicOr(p0.status, 1);
    }

    // ---------------------------------------------------------------- teardown
    ptx::tc_fence_before();
    if constexpr (kCG == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == EPI_WARPS + 1) ptx::tmem_dealloc<kCG>(tmem_base, 512);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(f);
    });
    return fn;
}

template <int kCG, bool kAMN, bool kRagged, bool kStream = false>
int launch_variant(GemmGroup& g, int max_sms, cudaStream_t stream) {
    using C = Cfg<kCG>;
    auto kern = gemm_tc_kernel<kCG, kAMN, kRagged, kStream>;
    long long tiles = 0;
    for (int i = 0; i < g.count; ++i) {
        tiles += (long long)((g.p[i].M + BM * kCG - 1) / (BM * kCG)) * ((g.p[i].N + BN - 1) / BN);
        g.tile_end[i] = (int)tiles;
    }
    long long want = tiles * kCG;
    int grid = (int)(want < max_sms ? want : max_sms);
    grid -= grid % kCG;
    if (grid < kCG) grid = kCG;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    // CTA pairs per cluster (experiment, MPLU_GEMM_PAIRS_PER_CLUSTER = 1 / 2 / 4): pairs of one cluster take consecutive tiles
    // = the same B columns and neighbouring A rows at the same time
    int pairs = 1;
    if (kCG == 2) {
        static const int env_pairs = [] { const char* e = getenv("MPLU_GEMM_PAIRS_PER_CLUSTER"); return e ? atoi(e) : 1; }();
        pairs = (env_pairs == 2 || env_pairs == 4) ? env_pairs : 1;
        while (pairs > 1 && grid % (kCG * pairs) != 0) grid -= kCG;  // whole clusters only
        if (grid < kCG * pairs) { pairs = 1; grid = grid < kCG ? kCG : grid; }
        cfg.gridDim = dim3(grid);
    }
    attrs[0].val.clusterDim.x = kCG * pairs;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = g.p[0].pdl ? 2 : 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, g);
}

template <int kCG, bool kAMN, bool kRagged, bool kStream = false>
int set_smem_attr() {
    return (int)cudaFuncSetAttribute(gemm_tc_kernel<kCG, kAMN, kRagged, kStream>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg<kCG>::SMEM_BYTES);
}

}  // namespace

int gemm_tc_init() {
    int rc = 0;
    if ((rc = set_smem_attr<1, true, false>())) return rc;
    if ((rc = set_smem_attr<2, true, false>())) return rc;
    if ((rc = set_smem_attr<1, false, false>())) return rc;
    if ((rc = set_smem_attr<2, false, false>())) return rc;
    if ((rc = set_smem_attr<1, true, false, true>())) return rc;
    if ((rc = set_smem_attr<2, true, false, true>())) return rc;
    if ((rc = set_smem_attr<1, true, true>())) return rc;
    if ((rc = set_smem_attr<2, true, true>())) return rc;
    if ((rc = set_smem_attr<1, false, true>())) return rc;
    if ((rc = set_smem_attr<2, false, true>())) return rc;
    return get_encode_fn() ? 0 : -1;
}

int make_tmap_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    uint32_t box_cols) {
    EncodeFn enc = get_encode_fn();
    if (!enc) return -1;
    cuuint64_t dims[2] = {rows, cols};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_rows, box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

int make_tmap_plain(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols) {
    EncodeFn enc = get_encode_fn();
    if (!enc || (elem_bytes != 2 && elem_bytes != 4)) return -1;
    cuuint64_t dims[2] = {rows, cols};
    cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {box_rows, box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

void gemm_box_shapes(int variant, uint32_t* a_box_rows, uint32_t* a_box_cols, uint32_t* b_box_rows,
                     uint32_t* b_box_cols) {
    const bool cg2 = (variant == GEMM_CG2_AMN || variant == GEMM_CG2_AK);
    const bool amn = (variant == GEMM_CG1_AMN || variant == GEMM_CG2_AMN);
    if (amn) { *a_box_rows = 64; *a_box_cols = BK; }  // 64 m (contiguous) x 64 k, two boxes per stage
    else     { *a_box_rows = BK; *a_box_cols = BM; }  // 64 k (contiguous) x 128 m
    *b_box_rows = BK;
    *b_box_cols = cg2 ? BN / 2 : BN;
}

int launch_gemm_group(int variant, GemmGroup& g, int max_sms, cudaStream_t stream) {
    if (g.count <= 0 || g.count > kMaxGroup) return (int)cudaErrorInvalidValue;
    bool ragged = false;
    for (int i = 0; i < g.count; ++i) {
        const GemmParams& p = g.p[i];
        if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.K % BK != 0 || p.bf16 != g.p[0].bf16) return (int)cudaErrorInvalidValue;
        ragged = ragged || (p.M % 32) || (p.N % 32) || (p.H && (p.h_cols % 32) && p.h_cols < p.N);
    }
    if (max_sms <= 0) {  // the current device's SM count (an attribute query, cheap; never cached across devices)
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // aligned fast path: whole 32 x 32 epilogue chunks (the factorization only ever issues these)
    if (!ragged) {
        bool stream_c = true;  // only when every problem of the group asks for it
        for (int i = 0; i < g.count; ++i) stream_c = stream_c && g.p[i].stream_c;
        if (stream_c && variant == GEMM_CG1_AMN) return launch_variant<1, true, false, true>(g, max_sms, stream);
        if (stream_c && variant == GEMM_CG2_AMN) return launch_variant<2, true, false, true>(g, max_sms, stream);
        switch (variant) {
            case GEMM_CG1_AMN: return launch_variant<1, true, false>(g, max_sms, stream);
            case GEMM_CG2_AMN: return launch_variant<2, true, false>(g, max_sms, stream);
            case GEMM_CG1_AK: return launch_variant<1, false, false>(g, max_sms, stream);
            case GEMM_CG2_AK: return launch_variant<2, false, false>(g, max_sms, stream);
            default: return (int)cudaErrorInvalidValue;
        }
    }
    switch (variant) {
        case GEMM_CG1_AMN: return launch_variant<1, true, true>(g, max_sms, stream);
        case GEMM_CG2_AMN: return launch_variant<2, true, true>(g, max_sms, stream);
        case GEMM_CG1_AK: return launch_variant<1, false, true>(g, max_sms, stream);
        case GEMM_CG2_AK: return launch_variant<2, false, true>(g, max_sms, stream);
        default: return (int)cudaErrorInvalidValue;
    }
}

int launch_gemm_tc2(int variant, const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmParams& p,
                    const CUtensorMap* tmA1, const CUtensorMap* tmB1, const GemmParams* p1_or_null, int max_sms,
                    cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0) return 0;
    GemmGroup g;
    g.count = 1;
    g.tmA[0] = *tmA; g.tmB[0] = *tmB; g.p[0] = p;
    if (p1_or_null && p1_or_null->M > 0 && p1_or_null->N > 0) {
        g.count = 2;
        g.tmA[1] = *tmA1; g.tmB[1] = *tmB1; g.p[1] = *p1_or_null;
    }
    return launch_gemm_group(variant, g, max_sms, stream);
}

int launch_gemm_tc(int variant, const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmParams& p, int max_sms,
                   cudaStream_t stream) {
    return launch_gemm_tc2(variant, tmA, tmB, p, nullptr, nullptr, nullptr, max_sms, stream);
}

}  // namespace mplu

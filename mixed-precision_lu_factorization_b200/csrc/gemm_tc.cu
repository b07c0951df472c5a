// Trailing-update GEMM on the 5th-generation tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma kind::f16 issued by one thread
//   -> fp32 accumulators in TMEM (2 x 256 columns, double buffered) -> tcgen05.ld epilogue that applies
//   C <- C + alpha*acc in fp32 (optionally taking the addend from the original fp64 matrix on first touch) and
//   emits the scaled 16-bit shadow the next panel / TRSM will consume.
//
// Replaces the rank-r cublasDgemm / cublasDtrsm pair of the reference (/root/reference/MPF.cu:215-239).
//
// Warp roles (320 threads): warps 0-7 epilogue (TMEM lanes 32*(w%4).., column half w/4), warp 8 TMA producer,
// warp 9 TMEM allocator + MMA issuer.  Persistent: each CTA (or CTA pair, cta_group::2) walks tiles t, t+G, t+2G, ...
#include "gemm_tc.h"
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

// ---- epilogue helpers: one 32-column chunk of one row per thread.  kFull = the whole 32x32 patch is in range, so
// the hot path carries no per-element predicates or branches (the first version of this epilogue was 13.7k SASS
// instructions and instruction-fetch bound: profiles/r01_gemm_epilogue_v1.txt).
template <bool kFull>
__device__ __forceinline__ void epi_load(float (&dst)[32], const GemmParams& p, int row, int col0) {
    if (p.Cin64) {
        const double* src = p.Cin64 + row + (long long)col0 * p.ldc64;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            dst[j] = (kFull || (row < p.M && col0 + j < p.N)) ? static_cast<float>(__ldg(src + (long long)j * p.ldc64)) : 0.f;
    } else {
        const float* src = p.Cin + row + (long long)col0 * p.ldcin;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            dst[j] = (kFull || (row < p.M && col0 + j < p.N)) ? src[(long long)j * p.ldcin] : 0.f;
    }
}

template <bool kFull>
__device__ __forceinline__ void epi_store(const uint32_t (&v)[32], const float (&cin)[32], const GemmParams& p,
                                          float alpha, float hs, float hmax, int row, int wrow0, int col0, bool& ovf) {
    float out[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = fmaf(alpha, __uint_as_float(v[j]), cin[j]);
    if (p.C) {
        float* dst = p.C + row + (long long)col0 * p.ldc;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (kFull || (row < p.M && col0 + j < p.N)) dst[(long long)j * p.ldc] = out[j];
    }
    if (p.H && (wrow0 < p.h_rows || col0 < p.h_cols)) {  // warp-uniform reject of chunks outside the shadow region
        const bool row_in = row < p.h_rows;
        if (p.bf16) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.H) + row + (long long)col0 * p.ldh;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float hv = out[j] * hs;
                if ((kFull || (row < p.M && col0 + j < p.N)) && (row_in || col0 + j < p.h_cols))
                    dst[(long long)j * p.ldh] = __float2bfloat16_rn(hv);
            }
        } else {
            __half* dst = reinterpret_cast<__half*>(p.H) + row + (long long)col0 * p.ldh;
            float mx = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float hv = out[j] * hs;
                if ((kFull || (row < p.M && col0 + j < p.N)) && (row_in || col0 + j < p.h_cols)) {
                    dst[(long long)j * p.ldh] = __float2half_rn(hv);
                    mx = fmaxf(mx, fabsf(hv));
                    ovf |= (hv != hv);
                }
            }
            ovf |= (mx > hmax);
        }
    }
}

template <int kCG, bool kAMN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
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
    uint32_t cta_rank = 0;
    if constexpr (kCG == 2) cta_rank = ptx::cluster_ctarank();

    if (warp == EPI_WARPS && lane == 0) {
        ptx::prefetch_tmap(&tmA);
        ptx::prefetch_tmap(&tmB);
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

    const int num_m = (p.M + BM * kCG - 1) / (BM * kCG);
    const int num_n = (p.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_kb = p.K / BK;
    const int first_tile = blockIdx.x / kCG;
    const int tile_step = gridDim.x / kCG;

    if (warp == EPI_WARPS) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step) {
                int mt, nt;
                tile_coords(t, num_m, num_n, mt, nt);
                const int m0 = mt * BM * kCG + cta_rank * BM;
                const int n0 = nt * BN + cta_rank * C::LOAD_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1);
                    if constexpr (kCG == 1) {
                        ptx::mbar_arrive_expect_tx(&full[stage], C::A_BYTES + C::B_BYTES);
                    } else {
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * (C::A_BYTES + C::B_BYTES));
                        else ptx::mbar_arrive_cluster(&full[stage], 0);
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
            const uint32_t idesc = make_idesc_f16(BM * kCG, BN, p.bf16 != 0, kAMN, false);
            uint32_t stage = 0, phase = 0, iter = 0;
            for (int t = first_tile; t < num_tiles; t += tile_step, ++iter) {
                const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
                ptx::mbar_wait(&tempty[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
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
                            ptx::umma_f16<kCG>(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit<kCG>(&empty[stage]);
                        if (kb == num_kb - 1) ptx::umma_commit<kCG>(&tfull[as]);
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
        // The addend C is software-prefetched one chunk ahead (also across tiles, i.e. while the MMAs of this
        // tile are still running) so that each warp keeps 2 x 32 x 128 B of loads in flight.
        float alpha = p.alpha;
        if (p.alpha_p1) alpha *= __ldg(p.alpha_p1);
        if (p.alpha_p2) alpha *= __ldg(p.alpha_p2);
        float hs = p.hscale;
        if (p.hscale_p) hs *= __ldg(p.hscale_p);
        const float hmax = p.bf16 ? 3.0e38f : 65504.f;
        const uint32_t q = warp & 3, half = warp >> 2;
        const bool has_cin = (p.Cin != nullptr) || (p.Cin64 != nullptr);
        bool ovf = false;

        float cinA[32], cinB[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { cinA[j] = 0.f; cinB[j] = 0.f; }
        uint32_t iter = 0;
        int t = first_tile;
        int mt = 0, nt = 0;
        if (t < num_tiles) {
            tile_coords(t, num_m, num_n, mt, nt);
            if (has_cin) {
                const int wrow0 = mt * BM * kCG + cta_rank * BM + q * 32, c0 = nt * BN + half * 128;
                if (wrow0 + 32 <= p.M && c0 + 32 <= p.N) epi_load<true>(cinA, p, wrow0 + lane, c0);
                else epi_load<false>(cinA, p, wrow0 + lane, c0);
            }
        }
        for (; t < num_tiles; t += tile_step, ++iter) {
            const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
            const int wrow0 = mt * BM * kCG + cta_rank * BM + q * 32;
            const int row = wrow0 + lane;
            const int colbase = nt * BN + half * 128;
            const bool full = (wrow0 + 32 <= p.M) && (colbase + 128 <= p.N);  // warp-uniform
            // next tile (for the cross-tile prefetch)
            const int tn = t + tile_step;
            int mtn = 0, ntn = 0;
            if (tn < num_tiles) tile_coords(tn, num_m, num_n, mtn, ntn);

            ptx::mbar_wait(&tfull[as], aphase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * BN + half * 128;
            uint32_t v[32];
#pragma unroll 1
            for (int c = 0; c < 128; c += 64) {
                // chunk c (addend in cinA); prefetch chunk c+32 into cinB
                ptx::tmem_ld_32x32(taddr + c, v);
                if (has_cin) {
                    if (full) epi_load<true>(cinB, p, row, colbase + c + 32);
                    else epi_load<false>(cinB, p, row, colbase + c + 32);
                }
                ptx::tmem_ld_wait();
                if (full) epi_store<true>(v, cinA, p, alpha, hs, hmax, row, wrow0, colbase + c, ovf);
                else epi_store<false>(v, cinA, p, alpha, hs, hmax, row, wrow0, colbase + c, ovf);
                // chunk c+32 (cinB); prefetch chunk c+64 -- or the next tile's first chunk -- into cinA
                ptx::tmem_ld_32x32(taddr + c + 32, v);
                if (has_cin) {
                    if (c == 0) {
                        if (full) epi_load<true>(cinA, p, row, colbase + 64);
                        else epi_load<false>(cinA, p, row, colbase + 64);
                    } else if (tn < num_tiles) {
                        const int wrow0n = mtn * BM * kCG + cta_rank * BM + q * 32, c0n = ntn * BN + half * 128;
                        if (wrow0n + 32 <= p.M && c0n + 32 <= p.N) epi_load<true>(cinA, p, wrow0n + lane, c0n);
                        else epi_load<false>(cinA, p, wrow0n + lane, c0n);
                    }
                }
                ptx::tmem_ld_wait();
                if (c == 64) {
                    // accumulator fully read: hand the TMEM stage back before the last stores
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (kCG == 1) ptx::mbar_arrive(&tempty[as]);
                        else ptx::mbar_arrive_cluster(&tempty[as], 0);
                    }
                }
                if (full) epi_store<true>(v, cinB, p, alpha, hs, hmax, row, wrow0, colbase + c + 32, ovf);
                else epi_store<false>(v, cinB, p, alpha, hs, hmax, row, wrow0, colbase + c + 32, ovf);
            }
            mt = mtn; nt = ntn;
        }
        if (p.status && __any_sync(0xffffffffu, ovf) && lane == 0) atomicOr(p.status, 1);
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

template <int kCG, bool kAMN>
int launch_variant(const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmParams& p, int max_sms,
                   cudaStream_t stream) {
    using C = Cfg<kCG>;
    auto kern = gemm_tc_kernel<kCG, kAMN>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_done = true;
    }
    const int num_m = (p.M + BM * kCG - 1) / (BM * kCG);
    const int num_n = (p.N + BN - 1) / BN;
    long long want = (long long)num_m * num_n * kCG;
    int grid = (int)(want < max_sms ? want : max_sms);
    grid -= grid % kCG;
    if (grid < kCG) grid = kCG;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = kCG;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kern, *tmA, *tmB, p);
}

}  // namespace

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

void gemm_box_shapes(int variant, uint32_t* a_box_rows, uint32_t* a_box_cols, uint32_t* b_box_rows,
                     uint32_t* b_box_cols) {
    const bool cg2 = (variant == GEMM_CG2_AMN || variant == GEMM_CG2_AK);
    const bool amn = (variant == GEMM_CG1_AMN || variant == GEMM_CG2_AMN);
    if (amn) { *a_box_rows = 64; *a_box_cols = BK; }  // 64 m (contiguous) x 64 k, two boxes per stage
    else     { *a_box_rows = BK; *a_box_cols = BM; }  // 64 k (contiguous) x 128 m
    *b_box_rows = BK;
    *b_box_cols = cg2 ? BN / 2 : BN;
}

int launch_gemm_tc(int variant, const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmParams& p, int max_sms,
                   cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0) return 0;
    if (p.K <= 0 || p.K % BK != 0) return (int)cudaErrorInvalidValue;
    if (max_sms <= 0) {
        static int sms = 0;
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        }
        max_sms = sms;
    }
    switch (variant) {
        case GEMM_CG1_AMN: return launch_variant<1, true>(tmA, tmB, p, max_sms, stream);
        case GEMM_CG2_AMN: return launch_variant<2, true>(tmA, tmB, p, max_sms, stream);
        case GEMM_CG1_AK: return launch_variant<1, false>(tmA, tmB, p, max_sms, stream);
        case GEMM_CG2_AK: return launch_variant<2, false>(tmA, tmB, p, max_sms, stream);
        default: return (int)cudaErrorInvalidValue;
    }
}

}  // namespace mplu

// Host-side interface of the tcgen05 trailing-update GEMM (see gemm_tc.cu).
//
//   out(m,n) = cin(m,n) + alpha * sum_k A(m,k) * B(k,n)          m < M, n < N, K % 64 == 0
//
// A, B are 16-bit (fp16 or bf16) column-major matrices described by TMA tensor maps that cover the WHOLE parent
// array; the sub-block a GEMM works on is selected with element origins (a_r0,a_c0)/(b_r0,b_c0).  This is the
// Schur-complement update A22 -= L21*U12 of the reference (cublasDgemm at /root/reference/MPF.cu:230-239) and, with
// an explicitly inverted triangular block as one operand, its TRSM (cublasDtrsm at MPF.cu:215-225).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mplu {

enum GemmTri : int { TRI_NONE = 0, TRI_A_LOWER = 1, TRI_A_UPPER = 2, TRI_B_UPPER = 3, TRI_B_LOWER = 4 };

enum GemmVariant : int {
    GEMM_CG1_AMN = 0,  // 1-CTA tiles 128x256, A column-major (M contiguous)
    GEMM_CG2_AMN = 1,  // CTA-pair tiles 256x256, A column-major
    GEMM_CG1_AK = 2,   // 1-CTA, A given as its transpose: parent array is K x M column-major (K contiguous)
    GEMM_CG2_AK = 3,   // CTA-pair, same A layout; for these two (a_r0,a_c0) = (k0, m0) in the transposed parent
};

// Where the caller's ORIGINAL fp64 matrix lives (device memory, written before every factorization): lets a captured
// schedule take the addend of a tile's FIRST update straight from A -- the fp64 -> fp32 cast fused into the update's
// loads (reference: the cast pass of /root/reference/MPF.cu:20-25,106-121) -- without baking A's address into the graph.
struct ARef {
    const double* A;
    long long lda;
};

struct GemmParams {
    int M, N, K;
    int a_r0, a_c0;  // origin of the A block inside its parent array: (row m0, col k0)
    int b_r0, b_c0;  // origin of the B block inside its parent array: (row k0, col n0)
    float* C;        // fp32 output, column-major, already offset to the block origin (may be null)
    long long ldc;
    const float* Cin;  // fp32 addend (may alias C; null = none)
    long long ldcin;
    const ARef* cin64;  // instead of Cin: addend = (float)A(cin64_r0 + m, cin64_c0 + n) of the original fp64 matrix
    int cin64_r0, cin64_c0;
    void* H;  // optional 16-bit shadow of the output, column-major, offset to the block origin
    long long ldh;
    int h_rows, h_cols;     // shadow is written where (m < h_rows || n < h_cols)
    float alpha;            // static factor ...
    const float* alpha_p1;  // ... times *alpha_p1 (device, null = 1)
    const float* alpha_p2;  // ... times *alpha_p2 (device, null = 1)
    float hscale;  // shadow = cvt16(out * hscale * *hscale_p)
    const float* hscale_p;
    int bf16;     // 0 = fp16 operands/shadow, 1 = bf16
    int* status;  // device word; bit 0 set when a shadow value overflowed the 16-bit range
    int pdl;      // 1 = launch with programmatic stream serialization (prologue overlaps the predecessor's tail)
    int stream_c; // 1 = C / Cin / H are accessed with the streaming (evict-first) cache policy
    int tri;      // triangular operand: the K range of a tile is cut to the part where that operand is non-zero
                  // (TRI_NONE, TRI_A_LOWER: A(m,k)=0 for k>m, TRI_A_UPPER: k<m, TRI_B_UPPER: B(k,n)=0 for k>n, TRI_B_LOWER: k<n)
};

// Build a 2-D TMA map (SWIZZLE_128B, 16-bit elements) over a column-major parent array with `rows` x `cols`
// elements and leading dimension `ld`; a box is box_rows (contiguous) x box_cols.
int make_tmap_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    uint32_t box_cols);

// The same without swizzle, 2- or 4-byte elements (shared -> global tile stores: the tile sits densely in shared memory).
int make_tmap_plain(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols);

// Box shapes each variant expects for its A and B maps.
void gemm_box_shapes(int variant, uint32_t* a_box_rows, uint32_t* a_box_cols, uint32_t* b_box_rows,
                     uint32_t* b_box_cols);

// Opt every kernel instantiation into its dynamic shared-memory size (once per device, before any launch / capture).
int gemm_tc_init();

// Launch on `stream` using at most `max_sms` SMs (0 = all).  Returns cudaError_t as int.
int launch_gemm_tc(int variant, const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmParams& p, int max_sms,
                   cudaStream_t stream);

// A launch carries up to kMaxGroup independent problems of the same variant and element type; the tiles of problem i+1
// follow those of problem i in the persistent tile order.
constexpr int kMaxGroup = 4;
struct alignas(64) GemmGroup {
    CUtensorMap tmA[kMaxGroup];
    CUtensorMap tmB[kMaxGroup];
    GemmParams p[kMaxGroup];
    int tile_end[kMaxGroup];  // prefix sums of the problems' tile counts (filled by the launcher)
    int count;
};
int launch_gemm_group(int variant, GemmGroup& g, int max_sms, cudaStream_t stream);

// Grouped launch: a second, independent problem (same variant and element type) rides in the same kernel; its tiles
// follow the first problem's.  p1 == nullptr (or empty) degenerates to launch_gemm_tc.
int launch_gemm_tc2(int variant, const CUtensorMap* tmA, const CUtensorMap* tmB, const GemmParams& p,
                    const CUtensorMap* tmA1, const CUtensorMap* tmB1, const GemmParams* p1, int max_sms,
                    cudaStream_t stream);

}  // namespace mplu

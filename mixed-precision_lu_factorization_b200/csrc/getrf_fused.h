// Fused GETRF of a diagonal block: ONE persistent launch runs the whole inverse-carrying recursion of lu.cu's
// Sched::getrf -- the 128x128 leaves (leaf.cuh) and every tcgen05 product between them -- as a step program, with a
// grid-wide barrier in global memory between dependent steps instead of a kernel boundary.  See getrf_fused.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mplu {

// the 16-bit arrays a product can take its operands from (one TMA map each as A operand and as B operand)
enum FusedMap : int { FM_WH = 0, FM_FH = 1, FM_LINV = 2, FM_UINV = 3, FM_T1 = 4, FM_T2 = 5, FM_COUNT = 6 };

// out(m,n) = (accumulate ? W(c_r0+m, c_c0+n) : 0) + alpha * sum_k A(m,k) B(k,n); same meaning as GemmParams (gemm_tc.h).
// M, N multiples of 128, K a multiple of 64.  The fp32 result goes to the working matrix W at (c_r0, c_c0), the 16-bit
// copy (scaled by *hscale_p) to array h_map at (h_r0, h_c0); both are written as whole 128x128 tiles (TMA stores).
struct alignas(16) FusedProblem {
    int M, N, K;
    int a_map, a_r0, a_c0;  // A block origin inside its parent array: (row m0, col k0)
    int b_map, b_r0, b_c0;  // B block origin: (row k0, col n0)
    int tri;                // GemmTri
    int accumulate;
    float alpha;            // times *alpha_p1 times *alpha_p2 (null = 1)
    int c_r0, c_c0;         // c_r0 < 0: no fp32 result
    int h_map, h_r0, h_c0;  // h_map < 0: no 16-bit result
    int pad_;
    const float* alpha_p1;
    const float* alpha_p2;
    const float* hscale_p;
};
static_assert(sizeof(FusedProblem) == 96, "FusedProblem layout");

enum FusedStepKind : int { FS_GEMM = 0, FS_LEAF = 1 };
struct alignas(16) FusedStep {
    int kind;
    int first_problem, num_problems;  // FS_GEMM: up to 4 independent products
    int tile_end[4];                  //          prefix sums of their 128x128 output tile counts
    int k0, blk, first_in_tile, valid, T;  // FS_LEAF: block origin, 128-block index, first leaf of tile T?, valid rows
};
static_assert(sizeof(FusedStep) == 48, "FusedStep layout");

struct alignas(64) FusedMaps {
    CUtensorMap a[FM_COUNT];  // operand loads, SWIZZLE_128B: boxes of 64 (rows, contiguous) x 64
    CUtensorMap b[FM_COUNT];  //                              boxes of 64 x 128
    CUtensorMap h[FM_COUNT];  // 16-bit result stores, unswizzled boxes of 128 x 128
    CUtensorMap c;            // fp32 result stores into W, unswizzled boxes of 128 x 128
};

struct FusedArgs {
    const void* program;  // device: num_steps FusedStep followed by num_problems FusedProblem
    int num_steps, num_problems;
    unsigned* barrier;    // device word, zero before the launch
    float* W; long long ldw;
    void* Linv16; void* Uinv16; long long ld16;  // band origins (row 0 of the tile's band, column 0 of the matrix)
    float* Linv32; float* Uinv32;
    float* inv_scales;
    int bf16;
    int* status;
    long long* dbg_clk;  // development aid: num_steps + 3 time stamps (null: none)
};

int getrf_fused_init();  // per-device kernel attributes
// grid of `num_ctas` (even, >= 2) CTAs in clusters of two, all of which must be able to be resident at the same time
int launch_getrf_fused(const FusedMaps& maps, const FusedArgs& args, int num_ctas, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------------
// Dataflow GETRF (getrf_flow.cu): the same diagonal block factored RIGHT-looking at 128-block granularity by one
// persistent launch WITHOUT grid barriers.  CTAs 0/1 run leaf after leaf; every other CTA is a helper that pulls 128x128
// tile products (panel solves with a leaf's inverses, rank-128 Schur updates, inverse merges) from one priority-ordered
// task list through an atomic queue head.  Dependencies are explicit: a task waits until up to four counters in global
// memory have reached their targets and bumps up to three counters once its result tile is globally visible.
struct alignas(16) FlowTask {
    uint16_t problem;      // index into the problems; 0xFFFF = stop marker
    uint8_t mt, nt;        // output tile of that problem
    uint8_t kb0, kb1;      // 64-wide k-blocks [kb0, kb1) (triangular operands: the non-zero part)
    uint16_t wait_ctr[4];  // 0xFFFF = unused
    uint16_t wait_val[4];
    uint16_t sig_ctr[3];   // 0xFFFF = unused
    uint16_t pad_[2];
};
static_assert(sizeof(FlowTask) == 32, "FlowTask layout");

struct alignas(16) FlowLeaf {
    int k0, blk, first_in_tile, valid, T;
    uint16_t wait_ctr, wait_val;  // the block has received all of its updates (0xFFFF: no wait)
    uint16_t sig_ctr, pad_;       // bumped by BOTH leaf CTAs when L\U and the block inverses are visible (target 2)
    int pad2_;
};
static_assert(sizeof(FlowLeaf) == 32, "FlowLeaf layout");

struct FlowArgs {
    const FlowLeaf* leaves; int num_leaves;
    const FusedProblem* problems; int num_problems;
    const FlowTask* tasks; int num_tasks;
    int num_main;         // tasks [0, num_main): leaf-to-leaf chain and updates; [num_main, num_tasks): inverse merges
    int merge_ctas;       // the last merge_ctas helper CTAs take merge tasks first, the others main tasks first
    unsigned* counters;   // device words, zero before the launch; [0] / [1] = queue heads of the two lists
    float* W; long long ldw;
    void* Linv16; void* Uinv16; long long ld16;
    float* Linv32; float* Uinv32;
    float* inv_scales;
    int bf16;
    int* status;
    long long* dbg;       // development aid: per leaf (start, end) and per task (dependencies met, signalled) in %globaltimer ns
};

int getrf_flow_init();  // per-device kernel attributes
// grid of `num_ctas` (even, >= 4) CTAs in clusters of two, all of which must be able to be resident at the same time
int launch_getrf_flow(const FusedMaps& maps, const FlowArgs& args, int num_ctas, cudaStream_t st);

}  // namespace mplu

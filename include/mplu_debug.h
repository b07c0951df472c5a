/* mplu_debug.h -- development / measurement aids exported by libmplu.so next to the product interface (include/mplu.h).
 * Nothing here is needed to use the solver; tools/ and tests/ call these through ctypes to time kernels in isolation,
 * read per-phase clock stamps and check intermediate results.  They are kept in the shipped library on purpose: the
 * numbers under profiles/ were produced by exactly the binary that is benchmarked. */
#ifndef MPLU_DEBUG_H
#define MPLU_DEBUG_H

#include "mplu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* mplu_diag_lu128 with clock64() stamps of the stand-alone leaf's phases written to d_clocks[0..31] */
int mplu_diag_lu128_timed(float *dW, long long ldw, float *dLinv, float *dUinv, long long *d_clocks, void *stream);

/* `reps` dependent launches of one GEMM shape captured into a CUDA graph and replayed: average device time per launch (us) */
int mplu_bench_gemm_chain(int variant, int M, int N, int K, int reps, int pdl, int shadow, int accumulate, int max_sms,
                          float *us_per_launch);

/* timeline marks of the factorization schedule (events around the lanes' phases): enable (takes effect at the next
 * schedule capture), then read (tag, ms since the first mark) pairs after a synchronised factorization */
void mplu_debug_marks_enable(mplu_context *ctx, int on);
int mplu_debug_timeline(mplu_context *ctx, int *tags, float *ms, int max);

/* per-step clock stamps of the fused GETRF launches (csrc/getrf_fused.cu): on = 1 step heads, 2 also sub-step stamps;
 * _profile: (kind, tiles, K | leaf origin, clock64 at the step's head) records of fused launch `launch`, then the phase
 * stamps of one of its leaves and one (-1, 0, ns of the launch, clock64 at the end) record; _raw: every stamp slot */
int mplu_debug_fused_profile_enable(mplu_context *ctx, int on);
int mplu_debug_fused_profile(mplu_context *ctx, int launch, long long *out, int max_records);
int mplu_debug_fused_raw(mplu_context *ctx, int launch, long long *out, int max_slots);

/* %globaltimer stamps of the `launch`-th dataflow GETRF launch (csrc/getrf_flow.cu) of the following factorizations
 * (launch < 0: off; takes effect at the next schedule capture).  _profile, after a synchronised factorization:
 * stamps[0 .. 2 L) = (start, end) ns of the L leaves, then four per task of the list (taken from the queue, dependencies
 * met, result signalled, CTA); tasks_out receives the 32-byte FlowTask records (kind / step in the padding).  Returns the
 * task count */
int mplu_debug_flow_profile_enable(mplu_context *ctx, int launch);
int mplu_debug_flow_profile(mplu_context *ctx, long long *stamps, int max_stamps, unsigned char *tasks_out, int max_task_bytes,
                            int *num_leaves);

/* the fp32 inverses of the unit-lower / upper factor of diagonal 128-block `blk` (column-major 128 x 128 each, device
 * or host destination), as the triangular solves use them */
int mplu_debug_block_inverses(mplu_context *ctx, int blk, float *Linv, float *Uinv);

#ifdef __cplusplus
}
#endif
#endif /* MPLU_DEBUG_H */

/*
 * diinn_b200_debug.h -- debug taps (bit-exact tests) and hardware probes of libdiinn_b200.so. Not part of the product
 * surface: nothing in the decode path calls these, and a reference-side binding does not need them.
 */
#ifndef DIINN_B200_DEBUG_H
#define DIINN_B200_DEBUG_H

#include "diinn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-axis nearest-exact source index and scaled relative coordinate, exactly the values
 * _make_pos_encoding (diinn.py:94-110) produces: ih[H_up], iw[W_up] int32; rel_h[H_up], rel_w[W_up] fp32.
 * (A stand-alone kernel built from the same device functions as the fused kernels; see diinn_debug_set_tap for the
 * values the fused stage-B kernel itself derives.) */
int diinn_debug_gather(diinn_handle* h, int H, int W, int H_up, int W_up, int32_t* ih, int32_t* iw,
                       float* rel_h, float* rel_w, void* stream);
/* Same for the query entry: idx[B*Q] = ih*W+iw, rel[B*Q*2], ratio[B*Q]. */
int diinn_debug_query_gather(diinn_handle* h, int B, int H, int W, const float* coord, const float* cell, int Q,
                             int32_t* idx, float* rel, float* ratio, void* stream);
/* Tap INSIDE the fused stage-B kernel: while set, every tensor-path diinn_decode also writes, for each output pixel,
 * (ih, iw, bits(rel_h), bits(rel_w)) as make_row() of csrc/stage_b_umma.cu derived them -- 4 x int32 at index
 * (pixel's channel-0 element offset in `out`), so for a contiguous (1,3,rows,W_up) band buffer entry r*W_up + c.
 * tap must hold (largest channel-0 offset + 1) x 4 int32. NULL switches the tap off. */
int diinn_debug_set_tap(diinn_handle* h, int32_t* tap);
/* LR-resolution hoisted pre-activations P (B*H*W, 1024) fp32 = [relu(K0 x) | K_i[:,256:] x + b_i, i=1..3]. */
int diinn_debug_stage_a(diinn_handle* h, const void* feat, int B, int C, int H, int W, float* P, void* workspace,
                        size_t workspace_bytes, int io_dtype, int compute, void* stream);
/* tcgen05 self-test: D(M x N fp32) = A(M x K bf16, row-major) * B(N x K bf16, row-major)^T through the same
 * TMA / UMMA-descriptor / TMEM plumbing the fused kernels use. M%128==0, N%256==0, K%64==0. cta_group 1|2;
 * 11|12 = the same with A, B holding fp16 and fp16 TMEM accumulators read back with tcgen05.ld.pack::16b. */
int diinn_debug_umma_gemm(diinn_handle* h, const void* A, const void* B, float* D, int M, int N, int K,
                          int cta_group, void* stream);
/* Tensor-pipe pace probe: `iters` back-to-back tcgen05.mma (M = 128*cta_group, N = n_cols, K = 16, bf16) on resident
 * shared-memory operands in n_ctas CTAs; cyc_per_mma[n_ctas / cta_group] (device) receives clock64 cycles per MMA.
 * noise: 16 extra warps per CTA hammer the idle TMEM half (bit 0), shared memory (bit 1) or the MUFU (bit 2) meanwhile. */
int diinn_debug_umma_pace(diinn_handle* h, int cta_group, int n_cols, int iters, int n_ctas, float* cyc_per_mma,
                          int noise, void* stream);
/* Host-only: which stage-B kernel a grid decode of rows [row0,row1) would launch on a device with sm_count SMs (no device, no
 * handle needed; compute = DIINN_COMPUTE_FP32 | _BF16 | _FP16). out12 = {select-MMA variant, phase table, canonical relative
 * coordinates, s_h, s_w, log2 of the patch width, K_sel, LR rows and columns of a pair's P box, work items, patch rows, CTA pairs
 * per patch row}. The first three must not depend on the row range (row tiles are bit-identical to the full decode); tested on
 * the CPU by tests/test_plan.py. */
int diinn_debug_plan_stage_b(int sm_count, int decoder_mode, int B, int H, int W, int H_up, int W_up, int row0, int row1,
                             int compute, int32_t* out12);
/* DIINN_TRACE=1 in the environment makes the fused stage-B kernel record clock64() at its pipeline events (leader
 * CTA of the first CTA pair, first 8 tiles); this copies the first n (<=1024) samples to host_out. */
int diinn_debug_read_trace(diinn_handle* h, int64_t* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* DIINN_B200_DEBUG_H */

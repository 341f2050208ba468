/*
 * diinn_b200.h -- C ABI of the B200-native (sm_100a) DIINN query decoder.
 *
 * One shared library (libdiinn_b200.so) replaces the hot path of the reference:
 *   ImplicitDecoder.forward / _make_pos_encoding / step (mode 3 -- the benchmarked wiring -- and modes 1, 2, 4;
 *   init_q=False), /root/reference/src/models/components/diinn.py:94-110 (coordinates), :163-173 (forward),
 *   :132-139 (dual-interactive K/Q MLP), :116-131,140-147 (the other wirings), reached from DIINN.forward (diinn.py:18),
 *   SRLitModule.forward (src/models/sr_module.py:104-105) and demo2.py:40.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every function returns DIINN_OK (0) or a negative diinn_status; nothing throws, nothing exits.
 *     diinn_last_error() returns a human-readable message for the last failure on that handle
 *     (or, with a NULL handle, of the calling thread's last failed diinn_create).
 *   - device pointers unless the name says "host"; the caller owns feat / out / workspace, the library owns
 *     the repacked weights and its TMA descriptors. No allocation happens inside decode/query once the
 *     handle has weights, so the calls are CUDA-graph capturable.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). All work is asynchronous on
 *     that stream except the *_host entry, which synchronises the stream before returning.
 *   - there is NO CPU fallback: a device that is not sm_100 yields DIINN_ERR_UNSUPPORTED_DEVICE.
 */
#ifndef DIINN_B200_H
#define DIINN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct diinn_handle diinn_handle;

typedef enum diinn_status {
  DIINN_OK = 0,
  DIINN_ERR_BAD_ARG = -1,
  DIINN_ERR_BAD_SHAPE = -2,
  DIINN_ERR_BAD_DTYPE = -3,
  DIINN_ERR_UNSUPPORTED_MODE = -4,   /* anything but mode in {1,2,3,4}, init_q in {0,1}, 64 channels, 4x256 hidden */
  DIINN_ERR_WORKSPACE_TOO_SMALL = -5,
  DIINN_ERR_CUDA = -6,
  DIINN_ERR_NO_WEIGHTS = -7,
  DIINN_ERR_UNSUPPORTED_DEVICE = -8
} diinn_status;

/* arithmetic path. Every tensor-core mode accumulates in fp32 (TMEM) and keeps P, biases, relu, sin, the coordinates and the
 * RGB projection in fp32; the modes differ in how activations, features and weights enter the MMAs.
 * Accuracy envelope (max-abs against the fp32 reference; tests/test_gpu_parity.py):
 *   default-init weights (|out| < 0.06)      FP32 <= 2e-6    FP16 ~3e-6    BF16 ~2e-5
 *   gain-scaled weights, O(1) activations    FP32 <= 1e-4    FP16 ~2e-3    BF16 ~1.6e-2  (K x3, Q x10: bf16 misses the 1e-2
 *   contract there, fp16 operands keep it -- which is why FP16 is the default 16-bit mode of the Python layer)
 * Sine arguments: MUFU.SIN's error grows as ~6e-8 |x|; DIINN_COMPUTE_FP32 range-reduces first (error ~5e-7 for |x| < 1e5).
 * Coordinates (diinn.py:94-110): gather indices are bit-exact in every mode; relative coordinates are the reference's fp32
 * values bit for bit, except that the 16-bit modes, on INTEGER scale factors with at most 16 phases (s_h * s_w <= 16), use the
 * exact value of the pixel's phase, (2p + 1)/s - 1 -- the number the reference's rounded coordinate grids scatter around by a
 * few ulps of a [-1, 1] coordinate times the axis length (5e-5 on a 339 x 510 map) -- so that Q.0's sines form a small table
 * (DESIGN.md section 4.1d). Effect on the image, fp32 oracle with either set of coordinates: < 1e-8 on default-init weights, up to
 * 8e-4 on the gain-scaled set at the 339 x 510 geometry (tests/test_canonical_coords.py) -- the reference's own sensitivity to
 * that rounding. DIINN_COMPUTE_FP32 / _FP32_SIMT never do this; DIINN_NO_CANON=1 in the environment turns it off. */
typedef enum diinn_compute {
  DIINN_COMPUTE_FP32 = 0, /* fp32 PRECISION on the tensor cores: every operand is an fp16 hi + lo pair (22 mantissa bits) and
                             every product three MMAs (hi.hi + lo.hi + hi.lo). init_q=True decodes fall back to _FP32_SIMT. */
  DIINN_COMPUTE_BF16 = 1, /* bf16 operands                                                                                 */
  DIINN_COMPUTE_FP16 = 2, /* fp16 operands (saturating conversions): same tensor rate as bf16, 8x less operand noise        */
  DIINN_COMPUTE_FP32_SIMT = 3 /* exact fp32 FMA on CUDA cores end to end (cross-check path, ~40x slower)                     */
} diinn_compute;

/* element type of the feat / out buffers. DIINN_IO_BF16_NHWC (the encoder hand-off): feat is bf16 in
 * channels-last memory order (B,H,W,64) -- what `x.to(torch.bfloat16, memory_format=torch.channels_last)` holds -- which IS
 * stage A's TMA layout, so the layout pass is skipped; out is bf16 NCHW as with DIINN_IO_BF16. Tensor paths only. */
typedef enum diinn_io_dtype { DIINN_IO_F32 = 0, DIINN_IO_BF16 = 1, DIINN_IO_BF16_NHWC = 2 } diinn_io_dtype;

/* Constructor arguments of the reference ImplicitDecoder (diinn.py:40). */
typedef struct diinn_config {
  int in_channels; /* 64 */
  int hidden;      /* 256 (hidden_dims = [256]*n_layers) */
  int n_layers;    /* 4 */
  int mode;        /* 3 (paper / benchmark wiring); 1 / 2: K chain fed by k instead of q (diinn.py:57-72,116-131);
                      4: mode 3 with a 3x3 reflect-padded last conv over the HR grid (diinn.py:73-90,140-147) */
  int init_q;      /* 0, or 1: sine gate on the unfolded features, first_layer + 576-wide Q.0 (diinn.py:48-51,113-115);
                      grid decode only (diinn_query* return DIINN_ERR_UNSUPPORTED_MODE) */
  int device;      /* CUDA device ordinal the handle lives on */
} diinn_config;

/* The 18 (init_q=True: 20) tensors of the reference state_dict, fp32, contiguous, reference layout:
 *   k_weight[0] (256,576)  k_weight[1..3] (256,832) [mode 1: (256,256)]  q_weight[0] (256,3)  q_weight[1..3] (256,256)
 *   *_bias (256)           last_weight (3,256) [mode 4: (3,256,3,3)]       last_bias (3)
 * on_device != 0: pointers are device pointers on the handle's device; otherwise host pointers. */
typedef struct diinn_weights_f32 {
  const float* k_weight[4];
  const float* k_bias[4];
  const float* q_weight[4];
  const float* q_bias[4];
  const float* last_weight;
  const float* last_bias;
  int on_device;
  /* init_q=True only (NULL otherwise): first_layer = Conv2d(3, 576, 1) of diinn.py:48-51, weight (576,3), bias (576);
   * q_weight[0] is then (256,576). */
  const float* first_weight;
  const float* first_bias;
} diinn_weights_f32;

/* ---- lifetime -------------------------------------------------------------------------------------- */
int diinn_create(diinn_handle** out, const diinn_config* cfg);
void diinn_destroy(diinn_handle* h);
const char* diinn_last_error(const diinn_handle* h);

/* Repack the reference-layout weights into the library's layouts (fp32 hoisted matrices for the CUDA-core
 * path; bf16 / fp16 / fp16 hi+lo, K-permuted, 128B-swizzle tiles for the tcgen05 paths). Replaces
 * ImplicitDecoder.__init__ / load_state_dict (diinn.py:40-92). */
int diinn_set_weights(diinn_handle* h, const diinn_weights_f32* w, void* stream);

/* ---- the hot path ---------------------------------------------------------------------------------- */
/* Scratch needed by diinn_decode for HR rows [row0,row1) of a (B,C,H,W)->(B,3,H_up,W_up) decode. */
size_t diinn_workspace_bytes(const diinn_handle* h, int B, int H, int W, int H_up, int W_up,
                             int row0, int row1, int compute);

/* ImplicitDecoder.forward(x, size, bsize) for HR rows [row0,row1) (diinn.py:163-173; bsize is pure scheduling,
 * diinn.py:149-160, and has no equivalent here because no per-pixel intermediate is ever materialised).
 *   feat : (B,C,H,W) contiguous NCHW, io_dtype
 *   out  : element (b,c,row,col) is written at out[b*out_batch_stride + c*out_chan_stride +
 *          (row-row0)*out_row_stride + col] (strides in elements, io_dtype); for a full contiguous
 *          (B,3,H_up,W_up) tensor pass out + row0*W_up, 3*H_up*W_up, H_up*W_up, W_up.
 * row0=0,row1=H_up decodes the whole image; row tiles are how the query grid shards across GPUs. */
int diinn_decode(diinn_handle* h, const void* feat, int B, int C, int H, int W, int H_up, int W_up,
                 int row0, int row1, void* out, int64_t out_batch_stride, int64_t out_chan_stride,
                 int64_t out_row_stride, void* workspace, size_t workspace_bytes, int io_dtype, int compute,
                 void* stream);

/* Fused decode + assembly across the GPUs of one NVLink/NVSwitch domain. Same as diinn_decode, but every pixel of rows
 * [row0,row1) is stored into n_peers peer-mapped image buffers at once (out_peers[i] = the address a plain decode would
 * get as `out` in rank i's buffer, i.e. already offset to row0; this GPU's own buffer is one of them), so when all ranks'
 * kernels have finished every rank holds the whole image and no collective is needed. If out_multicast is non-NULL
 * (an NVSwitch multicast mapping of the same buffers, fp32 only) one multimem.st per value replaces the n_peers stores.
 * The caller synchronises the ranks afterwards (e.g. a symmetric-memory barrier). */
int diinn_decode_multi(diinn_handle* h, const void* feat, int B, int C, int H, int W, int H_up, int W_up, int row0,
                       int row1, void* const* out_peers, int n_peers, void* out_multicast, int64_t out_batch_stride,
                       int64_t out_chan_stride, int64_t out_row_stride, void* workspace, size_t workspace_bytes,
                       int io_dtype, int compute, void* stream);

/* Same call with HOST buffers: copies feat H2D, decodes, copies the (B,3,row1-row0,W_up) band D2H into
 * out_host (contiguous), synchronises `stream`. Device scratch is owned and cached by the handle. This is the
 * call a CPU-side caller such as demo2.py:40 makes; bench.py's `e2e` times it. */
int diinn_decode_host(diinn_handle* h, const void* feat_host, int B, int C, int H, int W, int H_up, int W_up,
                      int row0, int row1, void* out_host, int io_dtype, int compute, void* stream);

/* Eval glue either side of the decode, fused into the store of the last epilogue:
 *   v = pred * scale + bias          (sr_module.py:123  `pred_hr * self.div + self.sub`, two rounded fp32 ops)
 *   v = clamp(v, lo, hi)             (sr_module.py:123  `.clamp_(0, 1)`)
 *   u8 = (uint8) clamp(v*255 + 0.5, 0, 255)   (torchvision save_image's quantisation, demo2.py:41)
 * With quantize_u8 the `out` buffers of decode / decode_multi / decode_host / query hold uint8 elements (strides stay in
 * elements), whatever io_dtype says about feat: 4x fewer bytes to assemble across GPUs or to copy back to the host.
 * Multicast stores (diinn_decode_multi with out_multicast) stay fp32-only. NULL restores the identity. */
typedef struct diinn_output_transform {
  int affine;       /* apply scale/bias */
  float scale, bias;
  int clamp;        /* apply lo/hi */
  float lo, hi;
  int quantize_u8;
} diinn_output_transform;
int diinn_set_output_transform(diinn_handle* h, const diinn_output_transform* t);

/* The reference's `bsize` argument (ImplicitDecoder.forward(x, size, bsize), diinn.py:163; 0 = None). For modes 1-3 it is
 * pure scheduling and ignored. In mode 4 it is part of the RESULT: batched_step (diinn.py:149-160) runs `step`, and with it
 * the 3x3 reflect-padded last conv, on column strips of bsize // H_up columns one at a time, so columns reflect at the
 * borders of their strip. diinn_decode reproduces that; like the reference it rejects a bsize that leaves a strip one
 * column wide (DIINN_ERR_BAD_SHAPE) and one below H_up (DIINN_ERR_BAD_ARG: the reference loops forever there). */
int diinn_set_bsize(diinn_handle* h, int64_t bsize);

/* PSNR as the reference evaluates it (calc_psnr, sr_module.py:21-38), computed on the device:
 *   dataset 0 = None (all pixels, all channels), 1 = 'benchmark' (shave = scale, Y conversion with
 *   (65.738, 129.057, 25.064)/256 when C > 1), 2 = 'div2k' (shave = scale + 6).
 * sr, hr: (B,C,H,W) contiguous device tensors of `dtype` (DIINN_IO_F32 / _BF16). Synchronises `stream`;
 * *psnr_host = -10 log10(mean(valid^2)). */
int diinn_psnr(diinn_handle* h, const void* sr, const void* hr, int dtype, int B, int C, int H, int W, int dataset,
               int scale, float rgb_range, double* psnr_host, void* stream);

/* Superset entry with the (feat, coord, cell) signature north_star names (LIIF.query_rgb's, liif.py:59):
 *   coord (B,Q,2) fp32 (h,w) in [-1,1], cell (B,Q,2) fp32, out (B,Q,3) io_dtype.
 * DIINN semantics per axis: idx = clamp(floor((c+1)*n/2)), rel = (c - centre[idx])*n,
 * ratio = cell_h*cell_w*H*W/4. On the regular HR grid it reproduces diinn_decode bit for bit. */
size_t diinn_query_workspace_bytes(const diinn_handle* h, int B, int H, int W, int Q, int compute);
int diinn_query(diinn_handle* h, const void* feat, int B, int C, int H, int W, const float* coord,
                const float* cell, int Q, void* out, void* workspace, size_t workspace_bytes, int io_dtype,
                int compute, void* stream);

/* "Next" row 1 of SURVEY.md section 8(f): the same query with LIIF's 4-neighbour local ensemble (LIIF.query_rgb,
 * liif.py:71-127): every query is decoded at its four shifted nearest LR cells (shift +-1/n + 1e-6, clamp, nearest
 * lookup) and the four RGB predictions are blended by the diagonally swapped areas |rel_h*rel_w| + 1e-9 inside the
 * last epilogue. Workspace: diinn_query_workspace_bytes(h, B, H, W, 4*Q, compute). */
int diinn_query_ensemble(diinn_handle* h, const void* feat, int B, int C, int H, int W, const float* coord,
                         const float* cell, int Q, void* out, void* workspace, size_t workspace_bytes, int io_dtype,
                         int compute, void* stream);

/* "Next" row 1 of SURVEY.md section 8(f), second half: LIIF-proper decoding. The handle (created with mode=3, init_q=0) takes
 * the weights of LIIF's own imnet = MLP(580, 3, [256]*4) (liif.py:19-26, mlp.py:5-15; state_dict keys imnet.layers.{0,2,4,6,8}):
 *   weight[0] (256,580)   weight[1..3] (256,256)   weight[4] (3,256)   bias[i] (256) / (3)      fp32, contiguous
 * whose 580 inputs are [unfolded features 576 | rel_coord 2 | rel_cell 2] (liif.py:105-111). Afterwards diinn_query is
 * LIIF.query_rgb(feat, coord, cell) with local_ensemble=False and diinn_query_ensemble the same with local_ensemble=True
 * (feat_unfold=True, cell_decode=True; liif.py:59-127): lookups by grid_sample(nearest) on the clamped coordinate, ReLU
 * layers, area blend. The feature part of the first Linear is evaluated once per LR cell by the same stage-A kernel, the
 * rest per query by the same stage-B kernel. diinn_decode* return DIINN_ERR_UNSUPPORTED_MODE on such a handle (LIIF.forward
 * queries the grid's own coordinates, liif.py:151-158); diinn_set_weights switches it back. Tensor paths only. */
typedef struct diinn_liif_weights_f32 {
  const float* weight[5];
  const float* bias[5];
  int on_device;
} diinn_liif_weights_f32;
int diinn_set_weights_liif(diinn_handle* h, const diinn_liif_weights_f32* w, void* stream);

/* Per-kernel device timing of the tcgen05 path, measured with CUDA events recorded on the caller's stream around
 * the three kernels of every diinn_decode (layout pass, stage A, stage B) while enabled. diinn_get_kernel_times
 * synchronises on the recorded events, returns the summed milliseconds and the number of decodes, and resets.
 * bench.py uses it for the roofline of the dominant kernel (stage B). Both calls, like diinn_decode_host, also read the
 * device-side consistency flag of the tcgen05 kernels and return DIINN_ERR_CUDA if one of them raised it. */
int diinn_set_profiling(diinn_handle* h, int enable);  /* enable: creates the event pool (2048 decodes per window) */
int diinn_get_kernel_times(diinn_handle* h, double* ms_layout, double* ms_stage_a, double* ms_stage_b,
                           int64_t* n_decodes);

/* Number of kernels this library launched on the handle since creation (bench.py's gpu_launches). */
int64_t diinn_launch_count(const diinn_handle* h);
const char* diinn_version(void);

/* Debug taps and hardware probes live in diinn_b200_debug.h (same library). */

#ifdef __cplusplus
}
#endif
#endif /* DIINN_B200_H */

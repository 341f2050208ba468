// Internal types shared by the .cu files of libdiinn_b200.so (not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/diinn_b200.h"
#include "../../include/diinn_b200_debug.h"

namespace diinn {

constexpr int kC = 64;          // encoder channels (in_channels, diinn.py:40)
constexpr int kUnfold = 576;    // 3x3 unfold of 64 channels (diinn.py:168)
constexpr int kD = 256;         // hidden width
constexpr int kLayers = 4;      // K/Q layer pairs
constexpr int kPCols = 1024;    // hoisted LR-resolution pre-activations per LR pixel: [k0 | kx1 | kx2 | kx3]

// ---------------------------------------------------------------------------------------------------------
// Coordinates (diinn.py:94-110). Per axis with n source and n_up destination samples the reference computes, in
// fp32 with separately rounded mul/add (no FMA):
//   c(i, n)  = fl(a_n + fl(b_n * i)),      a_n = fp32(-1 + 1/n), b_n = fp32(2/n)   (python doubles -> fp32)
//   idx(j)   = min(floorf(fl((j + 0.5) * scale)), n - 1),  scale = fp32(n) / fp32(n_up)   (nearest-exact)
//   rel(j)   = fl(fl(c(j, n_up) - c(idx(j), n)) * fp32(n))
// The host fills AxisParams with the rounded constants so host and device agree bit for bit.
// ---------------------------------------------------------------------------------------------------------
struct AxisParams {
  float a_in, b_in, a_up, b_up, scale, n_in_f;
  int n_in, n_up;
};

__host__ inline AxisParams make_axis(int n_in, int n_up) {
  AxisParams p;
  p.a_in = static_cast<float>(-1.0 + 1.0 / n_in);
  p.b_in = static_cast<float>(2.0 / n_in);
  p.a_up = static_cast<float>(-1.0 + 1.0 / n_up);
  p.b_up = static_cast<float>(2.0 / n_up);
  p.scale = static_cast<float>(n_in) / static_cast<float>(n_up);
  p.n_in_f = static_cast<float>(n_in);
  p.n_in = n_in;
  p.n_up = n_up;
  return p;
}

__device__ __forceinline__ int axis_index(const AxisParams& p, int j) {
  const float t = __fmul_rn(static_cast<float>(j) + 0.5f, p.scale);
  const int i = static_cast<int>(floorf(t));
  return min(i, p.n_in - 1);
}
__device__ __forceinline__ float axis_centre_in(const AxisParams& p, int i) {
  return __fadd_rn(p.a_in, __fmul_rn(p.b_in, static_cast<float>(i)));
}
__device__ __forceinline__ float axis_rel(const AxisParams& p, int j, int i) {
  const float cu = __fadd_rn(p.a_up, __fmul_rn(p.b_up, static_cast<float>(j)));
  return __fmul_rn(__fsub_rn(cu, axis_centre_in(p, i)), p.n_in_f);
}
// query entry: idx = clamp(floorf(fl(fl(c + 1) * fl(n/2)))), rel = fl(fl(c - centre) * n)
__device__ __forceinline__ int query_index(const AxisParams& p, float c) {
  const float t = __fmul_rn(__fadd_rn(c, 1.0f), p.n_in_f * 0.5f);
  const int i = static_cast<int>(floorf(t));
  return max(0, min(i, p.n_in - 1));
}
__device__ __forceinline__ float query_rel(const AxisParams& p, float c, int i) {
  return __fmul_rn(__fsub_rn(c, axis_centre_in(p, i)), p.n_in_f);
}

// local-ensemble lookup (LIIF.query_rgb, liif.py:88-104): shifted + clamped coordinate -> grid_sample(nearest,
// align_corners=False) index = nearbyint(((c_ + 1) * n - 1) / 2), every step rounded in fp32
__device__ __forceinline__ int ensemble_index(const AxisParams& p, float c, float shift, float lo, float hi) {
  const float cs = fminf(fmaxf(__fadd_rn(c, shift), lo), hi);
  const float t = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(cs, 1.0f), p.n_in_f), 1.0f), 0.5f);
  return max(0, min(__float2int_rn(t), p.n_in - 1));
}

// ---------------------------------------------------------------------------------------------------------
// How the fused kernels enumerate "HR query pixels". Two sources: the regular grid of forward(x, size) and the
// explicit (coord, cell) list of query().
// ---------------------------------------------------------------------------------------------------------
struct PixelSource {
  int mode;  // 0 = grid rows [row0,row1) x [0,W_up) per batch image, 1 = query list
  // grid
  AxisParams ax_h, ax_w;
  float ratio;
  int B, H, W, H_up, W_up, row0, row1;
  int lr_row0, lr_rows;  // P holds LR rows [lr_row0, lr_row0 + lr_rows) of every batch image
  // init_q=True (csrc/init_q.cu): the sine gate on x makes every K input a per-HR-pixel quantity, so P holds one row per
  // pixel of the launch: row = g - p_base in the fp32 path's linear pixel order, (b, row - row0, col) in stage B's. The
  // launch covers rows [row0,row1) of a band whose image offsets count from out_row0 (tensor path only).
  int per_pixel_p;
  long long p_base;
  int out_row0;
  // query
  const float* coord;  // (B,Q,2)
  const float* cell;   // (B,Q,2)
  int Q;
  float hw_f;  // fp32(H*W)
  // query list with LIIF's 4-neighbour local ensemble (diinn_query_ensemble): row g = 4*query + v, v = 2*(vx>0) + (vy>0);
  // the four rows of a query are blended by area in the last epilogue (liif.py:117-127)
  int ensemble;
  float sh_h[2], sh_w[2];  // fp32(v/n + 1e-6) for v = -1, +1
  float clamp_lo, clamp_hi;
  // query list decoded by LIIF's own imnet (diinn_set_weights_liif): lookups as LIIF.query_rgb does them with or without
  // the ensemble, rel_cell = cell * (H, W) instead of `ratio`
  int liif;
};

constexpr int kMaxPeers = 8;

struct OutSpec {
  void* ptr;
  int64_t batch_stride, chan_stride, row_stride;  // grid mode (elements); query mode: out[(b*Q+q)*3 + c]
  int io_dtype;
  // Fused assembly across the GPUs of one NVLink domain (diinn_decode_multi): every pixel is stored into the same
  // offset of n_peers peer-mapped image buffers (peers[] includes this GPU's own buffer), or once through an NVSwitch
  // multicast address (mc != nullptr, fp32 only). n_peers == 0: plain local store to ptr.
  void* peers[kMaxPeers];
  int n_peers;
  void* mc;
  // mode 4 (3x3 reflect-padded last conv over HR pixels, diinn.py:90): stage B does not project to RGB but dumps q_3 as
  // bf16, pixel-major (pixel offset = the channel-0 offset computed from batch_stride / row_stride, x 256); csrc/mode4.cu
  // then runs the convolution. nullptr everywhere else.
  __nv_bfloat16* q3;
  float* q3f;  // the same dump in fp32 (the fp32-precision tensor path, DIINN_COMPUTE_FP32); at most one of q3 / q3f is set
  // eval glue fused into the store (diinn_set_output_transform): bit 0 affine, bit 1 clamp, bit 2 uint8 quantisation
  int t_flags;
  float t_scale, t_bias, t_lo, t_hi;
};

__device__ __forceinline__ void store_elem(const OutSpec& o, void* base, int64_t off, float v) {
  if (o.t_flags & 4) {
    const float q = fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.f), 0.5f), 0.f), 255.f);
    static_cast<uint8_t*>(base)[off] = static_cast<uint8_t>(q);
  } else if (o.io_dtype == DIINN_IO_F32) {
    static_cast<float*>(base)[off] = v;
  } else {
    static_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v);
  }
}

// store one value of the output image at element offset `off` of every destination (after the optional eval glue)
__device__ __forceinline__ void store_out(const OutSpec& o, int64_t off, float v) {
  if (o.t_flags & 1) v = __fadd_rn(__fmul_rn(v, o.t_scale), o.t_bias);  // two rounded ops, as the reference's `x * div + sub`
  if (o.t_flags & 2) v = fminf(fmaxf(v, o.t_lo), o.t_hi);
  if (o.mc != nullptr) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(static_cast<float*>(o.mc) + off), "f"(v) : "memory");
  } else if (o.n_peers > 0) {
    for (int i = 0; i < o.n_peers; ++i) store_elem(o, o.peers[i], off, v);
  } else {
    store_elem(o, o.ptr, off, v);
  }
}

// small fp32 parameters every path needs in registers/constant bank (passed by value as kernel params)
struct SmallParams {
  float bq[kLayers][kD];   // Q biases (layer 0..3)
  float wq0[kD][4];        // per feature: Q.0 weight row + bias = (w_relh, w_relw, w_ratio, bq0)  -> one LDC.128
  float wl_t[kD][4];       // per feature: last_layer.weight column = (wl[0][f], wl[1][f], wl[2][f], 0)
  float bl[4];             // last_layer.bias (+pad)
  // wq0 again for feature PAIRS (f, f+1), component-major, for the packed fp32x2 (FFMA2) layer 0 of the tensor path:
  float2 wq0_p[kD / 2][4]; // (w_relh, w_relw, w_ratio, bq0) x (f, f+1)
  int q0_folded;           // stage B, grid decodes: wq0_p[.][3] already holds w_ratio * ratio + bq0 (the launcher folds it)
  int pad_[3];
};

// Optional fused epilogue of the library GEMM (gemm.cu) for the LR-resolution K chain of modes 1 / 2
// (lr_chain.cu): instead of D = A.B^T it does  P[m0 + r][256 layer + n] += acc  and hands the next layer its A operand,
// A_next[r][n] = bf16(relu(that)).
struct ChainEpilogue {
  float* P;                 // nullptr: plain GEMM, D is written
  __nv_bfloat16* A_next;    // nullptr for the last layer
  long long m0, rows;       // first P row of this chunk / valid rows in it
  int layer;
  int p16;                  // P holds fp16 rows (stage B's select-MMA variant): read-modify-write in fp16
};

struct Handle;  // defined in handle.h

}  // namespace diinn

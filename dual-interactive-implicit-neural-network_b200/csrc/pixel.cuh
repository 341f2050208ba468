// Per-pixel bookkeeping of the CUDA-core kernels (csrc/simt.cu, csrc/init_q.cu): which LR cell, which P row and which
// synthetic inputs (rel_h, rel_w, ratio) belong to the g-th HR query of a PixelSource.
#pragma once
#include "common.cuh"

namespace diinn {

// ---------------------------------------------------------------------------------------------------------
// per-pixel bookkeeping
// ---------------------------------------------------------------------------------------------------------
struct PixInfo {
  int p_idx;  // row of P
  float rel_h, rel_w, ratio;
  float area;  // ensemble rows only: |rel_h * rel_w| + 1e-9
  int b, oh, ow;  // grid: batch / HR row / HR col; query: b, q, unused
  int ih, iw;     // nearest LR cell
};

__device__ __forceinline__ PixInfo pixel_info(const PixelSource& s, int64_t g) {
  PixInfo pi;
  if (s.mode == 0) {
    const int nrows = s.row1 - s.row0;
    const int64_t per_img = static_cast<int64_t>(nrows) * s.W_up;
    const int b = static_cast<int>(g / per_img);
    const int rem = static_cast<int>(g - b * per_img);
    const int oh = s.row0 + rem / s.W_up;
    const int ow = rem % s.W_up;
    const int ih = axis_index(s.ax_h, oh), iw = axis_index(s.ax_w, ow);
    pi.p_idx = (b * s.lr_rows + (ih - s.lr_row0)) * s.W + iw;
    pi.rel_h = axis_rel(s.ax_h, oh, ih);
    pi.rel_w = axis_rel(s.ax_w, ow, iw);
    pi.ratio = s.ratio;
    pi.b = b;
    pi.oh = oh;
    pi.ow = ow;
    pi.ih = ih, pi.iw = iw;
  } else {
    const int64_t q = s.ensemble ? (g >> 2) : g;  // query index in (B*Q)
    const int v = static_cast<int>(g & 3);
    const int b = static_cast<int>(q / s.Q);
    const float ch = s.coord[q * 2], cw = s.coord[q * 2 + 1];
    int ih, iw;
    if (s.ensemble) {
      ih = ensemble_index(s.ax_h, ch, s.sh_h[v >> 1], s.clamp_lo, s.clamp_hi);
      iw = ensemble_index(s.ax_w, cw, s.sh_w[v & 1], s.clamp_lo, s.clamp_hi);
    } else {
      ih = query_index(s.ax_h, ch), iw = query_index(s.ax_w, cw);
    }
    pi.p_idx = (b * s.H + ih) * s.W + iw;
    pi.rel_h = query_rel(s.ax_h, ch, ih);
    pi.rel_w = query_rel(s.ax_w, cw, iw);
    pi.ratio = __fmul_rn(__fmul_rn(__fmul_rn(s.cell[q * 2], s.cell[q * 2 + 1]), s.hw_f), 0.25f);
    pi.area = __fadd_rn(fabsf(__fmul_rn(pi.rel_h, pi.rel_w)), 1e-9f);
    pi.b = b;
    pi.oh = static_cast<int>(q - static_cast<int64_t>(b) * s.Q);
    pi.ow = 0;
    pi.ih = ih, pi.iw = iw;
  }
  if (s.per_pixel_p) pi.p_idx = static_cast<int>(g - s.p_base);  // init_q=True: P has one row per pixel of the chunk
  return pi;
}

}  // namespace diinn

// Decoder mode 4 (diinn.py:81-90,140-147): the mode-3 dual-interactive stack, but `last_layer` is a 3x3 convolution
// with reflect padding over the HR pixel grid instead of a 1x1 projection -- the only cross-pixel coupling at HR
// resolution anywhere in the decoder. Stage B (either path) stops at q_3 and dumps it pixel-major, this kernel does
//   out[b, c, y, x] = bl[c] + sum_{ky,kx} sum_f Wl[c, f, ky, kx] * q_3[b, refl(y+ky-1), refl(x+kx-1), f]
// for the rows of the caller's band; the dump carries one halo row on each side (clipped at the image border, where the
// reflection folds back into the band). CUDA cores: 6 912 FMA per pixel against 786 432 tensor FLOP of stage B.
//
// bsize (diinn.py:149-160): the reference's batched_step runs `step` -- and with it this convolution and its reflect
// padding -- on column strips [ql, ql + bsize // H_up) one at a time, so in mode 4 (and only there) bsize changes the
// result: columns reflect at the borders of their strip, not of the image. `strip` reproduces that (strip = W_up: none).
#include "handle.h"

namespace diinn {

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// torch's 'reflect' padding by one element on [lo, hi): lo-1 -> lo+1, hi -> hi-2 (needs hi - lo >= 2)
__device__ __forceinline__ int reflect1(int i, int lo, int hi) { return i < lo ? 2 * lo - i : (i >= hi ? 2 * hi - 2 - i : i); }

constexpr int kConvRows = 4;  // output rows per thread: every weight load feeds 3 x 4 FMAs

// One thread per column x, kConvRows consecutive rows. q3: (B, qrows, W_up, 256) of T, holding HR rows [qr0, qr0 + qrows).
// wl4: (9 taps, 256, 4) fp32 = (Wl[0], Wl[1], Wl[2], 0) per (tap, feature): one address for the whole warp, i.e. one
// broadcast L1 load per 12 FMAs. Accumulation order per pixel: taps (ky, kx) row-major, features ascending, bias last.
template <typename T>
__global__ void __launch_bounds__(128) last_conv3x3_reflect_kernel(const T* __restrict__ q3, const float4* __restrict__ wl4,
                                                                  float b0, float b1, float b2, int H_up, int W_up, int strip,
                                                                  int qr0, int qrows, int row0, int row1, OutSpec out) {
  const int x_raw = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y0 = row0 + (blockIdx.y * 4 + (threadIdx.x >> 5)) * kConvRows;
  const int b = blockIdx.z;
  if (x_raw >= W_up || y0 >= row1) return;
  const int x = x_raw;
  const int s0 = (x / strip) * strip;                    // this column's strip [s0, s1)
  const int s1 = min(s0 + strip, W_up);
  float acc[kConvRows][3];
#pragma unroll
  for (int p = 0; p < kConvRows; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0.f;
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int xx = reflect1(x + dx, s0, s1);
    const T* src[kConvRows];
#pragma unroll
    for (int p = 0; p < kConvRows; ++p) {
      const int y = min(y0 + p, row1 - 1);                // rows past the band repeat its last row (not stored)
      const int yy = reflect1(y + dy, 0, H_up);
      src[p] = q3 + (static_cast<size_t>(b * qrows + (yy - qr0)) * W_up + xx) * kD;
    }
    const float4* w = wl4 + tap * kD;
#pragma unroll 2
    for (int f = 0; f < kD; f += 8) {
      float v[kConvRows][8];
#pragma unroll
      for (int p = 0; p < kConvRows; ++p) load8<T>(src[p] + f, v[p]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 wv = __ldg(w + f + i);
#pragma unroll
        for (int p = 0; p < kConvRows; ++p) {
          acc[p][0] = fmaf(wv.x, v[p][i], acc[p][0]);
          acc[p][1] = fmaf(wv.y, v[p][i], acc[p][1]);
          acc[p][2] = fmaf(wv.z, v[p][i], acc[p][2]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < kConvRows; ++p) {
    const int y = y0 + p;
    if (y >= row1) break;
    const int64_t off = b * out.batch_stride + static_cast<int64_t>(y - row0) * out.row_stride + x;
    store_out(out, off, acc[p][0] + b0);
    store_out(out, off + out.chan_stride, acc[p][1] + b1);
    store_out(out, off + 2 * out.chan_stride, acc[p][2] + b2);
  }
}

int launch_last_conv3x3(Handle* h, const void* q3, bool q3_is_f32, int B, int H_up, int W_up, int strip, int qr0, int qrows,
                        int row0, int row1, const OutSpec& out, cudaStream_t s) {
  dim3 grid((W_up + 31) / 32, (row1 - row0 + 4 * kConvRows - 1) / (4 * kConvRows), B);
  const float4* w = reinterpret_cast<const float4*>(h->WL4);
  const float b0 = h->small.bl[0], b1 = h->small.bl[1], b2 = h->small.bl[2];
  if (q3_is_f32)
    last_conv3x3_reflect_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(q3), w, b0, b1, b2, H_up, W_up, strip,
                                                            qr0, qrows, row0, row1, out);
  else
    last_conv3x3_reflect_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(q3), w, b0, b1, b2, H_up,
                                                                    W_up, strip, qr0, qrows, row0, row1, out);
  h->launches += 1;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

}  // namespace diinn

// Decoder mode 4 (diinn.py:81-90,140-147): the mode-3 dual-interactive stack, but `last_layer` is a 3x3 convolution
// with reflect padding over the HR pixel grid instead of a 1x1 projection -- the only cross-pixel coupling at HR
// resolution anywhere in the decoder. Stage B (either path) stops at q_3 and dumps it pixel-major, this kernel does
//   out[b, c, y, x] = bl[c] + sum_{ky,kx} sum_f Wl[c, f, ky, kx] * q_3[b, refl(y+ky-1), refl(x+kx-1), f]
// for the rows of the caller's band; the dump carries one halo row on each side (clipped at the image border, where the
// reflection folds back into the band).
//
// Tensor path (bf16 q_3 dump), two kernels. (1) last_conv_project_kernel: the convolution is linear, so every pixel is first
// projected onto all 27 (tap, channel) outputs, T[p] = W27 q_3[p] -- a (pixels x 256) x (256 x 27) GEMM whose N is far too
// small for a tcgen05 tile of its own, done with warp-level mma.sync.m16n8k16 (bf16 operands, fp32 accumulation; the fp32
// weights enter as a bf16 hi + lo pair in two N tiles, so they keep ~16 mantissa bits). It reads every q_3 row exactly
// once, coalesced: HBM-bound on the 512 B/px dump. (2) last_conv_gather_kernel: out[y,x,c] = bl[c] + sum_tap
// T[tap, c][refl(y+dy), refl(x+dx)] -- 27 coalesced plane reads per pixel. The fp32 path keeps the direct CUDA-core
// convolution below (last_conv3x3_reflect_kernel<float>), 6 912 exact fp32 FMA per pixel.
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

// ---------------------------------------------------------------------------------------------------------
// (1) T(27 planes x n_px fp32) = (q3(n_px x 256 bf16) . W27^T)^T, plane tap*3 + c. One warp per 32 pixels per iteration.
//
// mma.sync.m16n8k16 fragments (g = lane / 4, t = lane % 4): A regs (row g | g+8, k 2t..2t+1 | 2t+8..2t+9), B regs
// (k 2t..2t+1 | 2t+8..2t+9, n g), C regs (row g | g+8, col 2t, 2t+1). The sum over k is order-free, so inside a 64-feature
// block lane t takes features [8t, 8t+8) and [32+8t, 32+8t+8) of its rows -- two 16-byte loads, whose pairs (2s, 2s+1)
// serve as the low / high k pair of step s = 0..3; wfrag (pack.cu) holds B in exactly that permutation.
// N tiles 0..3 = bf16(W), 4..7 = bf16(W - bf16(W)) with identical column maps: hi + lo meet in the same accumulator slot.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128, 3) last_conv_project_kernel(const __nv_bfloat16* __restrict__ q3,
                                                               const uint2* __restrict__ wfrag, float* __restrict__ T,
                                                               int64_t n_px) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p0 = warp * 32; p0 < n_px; p0 += n_warps * 32) {
    float acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
    const __nv_bfloat16* rowp[2][2];  // [m tile][row g | g+8]; rows past the end repeat the last pixel (never stored)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int64_t p = p0 + mt * 16 + hr * 8 + g;
        rowp[mt][hr] = q3 + (p < n_px ? p : n_px - 1) * kD + t * 8;
      }
#pragma unroll 1
    for (int kb = 0; kb < 4; ++kb) {
      uint4 a[2][2][2];  // [m tile][row half][feature half]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          a[mt][hr][0] = __ldg(reinterpret_cast<const uint4*>(rowp[mt][hr] + kb * 64));
          a[mt][hr][1] = __ldg(reinterpret_cast<const uint4*>(rowp[mt][hr] + kb * 64 + 32));
        }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint2 b = __ldg(wfrag + ((kb * 4 + s) * 8 + j) * 32 + lane);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t* lo0 = &a[mt][0][0].x;  // row g,   features 8t..
            const uint32_t* lo1 = &a[mt][1][0].x;  // row g+8
            const uint32_t* hi0 = &a[mt][0][1].x;  // row g,   features 32+8t..
            const uint32_t* hi1 = &a[mt][1][1].x;
            mma_bf16_16816(acc[mt][j], lo0[s], lo1[s], hi0[s], hi1[s], b.x, b.y);
          }
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int64_t p = p0 + mt * 16 + hr * 8 + g;
        if (p >= n_px) continue;
        // planar T: one store instruction covers 4 columns x 8 consecutive pixels = four full 32-byte sectors
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 8 * j + 2 * t + e;
            if (col < 27) T[col * n_px + p] = acc[mt][j][2 * hr + e] + acc[mt][j + 4][2 * hr + e];
          }
      }
  }
}

// (2) one thread per output pixel: bias + the nine (tap, channel) triples of its reflected neighbours
__global__ void __launch_bounds__(128) last_conv_gather_kernel(const float* __restrict__ T, float b0, float b1, float b2,
                                                              int H_up, int W_up, int strip, int qr0, int qrows, int row0,
                                                              int row1, OutSpec out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = row0 + blockIdx.y * 4 + (threadIdx.x >> 5);
  const int b = blockIdx.z;
  if (x >= W_up || y >= row1) return;
  const int s0 = (x / strip) * strip;  // this column's strip [s0, s1)
  const int s1 = min(s0 + strip, W_up);
  const int64_t n_px = static_cast<int64_t>(gridDim.z) * qrows * W_up;  // plane stride of T
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = reflect1(y + tap / 3 - 1, 0, H_up), xx = reflect1(x + tap % 3 - 1, s0, s1);
    const float* tp = T + (tap * 3) * n_px + (static_cast<size_t>(b * qrows + (yy - qr0)) * W_up + xx);
    a0 += __ldg(tp), a1 += __ldg(tp + n_px), a2 += __ldg(tp + 2 * n_px);  // a warp reads 32 neighbouring pixels of one plane
  }
  const int64_t off = b * out.batch_stride + static_cast<int64_t>(y - row0) * out.row_stride + x;
  store_out(out, off, a0 + b0);
  store_out(out, off + out.chan_stride, a1 + b1);
  store_out(out, off + 2 * out.chan_stride, a2 + b2);
}

int launch_last_conv_umma_path(Handle* h, const __nv_bfloat16* q3, float* T, int B, int H_up, int W_up, int strip, int qr0,
                               int qrows, int row0, int row1, const OutSpec& out, cudaStream_t s) {
  const int64_t n_px = static_cast<int64_t>(B) * qrows * W_up;
  const int64_t want = (n_px + 127) / 128;  // 4 warps x 32 pixels per CTA and iteration
  const int64_t cap = static_cast<int64_t>(h->sm_count > 0 ? h->sm_count : 148) * 3;
  last_conv_project_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 128, 0, s>>>(
      q3, reinterpret_cast<const uint2*>(h->WL27frag), T, n_px);
  dim3 grid((W_up + 31) / 32, (row1 - row0 + 3) / 4, B);
  last_conv_gather_kernel<<<grid, 128, 0, s>>>(T, h->small.bl[0], h->small.bl[1], h->small.bl[2], H_up, W_up, strip, qr0, qrows,
                                               row0, row1, out);
  h->launches += 2;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
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

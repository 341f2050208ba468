// fp32 CUDA-core path of the DIINN query decoder (DIINN_COMPUTE_FP32_SIMT) plus the coordinate / index kernels shared
// by every path. True fp32 FMA arithmetic end to end: this is the path that meets the fp32 tolerance (1e-4) with
// margin on any weight set; the throughput path is the tcgen05 one (stage_a_umma.cu / stage_b_umma.cu).
//
// Stage A (once per LR pixel):  P[l] = [relu(K0 x_l + b0) | K_i[:,256:] x_l + b_i (i=1..3)]      diinn.py:133,136
// Stage B (once per HR pixel):  q0 = P[l][0:256] * sin(Q0 s_p + bq0)                              diinn.py:134
//                               k_i = relu(K_i[:, :256] q_{i-1} + P[l][256i : 256i+256])          diinn.py:136
//                               q_i = k_i * sin(Q_i q_{i-1} + bq_i)                               diinn.py:137
//                               rgb = last q_3 + bl                                               diinn.py:138
#include "handle.h"
#include "pixel.cuh"

namespace diinn {

__device__ __forceinline__ int64_t total_pixels(const PixelSource& s) {
  return s.mode == 0 ? static_cast<int64_t>(s.B) * (s.row1 - s.row0) * s.W_up
                     : static_cast<int64_t>(s.B) * s.Q * (s.ensemble ? 4 : 1);
}

__device__ __forceinline__ void store_rgb(const OutSpec& o, const PixelSource& s, const PixInfo& pi, int64_t g,
                                          int c, float v) {
  int64_t off;
  if (s.mode == 0)
    off = pi.b * o.batch_stride + c * o.chan_stride + static_cast<int64_t>(pi.oh - s.row0) * o.row_stride + pi.ow;
  else
    off = (s.ensemble ? (g >> 2) : g) * 3 + c;
  store_out(o, off, v);
}

// ---------------------------------------------------------------------------------------------------------
// coordinate / index taps (bit-exact against _make_pos_encoding, diinn.py:94-110)
// ---------------------------------------------------------------------------------------------------------
__global__ void axis_tables_kernel(AxisParams ah, AxisParams aw, int32_t* ih, int32_t* iw, float* rel_h,
                                   float* rel_w) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < ah.n_up) {
    const int i = axis_index(ah, j);
    ih[j] = i;
    rel_h[j] = axis_rel(ah, j, i);
  }
  if (j < aw.n_up) {
    const int i = axis_index(aw, j);
    iw[j] = i;
    rel_w[j] = axis_rel(aw, j, i);
  }
}

int launch_axis_tables(Handle* h, const AxisParams& ah, const AxisParams& aw, int32_t* ih, int32_t* iw,
                       float* rel_h, float* rel_w, cudaStream_t s) {
  const int n = ah.n_up > aw.n_up ? ah.n_up : aw.n_up;
  axis_tables_kernel<<<(n + 255) / 256, 256, 0, s>>>(ah, aw, ih, iw, rel_h, rel_w);
  h->launches += 1;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

__global__ void query_gather_kernel(PixelSource src, int32_t* idx, float* rel, float* ratio) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= total_pixels(src)) return;
  const PixInfo pi = pixel_info(src, g);
  idx[g] = pi.p_idx - pi.b * src.H * src.W;
  rel[g * 2] = pi.rel_h;
  rel[g * 2 + 1] = pi.rel_w;
  ratio[g] = pi.ratio;
}

int launch_query_gather(Handle* h, const PixelSource& src, int32_t* idx, float* rel, float* ratio, cudaStream_t s) {
  const int64_t n = static_cast<int64_t>(src.B) * src.Q;
  query_gather_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, idx, rel, ratio);
  h->launches += 1;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Stage A, fp32: implicit-im2col SGEMM  P(M x 1024) = X(M x 576) * WA32^T + bA, ReLU on the first 256 columns.
// 64x64 block tile, 16-deep K tiles, 256 threads each owning a 4x4 micro-tile.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) stage_a_fp32_kernel(const T* __restrict__ feat, const float* __restrict__ WA,
                                                           const float* __restrict__ bA, float* __restrict__ P,
                                                           int B, int H, int W, int lr_row0, int lr_rows) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int M = B * lr_rows * W;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // A loader: thread -> pixel (tid & 63), k rows (tid >> 6) + 4*j
  const int am = tid & 63;
  const int m = m0 + am;
  int ab = 0, ahh = 0, aww = 0;
  const bool mvalid = m < M;
  if (mvalid) {
    ab = m / (lr_rows * W);
    const int rem = m - ab * lr_rows * W;
    ahh = lr_row0 + rem / W;
    aww = rem % W;
  }
  const int bn = tid >> 2, bk = (tid & 3) * 4;  // B loader: row n0+bn, k offset bk..bk+3

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < kUnfold; k0 += 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = k0 + (tid >> 6) + 4 * j;
      const int c = kk / 9, tap = kk - c * 9;
      const int hh = ahh + tap / 3 - 1, ww = aww + tap % 3 - 1;
      float v = 0.f;
      if (mvalid && hh >= 0 && hh < H && ww >= 0 && ww < W)
        v = static_cast<float>(feat[((static_cast<size_t>(ab) * kC + c) * H + hh) * W + ww]);
      As[(tid >> 6) + 4 * j][am] = v;
    }
    {
      const float4 v = *reinterpret_cast<const float4*>(WA + static_cast<size_t>(n0 + bn) * kUnfold + k0 + bk);
      Bs[bk + 0][bn] = v.x;
      Bs[bk + 1][bn] = v.y;
      Bs[bk + 2][bn] = v.z;
      Bs[bk + 3][bn] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
    float4 o;
    float* po = &o.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      float v = acc[i][j] + bA[n];
      if (n < kD) v = fmaxf(v, 0.f);
      po[j] = v;
    }
    *reinterpret_cast<float4*>(P + static_cast<size_t>(mm) * kPCols + n0 + tx * 4) = o;
  }
}

int launch_stage_a_fp32(Handle* h, const void* feat, int io_dtype, int B, int H, int W, int lr_row0, int lr_rows,
                        float* P, cudaStream_t s) {
  const int M = B * lr_rows * W;
  dim3 grid((M + 63) / 64, kPCols / 64);
  if (io_dtype == DIINN_IO_F32)
    stage_a_fp32_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(feat), h->WA32, h->bA, P, B, H, W,
                                                    lr_row0, lr_rows);
  else
    stage_a_fp32_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(feat), h->WA32, h->bA,
                                                            P, B, H, W, lr_row0, lr_rows);
  h->launches += 1;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Stage B, fp32.
// ---------------------------------------------------------------------------------------------------------
// layer 0: q0[p][f] = P[l(p)][f] * sin(wq0[f] . (rel_h, rel_w, ratio) + bq0[f])
__global__ void __launch_bounds__(256) layer0_fp32_kernel(PixelSource src, const __grid_constant__ SmallParams sp,
                                                          const float* __restrict__ P, float* __restrict__ q,
                                                          int64_t g0, int64_t g1) {
  __shared__ int s_idx[64];
  __shared__ float s_syn[64][3];
  const int64_t base = g0 + static_cast<int64_t>(blockIdx.x) * 64;
  if (threadIdx.x < 64) {
    const int64_t g = base + threadIdx.x;
    if (g < g1) {
      const PixInfo pi = pixel_info(src, g);
      s_idx[threadIdx.x] = pi.p_idx;
      s_syn[threadIdx.x][0] = pi.rel_h;
      s_syn[threadIdx.x][1] = pi.rel_w;
      s_syn[threadIdx.x][2] = pi.ratio;
    }
  }
  __syncthreads();
  const int f = threadIdx.x;
  const float w0 = sp.wq0[f][0], w1 = sp.wq0[f][1], w2 = sp.wq0[f][2], b = sp.bq[0][f];
  const int n = static_cast<int>((g1 - base) < 64 ? (g1 - base) : 64);
  for (int p = 0; p < n; ++p) {
    // same accumulation order as a dot product over (rel_h, rel_w, ratio) followed by the bias
    float t = __fmul_rn(w0, s_syn[p][0]);
    t = fmaf(w1, s_syn[p][1], t);
    t = fmaf(w2, s_syn[p][2], t);
    t += b;
    const float k0 = P[static_cast<size_t>(s_idx[p]) * kPCols + f];
    q[static_cast<size_t>(base - g0 + p) * kD + f] = k0 * sinf(t);
  }
}

// layers 1..3: [K-part | Q-part] SGEMM over q_in (chunk x 256) with the fused dual-interactive epilogue.
// Block tile: 64 pixels x (32 K features + the 32 matching Q features).
__global__ void __launch_bounds__(256) layer_fp32_kernel(PixelSource src, const float* __restrict__ q_in,
                                                         const float* __restrict__ WB,  // (512,256) of this layer
                                                         const float* __restrict__ bq,  // (256) of this layer
                                                         const float* __restrict__ P, int p_col0,
                                                         float* __restrict__ q_out, int64_t g0, int64_t g1) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  __shared__ int s_idx[64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t base = g0 + static_cast<int64_t>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 32;
  const int rows = static_cast<int>((g1 - base) < 64 ? (g1 - base) : 64);
  if (tid < 64 && tid < rows) s_idx[tid] = pixel_info(src, base + tid).p_idx;

  const int lm = tid >> 2, lk = (tid & 3) * 4;  // loader: row lm, k offset lk..lk+3 (float4 along k)
  const int brow = lm < 32 ? n0 + lm : kD + n0 + (lm - 32);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < kD; k0 += 16) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lm < rows) a = *reinterpret_cast<const float4*>(q_in + static_cast<size_t>(base - g0 + lm) * kD + k0 + lk);
    As[lk + 0][lm] = a.x;
    As[lk + 1][lm] = a.y;
    As[lk + 2][lm] = a.z;
    As[lk + 3][lm] = a.w;
    const float4 b = *reinterpret_cast<const float4*>(WB + static_cast<size_t>(brow) * kD + k0 + lk);
    Bs[lk + 0][lm] = b.x;
    Bs[lk + 1][lm] = b.y;
    Bs[lk + 2][lm] = b.z;
    Bs[lk + 3][lm] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
      bv[0] = Bs[k][tx * 2];
      bv[1] = Bs[k][tx * 2 + 1];
      bv[2] = Bs[k][32 + tx * 2];
      bv[3] = Bs[k][32 + tx * 2 + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i;
    if (r >= rows) continue;
    const float* prow = P + static_cast<size_t>(s_idx[r]) * kPCols + p_col0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = n0 + tx * 2 + j;
      const float k = fmaxf(acc[i][j] + prow[f], 0.f);
      const float sq = sinf(acc[i][2 + j] + bq[f]);
      q_out[static_cast<size_t>(base - g0 + r) * kD + f] = k * sq;
    }
  }
}

// rgb = last q3 + bl : one warp per pixel
__global__ void __launch_bounds__(256) last_fp32_kernel(PixelSource src, OutSpec out,
                                                        const __grid_constant__ SmallParams sp,
                                                        const float* __restrict__ q, int64_t g0, int64_t g1) {
  const int lane = threadIdx.x & 31;
  const int64_t g_raw = g0 + static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const bool ens = src.mode == 1 && src.ensemble;
  if (g_raw >= g1 && !ens) return;
  const int64_t g = g_raw < g1 ? g_raw : g1 - 1;  // ensemble blocks keep all warps alive for the block barrier
  const float* row = q + static_cast<size_t>(g - g0) * kD;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
  for (int j = 0; j < kD / 32; ++j) {
    const int f = lane + 32 * j;
    const float v = row[f];
    a0 = fmaf(sp.wl_t[f][0], v, a0);
    a1 = fmaf(sp.wl_t[f][1], v, a1);
    a2 = fmaf(sp.wl_t[f][2], v, a2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if (src.mode == 1 && src.ensemble) {
    // the block's 8 warps hold the 4 variants of 2 queries (g0 and every chunk are multiples of 8): blend by area with
    // the diagonal swap, accumulating in the reference's order (liif.py:117-127)
    __shared__ float s_v[8][4];
    if (lane == 0) {
      s_v[threadIdx.x >> 5][0] = a0 + sp.bl[0];
      s_v[threadIdx.x >> 5][1] = a1 + sp.bl[1];
      s_v[threadIdx.x >> 5][2] = a2 + sp.bl[2];
      s_v[threadIdx.x >> 5][3] = pixel_info(src, g).area;
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if (lane == 0 && (w & 3) == 0 && g_raw < g1) {
      const float tot = ((s_v[w][3] + s_v[w + 1][3]) + s_v[w + 2][3]) + s_v[w + 3][3];
      const PixInfo pi = pixel_info(src, g);
      for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
        for (int v = 0; v < 4; ++v) acc += s_v[w + v][c] * (s_v[w + 3 - v][3] / tot);
        store_rgb(out, src, pi, g, c, acc);
      }
    }
    return;
  }
  if (lane == 0) {
    const PixInfo pi = pixel_info(src, g);
    store_rgb(out, src, pi, g, 0, a0 + sp.bl[0]);
    store_rgb(out, src, pi, g, 1, a1 + sp.bl[1]);
    store_rgb(out, src, pi, g, 2, a2 + sp.bl[2]);
  }
}

// layers 1..3 and the RGB projection (or mode 4's q_3 dump) for the pixels [g0,g1) whose q_0 sits in qbuf0
int run_layers_fp32(Handle* h, const PixelSource& src, const OutSpec& out, const float* P, float* qbuf0, float* qbuf1,
                    int64_t g0, int64_t g1, cudaStream_t s, float* q3_dump) {
  const unsigned nblk = static_cast<unsigned>((g1 - g0 + 63) / 64);
  const float* bq_dev = h->bq_dev;
  float* qi = qbuf0;
  float* qo = qbuf1;
  for (int li = 1; li < kLayers; ++li) {
    // mode 4: the last layer writes q_3 of the chunk's pixels straight into the dump (grid order = its pixel-major
    // order); the 3x3 conv (csrc/mode4.cu) replaces last_fp32_kernel once every chunk is done
    float* dst = (q3_dump != nullptr && li == kLayers - 1) ? q3_dump + static_cast<size_t>(g0) * kD : qo;
    layer_fp32_kernel<<<dim3(nblk, kD / 32), 256, 0, s>>>(src, qi, h->WB32 + static_cast<size_t>(li - 1) * 512 * kD,
                                                          bq_dev + li * kD, P, li * kD, dst, g0, g1);
    float* t = qi;
    qi = qo;
    qo = t;
  }
  if (q3_dump == nullptr)
    last_fp32_kernel<<<static_cast<unsigned>((g1 - g0 + 7) / 8), 256, 0, s>>>(src, out, h->small, qi, g0, g1);
  h->launches += q3_dump == nullptr ? 4 : 3;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

int run_stage_b_fp32(Handle* h, const PixelSource& src, const OutSpec& out, const float* P, float* qbuf0,
                     float* qbuf1, int64_t chunk, cudaStream_t s, float* q3_dump) {
  const int64_t total = src.mode == 0 ? static_cast<int64_t>(src.B) * (src.row1 - src.row0) * src.W_up
                                      : static_cast<int64_t>(src.B) * src.Q * (src.ensemble ? 4 : 1);
  for (int64_t g0 = 0; g0 < total; g0 += chunk) {
    const int64_t g1 = g0 + chunk < total ? g0 + chunk : total;
    layer0_fp32_kernel<<<static_cast<unsigned>((g1 - g0 + 63) / 64), 256, 0, s>>>(src, h->small, P, qbuf0, g0, g1);
    h->launches += 1;
    int rc = run_layers_fp32(h, src, out, P, qbuf0, qbuf1, g0, g1, s, q3_dump);
    if (rc) return rc;
  }
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

}  // namespace diinn

// init_q=True (diinn.py:48-51,113-115): before anything else `step` passes the synthetic input through
//   first_layer = Conv2d(3, 576, 1) + sin          s_p  = sin(Wf (rel_h, rel_w, ratio)_p + bf)        (576 per HR pixel)
// and gates the gathered unfolded features with it, x <- s_p * x_l(p). Q.0 then reads s_p (576 -> 256) instead of the
// 3-vector, and every x-facing block of the K layers sees the gated x: nothing that multiplies x depends on the LR
// cell alone any more, so stage A's hoist to LR resolution does not apply. What survives is the
// split of each K layer into an x-facing and a q/k-facing block, evaluated per chunk of HR pixels:
//
//   gate     S  = s_p,  XG = s_p * x_l(p)                          (chunk x 576 each; x gathered with the nearest-exact
//                                                                   index, zero-padded 3x3 neighbourhood)
//   GEMM     PX = XG . WA^T   (chunk x 1024: the x-facing blocks of K.0..3)      QS = S . Q0^T   (chunk x 256)
//   assemble PX += bk;  PX[:, :256] = relu(.) = k_0;  q_0: PX[:, :256] *= sin(QS + bq_0)
//   (modes 1 / 2: the k-fed chain of csrc/lr_chain.cu, now over HR pixels:  PX_i += WH_i . relu(PX_{i-1}), before q_0)
//   stage B  layers 1..3 and the last layer exactly as without init_q, reading ONE PX ROW PER PIXEL (kPix variants of
//            stage_b_umma_kernel / per_pixel_p in the fp32 kernels) and taking q_0 as given.
//
// fp32 path: CUDA-core SGEMMs in the reference's channel order (c*9 + tap), separate assemble / q_0 kernels. Tensor path
// (run_initq_umma): bf16 gate operands in tap-major order (tap*64 + c, so the gate reads whole 128-byte channel vectors of
// the NHWC feature copy and stage A's weight tiles serve unchanged); both products run on the STAGE-A KERNEL in matrix mode
// (stage_a_umma.cu: persistent, 6-stage TMA ring, TMA-store epilogue) with bias, ReLU and the q_0 product fused into the
// epilogue; PX is written once, as fp16 rows (2 KB per pixel), QS as fp32. Executed arithmetic: 2 264 064 FLOP per HR pixel
// (nothing is shared between pixels). What still round-trips through L2 / HBM per pixel: S, XG (1.1 KB each), PX (2 KB),
// QS (1 KB) -- a single persistent kernel that keeps XG in shared memory is the remaining step.
#include <cuda_fp16.h>

#include <cstdlib>

#include "handle.h"
#include "pixel.cuh"

namespace diinn {

constexpr int64_t kInitQChunkUmma = 148 * 128 * 8;  // HR pixels per tensor-path chunk: eight waves of stage-B tiles (measured on
                                                    // c3: 2 waves 11.4 ms, 4: 10.4, 8: 8.9, 16: 9.6 -- launch tails vs L2 residency of PX)
constexpr int64_t kInitQChunkFp32 = 1 << 15;        // HR pixels per fp32-path chunk

// ---------------------------------------------------------------------------------------------------------
// gate: 192 threads, 3 channels each, 32 pixels per batch. Column j of S / XG is unfolded channel
//   kTapMajor ? (tap = j / 64, c = j % 64) : (c = j / 9, tap = j % 9)        -- reference channel index k = c*9 + tap.
// Rows [g1 - g0, rows_pad) are zero-filled (padding up to the GEMM's M tile).
// ---------------------------------------------------------------------------------------------------------
struct FeatNCHW {      // the caller's (B,64,H,W) tensor
  const void* ptr;
  int bf16;
  __device__ __forceinline__ float at(int b, int c, int hh, int ww, int H, int W) const {
    const size_t i = ((static_cast<size_t>(b) * kC + c) * H + hh) * W + ww;
    return bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(ptr)[i]) : static_cast<const float*>(ptr)[i];
  }
};
struct FeatNHWC {      // (B, frows, W, 64) bf16 holding LR rows [fr0, fr0 + frows)
  const __nv_bfloat16* ptr;
  int fr0, frows;
  __device__ __forceinline__ float at(int b, int c, int hh, int ww, int H, int W) const {
    return __bfloat162float(ptr[((static_cast<size_t>(b) * frows + (hh - fr0)) * W + ww) * kC + c]);
  }
};

__device__ __forceinline__ void store_gate(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_gate(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename Feat, typename TO, bool kTapMajor, bool kFastSin>
__global__ void __launch_bounds__(192) initq_gate_kernel(PixelSource src, Feat feat, const float4* __restrict__ wf4,
                                                         TO* __restrict__ S, TO* __restrict__ XG, int64_t g0, int64_t g1,
                                                         int64_t rows_pad) {
  // batches of 32 rows: the first warp works out the 32 pixels' bookkeeping (index divisions, nearest-exact cell, relative
  // coordinates), then every thread runs its three channels -- first_layer rows held in registers -- over the batch
  __shared__ float s_syn[32][3];
  __shared__ int s_loc[32][3];  // batch image, LR row, LR col
  int c[3], dh[3], dw[3];
  float4 w[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int j = threadIdx.x + 192 * i;
    c[i] = kTapMajor ? (j & 63) : j / 9;
    const int tap = kTapMajor ? (j >> 6) : j % 9;
    dh[i] = tap / 3 - 1, dw[i] = tap % 3 - 1;
    w[i] = __ldg(wf4 + c[i] * 9 + tap);
  }
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * 32; base < rows_pad; base += static_cast<int64_t>(gridDim.x) * 32) {
    __syncthreads();  // the previous batch has been consumed
    if (threadIdx.x < 32 && g0 + base + threadIdx.x < g1) {
      const PixInfo pi = pixel_info(src, g0 + base + threadIdx.x);
      s_syn[threadIdx.x][0] = pi.rel_h, s_syn[threadIdx.x][1] = pi.rel_w, s_syn[threadIdx.x][2] = pi.ratio;
      s_loc[threadIdx.x][0] = pi.b, s_loc[threadIdx.x][1] = pi.ih, s_loc[threadIdx.x][2] = pi.iw;
    }
    __syncthreads();
    const int n = static_cast<int>(rows_pad - base < 32 ? rows_pad - base : 32);
    for (int r = 0; r < n; ++r) {
      const int64_t row = base + r;
      TO* srow = S + row * kUnfold;
      TO* xrow = XG + row * kUnfold;
      if (g0 + row >= g1) {  // padding up to the GEMM's M tile
#pragma unroll
        for (int i = 0; i < 3; ++i) store_gate(srow + threadIdx.x + 192 * i, 0.f), store_gate(xrow + threadIdx.x + 192 * i, 0.f);
        continue;
      }
      const float rel_h = s_syn[r][0], rel_w = s_syn[r][1], ratio = s_syn[r][2];
      const int b = s_loc[r][0], ih = s_loc[r][1], iw = s_loc[r][2];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        // same accumulation order as layer 0 of the init_q=False path: a 3-term dot product, then the bias
        float t = __fmul_rn(w[i].x, rel_h);
        t = fmaf(w[i].y, rel_w, t);
        t = fmaf(w[i].z, ratio, t);
        t += w[i].w;
        const float sv = kFastSin ? __sinf(t) : sinf(t);
        const int hh = ih + dh[i], ww = iw + dw[i];
        const float xv = (hh >= 0 && hh < src.H && ww >= 0 && ww < src.W) ? feat.at(b, c[i], hh, ww, src.H, src.W) : 0.f;
        store_gate(srow + threadIdx.x + 192 * i, sv);
        store_gate(xrow + threadIdx.x + 192 * i, sv * xv);
      }
    }
  }
}

// Tensor-path gate, vectorised: 288 threads = 4 rows x 72 groups of 8 consecutive tap-major columns (one tap, channels
// [c0, c0 + 8)): one 16-byte load of the NHWC feature vector, eight sines, two 16-byte stores (S, XG) per thread and row.
// Same arithmetic per element as initq_gate_kernel<.., true, true>.
__global__ void __launch_bounds__(288) initq_gate_vec_kernel(PixelSource src, FeatNHWC feat, const float4* __restrict__ wf4,
                                                             __nv_bfloat16* __restrict__ S, __nv_bfloat16* __restrict__ XG,
                                                             int64_t g1, int64_t rows_pad) {
  __shared__ float s_syn[32][3];
  __shared__ int s_loc[32][3];
  const int grp = threadIdx.x % 72, rsub = threadIdx.x / 72;
  const int j0 = grp * 8, tap = j0 >> 6, c0 = j0 & 63;
  const int dh = tap / 3 - 1, dw = tap % 3 - 1;
  float4 w[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) w[e] = __ldg(wf4 + (c0 + e) * 9 + tap);
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * 32; base < rows_pad; base += static_cast<int64_t>(gridDim.x) * 32) {
    __syncthreads();
    if (threadIdx.x < 32 && base + threadIdx.x < g1) {
      const PixInfo pi = pixel_info(src, base + threadIdx.x);
      s_syn[threadIdx.x][0] = pi.rel_h, s_syn[threadIdx.x][1] = pi.rel_w, s_syn[threadIdx.x][2] = pi.ratio;
      s_loc[threadIdx.x][0] = pi.b, s_loc[threadIdx.x][1] = pi.ih, s_loc[threadIdx.x][2] = pi.iw;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = rsub; r < 32; r += 4) {
      const int64_t row = base + r;
      if (row >= rows_pad) break;
      uint4 so = make_uint4(0u, 0u, 0u, 0u), xo = so;
      if (row < g1) {
        const float rel_h = s_syn[r][0], rel_w = s_syn[r][1], ratio = s_syn[r][2];
        const int b = s_loc[r][0], hh = s_loc[r][1] + dh, ww = s_loc[r][2] + dw;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (hh >= 0 && hh < src.H && ww >= 0 && ww < src.W) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(
              feat.ptr + ((static_cast<size_t>(b) * feat.frows + (hh - feat.fr0)) * src.W + ww) * kC + c0));
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) x[2 * i] = __uint_as_float(uw[i] << 16), x[2 * i + 1] = __uint_as_float(uw[i] & 0xffff0000u);
        }
        uint32_t sp[4], xp[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float sv[2];
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            float t = __fmul_rn(w[e + d].x, rel_h);
            t = fmaf(w[e + d].y, rel_w, t);
            t = fmaf(w[e + d].z, ratio, t);
            t += w[e + d].w;
            sv[d] = __sinf(t);
          }
          __nv_bfloat162 s2 = __floats2bfloat162_rn(sv[0], sv[1]);
          __nv_bfloat162 x2 = __floats2bfloat162_rn(sv[0] * x[e], sv[1] * x[e + 1]);
          sp[e >> 1] = *reinterpret_cast<uint32_t*>(&s2);
          xp[e >> 1] = *reinterpret_cast<uint32_t*>(&x2);
        }
        so = make_uint4(sp[0], sp[1], sp[2], sp[3]);
        xo = make_uint4(xp[0], xp[1], xp[2], xp[3]);
      }
      *reinterpret_cast<uint4*>(S + row * kUnfold + j0) = so;
      *reinterpret_cast<uint4*>(XG + row * kUnfold + j0) = xo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// fp32: C(M x N, ldc) = A(M x K) . B(N x K)^T, 64x64 tiles, 256 threads, 4x4 per thread. K % 16 == 0, N % 64 == 0.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                       float* __restrict__ Cm, int64_t M, int K, int ldc) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      const int64_t m = m0 + r;
      As[k][r] = m < M ? A[m * K + k0 + k] : 0.f;
      Bs[k][r] = Bm[static_cast<size_t>(n0 + r) * K + k0 + k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i], b[i] = Bs[k][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) Cm[m * ldc + n0 + tx * 4 + j] = acc[i][j];
  }
}

// ---------------------------------------------------------------------------------------------------------
// assemble (flags bit 0): PX[r][n] += bA[n], ReLU on n < 256.   q_0 (flags bit 1): PX[r][f] *= sin(QS[r][f] + bq0[f]).
// One thread per 4 columns; rows < M only.
// ---------------------------------------------------------------------------------------------------------
template <bool kFastSin>
__global__ void __launch_bounds__(256) initq_assemble_kernel(float* __restrict__ PX, const float* __restrict__ QS,
                                                             const float* __restrict__ bA, const float* __restrict__ bq0,
                                                             int64_t M, int flags) {
  const int64_t total = M * (kPCols / 4);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / (kPCols / 4);
    const int n = static_cast<int>(e % (kPCols / 4)) * 4;
    if (!(flags & 1) && n >= kD) continue;  // the q_0 pass touches block 0 only
    float4 v = *reinterpret_cast<float4*>(PX + r * kPCols + n);
    if (flags & 1) {
      const float4 b = *reinterpret_cast<const float4*>(bA + n);
      v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
      if (n < kD) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    }
    if ((flags & 2) && n < kD) {
      const float4 t = *reinterpret_cast<const float4*>(QS + r * kD + n);
      const float4 b = *reinterpret_cast<const float4*>(bq0 + n);
      const float a0 = t.x + b.x, a1 = t.y + b.y, a2 = t.z + b.z, a3 = t.w + b.w;
      v.x *= kFastSin ? __sinf(a0) : sinf(a0);
      v.y *= kFastSin ? __sinf(a1) : sinf(a1);
      v.z *= kFastSin ? __sinf(a2) : sinf(a2);
      v.w *= kFastSin ? __sinf(a3) : sinf(a3);
    }
    *reinterpret_cast<float4*>(PX + r * kPCols + n) = v;
  }
}

// q_0 of the fp32 path leaves PX alone (the chain of modes 1 / 2 is finished by then) and fills the activation buffer
__global__ void __launch_bounds__(256) initq_q0_fp32_kernel(const float* __restrict__ PX, const float* __restrict__ QS,
                                                            const float* __restrict__ bq0, float* __restrict__ q, int64_t M) {
  const int64_t total = M * kD;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / kD;
    const int f = static_cast<int>(e % kD);
    q[e] = PX[r * kPCols + f] * sinf(QS[e] + bq0[f]);
  }
}

static unsigned capped_blocks(const Handle* h, int64_t want) {
  const int64_t cap = static_cast<int64_t>(h->sm_count > 0 ? h->sm_count : 148) * 8;
  return static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
InitQPlan plan_initq(int B, int W_up, int rows, int compute, int mode, size_t off) {
  InitQPlan p;
  const bool fp32 = compute == DIINN_COMPUTE_FP32_SIMT;
  const int64_t total = static_cast<int64_t>(B) * rows * W_up;
  if (fp32) {
    p.chunk = total < kInitQChunkFp32 ? total : kInitQChunkFp32;
    p.rows_pad = p.chunk;
  } else {
    static int64_t chunk_px = 0;  // DIINN_INITQ_CHUNK=<pixels>: measurement override
    if (chunk_px == 0) {
      const char* e = getenv("DIINN_INITQ_CHUNK");
      chunk_px = (e && atoll(e) > 0) ? atoll(e) : kInitQChunkUmma;
    }
    int64_t r = chunk_px / (static_cast<int64_t>(B) * W_up);
    // whole 8-row stage-B patches, so that only the band's last chunk has a partial tile row; very wide images (c4: 7 680
    // columns) take one patch row per chunk as long as that stays within 4 chunks' worth of scratch
    if (r >= 8) r -= r % 8;
    else if (static_cast<int64_t>(B) * W_up * 8 <= 4 * chunk_px) r = 8;
    r = r < 1 ? 1 : (r > rows ? rows : r);
    p.chunk_rows = static_cast<int>(r);
    p.chunk = static_cast<int64_t>(B) * r * W_up;
    p.rows_pad = (p.chunk + 255) / 256 * 256;  // M tile of the CTA-pair GEMM
  }
  const size_t el = fp32 ? sizeof(float) : sizeof(__nv_bfloat16);
  p.off_S = off;
  off += align_up(static_cast<size_t>(p.rows_pad) * kUnfold * el);
  p.off_XG = off;
  off += align_up(static_cast<size_t>(p.rows_pad) * kUnfold * el);
  p.off_PX = off;
  off += align_up(static_cast<size_t>(p.rows_pad) * kPCols * (fp32 ? sizeof(float) : sizeof(__half)));  // tensor path: fp16 rows
  p.off_QS = off;
  off += align_up(static_cast<size_t>(p.rows_pad) * kD * sizeof(float));
  if (fp32) {
    p.off_q0 = off;
    off += align_up(static_cast<size_t>(p.chunk) * kD * sizeof(float));
    p.off_q1 = off;
    off += align_up(static_cast<size_t>(p.chunk) * kD * sizeof(float));
  } else if (mode == 1 || mode == 2) {
    p.off_chain = off;
    off += align_up(lr_chain_scratch_bytes(p.chunk));
  }
  p.end = off;
  return p;
}

int run_initq_fp32(Handle* h, const void* feat, int io_dtype, const PixelSource& src_in, const OutSpec& out, char* ws,
                   const InitQPlan& pl, cudaStream_t s, float* q3_dump) {
  const int64_t total = static_cast<int64_t>(src_in.B) * (src_in.row1 - src_in.row0) * src_in.W_up;
  float* S = reinterpret_cast<float*>(ws + pl.off_S);
  float* XG = reinterpret_cast<float*>(ws + pl.off_XG);
  float* PX = reinterpret_cast<float*>(ws + pl.off_PX);
  float* QS = reinterpret_cast<float*>(ws + pl.off_QS);
  float* q0 = reinterpret_cast<float*>(ws + pl.off_q0);
  float* q1 = reinterpret_cast<float*>(ws + pl.off_q1);
  const bool chain = h->cfg.mode == 1 || h->cfg.mode == 2;
  const FeatNCHW f{feat, io_dtype != DIINN_IO_F32};
  const float4* wf4 = reinterpret_cast<const float4*>(h->WF4);
  for (int64_t g0 = 0; g0 < total; g0 += pl.chunk) {
    const int64_t g1 = g0 + pl.chunk < total ? g0 + pl.chunk : total;
    const int64_t M = g1 - g0;
    PixelSource src = src_in;
    src.per_pixel_p = 1, src.p_base = g0;
    initq_gate_kernel<FeatNCHW, float, false, false><<<capped_blocks(h, (M + 31) / 32), 192, 0, s>>>(src, f, wf4, S, XG, g0, g1, M);
    const unsigned mb = static_cast<unsigned>((M + 63) / 64);
    sgemm_nt_kernel<<<dim3(mb, kPCols / 64), 256, 0, s>>>(XG, h->WA32, PX, M, kUnfold, kPCols);
    sgemm_nt_kernel<<<dim3(mb, kD / 64), 256, 0, s>>>(S, h->WQ0_32, QS, M, kUnfold, kD);
    initq_assemble_kernel<false><<<capped_blocks(h, (M * (kPCols / 4) + 255) / 256), 256, 0, s>>>(PX, QS, h->bA, h->bq_dev, M, 1);
    h->launches += 4;
    DIINN_CUDA_OK(h, cudaGetLastError());
    int rc;
    if (chain && (rc = run_lr_chain_fp32(h, PX, M, s))) return rc;
    initq_q0_fp32_kernel<<<capped_blocks(h, (M * kD + 255) / 256), 256, 0, s>>>(PX, QS, h->bq_dev, q0, M);
    h->launches += 1;
    if ((rc = run_layers_fp32(h, src, out, PX, q0, q1, g0, g1, s, q3_dump))) return rc;
  }
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

// q_0 of the chain modes on the tensor path: PX[r][f] (fp16) *= sin(QS[r][f]), f < 256 (QS already holds Q.0 s + bq_0)
__global__ void __launch_bounds__(256) initq_q0_p16_kernel(__half* __restrict__ PX, const float* __restrict__ QS, int64_t M) {
  const int64_t total = M * (kD / 4);
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / (kD / 4);
    const int n = static_cast<int>(e % (kD / 4)) * 4;
    uint2* p = reinterpret_cast<uint2*>(PX + r * kPCols + n);
    const uint2 u = *p;
    const float4 t = *reinterpret_cast<const float4*>(QS + r * kD + n);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    const __half2 lo = __floats2half2_rn(a.x * __sinf(t.x), a.y * __sinf(t.y));
    const __half2 hi = __floats2half2_rn(b.x * __sinf(t.z), b.y * __sinf(t.w));
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&lo), o.y = *reinterpret_cast<const uint32_t*>(&hi);
    *p = o;
  }
}

// Tensor path, per chunk of HR pixels (whole patch rows of stage B):
//   gate      S, XG (bf16, K tap-major)                                   initq_gate_vec_kernel
//   QS      = Q.0 S + bq_0 (fp32)                                         stage-A kernel in matrix mode, one N-block
//   PX      = [relu(K_0 XG + b_0) * sin(QS) | K_i[:, 256:] XG + b_i]      stage-A kernel in matrix mode, fp16 rows (2 KB / pixel);
//             (modes 1 / 2: without the sine, then the k-fed chain over the fp16 rows, then q_0)
//   stage B   kPix instantiation: one fp16 PX row per pixel, q_0 given
// The 576-wide products use bf16 operands whatever the decode's operand format (the gate is written once, as bf16).
int run_initq_umma(Handle* h, const __nv_bfloat16* nhwc, int fr0, int frows, const PixelSource& src_in, const OutSpec& out,
                   char* ws, const InitQPlan& pl, int fmt, cudaStream_t s) {
  __nv_bfloat16* S = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_S);
  __nv_bfloat16* XG = reinterpret_cast<__nv_bfloat16*>(ws + pl.off_XG);
  __half* PX = reinterpret_cast<__half*>(ws + pl.off_PX);
  float* QS = reinterpret_cast<float*>(ws + pl.off_QS);
  const bool chain = h->cfg.mode == 1 || h->cfg.mode == 2;
  const FeatNHWC f{nhwc, fr0, frows};
  const float4* wf4 = reinterpret_cast<const float4*>(h->WF4);
  for (int a = src_in.row0; a < src_in.row1; a += pl.chunk_rows) {
    const int b = a + pl.chunk_rows < src_in.row1 ? a + pl.chunk_rows : src_in.row1;
    const int64_t M = static_cast<int64_t>(src_in.B) * (b - a) * src_in.W_up;
    const int64_t Mp = (M + 255) / 256 * 256;
    PixelSource src = src_in;
    src.row0 = a, src.row1 = b;
    src.per_pixel_p = 1, src.p_base = 0, src.out_row0 = src_in.row0;
    initq_gate_vec_kernel<<<capped_blocks(h, (Mp + 31) / 32), 288, 0, s>>>(src, f, wf4, S, XG, M, Mp);
    h->launches += 1;
    DIINN_CUDA_OK(h, cudaGetLastError());
    int rc;
    if ((rc = launch_stage_a_matrix(h, S, Mp, 1, nullptr, QS, false, s))) return rc;
    if ((rc = launch_stage_a_matrix(h, XG, Mp, 0, chain ? nullptr : QS, PX, true, s))) return rc;
    if (chain) {  // modes 1 / 2: q_0 only once the k-fed chain has read k_0
      if ((rc = run_lr_chain_umma(h, reinterpret_cast<float*>(PX), M, ws + pl.off_chain, s, true))) return rc;
      initq_q0_p16_kernel<<<capped_blocks(h, (M * (kD / 4) + 255) / 256), 256, 0, s>>>(PX, QS, M);
      h->launches += 1;
      DIINN_CUDA_OK(h, cudaGetLastError());
    }
    if ((rc = launch_stage_b_umma(h, src, out, PX, 0, fmt, s))) return rc;
  }
  return DIINN_OK;
}

}  // namespace diinn

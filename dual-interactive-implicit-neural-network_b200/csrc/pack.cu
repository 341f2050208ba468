// Weight repacking (reference state_dict layout -> library layouts) and the NCHW -> NHWC bf16 feature pass.
//
// Reference layout (diinn.py:73-80,92):
//   K.0 (256,576)   K.i (256,832) with input channels [0,256) <- q and [256,832) <- x   (torch.cat([q,x]), diinn.py:136)
//   Q.0 (256,3)     Q.i (256,256)      last (3,256)
//   x channel c*9 + kh*3 + kw = feat[c, h+kh-1, w+kw-1]                                   (F.unfold, diinn.py:168)
// Mode 2 (diinn.py:65-72,124-131) has the same shapes with K.i's first 256 input channels fed by k instead of q; mode 1
// (diinn.py:57-64,116-123) has K.i (256,256) fed by k alone. In both the K chain never sees a per-HR-pixel quantity, so
// it is evaluated per LR pixel: stage A's matrix keeps the x-facing blocks (zero for mode 1), the k-facing 256x256 blocks
// go to WH for the LR chain (run_lr_chain_*), and stage B's K rows are ZERO so that its relu(0 + P[l]) returns k_i.
//
// Because every term multiplying x depends only on the LR pixel, the four x-facing blocks are stacked into one
// (1024 x 576) matrix evaluated once per LR pixel ("stage A"); the per-HR-pixel work ("stage B") keeps only the
// 256x256 q-facing blocks of K.1..3 next to Q.1..3.
#include <cuda_fp16.h>

#include "handle.h"

namespace diinn {

// the three 16-bit renderings of one fp32 weight: bf16, fp16 (saturating), fp16 residual after the fp16 part
__device__ __forceinline__ void put_formats(float v, size_t idx, uint16_t* __restrict__ bf, uint16_t* __restrict__ hi,
                                            uint16_t* __restrict__ lo) {
  bf[idx] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
  const float vc = fminf(fmaxf(v, -65504.f), 65504.f);
  const __half hh = __float2half_rn(vc);
  hi[idx] = __half_as_ushort(hh);
  lo[idx] = __half_as_ushort(__float2half_rn(vc - __half2float(hh)));
}

struct RefPtrs {
  const float* kw[4];
  const float* kb[4];
  const float* qw[4];
  const float* qb[4];
  const float* lw;
  const float* lb;
};

__global__ void pack_stage_a_kernel(RefPtrs r, int mode, float* __restrict__ WA32, float* __restrict__ bA,
                                    uint16_t* __restrict__ WA_bf, uint16_t* __restrict__ WA_hi,
                                    uint16_t* __restrict__ WA_lo) {
  const int n = blockIdx.x;  // 0..1023
  const int layer = n >> 8, row = n & 255;
  const int stride = layer == 0 ? kUnfold : kD + kUnfold;
  const int off = layer == 0 ? 0 : kD;
  const bool has_x = layer == 0 || mode != 1;  // mode 1: K.1..3 take k only
  const float* src = r.kw[layer] + static_cast<size_t>(row) * stride + off;
  for (int k = threadIdx.x; k < kUnfold; k += blockDim.x) {
    const float v = has_x ? src[k] : 0.f;
    WA32[static_cast<size_t>(n) * kUnfold + k] = v;
    const int c = k / 9, tap = k % 9;
    // (n-block, tap, row, c)
    put_formats(v, ((static_cast<size_t>(layer) * 9 + tap) * 256 + row) * kC + c, WA_bf, WA_hi, WA_lo);
  }
  if (threadIdx.x == 0) bA[n] = r.kb[layer][row];
}

__global__ void pack_stage_b_kernel(RefPtrs r, int mode, float* __restrict__ WB32, uint16_t* __restrict__ WB_bf,
                                    uint16_t* __restrict__ WB_hi, uint16_t* __restrict__ WB_lo, float* __restrict__ WH32,
                                    __nv_bfloat16* __restrict__ WH16) {
  const int n = blockIdx.x;   // 0..511
  const int li = blockIdx.y;  // 0..2 -> reference layer li+1
  const bool is_q = n >= kD;
  const int row = is_q ? n - kD : n;
  const float* src = is_q ? r.qw[li + 1] + static_cast<size_t>(row) * kD
                          : r.kw[li + 1] + static_cast<size_t>(row) * (mode == 1 ? kD : kD + kUnfold);
  const bool to_chain = !is_q && (mode == 1 || mode == 2);  // k-facing block: LR chain instead of stage B
  const int half = row >> 7;                              // which 128-feature half of the layer output
  const int tile_row = (is_q ? 128 : 0) + (row & 127);    // K-part rows [0,128), Q-part rows [128,256)
  for (int k = threadIdx.x; k < kD; k += blockDim.x) {
    float v = src[k];
    if (to_chain) {
      WH32[(static_cast<size_t>(li) * kD + row) * kD + k] = v;
      WH16[(static_cast<size_t>(li) * kD + row) * kD + k] = __float2bfloat16_rn(v);
      v = 0.f;
    }
    WB32[(static_cast<size_t>(li) * 512 + n) * kD + k] = v;
    const int kc = k >> 6, e = k & 63;
    put_formats(v, ((((static_cast<size_t>(li) * 2 + half) * 4 + kc) * 256) + tile_row) * 64 + e, WB_bf, WB_hi, WB_lo);
  }
}

// init_q=True, tensor path: Q.0 (256,576) in stage A's tile order (tap, row, channel), bf16 -- the B operand of the matrix-mode
// stage A that evaluates Q.0 on the gate (the x-facing K blocks reuse WA16 as they are)
__global__ void pack_initq_kernel(const float* __restrict__ q0w, __nv_bfloat16* __restrict__ WQ0A16) {
  const int n = blockIdx.x;  // 0..255: row of Q.0
  const float* src = q0w + static_cast<size_t>(n) * kUnfold;
  for (int j = threadIdx.x; j < kUnfold; j += blockDim.x)  // j = tap*64 + c  <-  reference k = c*9 + tap
    WQ0A16[(static_cast<size_t>(j >> 6) * kD + n) * kC + (j & 63)] = __float2bfloat16_rn(src[(j & 63) * 9 + (j >> 6)]);
}

int pack_weights(Handle* h, const diinn_weights_f32* w, cudaStream_t s) {
  RefPtrs r{};
  float* staging = nullptr;
  const int mode = h->cfg.mode;
  const bool init_q = h->cfg.init_q != 0;
  if (init_q && (!w->first_weight || !w->first_bias))
    return fail(h, DIINN_ERR_BAD_ARG, "init_q=True needs first_weight (576,3) and first_bias (576) (diinn.py:48-51)");
  const size_t kw_i = mode == 1 ? 256 * 256 : 256 * 832;
  const size_t sizes_kw[4] = {256 * 576, kw_i, kw_i, kw_i};
  const size_t sizes_qw[4] = {static_cast<size_t>(init_q ? 256 * 576 : 256 * 3), 256 * 256, 256 * 256, 256 * 256};
  size_t total = 0;
  for (int i = 0; i < 4; ++i) total += sizes_kw[i] + sizes_qw[i] + 512;
  total += 9 * 3 * 256 + 4;
  // host weights are staged on the device for the packing kernels; the guard frees the staging on every exit path
  struct StagingGuard {
    float** p;
    ~StagingGuard() {
      if (*p) cudaFree(*p);
    }
  } staging_guard{&staging};
  if (!w->on_device) {
    DIINN_CUDA_OK(h, cudaMalloc(&staging, total * sizeof(float)));
    float* p = staging;
    cudaError_t up_err = cudaSuccess;
    auto up = [&](const float* src, size_t n) -> const float* {
      const cudaError_t e = cudaMemcpyAsync(p, src, n * sizeof(float), cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess && up_err == cudaSuccess) up_err = e;
      const float* d = p;
      p += n;
      return d;
    };
    for (int i = 0; i < 4; ++i) {
      r.kw[i] = up(w->k_weight[i], sizes_kw[i]);
      r.kb[i] = up(w->k_bias[i], 256);
      r.qw[i] = up(w->q_weight[i], sizes_qw[i]);
      r.qb[i] = up(w->q_bias[i], 256);
    }
    r.lw = up(w->last_weight, mode == 4 ? 768 * 9 : 768);
    r.lb = up(w->last_bias, 3);
    DIINN_CUDA_OK(h, up_err);
  } else {
    for (int i = 0; i < 4; ++i) {
      r.kw[i] = w->k_weight[i];
      r.kb[i] = w->k_bias[i];
      r.qw[i] = w->q_weight[i];
      r.qb[i] = w->q_bias[i];
    }
    r.lw = w->last_weight;
    r.lb = w->last_bias;
  }
  if (!h->WA32) {
    DIINN_CUDA_OK(h, cudaMalloc(&h->WA32, sizeof(float) * kPCols * kUnfold));
    DIINN_CUDA_OK(h, cudaMalloc(&h->bA, sizeof(float) * kPCols));
    DIINN_CUDA_OK(h, cudaMalloc(&h->WB32, sizeof(float) * 3 * 512 * kD));
    DIINN_CUDA_OK(h, cudaMalloc(&h->bq_dev, sizeof(float) * kLayers * kD));
    for (int f = 0; f < 2; ++f) {
      DIINN_CUDA_OK(h, cudaMalloc(&h->WA16[f], sizeof(uint16_t) * kPCols * kUnfold));
      DIINN_CUDA_OK(h, cudaMalloc(&h->WB16[f], sizeof(uint16_t) * 3 * 512 * kD));
    }
    DIINN_CUDA_OK(h, cudaMalloc(&h->WA16lo, sizeof(uint16_t) * kPCols * kUnfold));
    DIINN_CUDA_OK(h, cudaMalloc(&h->WB16lo, sizeof(uint16_t) * 3 * 512 * kD));
    DIINN_CUDA_OK(h, cudaMalloc(&h->WH32, sizeof(float) * 3 * kD * kD));
    DIINN_CUDA_OK(h, cudaMalloc(&h->WH16, sizeof(__nv_bfloat16) * 3 * kD * kD));
  }
  pack_stage_a_kernel<<<kPCols, 192, 0, s>>>(r, mode, h->WA32, h->bA, h->WA16[0], h->WA16[1], h->WA16lo);
  pack_stage_b_kernel<<<dim3(512, 3), 128, 0, s>>>(r, mode, h->WB32, h->WB16[0], h->WB16[1], h->WB16lo, h->WH32, h->WH16);
  h->launches += 2;
  DIINN_CUDA_OK(h, cudaGetLastError());

  // small fp32 parameters -> host struct (they travel to the kernels as by-value parameters / constant bank)
  SmallParams& sp = h->small;
  const cudaMemcpyKind kind = w->on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
  for (int i = 0; i < 4; ++i)
    DIINN_CUDA_OK(h, cudaMemcpyAsync(sp.bq[i], w->q_bias[i], sizeof(float) * kD, kind, s));
  static thread_local float tmp_q0[kD * 3];  // init_q=True: Q.0 is (256,576) and layer 0 never reads wq0 (left zero)
  if (init_q) memset(tmp_q0, 0, sizeof(tmp_q0));
  else DIINN_CUDA_OK(h, cudaMemcpyAsync(tmp_q0, w->q_weight[0], sizeof(float) * kD * 3, kind, s));
  static thread_local float tmp_wf[kUnfold * 3], tmp_bf[kUnfold];
  if (init_q) {
    DIINN_CUDA_OK(h, cudaMemcpyAsync(tmp_wf, w->first_weight, sizeof(tmp_wf), kind, s));
    DIINN_CUDA_OK(h, cudaMemcpyAsync(tmp_bf, w->first_bias, sizeof(tmp_bf), kind, s));
  }
  static thread_local float tmp_wl[9 * 3 * kD];
  DIINN_CUDA_OK(h, cudaMemcpyAsync(tmp_wl, w->last_weight, sizeof(float) * 3 * kD * (mode == 4 ? 9 : 1), kind, s));
  DIINN_CUDA_OK(h, cudaMemcpyAsync(sp.bl, w->last_bias, sizeof(float) * 3, kind, s));
  for (int i = 0; i < 4; ++i)
    DIINN_CUDA_OK(h, cudaMemcpyAsync(h->bA_host + i * kD, w->k_bias[i], sizeof(float) * kD, kind, s));
  DIINN_CUDA_OK(h, cudaStreamSynchronize(s));
  sp.bl[3] = 0.f;
  for (int f = 0; f < kD; ++f) {
    sp.wq0[f][0] = tmp_q0[f * 3 + 0];
    sp.wq0[f][1] = tmp_q0[f * 3 + 1];
    sp.wq0[f][2] = tmp_q0[f * 3 + 2];
    sp.wq0[f][3] = sp.bq[0][f];
    // (mode 4's 3x3 last layer does not use wl_t: its (3,256,3,3) weight goes to WL4 below)
    sp.wl_t[f][0] = mode == 4 ? 0.f : tmp_wl[f];
    sp.wl_t[f][1] = mode == 4 ? 0.f : tmp_wl[kD + f];
    sp.wl_t[f][2] = mode == 4 ? 0.f : tmp_wl[2 * kD + f];
    sp.wl_t[f][3] = 0.f;
  }
  for (int f = 0; f < kD; f += 2)
    for (int c = 0; c < 4; ++c) sp.wq0_p[f >> 1][c] = make_float2(sp.wq0[f][c], sp.wq0[f + 1][c]);
  if (init_q) {
    static thread_local float wf4[kUnfold * 4];
    for (int k = 0; k < kUnfold; ++k) {
      for (int c = 0; c < 3; ++c) wf4[k * 4 + c] = tmp_wf[k * 3 + c];
      wf4[k * 4 + 3] = tmp_bf[k];
    }
    if (!h->WF4) {
      DIINN_CUDA_OK(h, cudaMalloc(&h->WF4, sizeof(wf4)));
      DIINN_CUDA_OK(h, cudaMalloc(&h->WQ0_32, sizeof(float) * kD * kUnfold));
      DIINN_CUDA_OK(h, cudaMalloc(&h->WQ0A16, sizeof(__nv_bfloat16) * kD * kUnfold));
    }
    DIINN_CUDA_OK(h, cudaMemcpyAsync(h->WF4, wf4, sizeof(wf4), cudaMemcpyHostToDevice, s));
    DIINN_CUDA_OK(h, cudaMemcpyAsync(h->WQ0_32, r.qw[0], sizeof(float) * kD * kUnfold, cudaMemcpyDeviceToDevice, s));
    pack_initq_kernel<<<kD, 192, 0, s>>>(h->WQ0_32, h->WQ0A16);
    h->launches += 1;
    DIINN_CUDA_OK(h, cudaGetLastError());
  }
  if (mode == 4) {
    // (c, f, ky, kx) -> (tap = ky*3 + kx, f, c) padded to float4
    static thread_local float wl4[9 * kD * 4];
    for (int tap = 0; tap < 9; ++tap)
      for (int f = 0; f < kD; ++f) {
        for (int c = 0; c < 3; ++c) wl4[(tap * kD + f) * 4 + c] = tmp_wl[(c * kD + f) * 9 + tap];
        wl4[(tap * kD + f) * 4 + 3] = 0.f;
      }
    if (!h->WL4) DIINN_CUDA_OK(h, cudaMalloc(&h->WL4, sizeof(wl4)));
    DIINN_CUDA_OK(h, cudaMemcpyAsync(h->WL4, wl4, sizeof(wl4), cudaMemcpyHostToDevice, s));
    // tensor path: W27[n = tap*3 + c][f] as mma.sync.m16n8k16 B fragments in the K permutation of last_conv_project_kernel
    auto bf16_bits = [](float v) -> uint32_t {  // round to nearest even (the weights are finite)
      uint32_t u;
      memcpy(&u, &v, 4);
      return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
    };
    auto bf16_val = [](uint32_t b) -> float {
      const uint32_t u = b << 16;
      float v;
      memcpy(&v, &u, 4);
      return v;
    };
    static thread_local uint32_t frag[4 * 4 * 8 * 32 * 2];
    for (int kb = 0; kb < 4; ++kb)
      for (int st = 0; st < 4; ++st)
        for (int j = 0; j < 8; ++j)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            const int n = 8 * (j & 3) + g;  // output column tap*3 + c; columns 27..31 are padding
            uint32_t regs[2];
            for (int half = 0; half < 2; ++half) {
              const int f = kb * 64 + half * 32 + t * 8 + 2 * st;
              uint32_t pk[2];
              for (int e = 0; e < 2; ++e) {
                const float wv = n < 27 ? tmp_wl[((n % 3) * kD + f + e) * 9 + n / 3] : 0.f;
                const uint32_t hi = bf16_bits(wv);
                pk[e] = j < 4 ? hi : bf16_bits(wv - bf16_val(hi));
              }
              regs[half] = pk[0] | (pk[1] << 16);
            }
            uint32_t* dst = frag + ((((kb * 4 + st) * 8 + j) * 32) + lane) * 2;
            dst[0] = regs[0], dst[1] = regs[1];
          }
    if (!h->WL27frag) DIINN_CUDA_OK(h, cudaMalloc(&h->WL27frag, sizeof(frag)));
    DIINN_CUDA_OK(h, cudaMemcpyAsync(h->WL27frag, frag, sizeof(frag), cudaMemcpyHostToDevice, s));
  }
  DIINN_CUDA_OK(h, cudaMemcpyAsync(h->bq_dev, sp.bq, sizeof(float) * kLayers * kD, cudaMemcpyHostToDevice, s));
  {
    // select-MMA variant of stage B: the Q-branch tiles of B_sel. Per (layer 1..3, half, 64-feature block) a K_sel x 64 fp16
    // tile whose every row is fp16(bq): a row of A_sel is one-hot, so whichever slot it selects the Q-branch accumulator
    // starts at the bias (rounded to fp16: an absolute 2^-12 |bq| in the sine argument, far below the fp16 operand noise).
    static thread_local uint16_t tab[3 * 2 * 2 * 32 * 64];
    for (int v = 0; v < 2; ++v) {
      const int ks = v == 0 ? 16 : 32;
      memset(tab, 0, sizeof(tab));
      for (int lh = 0; lh < 6; ++lh)
        for (int fb = 0; fb < 2; ++fb)
          for (int e = 0; e < 64; ++e) {
            const float b = sp.bq[lh / 2 + 1][(lh & 1) * 128 + fb * 64 + e];
            const float bc = b < -65504.f ? -65504.f : (b > 65504.f ? 65504.f : b);
            const uint16_t hi = __half_as_ushort(__float2half_rn(bc));
            uint16_t* t = tab + (static_cast<size_t>(lh * 2 + fb) * ks) * 64;
            for (int r = 0; r < ks; ++r) t[r * 64 + e] = hi;
          }
      const size_t bytes = sizeof(uint16_t) * 3 * 2 * 2 * ks * 64;
      if (!h->WSel16[v]) DIINN_CUDA_OK(h, cudaMalloc(&h->WSel16[v], bytes));
      DIINN_CUDA_OK(h, cudaMemcpyAsync(h->WSel16[v], tab, bytes, cudaMemcpyHostToDevice, s));
      DIINN_CUDA_OK(h, cudaStreamSynchronize(s));  // tab is reused
    }
  }
  DIINN_CUDA_OK(h, cudaStreamSynchronize(s));

  // TMA descriptors over the bf16 tiles (rows of 64 bf16 = 128 B, 128B swizzle applied by TMA on the way in)
  int rc;
  // (same element size and no arithmetic in a TMA copy: the bf16 descriptor type moves fp16 bits unchanged)
  for (int cg = 0; cg < 2; ++cg) {
    const uint32_t box_rows = cg == 0 ? 256 : 128;  // CTA pairs split every 256-row tile by N halves
    for (int f = 0; f < 2; ++f) {
      if ((rc = make_tmap_2d_bf16(h, &h->tmapWA[f][cg], h->WA16[f], 64, 4 * 9 * 256, 64, box_rows))) return rc;
      if ((rc = make_tmap_2d_bf16(h, &h->tmapWB[f][cg], h->WB16[f], 64, 3 * 2 * 4 * 256, 64, box_rows))) return rc;
    }
    if ((rc = make_tmap_2d_bf16(h, &h->tmapWAlo[cg], h->WA16lo, 64, 4 * 9 * 256, 64, box_rows))) return rc;
    if ((rc = make_tmap_2d_bf16(h, &h->tmapWBlo[cg], h->WB16lo, 64, 3 * 2 * 4 * 256, 64, box_rows))) return rc;
  }
  for (int v = 0; v < 2; ++v) {
    const uint32_t ks = v == 0 ? 16 : 32;
    if ((rc = make_tmap_2d_bf16(h, &h->tmapSelB[v], h->WSel16[v], 64, 3 * 2 * 2 * ks, 64, ks))) return rc;
  }
  if (init_q)
    for (int cg = 0; cg < 2; ++cg)
      if ((rc = make_tmap_2d_bf16(h, &h->tmapWQ0A[cg], h->WQ0A16, 64, 9 * 256, 64, cg == 0 ? 256 : 128))) return rc;
  h->has_weights = true;
  return DIINN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// feat (B,64,H,W) NCHW fp32|bf16, LR rows [r0,r1) -> (B, r1-r0, W, 64) in the 16-bit operand format FMT, the layout
// stage A's TMA im2col boxes read (one 128-byte row per LR pixel and tap); the split format also writes the fp16
// residual plane. Coalesced both ways through a padded smem tile.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int FMT>
__global__ void feat_to_nhwc_kernel(const T* __restrict__ feat, uint32_t* __restrict__ dst, uint32_t* __restrict__ dst_lo,
                                    int H, int W, int r0, int rows) {
  __shared__ float tile[kC][33];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // stage A may be scheduled while we drain (PDL)
  const int w0 = blockIdx.x * 32;
  const int row = blockIdx.y;  // relative to r0
  const int b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 256 threads: 8 warps
  for (int c = ty; c < kC; c += 8) {
    const int w = w0 + tx;
    float v = 0.f;
    if (w < W) v = static_cast<float>(feat[((static_cast<size_t>(b) * kC + c) * H + (r0 + row)) * W + w]);
    tile[c][tx] = v;
  }
  __syncthreads();
  // each thread writes 2 channels of one pixel: 32 threads cover one pixel's 64 channels (128 B)
  for (int p = ty; p < 32; p += 8) {
    const int w = w0 + p;
    if (w < W) {
      const float a = tile[2 * tx][p], c = tile[2 * tx + 1][p];
      const size_t o = ((static_cast<size_t>(b) * rows + row) * W + w) * (kC / 2) + tx;
      if constexpr (FMT == kFmtBf16) {
        const __nv_bfloat162 v = __floats2bfloat162_rn(a, c);
        dst[o] = *reinterpret_cast<const uint32_t*>(&v);
      } else {
        const float ac = fminf(fmaxf(a, -65504.f), 65504.f), cc = fminf(fmaxf(c, -65504.f), 65504.f);
        const __half2 v = __floats2half2_rn(ac, cc);
        dst[o] = *reinterpret_cast<const uint32_t*>(&v);
        if constexpr (FMT == kFmtSplit) {
          const float2 back = __half22float2(v);
          const __half2 l = __floats2half2_rn(ac - back.x, cc - back.y);
          dst_lo[o] = *reinterpret_cast<const uint32_t*>(&l);
        }
      }
    }
  }
}

int launch_feat_to_nhwc(Handle* h, const void* feat, int io_dtype, int fmt, int B, int H, int W, int r0, int r1, void* dst,
                        void* dst_lo, cudaStream_t s) {
  dim3 grid((W + 31) / 32, r1 - r0, B);
  uint32_t* d = static_cast<uint32_t*>(dst);
  uint32_t* dl = static_cast<uint32_t*>(dst_lo);
  if (fmt == kFmtSplit && !dst_lo) return fail(h, DIINN_ERR_BAD_ARG, "layout pass: the split format needs the residual plane");
#define DIINN_LAYOUT(T, FMTv) feat_to_nhwc_kernel<T, FMTv><<<grid, 256, 0, s>>>(static_cast<const T*>(feat), d, dl, H, W, r0, r1 - r0)
  if (io_dtype == DIINN_IO_F32) {
    if (fmt == kFmtBf16) DIINN_LAYOUT(float, kFmtBf16);
    else if (fmt == kFmtF16) DIINN_LAYOUT(float, kFmtF16);
    else DIINN_LAYOUT(float, kFmtSplit);
  } else {
    if (fmt == kFmtBf16) DIINN_LAYOUT(__nv_bfloat16, kFmtBf16);
    else if (fmt == kFmtF16) DIINN_LAYOUT(__nv_bfloat16, kFmtF16);
    else DIINN_LAYOUT(__nv_bfloat16, kFmtSplit);
  }
#undef DIINN_LAYOUT
  h->launches += 1;
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

}  // namespace diinn

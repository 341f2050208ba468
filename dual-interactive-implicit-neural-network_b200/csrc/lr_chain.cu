// K chain of decoder modes 1 and 2 at LR resolution (diinn.py:116-131).
//
//   mode 1: k_i = relu(K_i k_{i-1} + b_i)                 mode 2: k_i = relu(K_i [k_{i-1}; x] + b_i)
//
// Neither depends on the HR query, so k_i is a function of the LR pixel alone. Stage A has already left
//   P[l] = [ relu(K_0 x + b_0) | K_1[:,256:] x + b_1 | K_2[:,256:] x + b_2 | K_3[:,256:] x + b_3 ]      (x-part zero in mode 1)
// and this pass adds the k-facing 256x256 blocks in layer order:  P_i += WH_i . relu(P_{i-1}),  i = 1..3
// (block 0 is stored post-ReLU and relu is idempotent, so the same load works for every i). Stage B then runs
// unchanged with ZERO K rows: relu(0 + P_i) = k_i, q_i = k_i * sin(Q_i q_{i-1} + bq_i).
//
// fp32 path: a plain tiled SGEMM on CUDA cores (exact fp32 FMA). Tensor path: per chunk of <= 65536 LR pixels,
// one relu + bf16 conversion of block 0, then per layer the library's tcgen05 GEMM (gemm.cu, 128x256 tiles,
// fp32 accumulation) with a fused epilogue: P_i += acc, and the next layer's A operand bf16(relu(P_i)) written
// alongside (ChainEpilogue). Modes 1 / 2 are outside the benchmarked configuration.
#include <cuda_fp16.h>

#include "handle.h"

namespace diinn {

constexpr int64_t kChainChunk = 65536;  // LR pixels per tensor-path pass (multiple of 256): the A operand ping-pong (64 MB) stays in L2

// ---- fp32: P[m][256 i + n] += sum_k relu(P[m][256 (i-1) + k]) * WH[n][k], 64x64 tiles, 256 threads, 4x4 per thread
__global__ void __launch_bounds__(256) lr_chain_fp32_kernel(float* __restrict__ P, const float* __restrict__ WH,
                                                            int64_t M, int layer) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  const float* Pin = P + (layer - 1) * kD;
  for (int k0 = 0; k0 < kD; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      const int r = e >> 4, k = e & 15;
      const int64_t m = m0 + r;
      As[k][r] = m < M ? fmaxf(Pin[m * kPCols + k0 + k], 0.f) : 0.f;
      Bs[k][r] = WH[static_cast<size_t>(n0 + r) * kD + k0 + k];
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
  float* Pout = P + layer * kD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) Pout[m * kPCols + n0 + tx * 4 + j] += acc[i][j];
  }
}

int run_lr_chain_fp32(Handle* h, float* P, int64_t M, cudaStream_t s) {
  for (int layer = 1; layer <= 3; ++layer) {
    dim3 grid(static_cast<unsigned>((M + 63) / 64), kD / 64);
    lr_chain_fp32_kernel<<<grid, 256, 0, s>>>(P, h->WH32 + static_cast<size_t>(layer - 1) * kD * kD, M, layer);
    h->launches += 1;
  }
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

// ---- tensor path
// A16[r][k] = bf16(relu(P[m0 + r][256 (layer-1) + k])), zero rows for r >= rows (padding up to the GEMM's M tile).
// kP16: P holds fp16 rows (the select-MMA variant of stage B).
template <bool kP16>
__global__ void __launch_bounds__(256) lr_chain_prep_kernel(const void* __restrict__ Pv, __nv_bfloat16* __restrict__ A16,
                                                            int64_t m0, int64_t rows, int64_t rows_pad, int layer) {
  const int64_t total = rows_pad * (kD / 4);
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = g / (kD / 4);
    const int k = static_cast<int>(g % (kD / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      const int64_t off = (m0 + r) * kPCols + (layer - 1) * kD + k;
      if constexpr (kP16) {
        const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __half*>(Pv) + off);
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        v = *reinterpret_cast<const float4*>(static_cast<const float*>(Pv) + off);
      }
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
    __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&lo);
    pk.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(A16 + r * kD + k) = pk;
  }
}

static int64_t chain_rows_pad(int64_t M) {
  const int64_t c = M < kChainChunk ? M : kChainChunk;
  return (c + 255) / 256 * 256;
}

size_t lr_chain_scratch_bytes(int64_t M) {
  const int64_t rp = chain_rows_pad(M);
  return 2 * static_cast<size_t>(rp) * kD * sizeof(__nv_bfloat16);  // A operand ping-pong
}

int run_lr_chain_umma(Handle* h, float* P, int64_t M, void* scratch, cudaStream_t s, bool p16) {
  const int64_t rp_max = chain_rows_pad(M);
  __nv_bfloat16* A16[2] = {static_cast<__nv_bfloat16*>(scratch), static_cast<__nv_bfloat16*>(scratch) + rp_max * kD};
  const int blocks_cap = (h->sm_count > 0 ? h->sm_count : 148) * 8;
  for (int64_t m0 = 0; m0 < M; m0 += kChainChunk) {
    const int64_t rows = (M - m0 < kChainChunk) ? M - m0 : kChainChunk;
    const int64_t rp = (rows + 255) / 256 * 256;
    // layer 1's A operand from stage A's block 0; the GEMM epilogue of layer i then produces layer i+1's
    const int64_t nb = (rp * (kD / 4) + 255) / 256;
    const unsigned nblk = static_cast<unsigned>(nb < blocks_cap ? nb : blocks_cap);
    if (p16) lr_chain_prep_kernel<true><<<nblk, 256, 0, s>>>(P, A16[0], m0, rows, rp, 1);
    else lr_chain_prep_kernel<false><<<nblk, 256, 0, s>>>(P, A16[0], m0, rows, rp, 1);
    h->launches += 1;
    for (int layer = 1; layer <= 3; ++layer) {
      ChainEpilogue ce{};
      ce.P = P, ce.A_next = layer < 3 ? A16[layer & 1] : nullptr, ce.m0 = m0, ce.rows = rows, ce.layer = layer;
      ce.p16 = p16 ? 1 : 0;
      int rc = launch_umma_selftest(h, A16[(layer - 1) & 1], h->WH16 + static_cast<size_t>(layer - 1) * kD * kD, nullptr,
                                    static_cast<int>(rp), kD, kD, 2, s, &ce);
      if (rc) return rc;
    }
  }
  DIINN_CUDA_OK(h, cudaGetLastError());
  return DIINN_OK;
}

}  // namespace diinn

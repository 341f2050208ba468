// Eval glue on the device: PSNR exactly as the reference evaluates it (calc_psnr, src/models/sr_module.py:21-38).
// HBM-bound: reads sr and hr once (2 x B*C*H*W elements), one fp64 atomic per CTA.
#include <cmath>

#include "common.cuh"
#include "handle.h"

namespace diinn {

template <typename T>
__device__ __forceinline__ float ld_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float ld_as_float<float>(const float* p, int64_t i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}

// One thread per valid pixel (b, y, x) of the shaved window; gray != 0: the C = 3 channels collapse into luma first
// (diff.mul(convert).sum(dim=1), fp32, channel order), else every channel contributes its own squared difference.
template <typename T>
__global__ void __launch_bounds__(256) psnr_sse_kernel(const T* __restrict__ sr, const T* __restrict__ hr, int B, int C,
                                                       int H, int W, int shave, int gray, float rgb_range,
                                                       double* __restrict__ acc) {
  const int vh = H - 2 * shave, vw = W - 2 * shave;
  const int64_t total = static_cast<int64_t>(B) * vh * vw;
  const float conv[3] = {65.738f / 256.f, 129.057f / 256.f, 25.064f / 256.f};
  double local = 0.0;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(g % vw) + shave;
    const int y = static_cast<int>((g / vw) % vh) + shave;
    const int b = static_cast<int>(g / (static_cast<int64_t>(vw) * vh));
    const int64_t base = (static_cast<int64_t>(b) * C * H + y) * W + x;
    if (gray) {
      float d = 0.f;
      for (int c = 0; c < 3; ++c) {
        const int64_t i = base + static_cast<int64_t>(c) * H * W;
        d = __fadd_rn(d, __fmul_rn(__fdiv_rn(__fsub_rn(ld_as_float(sr, i), ld_as_float(hr, i)), rgb_range), conv[c]));
      }
      local += static_cast<double>(d) * d;
    } else {
      for (int c = 0; c < C; ++c) {
        const int64_t i = base + static_cast<int64_t>(c) * H * W;
        const float d = __fdiv_rn(__fsub_rn(ld_as_float(sr, i), ld_as_float(hr, i)), rgb_range);
        local += static_cast<double>(d) * d;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += part[i];
    atomicAdd(acc, s);
  }
}

}  // namespace diinn

using namespace diinn;

extern "C" int diinn_psnr(diinn_handle* h, const void* sr, const void* hr, int dtype, int B, int C, int H, int W,
                          int dataset, int scale, float rgb_range, double* psnr_host, void* stream) {
  if (!h) return DIINN_ERR_BAD_ARG;
  if (!sr || !hr || !psnr_host) return fail(h, DIINN_ERR_BAD_ARG, "null pointer");
  if (dtype != DIINN_IO_F32 && dtype != DIINN_IO_BF16) return fail(h, DIINN_ERR_BAD_DTYPE, "dtype");
  if (B < 1 || C < 1 || H < 1 || W < 1 || scale < 0 || !(rgb_range > 0.f)) return fail(h, DIINN_ERR_BAD_SHAPE, "bad shape");
  if (dataset < 0 || dataset > 2) return fail(h, DIINN_ERR_BAD_ARG, "dataset must be 0 (None), 1 (benchmark) or 2 (div2k)");
  int shave = 0, gray = 0;
  if (dataset == 1) {
    shave = scale;
    gray = C > 1;
    if (gray && C != 3) return fail(h, DIINN_ERR_BAD_SHAPE, "luma conversion needs 3 channels");
  } else if (dataset == 2) {
    shave = scale + 6;
  }
  // the reference slices [shave:-shave]: with shave == 0 that is an EMPTY slice (mean of nothing = nan); mirror it
  if (dataset != 0 && (shave == 0 || H - 2 * shave < 1 || W - 2 * shave < 1)) {
    *psnr_host = std::nan("");
    return DIINN_OK;
  }
  cudaSetDevice(h->cfg.device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!h->psnr_acc) DIINN_CUDA_OK(h, cudaMalloc(&h->psnr_acc, sizeof(double)));
  DIINN_CUDA_OK(h, cudaMemsetAsync(h->psnr_acc, 0, sizeof(double), s));
  const int64_t px = static_cast<int64_t>(B) * (H - 2 * shave) * (W - 2 * shave);
  int64_t blocks = (px + 255) / 256;
  const int64_t cap = static_cast<int64_t>(h->sm_count > 0 ? h->sm_count : 148) * 8;  // grid-stride, 8 CTAs per SM
  if (blocks > cap) blocks = cap;
  if (dtype == DIINN_IO_F32)
    psnr_sse_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const float*>(sr), static_cast<const float*>(hr),
                                                                        B, C, H, W, shave, gray, rgb_range, h->psnr_acc);
  else
    psnr_sse_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(sr), static_cast<const __nv_bfloat16*>(hr), B, C, H, W, shave, gray, rgb_range,
        h->psnr_acc);
  DIINN_CUDA_OK(h, cudaGetLastError());
  ++h->launches;
  double sse = 0.0;
  DIINN_CUDA_OK(h, cudaMemcpyAsync(&sse, h->psnr_acc, sizeof(double), cudaMemcpyDeviceToHost, s));
  DIINN_CUDA_OK(h, cudaStreamSynchronize(s));
  const double n = static_cast<double>(px) * (gray ? 1 : C);
  *psnr_host = -10.0 * std::log10(sse / n);
  return DIINN_OK;
}

// tcgen05 self-test GEMM: D(MxN fp32) = A(MxK bf16) * B(NxK bf16)^T.
// Exercises exactly the plumbing the fused decoder kernels rely on -- TMA tensor maps with 128B swizzle, the K-major
// SWIZZLE_128B UMMA shared-memory descriptor and its +32 B K-step, the kind::f16 instruction descriptor, TMEM
// allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b, and (cta_group 2) the CTA-pair variants: leader-CTA
// barrier for both CTAs' TMA loads, multicast commit, M=256 split across the pair, B split by N halves.
// One 128(x2) x 256 output tile per CTA (pair); 4-stage K pipeline of 64-element chunks.
#include "handle.h"
#include "ptx.cuh"

namespace diinn {
using namespace ptx;

template <int CG>
__global__ void __launch_bounds__(192, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     float* __restrict__ D, int M, int N, int K) {
  constexpr int STAGES = 4;
  constexpr int A_BYTES = 128 * 128;
  constexpr int B_ROWS = 256 / CG;
  constexpr int B_BYTES = B_ROWS * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int row0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * 256;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_ptr, 256);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int nk = K / 64;

  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < nk; ++kb) {
      const int st = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty[st], ph ^ 1);
      if (leader) mbar_arrive_expect_tx(&full[st], (A_BYTES + B_BYTES) * CG);
      if constexpr (CG == 1) {
        tma_load_2d(sA + st * A_BYTES, &tmA, &full[st], kb * 64, row0);
        tma_load_2d(sB + st * B_BYTES, &tmB, &full[st], kb * 64, n0);
      } else {
        tma_load_2d_2sm(sA + st * A_BYTES, &tmA, &full[st], kb * 64, row0);
        tma_load_2d_2sm(sB + st * B_BYTES, &tmB, &full[st], kb * 64, n0 + static_cast<int>(cta_rank) * B_ROWS);
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, 256);
    for (int kb = 0; kb < nk; ++kb) {
      const int st = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full[st], ph);
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA + st * A_BYTES), b0 = smem_u32(sB + st * B_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16<CG>(tmem_base, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
      umma_commit<CG>(&empty[st]);
    }
    umma_commit<CG>(tmem_full);
  } else if (warp >= 2) {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = row0 + q * 32 + lane;
    float* drow = D + static_cast<size_t>(row) * N + n0;
    for (int c0 = 0; c0 < 256; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (row < M) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(drow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, 256);
}

int launch_umma_selftest(Handle* h, const void* A, const void* B, float* D, int M, int N, int K, int cta_group,
                         cudaStream_t s) {
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_bf16(h, &tmA, A, K, M, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(h, &tmB, B, K, N, 64, 256 / cta_group))) return rc;
  const size_t smem = 4 * (128 * 128 + 256 / cta_group * 128) + 128 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(M / 128, N / 256, 1);
  cfg.blockDim = dim3(192, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cta_group == 1) {
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, umma_selftest_kernel<1>, tmA, tmB, D, M, N, K));
  } else {
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(umma_selftest_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, umma_selftest_kernel<2>, tmA, tmB, D, M, N, K));
  }
  h->launches += 1;
  return DIINN_OK;
}

}  // namespace diinn

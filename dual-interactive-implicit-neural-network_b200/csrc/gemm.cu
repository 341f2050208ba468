// tcgen05 self-test GEMM: D(MxN fp32) = A(MxK bf16) * B(NxK bf16)^T.
// Exercises exactly the plumbing the fused decoder kernels rely on -- TMA tensor maps with 128B swizzle, the K-major
// SWIZZLE_128B UMMA shared-memory descriptor and its +32 B K-step, the kind::f16 instruction descriptor, TMEM
// allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b, and (cta_group 2) the CTA-pair variants: leader-CTA
// barrier for both CTAs' TMA loads, multicast commit, M=256 split across the pair, B split by N halves.
// One 128(x2) x 256 output tile per CTA (pair); 4-stage K pipeline of 64-element chunks.
#include <cuda_fp16.h>

#include "handle.h"
#include "ptx.cuh"

namespace diinn {
using namespace ptx;

// three operand stages (3 x 32 KB with CTA pairs): with 256 TMEM columns per CTA two CTAs still share an SM, so one tile's
// epilogue runs under the other's loads and MMAs (the LR chain of modes 1 / 2 is epilogue / HBM bound), and the K loop
// (9 steps for init_q's 576-wide operands) has two loads in flight behind the MMAs instead of one
constexpr int kGemmStages = 3;

template <int CG>
__global__ void __launch_bounds__(192, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     float* __restrict__ D, int M, int N, int K, int f16acc, const ChainEpilogue ce) {
  constexpr int STAGES = kGemmStages;
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
  // LR chain launches are programmatic dependents of one another (PDL): the prologue above overlaps the predecessor's tail
  // and nothing it wrote is touched before this wait (a no-op for ordinary launches)
  grid_dep_launch();
  grid_dep_wait();
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
    const uint32_t idesc = f16acc ? umma_idesc_f16_acc16(128 * CG, 256) : umma_idesc_bf16(128 * CG, 256);
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
    if (f16acc) {
      // fp16 accumulators: one value per TMEM column, read two columns per register with .pack::16b
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t v[16];
        tmem_ld32_pack16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __half2 hv = *reinterpret_cast<const __half2*>(&v[j]);
            drow[c0 + 2 * j] = __low2float(hv);
            drow[c0 + 2 * j + 1] = __high2float(hv);
          }
        }
      }
    } else if (ce.P != nullptr && ce.p16) {
      // LR chain over the fp16 P of the select-MMA variant: P_layer = fp16(P_layer + acc), next A operand = bf16(relu(.)) of
      // the unrounded sum. 64 columns (128 B) per pass, loads issued before the accumulators are read.
      const bool valid = row < ce.rows;
      __half* prow = reinterpret_cast<__half*>(ce.P) + (ce.m0 + row) * kPCols + ce.layer * kD + n0;
      __nv_bfloat16* arow = ce.A_next ? ce.A_next + static_cast<size_t>(row) * kD + n0 : nullptr;
      for (int c0 = 0; c0 < 256; c0 += 64) {
        uint4 pv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pv[j] = valid ? *reinterpret_cast<const uint4*>(prow + c0 + 8 * j) : make_uint4(0u, 0u, 0u, 0u);
        uint32_t v[64];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0 + 16 * g, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * g]));
        tmem_ld_wait();
        uint32_t pk[32], ph[32];
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(pv);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a, b;
          unpack_f16x2(pw[j], a, b);
          a += __uint_as_float(v[2 * j]), b += __uint_as_float(v[2 * j + 1]);
          ph[j] = pack_f16x2_sat(a, b);
          pk[j] = pack_bf16x2(fmaxf(a, 0.f), fmaxf(b, 0.f));                   // padding rows of A_next stay zero
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(prow + c0 + 8 * j) = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
        }
        if (arow && row < M) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(arow + c0 + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
    } else if (ce.P != nullptr) {
      // LR chain: P_layer += acc, next layer's A operand = bf16(relu(P_layer)); a thread owns one row (64 B per step)
      const bool valid = row < ce.rows;
      float* prow = ce.P + (ce.m0 + row) * kPCols + ce.layer * kD + n0;
      __nv_bfloat16* arow = ce.A_next ? ce.A_next + static_cast<size_t>(row) * kD + n0 : nullptr;
      // 64 columns per pass: all 16 P loads of a pass are issued before the accumulators are read, so their latency
      // overlaps (a thread's 64 B pieces are 4 KB apart: nothing to coalesce, only latency to hide)
      for (int c0 = 0; c0 < 256; c0 += 64) {
        float4 pv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          pv[j] = valid ? *reinterpret_cast<const float4*>(prow + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[64];
#pragma unroll
        for (int g = 0; g < 4; ++g)
          tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0 + 16 * g, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * g]));
        tmem_ld_wait();
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float4 o = pv[j];
          if (valid) {
            o.x += __uint_as_float(v[4 * j]), o.y += __uint_as_float(v[4 * j + 1]);
            o.z += __uint_as_float(v[4 * j + 2]), o.w += __uint_as_float(v[4 * j + 3]);
            *reinterpret_cast<float4*>(prow + c0 + 4 * j) = o;
          }
          pk[2 * j] = pack_bf16x2(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f));          // padding rows of A_next stay zero
          pk[2 * j + 1] = pack_bf16x2(fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
        }
        if (arow && row < M) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(arow + c0 + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
    } else {
      // plain GEMM (the tcgen05 self-test / probe entry): D = acc
      for (int c0 = 0; c0 < 256; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (row < M) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(drow + c0 + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-pipe pace probe: `iters` back-to-back UMMAs (M = 128 x CG, N = n_cols, K = 16) on resident smem operands, no
// TMA, no epilogue. Reports clock64 cycles per MMA as seen by the issuing thread of every leader CTA. Used to separate
// "the tensor pipe is this fast here" from pipeline stalls when reading the fused kernels' traces.
// ---------------------------------------------------------------------------------------------------------
template <int CG>
__global__ void __launch_bounds__(640, 1) umma_pace_kernel(int n_cols, int iters, float* __restrict__ cyc_per_mma, int noise) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;              // 4 K-chunk tiles of [128 x 128 B]
  uint8_t* sB = smem + 4 * 16384;  // 4 tiles of [256/CG x 128 B]
  uint8_t* sN = sB + 4 * 32768;    // 16 KB scratch for the noise warps' stores
  uint64_t* bar = reinterpret_cast<uint64_t*>(sN + 16384);
  uint64_t* ring = bar + 1;  // 6 barriers that only ever receive commits (noise bit 4: fused-kernel loop structure)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ring + 6);
  volatile int* done = reinterpret_cast<volatile int*>(tmem_ptr + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = CG == 1 || cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) {
    // noise bit 3: pseudo-random bf16 operands in [-2,2) instead of zeros (data-dependent power)
    uint32_t x = (i + 1) * 2654435761u + blockIdx.x * 40503u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    const uint32_t lo = 0x3c00u | (x & 0x83ffu), hi = 0x3c00u | ((x >> 16) & 0x83ffu);
    reinterpret_cast<uint32_t*>(smem)[i] = (noise & 8) ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) *done = 0;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    for (int i = 0; i < 6; ++i) mbar_init(&ring[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_ptr, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0 && lane == 0 && leader) {
    const uint32_t idesc = umma_idesc_bf16(128 * CG, n_cols);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int kc = (i >> 2) & 3, k = i & 3, slot = (i >> 4) & 1;
      if ((noise & 16) && k == 0) {
        // what the fused kernels do per K-chunk: a barrier wait that passes immediately, then the fence
        mbar_wait(&ring[5], 1);
        tc_fence_after();
      }
      umma_bf16<CG>(tmem_base + slot * 256, umma_desc_sw128(a0 + kc * 16384 + k * 32),
                    umma_desc_sw128(b0 + kc * (32768 / CG) + k * 32), idesc, i ? 1u : 0u);
      if ((noise & 16) && k == 3) umma_commit<CG>(&ring[(i >> 2) % 5]);
    }
    if constexpr (CG == 2) umma_commit_2sm_local(bar); else umma_commit<1>(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (!(noise & 32)) cyc_per_mma[blockIdx.x / CG] = static_cast<float>(t1 - t0) / iters;
  }
  if (warp == 0 && lane == 0) {  // follower CTAs learn about completion through the teardown barrier only
    if (leader) *done = 1;
  }
  // noise warps 4..19 (the fused kernel's epilogue population): bit0 = tcgen05.ld of the idle TMEM half, bit1 = 128-bit
  // shared stores, bit2 = MUFU, until the leader's MMAs are done (followers: fixed trip count)
  if (warp >= 4 && (noise & 32)) {
    // TMEM read bandwidth probe: (noise >> 8) warps per lane quarter stream tcgen05.ld.32x32b.x16 back to back
    const int per_quarter = (noise >> 8) & 7;
    if (((warp - 4) >> 2) < per_quarter) {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      uint32_t v[16], w[16];
      float acc = 0.f;
      const long long t0 = clock64();
      if (noise & 128) {  // x32 columns of 16-bit data packed into 16 registers, two per step (= 64 columns per step)
        for (int it = 0; it < 512; ++it) {
          tmem_ld32_pack16(taddr + ((it * 64) & 255), v);
          tmem_ld32_pack16(taddr + ((it * 64 + 32) & 255), w);
          tmem_ld_wait();
          acc += __uint_as_float(v[it & 15]) + __uint_as_float(w[it & 15]);
        }
      } else if (noise & 64) {  // one x32 load instead of two x16 loads per step
        uint32_t u[32];
        for (int it = 0; it < 512; ++it) {
          tmem_ld32(taddr + ((it * 32) & 255), u);
          tmem_ld_wait();
          acc += __uint_as_float(u[it & 31]);
        }
      } else {
        for (int it = 0; it < 512; ++it) {
          tmem_ld16(taddr + ((it * 32) & 255), v);
          tmem_ld16(taddr + ((it * 32 + 16) & 255), w);
          tmem_ld_wait();
          acc += __uint_as_float(v[it & 15]) + __uint_as_float(w[it & 15]);
        }
      }
      const long long t1 = clock64();
      if (warp == 4 && lane == 0) cyc_per_mma[blockIdx.x / CG] = static_cast<float>(t1 - t0) / 1024.f;
      if (acc == 12345.678f) cyc_per_mma[0] = acc;
    }
  } else if (warp >= 4 && (noise & 7)) {
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 3) * 16;
    const uint32_t saddr = smem_u32(sN) + (warp - 4) * 1024 + lane * 16 + ((lane >> 3) & 1) * 0;
    float acc = 0.f;
    uint32_t v[16];
    for (int it = 0; it < (leader ? (1 << 30) : iters / 2); ++it) {
      if (noise & 1) {
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        acc += __uint_as_float(v[it & 15]);
      }
      if (noise & 2) st_shared_v4(saddr + ((it & 1) << 9), it, it, it, it);
      if (noise & 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += __sinf(acc + j);
      }
      if (leader && *done) break;
    }
    if (acc == 12345.678f) cyc_per_mma[0] = acc;
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, 512);
}

int launch_umma_pace(Handle* h, int cta_group, int n_cols, int iters, int n_ctas, float* cyc_per_mma, int noise,
                     cudaStream_t s) {
  const size_t smem = 4 * 16384 + 4 * 32768 + 16384 + 128 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_ctas, 1, 1);
  cfg.blockDim = dim3(640, 1, 1);
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
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(umma_pace_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, umma_pace_kernel<1>, n_cols, iters, cyc_per_mma, noise));
  } else {
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(umma_pace_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, umma_pace_kernel<2>, n_cols, iters, cyc_per_mma, noise));
  }
  h->launches += 1;
  return DIINN_OK;
}

int launch_umma_selftest(Handle* h, const void* A, const void* B, float* D, int M, int N, int K, int cta_group,
                         cudaStream_t s, const ChainEpilogue* chain) {
  const ChainEpilogue ce = chain ? *chain : ChainEpilogue{};
  // cta_group 1|2: bf16 operands, fp32 accumulators; 11|12: the same GEMM with fp16 operands and fp16 accumulators
  const int f16acc = cta_group >= 10 ? 1 : 0;
  cta_group = cta_group % 10;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_2d_bf16(h, &tmA, A, K, M, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(h, &tmB, B, K, N, 64, 256 / cta_group))) return rc;
  const size_t smem = kGemmStages * (128 * 128 + 256 / cta_group * 128) + 128 + 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(M / 128, N / 256, 1);
  cfg.blockDim = dim3(192, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ce.P != nullptr && h->pdl) ? 2 : 1;  // the LR chain's kernels only (each waits on its predecessor, see the kernel)
  if (cta_group == 1) {
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(umma_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, umma_selftest_kernel<1>, tmA, tmB, D, M, N, K, f16acc, ce));
  } else {
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(umma_selftest_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, umma_selftest_kernel<2>, tmA, tmB, D, M, N, K, f16acc, ce));
  }
  h->launches += 1;
  return DIINN_OK;
}

}  // namespace diinn

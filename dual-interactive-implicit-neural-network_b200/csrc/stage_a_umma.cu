// Stage A of the DIINN query decoder on tcgen05 tensor cores: everything in the K branch that multiplies the
// unfolded feature x_l depends only on the LR pixel l (diinn.py:133,136 with x = unfold3x3(feat), diinn.py:168), and
// "unfold 3x3 -> 1x1 conv" is a 3x3 zero-padded convolution. So per LR pixel, once:
//
//   P[l] (1024 fp32) = [ relu(K0 x_l + b0) | K_i[:, 256:] x_l + b_i , i = 1..3 ]      (implicit GEMM, K = 9 taps x 64 ch)
//
// A operand: feat as NHWC bf16; for tap (kh,kw) the K-chunk of LR pixel (h,w) is the 128-byte channel vector of pixel
// (h+kh-1, w+kw-1), fetched as one TMA 4-D box (64 ch x 16 w x 8 h x 1 b) per tap whose out-of-bounds zero fill IS the
// unfold's zero padding. B operand: the stacked weight matrix, K re-ordered to tap*64 + c (pack.cu), streamed from L2
// in [256 x 64] bf16 stages. Both operands travel through ONE ring of (A tap | W tap) stages: the tap tiles of a pixel
// tile are re-fetched for each of the four N-blocks (4x the L2->smem traffic, 31 B/clk per SM) so that no tile-sized A
// buffer has to be filled before a tile's first MMA -- the ring runs ahead across N-blocks and tiles alike. (The first
// version kept the whole 144 KB A tile resident and serialised its load against the MMAs.) D: two 256-column TMEM slots, one N-block each, drained by 4 epilogue warps that add the
// bias, apply ReLU to N-block 0, stage 32x32 fp32 blocks in 128B-swizzled shared memory and hand them to TMA stores
// (4-D box over P = (1024 cols, W, LR rows, B): the store clips rows/columns outside the image by itself).
//
// 256 threads: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA), warp 2 = TMEM allocator, warps 4..7 = epilogue.
// CG=2 runs CTA pairs (cta_group::2, M = 256 = two 8x16 LR patches, B split by N halves between the CTAs).
//
// Operand formats (FMT, handle.h): bf16, fp16, or the fp16 hi + lo SPLIT of the fp32-precision path, where every (n-block,
// tap) becomes three ring stages -- (x_hi, w_hi), (x_lo, w_hi), (x_hi, w_lo) -- accumulated into the same TMEM columns.
// kP16: P is written as fp16 (saturating), 2 KB per LR pixel -- the form stage B's select-MMA variant consumes as a tensor-core
// operand; the epilogue then packs 64 columns per staging round, so a TMA store moves a full 128-byte swizzle row per pixel
// and an N-block takes 4 store rounds instead of 8.
// Small launches (an 8-way row shard of a DIV2K image has 96 pixel tiles for 74 CTA pairs) split every tile's four
// N-blocks into 2 or 4 work items so that the last wave is not mostly idle (Geo::nsplit).
#include <cstdlib>
#include <cstring>

#include "handle.h"
#include "ptx.cuh"

namespace diinn {
using namespace ptx;

namespace sa {
constexpr int kPatchH = 8, kPatchW = 16;
constexpr int kTapBytes = 128 * 128;            // 16 KB: one tap of a 128-pixel tile (128 rows x 64 ch bf16)
constexpr int kRingBytes = 192 * 1024;
constexpr int kThreads = 256;
constexpr int kStoreBytes = 32 * 128;            // per epilogue warp: 32 rows x 32 fp32, 128B swizzle

template <int CG>
struct Cfg {
  static constexpr int kWRows = 256 / CG;                       // this CTA's share of an N-block's weight rows
  static constexpr int kWBytes = kWRows * 128;
  static constexpr int kStageBytes = kTapBytes + kWBytes;       // [A tap | W tap]: 32 KB (CG=2) or 48 KB (CG=1)
  static constexpr int kStages = kRingBytes / kStageBytes;      // 6 (CG=2) or 4 (CG=1)
};

struct BiasParams {
  float b[kPCols];
};

struct Smem {
  uint8_t store[4][kStoreBytes];  // must stay first: 1024-byte aligned (kRingBytes is a multiple of 1024)
  uint64_t w_full[6];
  uint64_t w_empty[6];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_ptr;
};
constexpr size_t kSmemBytes = kRingBytes + sizeof(Smem);
static_assert(kSmemBytes <= 232448, "exceeds 227 KB of dynamic shared memory");

struct Geo {
  int B, H, W;            // feature map
  int fr0;                // first LR row held by the NHWC copy
  int lr_row0, lr_rows;   // LR rows to produce (rows of P per image)
  int tiles_y, n_txp, n_work;  // n_work = pixel tiles x nsplit
  int nsplit;             // work items per pixel tile: each covers 4 / nsplit consecutive N-blocks (1, 2 or 4)
  int no_relu0;           // LIIF's imnet: block 0 is the feature part of the first Linear, its ReLU comes after the per-query part
  // Matrix mode (init_q=True, csrc/init_q.cu): the A operand is an explicit (rows x 576) 16-bit matrix -- one row per HR pixel,
  // K in tap-major order -- viewed as an "image" W pixels wide whose tap t is columns [64 t, 64 t + 64) of the same row
  // instead of a neighbouring pixel; `nblocks` 256-column N-blocks (4: the x-facing K blocks, 1: Q.0 on the gate).
  int matrix, nblocks;
  // matrix mode, N-block 0: the finished q_0 = relu(.) * sin(q0_arg[row][col]) (q0_arg = Q.0 s + bq_0, fp32 (rows, 256))
  const float* q0_arg;
};

template <int CG, int FMT, bool kP16>
__global__ void __launch_bounds__(kThreads, 1)
stage_a_umma_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmFlo,
                    const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
                    const __grid_constant__ CUtensorMap tmP, const __grid_constant__ BiasParams bias, const Geo g,
                    int* __restrict__ err_flag) {
  using C = Cfg<CG>;
  constexpr int kTerms = FMT == kFmtSplit ? 3 : 1;  // ring stages per (n-block, tap)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_ring = smem;
  Smem& sm = *reinterpret_cast<Smem*>(smem + kRingBytes);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = rank == 0;
  const int unit_id = blockIdx.x / CG, n_units = gridDim.x / CG;

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) atomicExch(err_flag, 2);
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmF);
    prefetch_tensormap(&tmW);
    prefetch_tensormap(&tmP);
    if constexpr (kTerms == 3) {
      prefetch_tensormap(&tmFlo);
      prefetch_tensormap(&tmWlo);
    }
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&sm.w_full[i], 1);
      mbar_init(&sm.w_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tmem_full[i], 1);
      mbar_init(&sm.tmem_empty[i], 4 * CG);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(&sm.tmem_ptr, 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_ptr;
  const int per_img = g.tiles_y * g.n_txp;
  const int nb_cnt = g.nblocks / g.nsplit;  // N-blocks per work item
  grid_dep_launch();  // stage B's CTAs may be scheduled as ours retire ...
  grid_dep_wait();    // ... and we read feat_nhwc only once the layout kernel has completed (PDL, see ptx.cuh)

  if (warp == 0) {
    {  // whole warp walks the loops (uniform registers); one elected lane issues the TMA / expect_tx instructions
      uint32_t it = 0;
      int t = 0;
      for (int work = unit_id; work < g.n_work; work += n_units, ++t) {
        const int tile = work / g.nsplit, part = work - tile * g.nsplit;
        const int b = tile / per_img;
        const int rem = tile - b * per_img;
        const int ty = rem / g.n_txp, txp = rem - ty * g.n_txp;
        const int h0 = g.lr_row0 + ty * kPatchH;
        const int w0 = (txp * CG + rank) * kPatchW;
        const int s_beg = part * nb_cnt * 9, s_end = s_beg + nb_cnt * 9;
        // (n-block, term, tap): the A tap tile is re-fetched per n-block and term (L2 hits). Split format: the small terms
        // (x_lo, w_hi) and (x_hi, w_lo) of all nine taps first, then (x_hi, w_hi) -- the tensor core truncates when it adds
        // an MMA's result to the accumulator, so the adds at full magnitude should be as few as possible (9, not 27).
#pragma unroll 1
        for (int sq = s_beg * kTerms; sq < s_end * kTerms; ++sq, ++it) {
          const int nbq = sq / (9 * kTerms), rq = sq - nbq * 9 * kTerms;
          const int term = rq / 9, tap = rq - term * 9;
          const int s36 = nbq * 9 + tap;
          const int c0 = g.matrix ? tap * kC : 0;
          const int c1 = g.matrix ? w0 : w0 + tap % 3 - 1, c2 = g.matrix ? h0 : h0 + tap / 3 - 1 - g.fr0;
          {
            const CUtensorMap* mf = (kTerms == 3 && term == 0) ? &tmFlo : &tmF;
            const CUtensorMap* mw = (kTerms == 3 && term == 1) ? &tmWlo : &tmW;
            const int st = it % C::kStages;
            mbar_wait(&sm.w_empty[st], ((it / C::kStages) & 1) ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&sm.w_full[st], C::kStageBytes * CG);
              uint8_t* dst = s_ring + st * C::kStageBytes;
              if constexpr (CG == 1) {
                tma_load_4d(dst, mf, &sm.w_full[st], c0, c1, c2, b);
                tma_load_2d(dst + kTapBytes, mw, &sm.w_full[st], 0, s36 * 256);
              } else {
                tma_load_4d_2sm(dst, mf, &sm.w_full[st], c0, c1, c2, b);
                tma_load_2d_2sm(dst + kTapBytes, mw, &sm.w_full[st], 0, s36 * 256 + rank * 128);
              }
            }
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) {  // whole warp loops, tcgen05 instructions under elect_one() (see stage_b_umma.cu)
      constexpr uint32_t idesc = FMT == kFmtBf16 ? umma_idesc_bf16(128 * CG, 256) : umma_idesc_f16(128 * CG, 256);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t it = 0, slot_use = 0;
      int t = 0;
      for (int work = unit_id; work < g.n_work; work += n_units, ++t) {
#pragma unroll 1
        for (int nb = 0; nb < nb_cnt; ++nb, ++slot_use) {
          const int slot = slot_use & 1;
          const uint32_t use = slot_use >> 1;  // how many times this slot was used before
          if constexpr (CG == 2) mbar_wait_cluster(&sm.tmem_empty[slot], (use & 1) ^ 1);
          else mbar_wait(&sm.tmem_empty[slot], (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_u + slot * 256;
#pragma unroll 1
          for (int tap = 0; tap < 9 * kTerms; ++tap, ++it) {
            const int st = it % C::kStages;
            mbar_wait(&sm.w_full[st], (it / C::kStages) & 1);
            tc_fence_after();
            const uint32_t a0 = smem_u32(s_ring + st * C::kStageBytes);
            const uint32_t b0 = a0 + kTapBytes;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16<CG>(d_tmem, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc,
                              (tap | k) != 0 ? 1u : 0u);
              umma_commit<CG>(&sm.w_empty[st]);
              if (tap == 9 * kTerms - 1) umma_commit<CG>(&sm.tmem_full[slot]);
            }
            __syncwarp();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int quarter = warp & 3;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t sbuf = smem_u32(sm.store[quarter]);
    const uint32_t srow = sbuf + lane * 128;  // this lane's tile row (TMEM lane) inside the warp's 32-row block
    uint32_t slot_use = 0;
    for (int work = unit_id; work < g.n_work; work += n_units) {
      const int tile = work / g.nsplit, part = work - tile * g.nsplit;
      const int b = tile / per_img;
      const int rem = tile - b * per_img;
      const int ty = rem / g.n_txp, txp = rem - ty * g.n_txp;
      const int h0 = ty * kPatchH + 2 * quarter;          // relative to lr_row0; this warp owns patch rows 2q, 2q+1
      const int w0 = (txp * CG + rank) * kPatchW;
#pragma unroll 1
      for (int nb = part * nb_cnt; nb < (part + 1) * nb_cnt; ++nb, ++slot_use) {
        const int slot = slot_use & 1;
        mbar_wait(&sm.tmem_full[slot], (slot_use >> 1) & 1);
        tc_fence_after();
        const uint32_t tslot = tmem_base + lane_bits + slot * 256;
        // matrix mode, N-block 0: q_0 = k_0 * sin(Q.0 s + bq_0), the sine argument read from this pixel's row of q0_arg
        const bool q0 = nb == 0 && g.q0_arg != nullptr;
        const float* q0row = q0 ? g.q0_arg + (static_cast<size_t>(h0 + (lane >> 4)) * g.W + w0 + (lane & 15)) * kD : nullptr;
        if constexpr (!kP16) {
#pragma unroll 1
          for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            float4 qa[8];
            if (q0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) qa[j] = __ldg(reinterpret_cast<const float4*>(q0row + c0) + j);
            }
            tmem_ld16(tslot + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
            tmem_ld16(tslot + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
            tmem_ld_wait();
            const int n0 = nb * 256 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = __uint_as_float(v[j]) + bias.b[n0 + j];
              if (nb == 0 && !g.no_relu0) x = fmaxf(x, 0.f);  // first 256 columns are k0 = relu(K0 x + b0)
              if (q0) x *= __sinf(reinterpret_cast<const float*>(qa)[j]);
              v[j] = __float_as_uint(x);
            }
            // the previous TMA store of this warp must have finished reading the staging block
            if (lane == 0) bulk_wait_group_read0();
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; ++u)
              st_shared_v4(srow + ((u ^ (lane & 7)) << 4), v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmP, sm.store[quarter], n0, w0, h0, b);
              bulk_commit_group();
            }
          }
        } else {
#pragma unroll 1
          for (int c0 = 0; c0 < 256; c0 += 64) {
            uint32_t v[64];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) tmem_ld16(tslot + c0 + 16 * q4, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * q4]));
            tmem_ld_wait();
            const int n0 = nb * 256 + c0;
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float x0 = __uint_as_float(v[j]) + bias.b[n0 + j], x1 = __uint_as_float(v[j + 1]) + bias.b[n0 + j + 1];
              if (nb == 0 && !g.no_relu0) x0 = fmaxf(x0, 0.f), x1 = fmaxf(x1, 0.f);
              v[j] = __float_as_uint(x0), v[j + 1] = __float_as_uint(x1);
            }
            if (q0) {  // (warp-uniform; 16 columns per load group keeps the register footprint flat)
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                float4 qa[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) qa[j] = __ldg(reinterpret_cast<const float4*>(q0row + c0 + 16 * g4) + j);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  v[16 * g4 + j] = __float_as_uint(__uint_as_float(v[16 * g4 + j]) * __sinf(reinterpret_cast<const float*>(qa)[j]));
              }
            }
#pragma unroll
            for (int j = 0; j < 64; j += 2) pk[j >> 1] = pack_f16x2_sat(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            if (lane == 0) bulk_wait_group_read0();
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 8; ++u)
              st_shared_v4(srow + ((u ^ (lane & 7)) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmP, sm.store[quarter], n0, w0, h0, b);  // tmP: fp16 map, box (64 cols, 16 w, 2 h)
              bulk_commit_group();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(&sm.tmem_empty[slot], 0);
          else mbar_arrive(&sm.tmem_empty[slot]);
        }
      }
    }
    if (lane == 0) bulk_wait_group0();
    __syncwarp();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

}  // namespace sa

namespace {

// grid sizing (N-block split) + launch, shared by the image and the matrix entry
int launch_sa_core(Handle* h, sa::Geo g, int n_tiles, int cta_group, int fmt, bool p16, const CUtensorMap& tmF,
                   const CUtensorMap& tmFlo, const CUtensorMap& tmW, const CUtensorMap& tmWlo, const CUtensorMap& tmP,
                   const sa::BiasParams& bias, cudaStream_t s) {
  using namespace sa;
  int* err_flag = h->err_flag;
  const int max_units = h->sm_count / cta_group;
  // N-block split: the split with the fewest (fractional) waves wins, ties go to the coarser one (less A re-fetch set-up)
  g.nsplit = 1;
  {
    static int env_ns = -1;
    if (env_ns < 0) {
      const char* e = getenv("DIINN_STAGE_A_NSPLIT");
      env_ns = e ? atoi(e) : 0;
    }
    if ((env_ns == 1 || env_ns == 2 || env_ns == 4) && env_ns <= g.nblocks) {
      g.nsplit = env_ns;
    } else {
      double best = 1e30;
      for (int ns = 1; ns <= g.nblocks; ns *= 2) {
        const int items = n_tiles * ns;
        const double waves = static_cast<double>((items + max_units - 1) / max_units) / ns;
        if (waves < best - 1e-9) best = waves, g.nsplit = ns;
      }
    }
  }
  g.n_work = n_tiles * g.nsplit;
  int units = max_units;
  if (units > g.n_work) units = g.n_work;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * cta_group, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cta_group;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = h->pdl ? 2 : 1;
#define DIINN_SA_LAUNCH(CGv, FMTv, P16v)                                                                                   \
  do {                                                                                                                     \
    DIINN_CUDA_OK(h, cudaFuncSetAttribute(stage_a_umma_kernel<CGv, FMTv, P16v>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          static_cast<int>(kSmemBytes)));                                                  \
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(&cfg, stage_a_umma_kernel<CGv, FMTv, P16v>, tmF, tmFlo, tmW, tmWlo, tmP, bias, g,   \
                                        err_flag));                                                                        \
  } while (0)
#define DIINN_SA_PICK(CGv)                                                  \
  do {                                                                      \
    if (fmt == kFmtSplit) DIINN_SA_LAUNCH(CGv, kFmtSplit, false);           \
    else if (fmt == kFmtBf16 && p16) DIINN_SA_LAUNCH(CGv, kFmtBf16, true);  \
    else if (fmt == kFmtBf16) DIINN_SA_LAUNCH(CGv, kFmtBf16, false);        \
    else if (p16) DIINN_SA_LAUNCH(CGv, kFmtF16, true);                      \
    else DIINN_SA_LAUNCH(CGv, kFmtF16, false);                              \
  } while (0)
  if (cta_group == 1) DIINN_SA_PICK(1);
  else DIINN_SA_PICK(2);
#undef DIINN_SA_PICK
#undef DIINN_SA_LAUNCH
  h->launches += 1;
  return DIINN_OK;
}

int stage_a_cta_group() {
  static int env_cg = -1;
  if (env_cg < 0) {
    const char* e = getenv("DIINN_CTA_GROUP_A");
    env_cg = (e && e[0] == '1') ? 1 : 2;
  }
  return env_cg;
}

}  // namespace

int launch_stage_a_umma(Handle* h, const void* feat_nhwc, const void* feat_lo, int fmt, int B, int H, int W, int fr0,
                        int frows, int lr_row0, int lr_rows, void* P, bool p16, cudaStream_t s) {
  using namespace sa;
  const int cta_group = stage_a_cta_group();
  if (fmt < 0 || fmt > kFmtSplit) return fail(h, DIINN_ERR_BAD_DTYPE, "stage A: unknown operand format");
  if (fmt == kFmtSplit && !feat_lo) return fail(h, DIINN_ERR_BAD_ARG, "stage A: the split format needs the residual plane");
  if (fmt == kFmtSplit && p16) return fail(h, DIINN_ERR_BAD_ARG, "stage A: the split format writes fp32 P");
  CUtensorMap tmF, tmFlo;
  const uint64_t dims[4] = {static_cast<uint64_t>(kC), static_cast<uint64_t>(W), static_cast<uint64_t>(frows),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {kC * 2ull, static_cast<uint64_t>(W) * kC * 2ull,
                               static_cast<uint64_t>(frows) * W * kC * 2ull};
  const uint32_t box[4] = {kC, kPatchW, kPatchH, 1};
  int rc = make_tmap_4d_bf16(h, &tmF, feat_nhwc, dims, strides, box);  // 16-bit elements: the type only matters to the MMA
  if (rc) return rc;
  if ((rc = make_tmap_4d_bf16(h, &tmFlo, fmt == kFmtSplit ? feat_lo : feat_nhwc, dims, strides, box))) return rc;
  CUtensorMap tmP;
  const uint64_t pdims[4] = {static_cast<uint64_t>(kPCols), static_cast<uint64_t>(W), static_cast<uint64_t>(lr_rows),
                             static_cast<uint64_t>(B)};
  const uint64_t pesz = p16 ? 2 : 4;
  const uint64_t pstrides[3] = {kPCols * pesz, static_cast<uint64_t>(W) * kPCols * pesz,
                                static_cast<uint64_t>(lr_rows) * W * kPCols * pesz};
  const uint32_t pbox[4] = {p16 ? 64u : 32u, kPatchW, 2, 1};   // 128 bytes of columns either way
  rc = p16 ? make_tmap_4d_bf16(h, &tmP, P, pdims, pstrides, pbox)   // (16-bit elements; TMA does no arithmetic)
           : make_tmap_4d_f32(h, &tmP, P, pdims, pstrides, pbox);
  if (rc) return rc;
  static_assert(sizeof(BiasParams) == sizeof(float) * kPCols, "bias block");
  const BiasParams& bias = *reinterpret_cast<const BiasParams*>(h->bA_host);
  Geo g{};
  g.B = B, g.H = H, g.W = W, g.fr0 = fr0, g.lr_row0 = lr_row0, g.lr_rows = lr_rows;
  const int tiles_x = (W + kPatchW - 1) / kPatchW;
  g.tiles_y = (lr_rows + kPatchH - 1) / kPatchH;
  g.n_txp = (tiles_x + cta_group - 1) / cta_group;
  g.no_relu0 = h->liif ? 1 : 0;
  g.nblocks = 4;
  const int ci = cta_group - 1;
  return launch_sa_core(h, g, B * g.tiles_y * g.n_txp, cta_group, fmt, p16, tmF, tmFlo, h->tmapWA[fmt == kFmtBf16 ? 0 : 1][ci],
                        h->tmapWAlo[ci], tmP, bias, s);
}

// Matrix mode (init_q=True): out (rows x 256 nblocks, fp32 or -- p16 -- fp16) = A (rows x 576, bf16, K tap-major) . W^T + bias, ReLU on N-block 0
// if relu0, and N-block 0 multiplied by sin(q0_arg) if given. rows % 256 == 0 (the caller pads). which: 0 = the stacked
// x-facing K blocks (N = 1024, bias b_k), 1 = Q.0 (N = 256, bias bq_0).
int launch_stage_a_matrix(Handle* h, const void* A16, int64_t rows, int which, const float* q0_arg, void* out, bool p16,
                          cudaStream_t s) {
  using namespace sa;
  const int cta_group = stage_a_cta_group();
  const int wm = kPatchW * cta_group;  // "image" width: one CTA (pair) tile per row of tiles
  if (rows <= 0 || rows % (kPatchH * wm) != 0) return fail(h, DIINN_ERR_BAD_SHAPE, "stage A (matrix): rows must be a multiple of the tile");
  const int nblocks = which == 0 ? 4 : 1;
  const int ci = cta_group - 1;
  CUtensorMap tmF, tmP;
  const uint64_t hrows = static_cast<uint64_t>(rows / wm);
  const uint64_t dims[4] = {static_cast<uint64_t>(kUnfold), static_cast<uint64_t>(wm), hrows, 1};
  const uint64_t strides[3] = {kUnfold * 2ull, static_cast<uint64_t>(wm) * kUnfold * 2ull, static_cast<uint64_t>(rows) * kUnfold * 2ull};
  const uint32_t box[4] = {kC, kPatchW, kPatchH, 1};
  int rc = make_tmap_4d_bf16(h, &tmF, A16, dims, strides, box);
  if (rc) return rc;
  const uint64_t ncols = static_cast<uint64_t>(nblocks) * 256;
  const uint64_t pdims[4] = {ncols, static_cast<uint64_t>(wm), hrows, 1};
  const uint64_t pesz = p16 ? 2 : 4;
  const uint64_t pstrides[3] = {ncols * pesz, static_cast<uint64_t>(wm) * ncols * pesz, static_cast<uint64_t>(rows) * ncols * pesz};
  const uint32_t pbox[4] = {p16 ? 64u : 32u, kPatchW, 2, 1};
  rc = p16 ? make_tmap_4d_bf16(h, &tmP, out, pdims, pstrides, pbox) : make_tmap_4d_f32(h, &tmP, out, pdims, pstrides, pbox);
  if (rc) return rc;
  static thread_local BiasParams bias;
  if (which == 0) memcpy(bias.b, h->bA_host, sizeof(float) * kPCols);
  else memcpy(bias.b, h->small.bq[0], sizeof(float) * kD);
  Geo g{};
  g.B = 1, g.H = static_cast<int>(hrows), g.W = wm, g.fr0 = 0, g.lr_row0 = 0, g.lr_rows = static_cast<int>(hrows);
  g.tiles_y = static_cast<int>(hrows / kPatchH);
  g.n_txp = 1;
  g.no_relu0 = which == 0 ? 0 : 1;
  g.matrix = 1, g.nblocks = nblocks, g.q0_arg = q0_arg;
  const CUtensorMap& tmW = which == 0 ? h->tmapWA[0][ci] : h->tmapWQ0A[ci];
  return launch_sa_core(h, g, g.tiles_y, cta_group, kFmtBf16, p16, tmF, tmF, tmW, tmW, tmP, bias, s);
}

}  // namespace diinn

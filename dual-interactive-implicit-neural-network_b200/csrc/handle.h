// diinn_handle: library-owned state (repacked weights, TMA descriptors, cached host-entry scratch).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>

#include <string>
#include <vector>

#include "common.cuh"

namespace diinn {

// operand formats of the tensor-core kernels (template parameter FMT of stage A / stage B)
constexpr int kFmtBf16 = 0;   // bf16 operands
constexpr int kFmtF16 = 1;    // fp16 operands (saturating conversions)
constexpr int kFmtSplit = 2;  // fp16 hi + lo split, three MMAs per product: fp32-level precision on the tensor pipe

struct Handle {
  diinn_config cfg{};
  int sm_count = 0;
  bool has_weights = false;
  bool liif = false;  // the weights are LIIF's imnet (diinn_set_weights_liif): query entries only, see api.cu
  std::string err;
  int64_t launches = 0;
  // optional per-kernel timing of the tcgen05 path (diinn_set_profiling): 4 events per decode
  bool profiling = false;
  bool pdl = true;  // programmatic dependent launch of stage A / stage B (DIINN_NO_PDL=1 at diinn_create turns it off)
  std::vector<cudaEvent_t> prof_events;  // created by diinn_set_profiling(1), 4 per decode, reused window after window
  size_t prof_used = 0;
  long long* trace_dev = nullptr;  // DIINN_TRACE=1: 1024 clock64 samples of the last stage-B launch
  int* err_flag = nullptr;         // device word the tcgen05 kernels raise on an internal consistency failure
  int4* tap = nullptr;             // diinn_debug_set_tap: per-pixel (ih, iw, rel) tap of the fused stage-B kernel

  // ---- fp32 CUDA-core path ----
  float* WA32 = nullptr;  // (1024, 576): rows [0,256) K.0; rows 256*i.. K.i[:,256:832]; reference k order c*9+tap
  float* bA = nullptr;    // (1024): K biases of layers 0..3
  float bA_host[kPCols] = {};  // host copy: travels to the tcgen05 stage-A kernel as a by-value parameter
  float* bq_dev = nullptr;  // (4, 256): Q biases (device copy of small.bq)
  float* WB32 = nullptr;  // (3, 512, 256): layer i=1..3: rows [0,256) K.i[:, :256], rows [256,512) Q.i
  // ---- tcgen05 paths ----
  // WA16[f]: (4 n-blocks, 9 taps, 256 rows, 64 ch), K index permuted to tap*64 + c; f = 0: bf16, f = 1: fp16 (= the hi part of
  // the split format); WA16lo: fp16 residual W - fp16(W). WB16[f] / WB16lo likewise: (3 layers, 2 halves, 4 k-chunks, 256
  // rows, 64) with rows [0,128) K-part, [128,256) Q-part. 16-bit storage; the element type only matters to the MMA.
  uint16_t* WA16[2] = {nullptr, nullptr};
  uint16_t* WA16lo = nullptr;
  uint16_t* WB16[2] = {nullptr, nullptr};
  uint16_t* WB16lo = nullptr;
  // TMA maps, [format][cta_group - 1]: 2-D (64, rows), box (64, 256) for single CTAs and (64, 128) for CTA pairs, 128B swizzle
  CUtensorMap tmapWA[2][2]{};
  CUtensorMap tmapWAlo[2]{};
  CUtensorMap tmapWB[2][2]{};
  CUtensorMap tmapWBlo[2]{};
  // select-MMA variant of stage B: the constant Q-branch tiles of B_sel, (3 layers, 2 halves, 2 feature blocks, K_sel rows, 64)
  // fp16 with bq_hi in row K_sel-2 and bq_lo in row K_sel-1, for K_sel = 16 ([0]) and 32 ([1])
  uint16_t* WSel16[2] = {nullptr, nullptr};
  CUtensorMap tmapSelB[2]{};
  // modes 1 / 2 (K chain entirely at LR resolution): the k-facing 256x256 blocks of K.1..3, row-major [n][k]
  float* WH32 = nullptr;          // (3, 256, 256)
  __nv_bfloat16* WH16 = nullptr;  // (3, 256, 256)
  float* WL4 = nullptr;           // mode 4: last_layer.weight as (9 taps, 256, 4) fp32 = (Wl[0], Wl[1], Wl[2], 0) per (tap, f)
  uint32_t* WL27frag = nullptr;   // mode 4, tensor path: the same weights as mma.sync B fragments, (4 kb, 4 s, 8 n tiles, 32
                                  // lanes) x uint2, bf16 hi (tiles 0..3) + lo (tiles 4..7) parts (csrc/mode4.cu)
  // init_q=True (csrc/init_q.cu): first_layer as (576, 4) fp32 = (w_relh, w_relw, w_ratio, bias) per unfolded channel,
  // Q.0 (256, 576) in the reference's channel order (fp32 path), and the tensor path's row-major bf16 operands in tap-major
  // channel order (k' = tap*64 + c): the x-facing K blocks (1024, 576) and Q.0 (256, 576)
  float* WF4 = nullptr;
  float* WQ0_32 = nullptr;
  __nv_bfloat16* WQ0A16 = nullptr;   // Q.0 (256,576) in stage A's (tap, row, channel) tile order: matrix-mode stage A on the gate
  CUtensorMap tmapWQ0A[2]{};         // [cta_group - 1]
  SmallParams small{};            // host copy; passed by value to kernels
  diinn_output_transform out_tf{};  // eval glue fused into the output store (all zero = identity)
  int64_t bsize = 0;                // diinn_set_bsize: the reference's query-chunk size (0 = None); only mode 4 reads it
  double* psnr_acc = nullptr;       // device accumulator of diinn_psnr

  // ---- cached device scratch for diinn_decode_host ----
  void* host_feat_dev = nullptr;
  size_t host_feat_bytes = 0;
  void* host_out_dev = nullptr;
  size_t host_out_bytes = 0;
  void* host_ws = nullptr;
  size_t host_ws_bytes = 0;
  // copy streams + events of the banded host pipeline (upload band k+1 / download band k-1 while band k decodes)
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[8] = {}, ev_dec[8] = {};
};

constexpr size_t kWsAlign = 1024;  // every region of the caller's workspace starts on this boundary
inline size_t align_up(size_t v) { return (v + kWsAlign - 1) / kWsAlign * kWsAlign; }

// workspace of one init_q=True decode (csrc/init_q.cu), carved behind the regions every decode has
struct InitQPlan {
  int64_t chunk = 0;     // HR pixels per chunk (tensor path: chunk_rows whole HR rows of every batch image)
  int chunk_rows = 0;
  int64_t rows_pad = 0;  // chunk rounded up to the GEMM's M tile
  size_t off_S = 0, off_XG = 0, off_PX = 0, off_QS = 0, off_q0 = 0, off_q1 = 0, off_chain = 0, end = 0;
};

inline int fail(Handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define DIINN_CUDA_OK(h, expr)                                                                       \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return ::diinn::fail((h), DIINN_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(_e));   \
  } while (0)

// ---- implemented across the .cu files -------------------------------------------------------------------
// pack.cu
int pack_weights(Handle* h, const diinn_weights_f32* w, cudaStream_t s);
// feat (B,64,H,W) NCHW fp32 | bf16, LR rows [r0,r1) -> (B, r1-r0, W, 64) 16-bit elements of operand format fmt (dst), plus
// the fp16 residual plane dst_lo for kFmtSplit
int launch_feat_to_nhwc(Handle* h, const void* feat, int io_dtype, int fmt, int B, int H, int W, int r0, int r1, void* dst,
                        void* dst_lo, cudaStream_t s);
// simt.cu
int launch_axis_tables(Handle* h, const AxisParams& ah, const AxisParams& aw, int32_t* ih, int32_t* iw,
                       float* rel_h, float* rel_w, cudaStream_t s);
int launch_query_gather(Handle* h, const PixelSource& src, int32_t* idx, float* rel, float* ratio, cudaStream_t s);
int launch_stage_a_fp32(Handle* h, const void* feat, int io_dtype, int B, int H, int W, int lr_row0, int lr_rows,
                        float* P, cudaStream_t s);
// LR-resolution K chain of modes 1 / 2: P[:, 256 i ..] += WH[i-1] . relu(P[:, 256 (i-1) ..]) for i = 1..3, M rows of P
int run_lr_chain_fp32(Handle* h, float* P, int64_t M, cudaStream_t s);
int run_lr_chain_umma(Handle* h, float* P, int64_t M, void* scratch, cudaStream_t s, bool p16 = false);  // p16: fp16 P rows
size_t lr_chain_scratch_bytes(int64_t M);
// mode 4: 3x3 reflect-padded last conv over the dumped q_3 (csrc/mode4.cu)
// tensor path: (pixels x 256) x (256 x 27) projection on mma.sync, then the 9-tap gather; T = (27, B*qrows*W_up) fp32 scratch
int launch_last_conv_umma_path(Handle* h, const __nv_bfloat16* q3, float* T, int B, int H_up, int W_up, int strip, int qr0,
                               int qrows, int row0, int row1, const OutSpec& out, cudaStream_t s);
int launch_last_conv3x3(Handle* h, const void* q3, bool q3_is_f32, int B, int H_up, int W_up, int strip, int qr0, int qrows,
                        int row0, int row1, const OutSpec& out, cudaStream_t s);
int run_stage_b_fp32(Handle* h, const PixelSource& src, const OutSpec& out, const float* P, float* qbuf0,
                     float* qbuf1, int64_t chunk, cudaStream_t s, float* q3_dump = nullptr);
int run_layers_fp32(Handle* h, const PixelSource& src, const OutSpec& out, const float* P, float* qbuf0, float* qbuf1,
                    int64_t g0, int64_t g1, cudaStream_t s, float* q3_dump);
// init_q=True (csrc/init_q.cu): gate, per-pixel x-facing GEMMs, then layers 1..3 with one P row per HR pixel
InitQPlan plan_initq(int B, int W_up, int rows, int compute, int mode, size_t off);
int run_initq_fp32(Handle* h, const void* feat, int io_dtype, const PixelSource& src, const OutSpec& out, char* ws,
                   const InitQPlan& pl, cudaStream_t s, float* q3_dump);
int run_initq_umma(Handle* h, const __nv_bfloat16* nhwc, int fr0, int frows, const PixelSource& src, const OutSpec& out,
                   char* ws, const InitQPlan& pl, int fmt, cudaStream_t s);
// stage_a_umma.cu / stage_b_umma.cu
// feat_nhwc: (B, frows, W, 64) 16-bit elements in the operand format fmt; feat_lo: the fp16 residual plane (kFmtSplit only)
// P: fp32 (B*lr_rows*W, 1024), or with p16 the same matrix in fp16 (what stage B's select-MMA variant consumes)
int launch_stage_a_matrix(Handle* h, const void* A16, int64_t rows, int which, const float* q0_arg, void* out, bool p16,
                          cudaStream_t s);
int launch_stage_a_umma(Handle* h, const void* feat_nhwc, const void* feat_lo, int fmt, int B, int H, int W, int fr0,
                        int frows, int lr_row0, int lr_rows, void* P, bool p16, cudaStream_t s);
// P: fp32 rows, or fp16 rows when stage_b_wants_p16(h, src, fmt) (the select-MMA variant, which feeds P to the tensor core)
int launch_stage_b_umma(Handle* h, const PixelSource& src, const OutSpec& out, const void* P, int cta_group, int fmt,
                        cudaStream_t s, int4* tap = nullptr);
bool stage_b_wants_p16(const Handle* h, const PixelSource& src, int fmt);
void plan_stage_b_probe(int sm_count, int decoder_mode, int B, int H, int W, int H_up, int W_up, int row0, int row1, int fmt,
                        int* out);
// gemm.cu
int launch_umma_selftest(Handle* h, const void* A, const void* B, float* D, int M, int N, int K, int cta_group,
                         cudaStream_t s, const ChainEpilogue* chain = nullptr);
int launch_umma_pace(Handle* h, int cta_group, int n_cols, int iters, int n_ctas, float* cyc_per_mma, int noise,
                     cudaStream_t s);
// tensor maps (api.cu)
int make_tmap_2d_bf16(Handle* h, CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                      uint32_t box_inner, uint32_t box_rows);
int make_tmap_4d_bf16(Handle* h, CUtensorMap* map, const void* base, const uint64_t dims[4],
                      const uint64_t strides_bytes[3], const uint32_t box[4]);
int make_tmap_4d_f32(Handle* h, CUtensorMap* map, const void* base, const uint64_t dims[4],
                     const uint64_t strides_bytes[3], const uint32_t box[4]);

}  // namespace diinn

struct diinn_handle : public diinn::Handle {};

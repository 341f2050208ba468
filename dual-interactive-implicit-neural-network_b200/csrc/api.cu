// extern "C" entry points of libdiinn_b200.so (declared in include/diinn_b200.h) and the host-side planning that
// sits between them and the kernels: argument validation, workspace carving, TMA descriptor encoding.
#include <cmath>
#include <cstring>
#include <exception>
#include <new>
#include <vector>

#include "handle.h"

using namespace diinn;

namespace {

thread_local std::string g_create_error;

// Every entry point runs on the handle's device and leaves the caller's current device as it found it (a process that
// drives several GPUs through PyTorch must not see torch.cuda.current_device() change under it).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// The tcgen05 kernels raise a device word when an internal invariant fails (operand buffers not 1024-byte aligned: the
// swizzled layouts would be read wrongly). Read it back at a point where the stream is synchronised anyway.
int check_err_flag(Handle* h) {
  int v = 0;
  if (h->err_flag && cudaMemcpy(&v, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess && v != 0) {
    cudaMemset(h->err_flag, 0, sizeof(int));
    return fail(h, DIINN_ERR_CUDA, "a tcgen05 kernel reported an internal consistency failure (code " + std::to_string(v) +
                                       "): its results are invalid");
  }
  return DIINN_OK;
}

// host twin of axis_index() in common.cuh (one rounded fp32 multiply, then floorf)
inline int host_axis_index(const AxisParams& p, int j) {
  volatile float t = (static_cast<float>(j) + 0.5f) * p.scale;
  int i = static_cast<int>(floorf(t));
  return i < p.n_in - 1 ? i : p.n_in - 1;
}

struct DecodePlan {
  int lr_row0 = 0, lr_rows = 0;  // LR rows P must hold
  int fr0 = 0, frows = 0;        // LR rows (with +-1 halo, clipped) of the NHWC bf16 copy
  size_t off_P = 0, off_q0 = 0, off_q1 = 0, off_nhwc = 0, nhwc_plane = 0, off_chain = 0, off_q3 = 0, off_T = 0, total = 0;
  InitQPlan iq;                  // init_q=True: replaces P / the activation chunks / the chain scratch
  int qr0 = 0, qr1 = 0;          // mode 4: HR rows whose q_3 is dumped (the band +- 1 halo row, clipped to the image)
  int64_t chunk = 0;
};

constexpr int64_t kFp32Chunk = 1 << 17;  // HR pixels per activation ping-pong pass of the fp32 path

inline bool is_simt(int compute) { return compute == DIINN_COMPUTE_FP32_SIMT; }
inline bool is_split(int compute) { return compute == DIINN_COMPUTE_FP32; }
// operand format of the tensor-core kernels for a compute mode
inline int fmt_of(int compute) {
  return compute == DIINN_COMPUTE_BF16 ? kFmtBf16 : compute == DIINN_COMPUTE_FP16 ? kFmtF16 : kFmtSplit;
}
// The fp32-precision tensor path covers every wiring except init_q=True, whose per-pixel x-facing GEMMs run on bf16 library
// GEMMs; there fp32 precision means the CUDA-core path.
inline int effective_compute(const Handle* h, int compute) {
  return (compute == DIINN_COMPUTE_FP32 && h && h->cfg.init_q) ? DIINN_COMPUTE_FP32_SIMT : compute;
}

DecodePlan plan_decode(int B, int H, int W, int H_up, int W_up, int row0, int row1, int compute, int mode,
                       int init_q = 0) {
  DecodePlan p;
  const AxisParams ah = make_axis(H, H_up);
  p.qr0 = row0, p.qr1 = row1;
  if (mode == 4) {  // stage A / stage B run over the band plus the rows its 3x3 reflect-padded last conv reads
    p.qr0 = row0 > 0 ? row0 - 1 : 0;
    p.qr1 = row1 < H_up ? row1 + 1 : H_up;
    row0 = p.qr0, row1 = p.qr1;
  }
  p.lr_row0 = host_axis_index(ah, row0);
  p.lr_rows = host_axis_index(ah, row1 - 1) - p.lr_row0 + 1;
  p.fr0 = p.lr_row0 > 0 ? p.lr_row0 - 1 : 0;
  const int fr1 = (p.lr_row0 + p.lr_rows + 1 < H) ? p.lr_row0 + p.lr_rows + 1 : H;
  p.frows = fr1 - p.fr0;
  size_t off = 0;
  const size_t nhwc_planes = is_split(compute) ? 2 : 1;  // the split format keeps an fp16 residual plane next to the hi plane
  if (init_q) {  // no LR-resolution P: everything is per HR pixel, chunk by chunk (csrc/init_q.cu)
    if (!is_simt(compute)) {
      p.off_nhwc = off;
      off += align_up(static_cast<size_t>(B) * p.frows * W * kC * sizeof(__nv_bfloat16));
    }
    p.iq = plan_initq(B, W_up, row1 - row0, compute, mode, off);
    off = p.iq.end;
  } else {
  p.off_P = off;
  off += align_up(static_cast<size_t>(B) * p.lr_rows * W * kPCols * sizeof(float));
  if (is_simt(compute)) {
    const int64_t total = static_cast<int64_t>(B) * (row1 - row0) * W_up;
    p.chunk = total < kFp32Chunk ? total : kFp32Chunk;
    p.off_q0 = off;
    off += align_up(static_cast<size_t>(p.chunk) * kD * sizeof(float));
    p.off_q1 = off;
    off += align_up(static_cast<size_t>(p.chunk) * kD * sizeof(float));
  } else {
    p.off_nhwc = off;
    p.nhwc_plane = align_up(static_cast<size_t>(B) * p.frows * W * kC * sizeof(__nv_bfloat16));
    off += nhwc_planes * p.nhwc_plane;
    if ((mode == 1 || mode == 2) && !is_split(compute)) {  // scratch of the LR-resolution K chain (split: fp32 chain in place)
      p.off_chain = off;
      off += align_up(lr_chain_scratch_bytes(static_cast<int64_t>(B) * p.lr_rows * W));
    }
  }
  }
  if (mode == 4) {
    p.off_q3 = off;
    const bool q3_f32 = is_simt(compute) || is_split(compute);
    off += align_up(static_cast<size_t>(B) * (p.qr1 - p.qr0) * W_up * kD * (q3_f32 ? sizeof(float) : sizeof(__nv_bfloat16)));
    if (!q3_f32) {  // the 27 (tap, channel) projections of every dumped pixel, planar fp32
      p.off_T = off;
      off += align_up(static_cast<size_t>(B) * (p.qr1 - p.qr0) * W_up * 27 * sizeof(float));
    }
  }
  p.total = off;
  return p;
}

inline bool chain_mode(const Handle* h) { return h->cfg.mode == 1 || h->cfg.mode == 2; }  // K chain at LR resolution

int check_common(Handle* h, int B, int C, int H, int W, int io_dtype, int compute) {
  if (!h) return DIINN_ERR_BAD_ARG;
  if (!h->has_weights) return fail(h, DIINN_ERR_NO_WEIGHTS, "diinn_set_weights has not been called");
  if (C != kC) return fail(h, DIINN_ERR_BAD_SHAPE, "feat must have 64 channels");
  if (B < 1 || H < 1 || W < 1) return fail(h, DIINN_ERR_BAD_SHAPE, "empty feature map");
  if (io_dtype != DIINN_IO_F32 && io_dtype != DIINN_IO_BF16 && io_dtype != DIINN_IO_BF16_NHWC)
    return fail(h, DIINN_ERR_BAD_DTYPE, "io_dtype");
  if (compute != DIINN_COMPUTE_FP32 && compute != DIINN_COMPUTE_BF16 && compute != DIINN_COMPUTE_FP16 &&
      compute != DIINN_COMPUTE_FP32_SIMT)
    return fail(h, DIINN_ERR_BAD_DTYPE, "compute must be DIINN_COMPUTE_FP32, _BF16, _FP16 or _FP32_SIMT");
  if (io_dtype == DIINN_IO_BF16_NHWC && (compute == DIINN_COMPUTE_FP32 || compute == DIINN_COMPUTE_FP32_SIMT))
    return fail(h, DIINN_ERR_BAD_DTYPE, "channels-last bf16 feature maps are taken by the 16-bit-operand paths only");
  return DIINN_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

namespace diinn {

int make_tmap_2d_bf16(Handle* h, CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                      uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(h, DIINN_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {inner * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, DIINN_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: " + std::to_string(r));
  return DIINN_OK;
}

static int make_tmap_4d(Handle* h, CUtensorMap* map, const void* base, CUtensorMapDataType dt, const uint64_t dims_[4],
                        const uint64_t strides_bytes[3], const uint32_t box_[4]);

int make_tmap_4d_bf16(Handle* h, CUtensorMap* map, const void* base, const uint64_t dims_[4],
                      const uint64_t strides_bytes[3], const uint32_t box_[4]) {
  return make_tmap_4d(h, map, base, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dims_, strides_bytes, box_);
}
int make_tmap_4d_f32(Handle* h, CUtensorMap* map, const void* base, const uint64_t dims_[4],
                     const uint64_t strides_bytes[3], const uint32_t box_[4]) {
  return make_tmap_4d(h, map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, dims_, strides_bytes, box_);
}

static int make_tmap_4d(Handle* h, CUtensorMap* map, const void* base, CUtensorMapDataType dt, const uint64_t dims_[4],
                        const uint64_t strides_bytes[3], const uint32_t box_[4]) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(h, DIINN_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4];
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) dims[i] = dims_[i], box[i] = box_[i];
  for (int i = 0; i < 3; ++i) strides[i] = strides_bytes[i];
  CUresult r = fn(map, dt, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, DIINN_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed: " + std::to_string(r));
  return DIINN_OK;
}

}  // namespace diinn

extern "C" {

const char* diinn_version(void) { return "diinn_b200 0.2 (sm_100a)"; }

int diinn_create(diinn_handle** out, const diinn_config* cfg) {
  if (!out || !cfg) {
    g_create_error = "null argument";
    return DIINN_ERR_BAD_ARG;
  }
  *out = nullptr;
  if (cfg->mode < 1 || cfg->mode > 4 || cfg->init_q < 0 || cfg->init_q > 1 || cfg->in_channels != kC ||
      cfg->hidden != kD || cfg->n_layers != kLayers) {
    g_create_error =
        "only mode in {1,2,3,4}, init_q in {0,1}, in_channels=64, hidden_dims=[256]*4 is implemented (diinn.py:40-92)";
    return DIINN_ERR_UNSUPPORTED_MODE;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return DIINN_ERR_UNSUPPORTED_DEVICE;
  }
  if (cfg->device < 0 || cfg->device >= ndev) {
    g_create_error = "bad device ordinal";
    return DIINN_ERR_BAD_ARG;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return DIINN_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                     "; this library is built for sm_100a (B200) only";
    return DIINN_ERR_UNSUPPORTED_DEVICE;
  }
  diinn_handle* h = new (std::nothrow) diinn_handle();
  if (!h) {
    g_create_error = "out of host memory";
    return DIINN_ERR_BAD_ARG;
  }
  h->cfg = *cfg;
  h->sm_count = prop.multiProcessorCount;
  DeviceGuard guard(cfg->device);
  // per-handle device scratch lives here so that decode / query never allocate (CUDA-graph capturable, one handle per
  // device or thread with no shared state)
  bool ok = cudaMalloc(&h->err_flag, sizeof(int)) == cudaSuccess && cudaMemset(h->err_flag, 0, sizeof(int)) == cudaSuccess;
  if (const char* np = getenv("DIINN_NO_PDL")) h->pdl = !(np[0] == '1');
  const char* tr = getenv("DIINN_TRACE");
  if (ok && tr && tr[0] == '1')
    ok = cudaMalloc(&h->trace_dev, 1024 * sizeof(long long)) == cudaSuccess &&
         cudaMemset(h->trace_dev, 0, 1024 * sizeof(long long)) == cudaSuccess;
  if (!ok) {
    g_create_error = std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    cudaFree(h->err_flag);
    cudaFree(h->trace_dev);
    delete h;
    return DIINN_ERR_CUDA;
  }
  *out = h;
  return DIINN_OK;
}

void diinn_destroy(diinn_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->cfg.device);
  cudaFree(h->WA32);
  cudaFree(h->bA);
  cudaFree(h->bq_dev);
  cudaFree(h->WB32);
  for (int f = 0; f < 2; ++f) {
    cudaFree(h->WA16[f]);
    cudaFree(h->WB16[f]);
  }
  cudaFree(h->WA16lo);
  cudaFree(h->WB16lo);
  cudaFree(h->WSel16[0]);
  cudaFree(h->WSel16[1]);
  cudaFree(h->err_flag);
  cudaFree(h->trace_dev);
  cudaFree(h->WH32);
  cudaFree(h->WL4);
  cudaFree(h->WL27frag);
  cudaFree(h->WF4);
  cudaFree(h->WQ0_32);
  cudaFree(h->WQ0A16);
  cudaFree(h->WH16);
  cudaFree(h->psnr_acc);
  cudaFree(h->host_feat_dev);
  cudaFree(h->host_out_dev);
  cudaFree(h->host_ws);
  if (h->s_h2d) {
    cudaStreamDestroy(h->s_h2d);
    cudaStreamDestroy(h->s_d2h);
    for (int i = 0; i < 8; ++i) {
      cudaEventDestroy(h->ev_h2d[i]);
      cudaEventDestroy(h->ev_dec[i]);
    }
  }
  for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
  delete h;
}

const char* diinn_last_error(const diinn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t diinn_launch_count(const diinn_handle* h) { return h ? h->launches : 0; }

int diinn_set_weights(diinn_handle* h, const diinn_weights_f32* w, void* stream) {
  if (!h || !w) return DIINN_ERR_BAD_ARG;
  for (int i = 0; i < 4; ++i)
    if (!w->k_weight[i] || !w->k_bias[i] || !w->q_weight[i] || !w->q_bias[i])
      return fail(h, DIINN_ERR_BAD_ARG, "null weight pointer");
  if (!w->last_weight || !w->last_bias) return fail(h, DIINN_ERR_BAD_ARG, "null weight pointer");
  DeviceGuard guard(h->cfg.device);
  h->liif = false;
  return pack_weights(h, w, static_cast<cudaStream_t>(stream));
}

// LIIF's imnet = MLP(580, 3, [256]*4) (liif.py:19-26, mlp.py:5-15) expressed in the decoder's own layouts, so that the same
// two kernels evaluate it:
//   Linear 0 (256,580): columns [0,576) multiply the unfolded features -> stage A block 0 (k_weight[0], bias b_0, NO ReLU
//                       there: Geo::no_relu0); columns 576..579 multiply (rel_h, rel_w, cell_h H, cell_w W) per query ->
//                       the four per-feature constants of stage B's layer 0 (SmallParams::wq0_p)
//   Linear i (256,256), i = 1..3: the q-facing block of K.i, with zero x-facing columns, so P block i = b_i for every LR cell
//   Q.1..3 = 0 (the kLiif instantiation of stage B ignores the gate), Linear 4 (3,256) = last_layer.
static int set_weights_liif_impl(diinn_handle* h, const diinn_liif_weights_f32* w, void* stream) {
  if (!h || !w) return DIINN_ERR_BAD_ARG;
  for (int i = 0; i < 5; ++i)
    if (!w->weight[i] || !w->bias[i]) return fail(h, DIINN_ERR_BAD_ARG, "null weight pointer");
  if (h->cfg.mode != 3 || h->cfg.init_q)
    return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "LIIF's imnet is hosted by a mode=3, init_q=0 handle");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int kIn = kUnfold + 4;
  std::vector<float> w0(static_cast<size_t>(kD) * kIn), wi(3 * static_cast<size_t>(kD) * kD), wl(3 * kD), b(4 * kD + 3);
  const cudaMemcpyKind kind = w->on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
  DIINN_CUDA_OK(h, cudaMemcpyAsync(w0.data(), w->weight[0], w0.size() * sizeof(float), kind, s));
  for (int i = 0; i < 3; ++i)
    DIINN_CUDA_OK(h, cudaMemcpyAsync(wi.data() + static_cast<size_t>(i) * kD * kD, w->weight[i + 1], sizeof(float) * kD * kD, kind, s));
  DIINN_CUDA_OK(h, cudaMemcpyAsync(wl.data(), w->weight[4], wl.size() * sizeof(float), kind, s));
  for (int i = 0; i < 4; ++i) DIINN_CUDA_OK(h, cudaMemcpyAsync(b.data() + i * kD, w->bias[i], sizeof(float) * kD, kind, s));
  DIINN_CUDA_OK(h, cudaMemcpyAsync(b.data() + 4 * kD, w->bias[4], sizeof(float) * 3, kind, s));
  DIINN_CUDA_OK(h, cudaStreamSynchronize(s));
  std::vector<float> k0(static_cast<size_t>(kD) * kUnfold), ki(3 * static_cast<size_t>(kD) * (kD + kUnfold), 0.f), q0(kD * 3),
      q0b(kD), zeros(static_cast<size_t>(kD) * kD, 0.f);
  for (int n = 0; n < kD; ++n) {
    memcpy(&k0[static_cast<size_t>(n) * kUnfold], &w0[static_cast<size_t>(n) * kIn], sizeof(float) * kUnfold);
    for (int c = 0; c < 3; ++c) q0[n * 3 + c] = w0[static_cast<size_t>(n) * kIn + kUnfold + c];
    q0b[n] = w0[static_cast<size_t>(n) * kIn + kUnfold + 3];
    for (int i = 0; i < 3; ++i)
      memcpy(&ki[(static_cast<size_t>(i) * kD + n) * (kD + kUnfold)], &wi[(static_cast<size_t>(i) * kD + n) * kD], sizeof(float) * kD);
  }
  diinn_weights_f32 v{};
  v.k_weight[0] = k0.data(), v.k_bias[0] = b.data();
  v.q_weight[0] = q0.data(), v.q_bias[0] = q0b.data();
  for (int i = 1; i < 4; ++i) {
    v.k_weight[i] = ki.data() + static_cast<size_t>(i - 1) * kD * (kD + kUnfold), v.k_bias[i] = b.data() + i * kD;
    v.q_weight[i] = zeros.data(), v.q_bias[i] = zeros.data();
  }
  v.last_weight = wl.data(), v.last_bias = b.data() + 4 * kD;
  v.on_device = 0;
  const int rc = pack_weights(h, &v, s);  // synchronises the stream before it returns: the host vectors may go
  h->liif = rc == DIINN_OK;
  return rc;
}

int diinn_set_weights_liif(diinn_handle* h, const diinn_liif_weights_f32* w, void* stream) {
  try {  // the host-side repacking allocates: nothing may throw across the C ABI
    return set_weights_liif_impl(h, w, stream);
  } catch (const std::exception& e) {
    return h ? fail(h, DIINN_ERR_BAD_ARG, std::string("diinn_set_weights_liif: ") + e.what()) : DIINN_ERR_BAD_ARG;
  }
}

size_t diinn_workspace_bytes(const diinn_handle* h, int B, int H, int W, int H_up, int W_up, int row0, int row1,
                             int compute) {
  if (B < 1 || H < 1 || W < 1 || H_up < 1 || W_up < 1 || row0 < 0 || row1 > H_up || row0 >= row1) return 0;
  return plan_decode(B, H, W, H_up, W_up, row0, row1, effective_compute(h, compute), h ? h->cfg.mode : 3,
                     h ? h->cfg.init_q : 0).total;
}

static int decode_impl(diinn_handle* h, const void* feat, int B, int C, int H, int W, int H_up, int W_up, int row0,
                       int row1, OutSpec o, void* workspace, size_t workspace_bytes, int compute, void* stream);

// The LR-resolution half of a tensor-path decode: layout pass (skipped for a channels-last bf16 map, which stage A reads in
// place as bf16 operands), stage A, and for modes 1 / 2 the K chain. nhwc / nhwc_lo / chain are workspace regions.
// p16: P is written as fp16 (what stage B's select-MMA variant consumes, stage_b_wants_p16); the K chain then updates it in fp16.
static int run_lr_stages(Handle* h, const void* feat, int io_dtype, int fmt, int B, int H, int W, int fr0, int frows,
                         int lr_row0, int lr_rows, float* P, bool p16, char* nhwc, char* nhwc_lo, char* chain, cudaStream_t s,
                         cudaEvent_t ev_after_layout = nullptr) {
  int rc;
  if (io_dtype == DIINN_IO_BF16_NHWC) {
    if (ev_after_layout) cudaEventRecord(ev_after_layout, s);
    rc = launch_stage_a_umma(h, feat, nullptr, kFmtBf16, B, H, W, 0, H, lr_row0, lr_rows, P, p16, s);
  } else {
    if ((rc = launch_feat_to_nhwc(h, feat, io_dtype, fmt, B, H, W, fr0, fr0 + frows, nhwc, nhwc_lo, s))) return rc;
    if (ev_after_layout) cudaEventRecord(ev_after_layout, s);
    rc = launch_stage_a_umma(h, nhwc, nhwc_lo, fmt, B, H, W, fr0, frows, lr_row0, lr_rows, P, p16, s);
  }
  if (rc) return rc;
  if (h->cfg.mode == 1 || h->cfg.mode == 2) {
    const int64_t M = static_cast<int64_t>(B) * lr_rows * W;
    rc = fmt == kFmtSplit ? run_lr_chain_fp32(h, P, M, s) : run_lr_chain_umma(h, P, M, chain, s, p16);
  }
  return rc;
}

int diinn_set_output_transform(diinn_handle* h, const diinn_output_transform* t) {
  if (!h) return DIINN_ERR_BAD_ARG;
  if (t && t->clamp && !(t->lo <= t->hi)) return fail(h, DIINN_ERR_BAD_ARG, "output transform: lo > hi");
  h->out_tf = t ? *t : diinn_output_transform{};
  return DIINN_OK;
}

int diinn_set_bsize(diinn_handle* h, int64_t bsize) {
  if (!h) return DIINN_ERR_BAD_ARG;
  if (bsize < 0) return fail(h, DIINN_ERR_BAD_ARG, "bsize must be >= 0 (0 = None)");
  h->bsize = bsize;
  return DIINN_OK;
}

// copy the handle's eval glue into the OutSpec of one call
static int apply_output_transform(Handle* h, OutSpec* o) {
  const diinn_output_transform& t = h->out_tf;
  o->t_flags = (t.affine ? 1 : 0) | (t.clamp ? 2 : 0) | (t.quantize_u8 ? 4 : 0);
  o->t_scale = t.scale, o->t_bias = t.bias, o->t_lo = t.lo, o->t_hi = t.hi;
  if (t.quantize_u8 && o->mc) return fail(h, DIINN_ERR_BAD_DTYPE, "multicast stores are fp32-only; uint8 output needs peer stores");
  return DIINN_OK;
}

int diinn_decode(diinn_handle* h, const void* feat, int B, int C, int H, int W, int H_up, int W_up, int row0,
                 int row1, void* out, int64_t out_batch_stride, int64_t out_chan_stride, int64_t out_row_stride,
                 void* workspace, size_t workspace_bytes, int io_dtype, int compute, void* stream) {
  OutSpec o{};
  o.ptr = out;
  o.batch_stride = out_batch_stride, o.chan_stride = out_chan_stride, o.row_stride = out_row_stride;
  o.io_dtype = io_dtype;
  return decode_impl(h, feat, B, C, H, W, H_up, W_up, row0, row1, o, workspace, workspace_bytes, compute, stream);
}

int diinn_decode_multi(diinn_handle* h, const void* feat, int B, int C, int H, int W, int H_up, int W_up, int row0,
                       int row1, void* const* out_peers, int n_peers, void* out_multicast, int64_t out_batch_stride,
                       int64_t out_chan_stride, int64_t out_row_stride, void* workspace, size_t workspace_bytes,
                       int io_dtype, int compute, void* stream) {
  if (!h) return DIINN_ERR_BAD_ARG;
  if (!out_peers || n_peers < 1 || n_peers > kMaxPeers) return fail(h, DIINN_ERR_BAD_ARG, "need 1..8 peer buffers");
  if (out_multicast && io_dtype != DIINN_IO_F32)
    return fail(h, DIINN_ERR_BAD_DTYPE, "multicast stores are implemented for fp32 images only");
  OutSpec o{};
  o.ptr = out_peers[0];
  o.batch_stride = out_batch_stride, o.chan_stride = out_chan_stride, o.row_stride = out_row_stride;
  o.io_dtype = io_dtype;
  o.n_peers = n_peers;
  o.mc = out_multicast;
  for (int i = 0; i < n_peers; ++i) {
    if (!out_peers[i]) return fail(h, DIINN_ERR_BAD_ARG, "null peer buffer");
    // every destination starts at the same row offset as a plain decode's `out` would
    o.peers[i] = static_cast<char*>(out_peers[i]);
  }
  return decode_impl(h, feat, B, C, H, W, H_up, W_up, row0, row1, o, workspace, workspace_bytes, compute, stream);
}

static int decode_impl(diinn_handle* h, const void* feat, int B, int C, int H, int W, int H_up, int W_up, int row0,
                       int row1, OutSpec o, void* workspace, size_t workspace_bytes, int compute, void* stream) {
  const int io_dtype = o.io_dtype;
  void* out = o.ptr;
  int rc = check_common(h, B, C, H, W, io_dtype, compute);
  if (rc) return rc;
  if (h->liif)
    return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "the handle holds LIIF's imnet (diinn_set_weights_liif): use diinn_query / "
                                               "diinn_query_ensemble with the grid's coordinates (liif.py:47-57,151-158)");
  compute = effective_compute(h, compute);
  if (io_dtype == DIINN_IO_BF16_NHWC) o.io_dtype = DIINN_IO_BF16;  // the image is bf16 NCHW either way
  if ((rc = apply_output_transform(h, &o))) return rc;
  if (!feat || !out) return fail(h, DIINN_ERR_BAD_ARG, "null feat/out");
  if (H_up < 1 || W_up < 1) return fail(h, DIINN_ERR_BAD_SHAPE, "size must be positive");
  if (row0 < 0 || row1 > H_up || row0 >= row1) return fail(h, DIINN_ERR_BAD_SHAPE, "bad row range");
  if (static_cast<int64_t>(B) * H * W >= (1ll << 31) / kPCols * 512)
    return fail(h, DIINN_ERR_BAD_SHAPE, "feature map too large");
  const DecodePlan plan = plan_decode(B, H, W, H_up, W_up, row0, row1, compute, h->cfg.mode, h->cfg.init_q);
  if (!workspace || workspace_bytes < plan.total)
    return fail(h, DIINN_ERR_WORKSPACE_TOO_SMALL,
                "workspace too small: need " + std::to_string(plan.total) + " bytes");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* P = reinterpret_cast<float*>(ws + plan.off_P);

  PixelSource src{};
  src.mode = 0;
  src.ax_h = make_axis(H, H_up);
  src.ax_w = make_axis(W, W_up);
  src.ratio = static_cast<float>((static_cast<double>(H) * W) / (static_cast<double>(H_up) * W_up));
  src.B = B, src.H = H, src.W = W, src.H_up = H_up, src.W_up = W_up, src.row0 = row0, src.row1 = row1;
  src.lr_row0 = plan.lr_row0, src.lr_rows = plan.lr_rows;

  // Mode 4: stage B runs over the band plus its halo rows and dumps q_3 there instead of projecting to RGB; the 3x3
  // reflect-padded last conv (csrc/mode4.cu) then writes the caller's band through the usual OutSpec.
  const bool mode4 = h->cfg.mode == 4;
  OutSpec ob = o;  // what stage B writes
  if (mode4) {
    if (H_up < 2 || W_up < 2) return fail(h, DIINN_ERR_BAD_SHAPE, "mode 4 (reflect padding) needs an output of at least 2x2");
    src.row0 = plan.qr0, src.row1 = plan.qr1;
    ob = OutSpec{};
    ob.io_dtype = o.io_dtype;
    ob.batch_stride = static_cast<int64_t>(plan.qr1 - plan.qr0) * W_up;  // pixel units: the dump is (B, rows, W_up, 256)
    ob.row_stride = W_up;
  }
  // batched_step (diinn.py:149-160) convolves column strips of bsize // H_up columns one by one, each with its own
  // reflect padding; a strip of width 1 cannot be reflect-padded (torch raises), a width of 0 never terminates there
  int strip = W_up;
  if (mode4 && h->bsize > 0) {
    const int64_t sw = h->bsize / H_up;
    if (sw < 1) return fail(h, DIINN_ERR_BAD_ARG, "bsize < H_up: the reference's batched_step makes no progress (diinn.py:155)");
    strip = sw < W_up ? static_cast<int>(sw) : W_up;
    if (strip == 1 || W_up % strip == 1)
      return fail(h, DIINN_ERR_BAD_SHAPE, "bsize leaves a column strip of width 1, which reflect padding rejects");
  }
  auto last_conv = [&](const void* q3, bool is_f32) {
    if (!is_f32)
      return launch_last_conv_umma_path(h, static_cast<const __nv_bfloat16*>(q3), reinterpret_cast<float*>(ws + plan.off_T), B,
                                        H_up, W_up, strip, plan.qr0, plan.qr1 - plan.qr0, row0, row1, o, s);
    return launch_last_conv3x3(h, q3, is_f32, B, H_up, W_up, strip, plan.qr0, plan.qr1 - plan.qr0, row0, row1, o, s);
  };

  const bool init_q = h->cfg.init_q != 0;
  if (is_simt(compute)) {
    if (init_q) {
      float* q3f = mode4 ? reinterpret_cast<float*>(ws + plan.off_q3) : nullptr;
      if ((rc = run_initq_fp32(h, feat, io_dtype, src, ob, ws, plan.iq, s, q3f))) return rc;
      return mode4 ? last_conv(q3f, true) : DIINN_OK;
    }
    if ((rc = launch_stage_a_fp32(h, feat, io_dtype, B, H, W, plan.lr_row0, plan.lr_rows, P, s))) return rc;
    if (chain_mode(h) && (rc = run_lr_chain_fp32(h, P, static_cast<int64_t>(B) * plan.lr_rows * W, s))) return rc;
    float* q3f = mode4 ? reinterpret_cast<float*>(ws + plan.off_q3) : nullptr;
    if ((rc = run_stage_b_fp32(h, src, ob, P, reinterpret_cast<float*>(ws + plan.off_q0),
                               reinterpret_cast<float*>(ws + plan.off_q1), plan.chunk, s, q3f)))
      return rc;
    return mode4 ? last_conv(q3f, true) : DIINN_OK;
  }
  // ---- tensor-core paths: layout pass -> stage A -> (LR chain) -> stage B -> (mode 4: last conv)
  const int fmt = fmt_of(compute);
  const bool split = fmt == kFmtSplit;
  if (mode4) {
    if (split) ob.q3f = reinterpret_cast<float*>(ws + plan.off_q3);
    else ob.q3 = reinterpret_cast<__nv_bfloat16*>(ws + plan.off_q3);
  }
  char* nhwc = ws + plan.off_nhwc;
  char* nhwc_lo = split ? nhwc + plan.nhwc_plane : nullptr;
  int ev = -1;  // profiling: 4 pre-created events per decode (diinn_set_profiling), none created here
  if (h->profiling && h->prof_used + 4 <= h->prof_events.size()) ev = static_cast<int>(h->prof_used), h->prof_used += 4;
  auto mark = [&](int i) {
    if (ev >= 0) cudaEventRecord(h->prof_events[ev + i], s);
  };
  if (init_q) {  // no stage A: the gate reads the NHWC copy (or the caller's channels-last tensor) per HR pixel
    mark(0);
    const __nv_bfloat16* src_nhwc = reinterpret_cast<const __nv_bfloat16*>(nhwc);
    int fr0 = plan.fr0, frows = plan.frows;
    if (io_dtype == DIINN_IO_BF16_NHWC) {
      src_nhwc = static_cast<const __nv_bfloat16*>(feat), fr0 = 0, frows = H;
    } else if ((rc = launch_feat_to_nhwc(h, feat, io_dtype, kFmtBf16, B, H, W, plan.fr0, plan.fr0 + plan.frows, nhwc, nullptr, s))) {
      return rc;  // (the gate and its library GEMMs take bf16 operands in every 16-bit compute mode)
    }
    mark(1);
    mark(2);
    rc = run_initq_umma(h, src_nhwc, fr0, frows, src, ob, ws, plan.iq, fmt, s);
    if (!rc && mode4) rc = last_conv(ob.q3, false);
    mark(3);
    return rc;
  }
  mark(0);
  // (encoder hand-off: a channels-last bf16 map is read in place as bf16 operands, whatever format stage B runs in -- a
  // bf16 feature map holds no more than bf16 precision anyway)
  const bool p16 = stage_b_wants_p16(h, src, fmt);
  if ((rc = run_lr_stages(h, feat, io_dtype, fmt, B, H, W, plan.fr0, plan.frows, plan.lr_row0, plan.lr_rows, P, p16, nhwc,
                          nhwc_lo, ws + plan.off_chain, s, ev >= 0 ? h->prof_events[ev + 1] : nullptr)))
    return rc;
  mark(2);
  rc = launch_stage_b_umma(h, src, ob, P, 0, fmt, s, h->tap);
  if (!rc && mode4) rc = split ? last_conv(ob.q3f, true) : last_conv(ob.q3, false);
  mark(3);
  return rc;
}

int diinn_decode_host(diinn_handle* h, const void* feat_host, int B, int C, int H, int W, int H_up, int W_up,
                      int row0, int row1, void* out_host, int io_dtype, int compute, void* stream) {
  int rc = check_common(h, B, C, H, W, io_dtype, compute);
  if (rc) return rc;
  if (!feat_host || !out_host) return fail(h, DIINN_ERR_BAD_ARG, "null feat/out");
  if (io_dtype == DIINN_IO_BF16_NHWC)
    return fail(h, DIINN_ERR_BAD_DTYPE, "the host entry takes NCHW feature maps (its row bands upload channel planes)");
  if (H_up < 1 || W_up < 1 || row0 < 0 || row1 > H_up || row0 >= row1)
    return fail(h, DIINN_ERR_BAD_SHAPE, "bad size / row range");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t esz = io_dtype == DIINN_IO_F32 ? 4 : 2;
  const size_t osz = h->out_tf.quantize_u8 ? 1 : esz;  // output element size (uint8 with the quantising eval glue)
  const size_t feat_bytes = static_cast<size_t>(B) * C * H * W * esz;
  const int nrows = row1 - row0;
  const size_t out_bytes = static_cast<size_t>(B) * 3 * nrows * W_up * osz;
  // Row bands: while band k decodes, band k+1's LR rows go up and band k-1's HR rows come down on two copy streams, so
  // a large image costs about max(PCIe, compute) instead of their sum. What stays exposed is the FIRST band's upload
  // and the LAST band's download, so large images get a thin first and last band (2/16 and 1/16 of the rows) around three
  // thick ones. Small images take one band. DIINN_HOST_BANDS="a,b,c,..." (relative weights) overrides the split.
  const int64_t px = static_cast<int64_t>(B) * nrows * W_up;
  int weights[7] = {1, 0, 0, 0, 0, 0, 0};
  int bands = 1;
  if (px >= (1 << 21)) {
    const int w5[5] = {2, 5, 5, 3, 1};
    bands = 5;
    for (int k = 0; k < 5; ++k) weights[k] = w5[k];
  } else if (px >= (1 << 17)) {  // e.g. one rank's row tile of an 8-way sharded DIV2K image: still worth overlapping
    bands = 3;
    weights[0] = weights[1] = weights[2] = 1;
  }
  if (const char* e = getenv("DIINN_HOST_BANDS")) {
    int n = 0;
    for (const char* p = e; *p && n < 7;) {
      const int v = atoi(p);
      if (v > 0) weights[n++] = v;
      while (*p && *p != ',') ++p;
      if (*p == ',') ++p;
    }
    if (n > 0) bands = n;
  }
  if (bands > nrows) {
    bands = 1;
    weights[0] = 1;
  }
  int wsum = 0;
  for (int k = 0; k < bands; ++k) wsum += weights[k];
  int edge[8];  // band k = HR rows [edge[k], edge[k+1])
  edge[0] = row0;
  for (int k = 0, acc = 0; k < bands; ++k) {
    acc += weights[k];
    edge[k + 1] = k + 1 == bands ? row1 : row0 + static_cast<int>(static_cast<int64_t>(nrows) * acc / wsum);
    // integer scale factor s <= 16: interior edges on multiples of s keep stage B's 8-row patches on whole LR cells (fewest
    // cells per patch: one select MMA per half slot and room for the phase table, stage_b_umma.cu); any split is bit-identical
    if (k + 1 < bands && H_up % H == 0 && H_up / H <= 16) {
      const int al = edge[k + 1] / (H_up / H) * (H_up / H);
      if (al > edge[k]) edge[k + 1] = al;
    }
    if (edge[k + 1] <= edge[k]) edge[k + 1] = edge[k] + 1 <= row1 ? edge[k] + 1 : row1;
  }
  edge[bands] = row1;
  size_t ws_bytes = 0;
  for (int k = 0; k < bands; ++k) {
    const int a = edge[k], b = edge[k + 1];
    if (a >= b) continue;
    const size_t n = diinn_workspace_bytes(h, B, H, W, H_up, W_up, a, b, compute);
    ws_bytes = n > ws_bytes ? n : ws_bytes;
  }
  auto grow = [&](void** p, size_t* have, size_t need) -> cudaError_t {
    if (*have >= need) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *have = need;
    return e;
  };
  DIINN_CUDA_OK(h, grow(&h->host_feat_dev, &h->host_feat_bytes, feat_bytes));
  DIINN_CUDA_OK(h, grow(&h->host_out_dev, &h->host_out_bytes, out_bytes));
  DIINN_CUDA_OK(h, grow(&h->host_ws, &h->host_ws_bytes, ws_bytes));
  if (!h->s_h2d) {
    DIINN_CUDA_OK(h, cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    DIINN_CUDA_OK(h, cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 8; ++i) {
      DIINN_CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
      DIINN_CUDA_OK(h, cudaEventCreateWithFlags(&h->ev_dec[i], cudaEventDisableTiming));
    }
  }
  // the copy streams start after whatever the caller queued on `s` before this call
  DIINN_CUDA_OK(h, cudaEventRecord(h->ev_dec[7], s));
  DIINN_CUDA_OK(h, cudaStreamWaitEvent(h->s_h2d, h->ev_dec[7], 0));
  DIINN_CUDA_OK(h, cudaStreamWaitEvent(h->s_d2h, h->ev_dec[7], 0));

  const AxisParams ah = make_axis(H, H_up);
  const size_t plane = static_cast<size_t>(H) * W * esz;      // one (b, c) plane of feat
  const size_t oplane = static_cast<size_t>(nrows) * W_up * osz;  // one (b, c) plane of the output band buffer
  int uploaded = 0;                                           // LR rows [0, uploaded) are already on the device
  bool started = false;
  for (int k = 0; k < bands; ++k) {
    const int a = edge[k], b = edge[k + 1];
    if (a >= b) continue;
    // LR rows this band reads (nearest-exact rows of [a,b) plus the 3x3 halo), minus what is already up. Mode 4 also
    // evaluates q_3 on one HR halo row each side of the band (plan_decode: qr0 / qr1), whose LR rows must be there too.
    const int ha = (h->cfg.mode == 4 && a > 0) ? a - 1 : a, hb = (h->cfg.mode == 4 && b < H_up) ? b + 1 : b;
    int lr0 = host_axis_index(ah, ha) - 1, lr1 = host_axis_index(ah, hb - 1) + 2;
    lr0 = lr0 < 0 ? 0 : lr0;
    lr1 = lr1 > H ? H : lr1;
    if (!started) uploaded = lr0, started = true;
    const int c0 = lr0 > uploaded ? lr0 : uploaded;
    if (lr1 > c0) {
      const size_t off = static_cast<size_t>(c0) * W * esz;
      DIINN_CUDA_OK(h, cudaMemcpy2DAsync(static_cast<char*>(h->host_feat_dev) + off, plane,
                                         static_cast<const char*>(feat_host) + off, plane,
                                         static_cast<size_t>(lr1 - c0) * W * esz, static_cast<size_t>(B) * C,
                                         cudaMemcpyHostToDevice, h->s_h2d));
      uploaded = lr1;
    }
    DIINN_CUDA_OK(h, cudaEventRecord(h->ev_h2d[k], h->s_h2d));
    DIINN_CUDA_OK(h, cudaStreamWaitEvent(s, h->ev_h2d[k], 0));
    char* oband = static_cast<char*>(h->host_out_dev) + static_cast<size_t>(a - row0) * W_up * osz;
    rc = diinn_decode(h, h->host_feat_dev, B, C, H, W, H_up, W_up, a, b, oband, static_cast<int64_t>(3) * nrows * W_up,
                      static_cast<int64_t>(nrows) * W_up, W_up, h->host_ws, h->host_ws_bytes, io_dtype, compute, s);
    if (rc) return rc;
    DIINN_CUDA_OK(h, cudaEventRecord(h->ev_dec[k], s));
    DIINN_CUDA_OK(h, cudaStreamWaitEvent(h->s_d2h, h->ev_dec[k], 0));
    DIINN_CUDA_OK(h, cudaMemcpy2DAsync(static_cast<char*>(out_host) + static_cast<size_t>(a - row0) * W_up * osz, oplane,
                                       oband, oplane, static_cast<size_t>(b - a) * W_up * osz,
                                       static_cast<size_t>(B) * 3, cudaMemcpyDeviceToHost, h->s_d2h));
  }
  DIINN_CUDA_OK(h, cudaStreamSynchronize(h->s_d2h));
  DIINN_CUDA_OK(h, cudaStreamSynchronize(s));
  return check_err_flag(h);
}

size_t diinn_query_workspace_bytes(const diinn_handle* h, int B, int H, int W, int Q, int compute) {
  if (B < 1 || H < 1 || W < 1 || Q < 1) return 0;
  // same carving as a full-image decode whose "grid" has B*Q pixels
  size_t off = align_up(static_cast<size_t>(B) * H * W * kPCols * sizeof(float));
  if (is_simt(compute)) {
    const int64_t total = static_cast<int64_t>(B) * Q;
    const int64_t chunk = total < kFp32Chunk ? total : kFp32Chunk;
    off += 2 * align_up(static_cast<size_t>(chunk) * kD * sizeof(float));
  } else {
    off += (is_split(compute) ? 2 : 1) * align_up(static_cast<size_t>(B) * H * W * kC * sizeof(__nv_bfloat16));
    if (h && (h->cfg.mode == 1 || h->cfg.mode == 2) && !is_split(compute))
      off += align_up(lr_chain_scratch_bytes(static_cast<int64_t>(B) * H * W));
  }
  return off;
}

static PixelSource make_query_source(int B, int H, int W, const float* coord, const float* cell, int Q,
                                     int ensemble = 0) {
  PixelSource src{};
  src.ensemble = ensemble;
  // python-double scalars of liif.py:79-80,92-94 rounded to fp32 when they meet the fp32 coordinate tensor
  for (int i = 0; i < 2; ++i) {
    const double v = i ? 1.0 : -1.0;
    src.sh_h[i] = static_cast<float>(v * (2.0 / H / 2.0) + 1e-6);
    src.sh_w[i] = static_cast<float>(v * (2.0 / W / 2.0) + 1e-6);
  }
  src.clamp_lo = static_cast<float>(-1 + 1e-6);
  src.clamp_hi = static_cast<float>(1 - 1e-6);
  src.mode = 1;
  src.ax_h = make_axis(H, H);
  src.ax_w = make_axis(W, W);
  src.B = B, src.H = H, src.W = W;
  src.lr_row0 = 0, src.lr_rows = H;
  src.coord = coord, src.cell = cell, src.Q = Q;
  src.hw_f = static_cast<float>(H * W);
  return src;
}

static int query_impl(diinn_handle* h, const void* feat, int B, int C, int H, int W, const float* coord,
                      const float* cell, int Q, int ensemble, void* out, void* workspace, size_t workspace_bytes,
                      int io_dtype, int compute, void* stream);

int diinn_query(diinn_handle* h, const void* feat, int B, int C, int H, int W, const float* coord, const float* cell,
                int Q, void* out, void* workspace, size_t workspace_bytes, int io_dtype, int compute, void* stream) {
  return query_impl(h, feat, B, C, H, W, coord, cell, Q, 0, out, workspace, workspace_bytes, io_dtype, compute, stream);
}

int diinn_query_ensemble(diinn_handle* h, const void* feat, int B, int C, int H, int W, const float* coord,
                         const float* cell, int Q, void* out, void* workspace, size_t workspace_bytes, int io_dtype,
                         int compute, void* stream) {
  return query_impl(h, feat, B, C, H, W, coord, cell, Q, 1, out, workspace, workspace_bytes, io_dtype, compute, stream);
}

static int query_impl(diinn_handle* h, const void* feat, int B, int C, int H, int W, const float* coord,
                      const float* cell, int Q, int ensemble, void* out, void* workspace, size_t workspace_bytes,
                      int io_dtype, int compute, void* stream) {
  int rc = check_common(h, B, C, H, W, io_dtype, compute);
  if (rc) return rc;
  if (!feat || !out || !coord || !cell) return fail(h, DIINN_ERR_BAD_ARG, "null pointer");
  if (h->cfg.mode == 4)
    return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "mode 4's 3x3 last conv is defined on the HR grid only: use diinn_decode");
  if (h->cfg.init_q)
    return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "init_q=True is implemented for the HR grid (diinn_decode) only");
  if (Q < 1) return fail(h, DIINN_ERR_BAD_SHAPE, "Q must be positive");
  const int E = ensemble ? 4 : 1;
  if (static_cast<int64_t>(Q) * E >= (1ll << 31) / 8) return fail(h, DIINN_ERR_BAD_SHAPE, "too many queries per call");
  const size_t need = diinn_query_workspace_bytes(h, B, H, W, Q * E, compute);
  if (!workspace || workspace_bytes < need)
    return fail(h, DIINN_ERR_WORKSPACE_TOO_SMALL, "workspace too small: need " + std::to_string(need) + " bytes");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* P = reinterpret_cast<float*>(ws);
  size_t off = align_up(static_cast<size_t>(B) * H * W * kPCols * sizeof(float));
  PixelSource src = make_query_source(B, H, W, coord, cell, Q, ensemble);
  src.liif = h->liif ? 1 : 0;
  OutSpec o{};
  o.ptr = out;
  o.io_dtype = io_dtype == DIINN_IO_BF16_NHWC ? DIINN_IO_BF16 : io_dtype;
  if ((rc = apply_output_transform(h, &o))) return rc;
  if (is_simt(compute)) {
    if (h->liif) return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "LIIF's imnet runs on the tensor paths (FP32 / FP16 / BF16) only");
    const int64_t total = static_cast<int64_t>(B) * Q * E;
    const int64_t chunk = total < kFp32Chunk ? total : kFp32Chunk;
    float* q0 = reinterpret_cast<float*>(ws + off);
    float* q1 = reinterpret_cast<float*>(ws + off + align_up(static_cast<size_t>(chunk) * kD * sizeof(float)));
    if ((rc = launch_stage_a_fp32(h, feat, io_dtype, B, H, W, 0, H, P, s))) return rc;
    if (chain_mode(h) && (rc = run_lr_chain_fp32(h, P, static_cast<int64_t>(B) * H * W, s))) return rc;
    return run_stage_b_fp32(h, src, o, P, q0, q1, chunk, s);
  }
  const int fmt = fmt_of(compute);
  const size_t plane = align_up(static_cast<size_t>(B) * H * W * kC * sizeof(__nv_bfloat16));
  char* nhwc = ws + off;
  char* nhwc_lo = fmt == kFmtSplit ? nhwc + plane : nullptr;
  char* chain = nhwc + (fmt == kFmtSplit ? 2 : 1) * plane;
  if ((rc = run_lr_stages(h, feat, io_dtype, fmt, B, H, W, 0, H, 0, H, P, false, nhwc, nhwc_lo, chain, s))) return rc;
  return launch_stage_b_umma(h, src, o, P, 0, fmt, s);
}

int diinn_debug_gather(diinn_handle* h, int H, int W, int H_up, int W_up, int32_t* ih, int32_t* iw, float* rel_h,
                       float* rel_w, void* stream) {
  if (!h || !ih || !iw || !rel_h || !rel_w) return DIINN_ERR_BAD_ARG;
  if (H < 1 || W < 1 || H_up < 1 || W_up < 1) return fail(h, DIINN_ERR_BAD_SHAPE, "bad shape");
  DeviceGuard guard(h->cfg.device);
  return launch_axis_tables(h, make_axis(H, H_up), make_axis(W, W_up), ih, iw, rel_h, rel_w,
                            static_cast<cudaStream_t>(stream));
}

int diinn_debug_query_gather(diinn_handle* h, int B, int H, int W, const float* coord, const float* cell, int Q,
                             int32_t* idx, float* rel, float* ratio, void* stream) {
  if (!h || !coord || !cell || !idx || !rel || !ratio) return DIINN_ERR_BAD_ARG;
  if (B < 1 || H < 1 || W < 1 || Q < 1) return fail(h, DIINN_ERR_BAD_SHAPE, "bad shape");
  DeviceGuard guard(h->cfg.device);
  return launch_query_gather(h, make_query_source(B, H, W, coord, cell, Q), idx, rel, ratio,
                             static_cast<cudaStream_t>(stream));
}

int diinn_debug_set_tap(diinn_handle* h, int32_t* tap) {
  if (!h) return DIINN_ERR_BAD_ARG;
  h->tap = reinterpret_cast<int4*>(tap);
  return DIINN_OK;
}

int diinn_debug_stage_a(diinn_handle* h, const void* feat, int B, int C, int H, int W, float* P, void* workspace,
                        size_t workspace_bytes, int io_dtype, int compute, void* stream) {
  int rc = check_common(h, B, C, H, W, io_dtype, compute);
  if (rc) return rc;
  if (!feat || !P) return fail(h, DIINN_ERR_BAD_ARG, "null pointer");
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_simt(compute)) return launch_stage_a_fp32(h, feat, io_dtype, B, H, W, 0, H, P, s);
  const int fmt = fmt_of(compute);
  const size_t plane = align_up(static_cast<size_t>(B) * H * W * kC * sizeof(__nv_bfloat16));
  const size_t need = (fmt == kFmtSplit ? 2 : 1) * plane;
  if (!workspace || workspace_bytes < need)
    return fail(h, DIINN_ERR_WORKSPACE_TOO_SMALL, "workspace too small: need " + std::to_string(need) + " bytes");
  char* nhwc = static_cast<char*>(workspace);
  if (io_dtype == DIINN_IO_BF16_NHWC)
    return launch_stage_a_umma(h, feat, nullptr, kFmtBf16, B, H, W, 0, H, 0, H, P, false, s);
  char* lo = fmt == kFmtSplit ? nhwc + plane : nullptr;
  if ((rc = launch_feat_to_nhwc(h, feat, io_dtype, fmt, B, H, W, 0, H, nhwc, lo, s))) return rc;
  return launch_stage_a_umma(h, nhwc, lo, fmt, B, H, W, 0, H, 0, H, P, false, s);
}

int diinn_debug_umma_pace(diinn_handle* h, int cta_group, int n_cols, int iters, int n_ctas, float* cyc_per_mma,
                          int noise, void* stream) {
  if (!h || !cyc_per_mma) return DIINN_ERR_BAD_ARG;
  if ((cta_group != 1 && cta_group != 2) || n_cols < 16 || n_cols > 256 || n_cols % 16 || iters < 1 || n_ctas < cta_group ||
      n_ctas % cta_group)
    return fail(h, DIINN_ERR_BAD_SHAPE, "bad pace-probe arguments");
  DeviceGuard guard(h->cfg.device);
  return launch_umma_pace(h, cta_group, n_cols, iters, n_ctas, cyc_per_mma, noise, static_cast<cudaStream_t>(stream));
}

constexpr size_t kProfDecodes = 2048;  // decodes per profiling window (4 events each), created when profiling is enabled

int diinn_set_profiling(diinn_handle* h, int enable) {
  if (!h) return DIINN_ERR_BAD_ARG;
  DeviceGuard guard(h->cfg.device);
  h->prof_used = 0;
  h->profiling = enable != 0;
  if (h->profiling && h->prof_events.empty()) {
    // all events are created HERE, never inside a decode (diinn_decode stays allocation-free and graph-capturable)
    h->prof_events.reserve(4 * kProfDecodes);
    for (size_t i = 0; i < 4 * kProfDecodes; ++i) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) break;
      h->prof_events.push_back(e);
    }
  } else if (!h->profiling) {
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    h->prof_events.clear();
  }
  return DIINN_OK;
}

int diinn_get_kernel_times(diinn_handle* h, double* ms_layout, double* ms_stage_a, double* ms_stage_b,
                           int64_t* n_decodes) {
  if (!h || !ms_layout || !ms_stage_a || !ms_stage_b || !n_decodes) return DIINN_ERR_BAD_ARG;
  DeviceGuard guard(h->cfg.device);
  *ms_layout = *ms_stage_a = *ms_stage_b = 0.0;
  *n_decodes = 0;
  const size_t n = h->prof_used / 4;
  for (size_t i = 0; i < n; ++i) {
    cudaEvent_t* e = &h->prof_events[4 * i];
    DIINN_CUDA_OK(h, cudaEventSynchronize(e[3]));
    float a = 0, b = 0, c = 0;
    DIINN_CUDA_OK(h, cudaEventElapsedTime(&a, e[0], e[1]));
    DIINN_CUDA_OK(h, cudaEventElapsedTime(&b, e[1], e[2]));
    DIINN_CUDA_OK(h, cudaEventElapsedTime(&c, e[2], e[3]));
    *ms_layout += a, *ms_stage_a += b, *ms_stage_b += c;
  }
  *n_decodes = static_cast<int64_t>(n);
  h->prof_used = 0;  // the events are reused by the next window
  return check_err_flag(h);
}

int diinn_debug_read_trace(diinn_handle* h, int64_t* host_out, int n) {
  if (!h || !host_out || n < 1 || n > 1024) return DIINN_ERR_BAD_ARG;
  if (!h->trace_dev) return fail(h, DIINN_ERR_BAD_ARG, "no trace: run a bf16 decode with DIINN_TRACE=1 first");
  DeviceGuard guard(h->cfg.device);
  DIINN_CUDA_OK(h, cudaDeviceSynchronize());
  DIINN_CUDA_OK(h, cudaMemcpy(host_out, h->trace_dev, sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
  return DIINN_OK;
}

int diinn_debug_plan_stage_b(int sm_count, int decoder_mode, int B, int H, int W, int H_up, int W_up, int row0, int row1,
                             int compute, int32_t* out12) {
  if (!out12 || sm_count < 1 || B < 1 || H < 1 || W < 1 || H_up < 1 || W_up < 1 || row0 < 0 || row1 > H_up || row0 >= row1 ||
      compute < DIINN_COMPUTE_FP32 || compute > DIINN_COMPUTE_FP16)
    return DIINN_ERR_BAD_ARG;
  plan_stage_b_probe(sm_count, decoder_mode, B, H, W, H_up, W_up, row0, row1, fmt_of(compute), out12);
  return DIINN_OK;
}

int diinn_debug_umma_gemm(diinn_handle* h, const void* A, const void* B, float* D, int M, int N, int K,
                          int cta_group, void* stream) {
  if (!h || !A || !B || !D) return DIINN_ERR_BAD_ARG;
  if (M < 128 || M % 128 || N < 256 || N % 256 || K < 64 || K % 64 ||
      (cta_group != 1 && cta_group != 2 && cta_group != 11 && cta_group != 12))
    return fail(h, DIINN_ERR_BAD_SHAPE, "need M%128==0, N%256==0, K%64==0, cta_group in {1,2,11,12}");
  if (cta_group % 10 == 2 && M % 256) return fail(h, DIINN_ERR_BAD_SHAPE, "cta_group 2 needs M%256==0");
  DeviceGuard guard(h->cfg.device);
  return launch_umma_selftest(h, A, B, D, M, N, K, cta_group, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

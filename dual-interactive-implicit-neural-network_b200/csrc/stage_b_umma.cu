// Stage B of the DIINN query decoder on tcgen05 tensor cores: the per-HR-pixel dual-interactive MLP
// (diinn.py:134-138 after hoisting the x-facing terms to LR resolution, see pack.cu / stage_a_umma.cu):
//
//   q0  = P[l][0:256] * sin(Wq0 s_p + bq0)                                        CUDA cores (K = 3)
//   k_i = relu(Wk_i[:, :256] q_{i-1} + P[l][256i:256i+256]),  q_i = k_i * sin(Wq_i q_{i-1} + bq_i),  i = 1..3
//   rgb = Wl q_3 + bl                                                              fused into layer 3's epilogue
//
// One persistent CTA per SM (CG=2: CTA pairs with cta_group::2 MMA, M = 256 across the pair, B operand split by N
// halves), 128 HR pixels (an 8x16 patch, or 128 consecutive queries) per CTA tile. Nothing per-pixel ever touches
// HBM except the 12 B/px output: activations live in two 64 KB shared-memory A-operand buffers (bf16, K-major,
// 128B swizzle) and the fp32 accumulators in TMEM.
//
// Per layer the [128 x 512] pre-activation is produced as two "half slots" of N = 256: TMEM columns
// [256h, 256h+128) = K-branch and [256h+128, 256h+256) = Q-branch of output features [128h, 128h+128). The
// epilogue of half h writes exactly K-chunks 2h, 2h+1 (64 features = one 128-byte swizzle row) of the next A
// operand and signals each chunk separately, so the next layer's MMAs start while the previous epilogue is still
// running, and layer 0 of the NEXT tile is interleaved into the epilogue warps during layer 3.
//
// Warp roles (640 threads): warp 0 = TMA producer (streams the 3 x 2 x 4 weight stages of [256 x 64] bf16 from L2),
// warp 1 = MMA issuer (leader CTA only), warp 2 = TMEM allocator, warp 3 = L2 prefetcher of upcoming tiles' P rows,
// warps 4..19 = 16 epilogue warps. ALL of them drain
// half slot 0, then half slot 1, of every layer: a warp owns one TMEM lane quarter and one 16-feature group of every
// 64-feature chunk, so a half slot is two steps per warp and every step completes one K-chunk of the next layer's A
// operand (the chunk that gates the next layer is one step behind the layer's last MMA). The control warpgroup gives
// its registers away (setmaxnreg) so each epilogue thread can hold two prefetched slices of P next to its accumulators.
//
// Operand formats (template FMT): 0 = bf16 operands, 1 = fp16 operands (11-bit mantissa: 8x less operand noise at the same
// tensor rate; conversions saturate at +-65504), 2 = fp16 hi + lo SPLIT -- every activation and weight is the sum of two
// fp16 numbers (22 mantissa bits) and each product is evaluated as a_hi w_hi + a_lo w_hi + a_hi w_lo (three MMAs into
// the same fp32 accumulator; the dropped a_lo w_lo term is 2^-22 relative). That is the fp32-PRECISION path on the
// tensor pipe (precision="fp32", DIINN_COMPUTE_FP32): activations stay in ONE 128 KB buffer (hi | lo halves), rewritten
// in place once all MMAs of a layer have retired, and the sine is range-reduced (Cody-Waite) before MUFU.SIN.
// Accumulation is fp32 in TMEM in every format.
//
// Select-MMA variant (template kSel; grid decodes whose CTA-pair patch touches at most 30 LR cells, i.e. scale factors from
// about x3.6 up -- c1, c2x4, c3, c4): the two ADDS of every epilogue element move into the tensor core. relu(accK + P[l]) needs
// the LR cell's hoisted row P[l]; sin(accQ + bq) needs the Q bias. Both are linear in a one-hot row vector: with
//   A_sel[row]  = e_slot(l(row))                                         (128 x K_sel fp16, K_sel = 16 or 32, built per tile)
//   B_sel       = [ P16 rows of the pair's LR patch ]  for the K-branch columns (one TMA 4-D box out of the fp16 P stage A
//                 writes for this variant), [ fp16(bq) in EVERY row ] for the Q-branch columns (constant table: whichever slot
//                 a row selects, it picks up the bias once)
// one or two extra K = 16 MMAs per half slot (MN-major B: a P row IS a row of N values) pre-load the accumulators with
// P[l] and bq, and the epilogue shrinks to relu(accK) * sin(accQ): no P loads, no bias loads, two adds fewer per pair.
//
// Phase table (template kTab; integer scale factors with at most 16 phases, 16-bit formats): every HR pixel of phase
// (p_h, p_w) has the relative coordinate (2p + 1)/s - 1, so sin(Wq0 s_p + bq0) is one of s_h s_w rows, evaluated once per CTA
// into shared memory; layer 0 is then q_0 = k_0 * table[phase] -- two LDS.128 and eight HMUL2 per step, no MUFU (Work::canon,
// layer0_step).
//
// Measured alternatives that did NOT pay: two 8-warp groups (one per half slot), four 128-column
// slots with per-slot groups, three slots (256|128|128), fp16 accumulators, packed FFMA2 for the RGB projection.
#include <cstdio>
#include <cstdlib>

#include "handle.h"
#include "ptx.cuh"

namespace diinn {
using namespace ptx;

namespace sb {
constexpr int kTileM = 128;
// a CTA tile is a patch of 128 HR pixels, 2^pw_log2 wide (Work::pw_log2): 8x16 by default, 4x32 / 2x64 / 16x8 when that
// leaves fewer (partly idle) waves for the launch -- e.g. the 170-row shard of an 8-way split DIV2K image: 19 waves
// instead of 20. A row's arithmetic does not depend on its tile, so the image is bit-identical whatever the shape.
constexpr int kPatchWLog2Default = 4;
constexpr int kActBytes = kTileM * kD * 2;       // 64 KB: one activation buffer (4 K-chunks x 16 KB)
constexpr int kChunkBytes = kTileM * 128;        // 16 KB: 128 rows x 128 B
constexpr int kSelBytes = 8 * 1024;              // select variant: one B_sel stage (2 x K_sel rows x 128 B)
constexpr int kASelBytes = 8 * 1024;             // select variant: one A_sel buffer (128 rows x 64 B, 64B swizzle); two of them
constexpr int kThreads = 640;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kRegsCtrl = 40, kRegsEpi = 104;

template <int CG, bool kSel>
struct Cfg {
  static constexpr int kStageRows = 256 / CG;
  static constexpr int kStageBytes = kStageRows * 128;
  static constexpr int kWBytes = (kSel ? 64 : 80) * 1024;      // weight ring: 5 stages per CTA of a pair, 4 with kSel
  static constexpr int kStages = kWBytes / kStageBytes;
  static constexpr int kSelOff = 2 * kActBytes + kWBytes;      // B_sel stage
  static constexpr int kASelOff = kSelOff + kSelBytes;         // A_sel buffers
  static constexpr int kCtrlOff = kSel ? kASelOff + 2 * kASelBytes : 2 * kActBytes + kWBytes;  // struct Smem
};

struct Smem {  // after the big buffers
  uint64_t w_full[6];
  uint64_t w_empty[6];
  uint64_t act_ready[2][4];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t a01_free;   // split format: the layer's MMAs no longer read K-chunks 0, 1 of the (in-place) activation buffer
  uint64_t sel_full, sel_empty;  // select variant: the B_sel stage
  uint32_t tmem_ptr;
  uint32_t pf_tile;    // index of the tile the weight producer has started (paces the L2 prefetcher warp)
  float partial[3][kTileM][3];  // RGB partial sums of the three non-reducing warp sets
};
template <int CG, bool kSel>
constexpr size_t smem_bytes() { return Cfg<CG, kSel>::kCtrlOff + sizeof(Smem); }
static_assert(smem_bytes<2, true>() <= 232448 && smem_bytes<2, false>() <= 232448, "exceeds 227 KB of dynamic shared memory");

// Layer-0 sine table (template kTab; integer scale factors with at most 16 phases, see Work::canon): up to 16 rows of
// 256 fp16 values sin(Wq0 s_phase + bq0), 512 B each, in two 4 KB halves. The select variant with K_sel = 16 leaves the
// upper half of its B_sel stage unused (the tile is 2 x 16 rows x 128 B), which holds phases 0..7; phases 8..15 (and, in
// the classic variant, all of them) live behind struct Smem.
constexpr int kTabHalf = 4096;
template <int CG, bool kSel, bool kTab>
constexpr size_t smem_bytes_tab() { return smem_bytes<CG, kSel>() + (kTab ? (kSel ? kTabHalf : 2 * kTabHalf) : 0); }
static_assert(sizeof(Smem) % 16 == 0, "the sine table behind struct Smem is read with 16-byte loads");
static_assert(smem_bytes_tab<2, true, true>() <= 232448 && smem_bytes_tab<2, false, true>() <= 232448, "exceeds 227 KB of dynamic shared memory");

struct Work {
  int n_work;          // work items per CTA pair (CG=2) / CTA (CG=1)
  int tiles_y, n_txp;  // grid mode
  int pw_log2;         // log2 of the patch width in pixels (3..6); patch height = 128 >> pw_log2
  int ksel;            // select variant: K_sel (16 or 32) >= box_r*box_c; slot = LR cell of the pair's patch
  int box_r, box_c;    // select variant: LR rows x columns of the TMA box that fetches a pair's P16 patch
  uint32_t sel_lbo, sel_sbo, sel_kstep;  // select variant: B_sel descriptor strides (bytes): 64-feature blocks, 8-row K groups,
                                         // and the start-address step of the second K = 16 MMA
  // Integer scale factors (H_up = s_h H, W_up = s_w W, 16-bit formats, grid decodes, s_h s_w <= 16): every HR pixel of
  // phase (p_h, p_w) = (oh - s_h ih, ow - s_w iw) has the same relative coordinate (2p + 1)/s - 1 in exact arithmetic; the
  // reference's fp32 values scatter around it by the rounding of its coordinate grids (5e-5 on a DIV2K image). With `canon`
  // every pixel uses that closed form (canon_rel), so sin(Wq0 s_p + bq0) takes s_h s_w distinct rows, which the kTab
  // instantiation evaluates ONCE per CTA into shared memory: layer 0 shrinks to q_0 = k_0 * table[phase] (no MUFU, no FMA).
  int canon, s_h, s_w;
  int4* tap;           // debug (diinn_debug_stage_b_rows): per output pixel (ih, iw, bits(rel_h), bits(rel_w)) as THIS kernel
                       // derives them, indexed by the pixel's channel-0 output offset; nullptr in product calls
};

struct RowCtx {
  const float* prow;  // P row of this pixel's LR cell (select variant: the fp16 row, see p16row())
  int slot;           // select variant: index of the LR cell inside the pair's patch box
  int phase;          // Work::canon: p_h * s_w + p_w
  float rel_h, rel_w, ratio;
  float area;  // ensemble rows: |rel_h * rel_w| + 1e-9
  float cell_w;  // kLiif only: rel_cell = cell * (H, W) (liif.py:107-110) travels as (ratio, cell_w)
  int64_t out_off;  // offset of channel 0
  bool valid;
};

// the relative coordinate of phase p of an integer scale factor s (Work::canon): fl(fl((2p + 1) / s) - 1)
__device__ __forceinline__ float canon_rel(int p, int s) {
  return __fadd_rn(__fdiv_rn(static_cast<float>(2 * p + 1), static_cast<float>(s)), -1.0f);
}

// kPix (init_q=True, csrc/init_q.cu): P holds one row per HR pixel of the launch's rows [row0,row1) -- pixel-major
// (b, row, col) -- instead of one per LR cell, and the launch covers a chunk of the band whose first row is out_row0.
// origin of a CTA pair's (or CTA's) patch and of its LR patch: shared by make_row (slots) and the producer (TMA coordinates)
template <int CG>
__device__ __forceinline__ void pair_origin(const PixelSource& s, const Work& wk, int ty, int txp, int& ih0, int& iw0) {
  const int pw = 1 << wk.pw_log2, ph = kTileM >> wk.pw_log2;
  ih0 = axis_index(s.ax_h, min(s.row0 + ty * ph, s.row1 - 1));
  iw0 = axis_index(s.ax_w, min(txp * CG * pw, s.W_up - 1));
}

template <int CG, bool kPix = false, bool kSel = false, bool kLiif = false>
__device__ __forceinline__ RowCtx make_row(const PixelSource& s, const OutSpec& o, const void* __restrict__ Pv,
                                           const Work& wk, int work, int rank, int r, bool write_tap = false) {
  RowCtx rc;
  const float* P = static_cast<const float*>(Pv);
  rc.slot = 0;
  rc.phase = 0;
  if (s.mode == 0) {
    const int per_img = wk.tiles_y * wk.n_txp;
    const int b = work / per_img;
    const int rem = work - b * per_img;
    const int ty = rem / wk.n_txp, txp = rem - ty * wk.n_txp;
    const int pw = 1 << wk.pw_log2, ph = kTileM >> wk.pw_log2;
    const int oh = s.row0 + ty * ph + (r >> wk.pw_log2);
    const int ow = (txp * CG + rank) * pw + (r & (pw - 1));
    rc.valid = oh < s.row1 && ow < s.W_up;
    const int ohc = min(oh, s.row1 - 1), owc = min(ow, s.W_up - 1);
    const int ih = axis_index(s.ax_h, ohc), iw = axis_index(s.ax_w, owc);
    if constexpr (kPix) {  // one fp16 row per HR pixel (float pointer, half the pitch: see load16h)
      rc.prow = P + (static_cast<size_t>(b * (s.row1 - s.row0) + (ohc - s.row0)) * s.W_up + owc) * (kPCols / 2);
    } else if constexpr (kSel) {  // fp16 P: half the row pitch in bytes (kept as a float pointer; see p16row())
      rc.prow = P + static_cast<size_t>((b * s.lr_rows + (ih - s.lr_row0)) * s.W + iw) * (kPCols / 2);
      int ih0, iw0;
      pair_origin<CG>(s, wk, ty, txp, ih0, iw0);
      rc.slot = (ih - ih0) * wk.box_c + (iw - iw0);
    } else {
      rc.prow = P + static_cast<size_t>((b * s.lr_rows + (ih - s.lr_row0)) * s.W + iw) * kPCols;
    }
#if DIINN_ABL & 1
    rc.prow = P;
#endif
    if (!kPix && wk.canon) {
      const int p_h = ohc - ih * wk.s_h, p_w = owc - iw * wk.s_w;
      rc.rel_h = canon_rel(p_h, wk.s_h);
      rc.rel_w = canon_rel(p_w, wk.s_w);
      rc.phase = p_h * wk.s_w + p_w;
    } else {
      rc.rel_h = axis_rel(s.ax_h, ohc, ih);
      rc.rel_w = axis_rel(s.ax_w, owc, iw);
    }
    rc.ratio = s.ratio;
    rc.out_off = b * o.batch_stride + static_cast<int64_t>(ohc - (kPix ? s.out_row0 : s.row0)) * o.row_stride + owc;
    if (wk.tap != nullptr && write_tap && rc.valid)
      wk.tap[rc.out_off] = make_int4(ih, iw, __float_as_int(rc.rel_h), __float_as_int(rc.rel_w));
  } else {
    const int64_t total = static_cast<int64_t>(s.B) * s.Q * (s.ensemble ? 4 : 1);
    const int64_t g = (static_cast<int64_t>(work) * CG + rank) * kTileM + r;
    rc.valid = g < total;
    const int64_t gc = rc.valid ? g : total - 1;
    const int64_t qi = s.ensemble ? (gc >> 2) : gc;  // query index; ensemble rows 4q..4q+3 are its four neighbours
    const int v = static_cast<int>(gc & 3);
    const int b = static_cast<int>(qi / s.Q);
    const float ch = __ldg(s.coord + qi * 2), cw = __ldg(s.coord + qi * 2 + 1);
    int ih, iw;
    if (s.ensemble) {
      ih = ensemble_index(s.ax_h, ch, s.sh_h[v >> 1], s.clamp_lo, s.clamp_hi);
      iw = ensemble_index(s.ax_w, cw, s.sh_w[v & 1], s.clamp_lo, s.clamp_hi);
    } else if constexpr (kLiif) {  // local_ensemble=False: no shift, but the same clamp + grid_sample lookup (liif.py:76,95-99)
      ih = ensemble_index(s.ax_h, ch, 0.f, s.clamp_lo, s.clamp_hi);
      iw = ensemble_index(s.ax_w, cw, 0.f, s.clamp_lo, s.clamp_hi);
    } else {
      ih = query_index(s.ax_h, ch), iw = query_index(s.ax_w, cw);
    }
    rc.prow = P + static_cast<size_t>((b * s.H + ih) * s.W + iw) * kPCols;
    rc.rel_h = query_rel(s.ax_h, ch, ih);
    rc.rel_w = query_rel(s.ax_w, cw, iw);
    if constexpr (kLiif) {
      rc.ratio = __fmul_rn(__ldg(s.cell + qi * 2), s.ax_h.n_in_f);
      rc.cell_w = __fmul_rn(__ldg(s.cell + qi * 2 + 1), s.ax_w.n_in_f);
    } else {
      rc.ratio = __fmul_rn(__fmul_rn(__fmul_rn(__ldg(s.cell + qi * 2), __ldg(s.cell + qi * 2 + 1)), s.hw_f), 0.25f);
    }
    rc.area = __fadd_rn(fabsf(__fmul_rn(rc.rel_h, rc.rel_w)), 1e-9f);
    if (s.ensemble) rc.valid = rc.valid && v == 0;  // lane 4j stores the blended query
    rc.out_off = qi * 3;
  }
  return rc;
}

// P is produced by stage A right before this kernel and is far larger than L2 for real images, so a tile's first
// touch of its P rows would be an HBM-latency load in the middle of the epilogue. The producer warp therefore pulls
// the P rows of the tile two work items ahead into L2 (whole 4 KB rows, one bulk prefetch per LR row segment).
// kEsz: bytes per P element (4, or 2 for the select variant's fp16 P)
template <int CG, int kEsz = 4>
__device__ __forceinline__ void prefetch_tile_rows(const PixelSource& s, const void* __restrict__ Pv, const Work& wk,
                                                   int work, int rank, int lane) {
  const char* P = static_cast<const char*>(Pv);
  constexpr size_t kRow = static_cast<size_t>(kPCols) * kEsz;  // bytes per P row
  if (s.mode == 0) {
    const int per_img = wk.tiles_y * wk.n_txp;
    const int b = work / per_img;
    const int rem = work - b * per_img;
    const int ty = rem / wk.n_txp, txp = rem - ty * wk.n_txp;
    const int pw = 1 << wk.pw_log2, ph = kTileM >> wk.pw_log2;
    const int oh0 = min(s.row0 + ty * ph, s.row1 - 1), oh1 = min(oh0 + ph - 1, s.row1 - 1);
    const int ow0 = min((txp * CG + rank) * pw, s.W_up - 1), ow1 = min(ow0 + pw - 1, s.W_up - 1);
    const int ih0 = axis_index(s.ax_h, oh0), ih1 = axis_index(s.ax_h, oh1);
    const int iw0 = axis_index(s.ax_w, ow0), iw1 = axis_index(s.ax_w, ow1);
    const int ncols = iw1 - iw0 + 1;
    const int segs = (ncols + 3) >> 2;  // <= 16 KB per prefetch
    const int n = (ih1 - ih0 + 1) * segs;
    for (int i = lane; i < n; i += 32) {
      const int ih = ih0 + i / segs, c0 = (i % segs) * 4;
      const int nc = min(4, ncols - c0);
      prefetch_l2_bulk(P + static_cast<size_t>((b * s.lr_rows + (ih - s.lr_row0)) * s.W + iw0 + c0) * kRow,
                       static_cast<uint32_t>(nc * kRow));
    }
  } else {
    const int64_t total = static_cast<int64_t>(s.B) * s.Q * (s.ensemble ? 4 : 1);
    for (int rr = lane; rr < kTileM; rr += 32) {
      const int64_t g = (static_cast<int64_t>(work) * CG + rank) * kTileM + rr;
      if (g >= total) break;
      const int64_t qi = s.ensemble ? (g >> 2) : g;
      const int v = static_cast<int>(g & 3);
      const int b = static_cast<int>(qi / s.Q);
      const float ch = __ldg(s.coord + qi * 2), cw = __ldg(s.coord + qi * 2 + 1);
      const int ih = s.ensemble ? ensemble_index(s.ax_h, ch, s.sh_h[v >> 1], s.clamp_lo, s.clamp_hi)
                     : s.liif   ? ensemble_index(s.ax_h, ch, 0.f, s.clamp_lo, s.clamp_hi) : query_index(s.ax_h, ch);
      const int iw = s.ensemble ? ensemble_index(s.ax_w, cw, s.sh_w[v & 1], s.clamp_lo, s.clamp_hi)
                     : s.liif   ? ensemble_index(s.ax_w, cw, 0.f, s.clamp_lo, s.clamp_hi) : query_index(s.ax_w, cw);
      prefetch_l2_bulk(P + static_cast<size_t>((b * s.H + ih) * s.W + iw) * kRow, static_cast<uint32_t>(kRow));
    }
  }
}

// kPix (init_q=True): the tile's own P rows -- 16 consecutive pixels x 2 KB (fp16) per patch row -- in 8 KB pieces, one per lane
template <int CG>
__device__ __forceinline__ void prefetch_tile_pixels(const PixelSource& s, const void* __restrict__ Pv, const Work& wk,
                                                     int work, int rank, int lane) {
  const char* P = static_cast<const char*>(Pv);
  const int per_img = wk.tiles_y * wk.n_txp;
  const int b = work / per_img;
  const int rem = work - b * per_img;
  const int ty = rem / wk.n_txp, txp = rem - ty * wk.n_txp;
  const int pw = 1 << wk.pw_log2, ph = kTileM >> wk.pw_log2;
  const int oh0 = s.row0 + ty * ph, ow0 = (txp * CG + rank) * pw;
  if (ow0 >= s.W_up) return;
  const int nrows = min(ph, s.row1 - oh0), ncols = min(pw, s.W_up - ow0);
  const int segs = (ncols + 3) >> 2;
  for (int i = lane; i < nrows * segs; i += 32) {
    const int r = i / segs, c0 = (i % segs) * 4;
    const int nc = min(4, ncols - c0);
    prefetch_l2_bulk(P + (static_cast<size_t>(b * (s.row1 - s.row0) + (oh0 + r - s.row0)) * s.W_up + ow0 + c0) * (kPCols * 2u),
                     static_cast<uint32_t>(nc) * kPCols * 2u);
  }
}

// 16-byte unit `unit` (0..7) of row r inside a [128 x 128 B] SWIZZLE_128B chunk
__device__ __forceinline__ uint32_t swz(uint32_t chunk_base, int r, int unit) {
  return chunk_base + r * 128 + ((unit ^ (r & 7)) << 4);
}

template <int CG>
__device__ __forceinline__ void signal(uint64_t* bar) {
  if constexpr (CG == 2) mbar_arrive_cluster(bar, 0); else mbar_arrive(bar);
}

// Timing ablations (tools/ablate_stage_b.py builds side libraries with -DDIINN_ABL=<mask>; results are WRONG by
// construction, only the run time is meaningful). bit 0: P slices read from one fixed row;
// bit 2: sine replaced by one FMUL; bit 3: no activation stores; bit 4: no TMEM loads.
#ifndef DIINN_ABL
#define DIINN_ABL 0
#endif
// packed fp32x2 (FADD2 / FMUL2 / FFMA2) in: 1 = layer 1..3 K/Q math, 2 = layer 0, 4 = RGB projection
// per-step epilogue timestamps for tools/trace_stage_b.py: the extra basic-block boundaries cost ~10 % (measured), so off
#ifndef DIINN_FINE_TRACE
#define DIINN_FINE_TRACE 0
#endif
#ifndef DIINN_TRACE_BUILD
#define DIINN_TRACE_BUILD DIINN_FINE_TRACE
#endif


// 16 fp32 values (64 B) of a P row as two 256-bit loads (LDG.E.256): a warp's 32 rows sit in ~8 LR cells 4 KB apart, so every
// load instruction costs ~8 L1 wavefronts whatever its width -- half the instructions of four 128-bit loads
__device__ __forceinline__ void load16(const float* __restrict__ p, float4 (&v)[4]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
#pragma unroll
  for (int j = 0; j < 2; ++j)
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[8 * j]), "=r"(u[8 * j + 1]), "=r"(u[8 * j + 2]), "=r"(u[8 * j + 3]), "=r"(u[8 * j + 4]),
                   "=r"(u[8 * j + 5]), "=r"(u[8 * j + 6]), "=r"(u[8 * j + 7])
                 : "l"(p + 8 * j));
}
// select variant: 16 fp16 values (32 B) of a P16 row, kept raw in v[0], v[1] until layer0_step unpacks them. ONE 256-bit load
// (LDG.E.256): the 32 rows of a warp sit in ~8 LR cells 2 KB apart, so every load instruction costs ~8 L1 wavefronts whatever
// its width -- half the instructions of two 128-bit loads, half the queue time (same-box A/B on c3: 1.91 -> 1.80 ms).
// Measured and rejected: all four k_0 slices of a tile fetched together one layer ahead, or fetched right before the epilogue
// warps' tensor-core waits (both ~1.92-1.97 ms: the load issue itself backs up the warps); see DESIGN.md section 4.1.
__device__ __forceinline__ void load16h(const uint16_t* __restrict__ p, float4 (&v)[4]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "l"(p));
}

// Sine of the Q branch (SineAct, diinn.py:21-26). MUFU.SIN works on x / 2pi rounded to fp32, so its absolute error grows as
// ~6e-8 |x| (1e-5 at |x| = 200): far inside the 16-bit-operand paths' budget for any sane argument. The split (fp32-
// precision) path first removes the whole turns with a two-constant Cody-Waite reduction, which keeps the error at the
// MUFU's own ~5e-7 for |x| up to ~1e5.
template <bool kReduce>
__device__ __forceinline__ float act_sin(float x) {
#if DIINN_ABL & 4
  return x * 0.5f;
#else
  if constexpr (kReduce) {
    const float k = __fadd_rn(__fmaf_rn(x, 0.15915494309189535f, 12582912.f), -12582912.f);  // rint(x / 2pi)
    x = __fmaf_rn(k, -6.28318548202514648f, x);      // fp32(2pi)
    x = __fmaf_rn(k, 1.74845553146951715e-7f, x);    // fp32(2pi) - 2pi
  }
  return __sinf(x);
#endif
}

// operand conversion of FMT: two fp32 -> one packed register of bf16 (FMT 0) or saturating fp16 (FMT 1, 2)
template <int FMT>
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  if constexpr (FMT == 0) return pack_bf16x2(lo, hi);
  else return pack_f16x2_sat(lo, hi);
}
// split formats: the fp16 residual of (a, b) after their fp16 part `hi16`
__device__ __forceinline__ uint32_t pack_residual(float a, float b, uint32_t hi16) {
  float ha, hb;
  unpack_f16x2(hi16, ha, hb);
  return pack_f16x2_sat(__fsub_rn(a, ha), __fsub_rn(b, hb));
}

// layer 0 for K-chunk kc, features [64kc + 16wg, +16) of row r -> act buffer. k0 = P[l][those features] (prefetched).
// Split formats write the fp16 residual into the lo half of the buffer (same swizzled position, kActBytes further on).
// kTab: sin(Wq0 s_p + bq0) comes out of the CTA's phase table (tab_row = this pixel's 512-byte row, tab_key = phase & 7: the
// 16-byte units of a row are XOR-swizzled by the phase so that the 8 phases a warp touches hit 8 different bank groups).
// canon (kSel, bf16 operands only): the sine is rounded to fp16 first, as the table holds it, so that row tiles which must
// compute (K_sel = 32: no room for the table) stay bit-identical to those which look it up.
template <int FMT, bool kPix = false, bool kSel = false, bool kLiif = false, bool kTab = false>
__device__ __forceinline__ void layer0_step(uint32_t act_base, int kc, int wg, int r, const RowCtx& rc,
                                            const SmallParams& sp, const float4 (&k0v)[4], uint32_t tab_row = 0,
                                            int tab_key = 0, bool canon = false) {
  constexpr bool kSplit = FMT == 2;
  const uint32_t chunk_base = act_base + kc * kChunkBytes;
  const int f0 = kc * 64 + wg * 16;
  float k0h[16];
  if constexpr (kSel && FMT != 1) {  // fp16 P: 16 values in the first 32 bytes (fp16 operands multiply them in half2, below)
    const uint32_t* raw = reinterpret_cast<const uint32_t*>(k0v);
#pragma unroll
    for (int j = 0; j < 8; ++j) unpack_f16x2(raw[j], k0h[2 * j], k0h[2 * j + 1]);
  }
  const float* k0 = kSel ? k0h : reinterpret_cast<const float*>(k0v);
  uint32_t pk[8], pl[8];
  if constexpr (kPix) {  // init_q=True: Q.0 reads the 576-wide gate, so q_0 was finished per pixel (fp16) by csrc/init_q.cu
    const uint32_t* raw = reinterpret_cast<const uint32_t*>(k0v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if constexpr (FMT == 1) {
        pk[j] = raw[j];
      } else {
        float a, b;
        unpack_f16x2(raw[j], a, b);
        pk[j] = pack_op<FMT>(a, b);
      }
    }
  } else if constexpr (kLiif) {
    // LIIF's imnet, first Linear(580, 256) + ReLU (mlp.py:9-12): the 576 feature columns were applied per LR cell by stage A
    // (k0 = W1[:, :576] x_l + b1, no ReLU there); the four coordinate columns (rel_h, rel_w, cell_h H, cell_w W) are per query
    const float2 rh = make_float2(rc.rel_h, rc.rel_h), rw = make_float2(rc.rel_w, rc.rel_w);
    const float2 chh = make_float2(rc.ratio, rc.ratio), cww = make_float2(rc.cell_w, rc.cell_w);
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float2* w = &sp.wq0_p[(f0 + j) >> 1][0];
      float2 t = __ffma2_rn(w[0], rh, make_float2(k0[j], k0[j + 1]));
      t = __ffma2_rn(w[1], rw, t);
      t = __ffma2_rn(w[2], chh, t);
      t = __ffma2_rn(w[3], cww, t);
      const float qx = fmaxf(t.x, 0.f), qy = fmaxf(t.y, 0.f);
      pk[j >> 1] = pack_op<FMT>(qx, qy);
      if constexpr (kSplit) pl[j >> 1] = pack_residual(qx, qy, pk[j >> 1]);
    }
  } else if constexpr (kTab) {
    uint32_t tv[8];
    const int u0 = kc * 8 + wg * 2;  // first 16-byte unit (8 features) of this step's 16 features
    ld_shared_v4(tab_row + (((u0) ^ tab_key) << 4), tv[0], tv[1], tv[2], tv[3]);
    ld_shared_v4(tab_row + (((u0 + 1) ^ tab_key) << 4), tv[4], tv[5], tv[6], tv[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if constexpr (kSel && FMT == 1) {
        const uint32_t k0raw = reinterpret_cast<const uint32_t*>(k0v)[j];
        asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pk[j]) : "r"(k0raw), "r"(tv[j]));
      } else {
        float sx, sy;
        unpack_f16x2(tv[j], sx, sy);
        const float2 q = __fmul2_rn(make_float2(k0[2 * j], k0[2 * j + 1]), make_float2(sx, sy));
        pk[j] = pack_op<FMT>(q.x, q.y);
      }
    }
  } else {
    // Layer 0 is bound by the FMA pipe (packed FFMA2 / FMUL2 and the fp16 -> fp32 unpacks all issue there), so:
    //  * grid decodes fold the per-image constant w_ratio * ratio + b into the bias on the host (sp.q0_folded): two FFMA2
    //    per feature pair instead of three;
    //  * the select variant with fp16 operands multiplies k_0 (already fp16) by the sine in half2 -- one F2FP + one HMUL2
    //    per pair instead of two unpacks + FMUL2 + F2FP (|k_0 sin| <= |k_0|: no overflow).
    const float2 rh = make_float2(rc.rel_h, rc.rel_h), rw = make_float2(rc.rel_w, rc.rel_w), ra = make_float2(rc.ratio, rc.ratio);
    const bool folded = sp.q0_folded != 0;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      // packed fp32x2 (two features per instruction)
      const float2* w = &sp.wq0_p[(f0 + j) >> 1][0];
      float2 t = __ffma2_rn(w[0], rh, w[3]);
      t = __ffma2_rn(w[1], rw, t);
      if (!folded) t = __ffma2_rn(w[2], ra, t);
      float2 sn = make_float2(act_sin<kSplit>(t.x), act_sin<kSplit>(t.y));
      if constexpr (!kSplit && !(kSel && FMT == 1)) {
        if (canon) unpack_f16x2(pack_f16x2_sat(sn.x, sn.y), sn.x, sn.y);
      }
      if constexpr (kSel && FMT == 1) {
        const uint32_t k0raw = reinterpret_cast<const uint32_t*>(k0v)[j >> 1];
        const uint32_t s16 = pack_f16x2_sat(sn.x, sn.y);
        asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(pk[j >> 1]) : "r"(k0raw), "r"(s16));
      } else {
        const float2 q = __fmul2_rn(make_float2(k0[j], k0[j + 1]), sn);
        pk[j >> 1] = pack_op<FMT>(q.x, q.y);
        if constexpr (kSplit) pl[j >> 1] = pack_residual(q.x, q.y, pk[j >> 1]);
      }
    }
  }
  st_shared_v4(swz(chunk_base, r, wg * 2), pk[0], pk[1], pk[2], pk[3]);
  st_shared_v4(swz(chunk_base, r, wg * 2 + 1), pk[4], pk[5], pk[6], pk[7]);
  if constexpr (kSplit) {
    st_shared_v4(swz(chunk_base + kActBytes, r, wg * 2), pl[0], pl[1], pl[2], pl[3]);
    st_shared_v4(swz(chunk_base + kActBytes, r, wg * 2 + 1), pl[4], pl[5], pl[6], pl[7]);
  }
}

// ---- epilogue of 16 features [128h + 64c + 16wg, +16) of one row, reference layer `layer` (1..3) ------------------
// raw accumulator registers of one step: 16 K + 16 Q columns of fp32
struct Raw {
  uint32_t v[32];
};

__device__ __forceinline__ void epi_load(uint32_t tslot, int col, Raw& raw) {  // issue only; caller waits
#if DIINN_ABL & 16
#pragma unroll
  for (int i = 0; i < 32; ++i) raw.v[i] = tslot + col + i;
  return;
#endif
  tmem_ld16(tslot + col, *reinterpret_cast<uint32_t(*)[16]>(&raw.v[0]));
  tmem_ld16(tslot + 128 + col, *reinterpret_cast<uint32_t(*)[16]>(&raw.v[16]));
}

// kx = the matching slice of P (prefetched). kLast: accumulate the RGB projection instead of writing the next A operand.
// kDump (mode 4): q_3 goes to HBM for the 3x3 last conv -- bf16 (q3row) from the 16-bit-operand formats, fp32 (q3row_f) from
// the split format.
// kPix: the P slice arrives as 16 fp16 values (per-pixel fp16 rows of init_q=True) instead of 16 fp32.
template <bool kLast, int FMT, bool kDump = false, bool kSel = false, bool kLiif = false, bool kPix = false>
__device__ __forceinline__ void epi_math(const Raw& raw, uint32_t out_base, int layer, int h, int c, int wg, int r,
                                         const SmallParams& sp, const float4 (&kxv)[4], float (&rgb)[3],
                                         __nv_bfloat16* q3row = nullptr, float* q3row_f = nullptr) {
  constexpr bool kSplit = FMT == 2;
  const int col = c * 64 + wg * 16;
  const int f0 = h * 128 + col;
  float kxh[16];
  if constexpr (kPix) {
    const uint32_t* raw = reinterpret_cast<const uint32_t*>(kxv);
#pragma unroll
    for (int j = 0; j < 8; ++j) unpack_f16x2(raw[j], kxh[2 * j], kxh[2 * j + 1]);
  }
  const float* kx = kPix ? kxh : reinterpret_cast<const float*>(kxv);
  uint32_t pk[8], pl[8];
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const float ak[2] = {__uint_as_float(raw.v[j]), __uint_as_float(raw.v[j + 1])};
    const float aq[2] = {__uint_as_float(raw.v[16 + j]), __uint_as_float(raw.v[16 + j + 1])};
    // packed fp32x2 adds / multiplies (Blackwell FADD2 / FMUL2); the select variant's accumulators already hold P[l] and bq
    float2 k2 = make_float2(ak[0], ak[1]), t2 = make_float2(aq[0], aq[1]);
    if constexpr (!kSel) {
      k2 = __fadd2_rn(k2, make_float2(kx[j], kx[j + 1]));
      t2 = __fadd2_rn(t2, *reinterpret_cast<const float2*>(&sp.bq[layer][f0 + j]));
    }
    k2.x = fmaxf(k2.x, 0.f), k2.y = fmaxf(k2.y, 0.f);
    // (kLiif: a plain ReLU layer of LIIF's imnet -- the Q-branch columns were multiplied by zero weights and are ignored)
    const float2 q2 = kLiif ? k2 : __fmul2_rn(k2, make_float2(act_sin<kSplit>(t2.x), act_sin<kSplit>(t2.y)));
    const float q[2] = {q2.x, q2.y};
    if constexpr (kLast && kDump) {
      if constexpr (kSplit) {
        if (q3row_f != nullptr) *reinterpret_cast<float2*>(q3row_f + f0 + j) = q2;
      } else {
        pk[j >> 1] = pack_bf16x2(q[0], q[1]);
      }
    } else if constexpr (kLast) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float4 w = *reinterpret_cast<const float4*>(&sp.wl_t[f0 + j + e][0]);
        rgb[0] = fmaf(w.x, q[e], rgb[0]);
        rgb[1] = fmaf(w.y, q[e], rgb[1]);
        rgb[2] = fmaf(w.z, q[e], rgb[2]);
      }
    }
    if constexpr (!kLast) {
      pk[j >> 1] = pack_op<FMT>(q[0], q[1]);
      if constexpr (kSplit) pl[j >> 1] = pack_residual(q[0], q[1], pk[j >> 1]);
    }
  }
  if constexpr (kLast && kDump && !kSplit) {
    if (q3row != nullptr) {  // rows outside the image / band have no store target
      uint4* dst = reinterpret_cast<uint4*>(q3row + f0);
      dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
  }
  if constexpr (!kLast) {
    const uint32_t chunk_base = out_base + (2 * h + c) * kChunkBytes;
#if DIINN_ABL & 8
    if (pk[0] == 0x12345678u) st_shared_v4(swz(chunk_base, r, wg * 2), pk[0], pk[1] ^ pk[2] ^ pk[3] ^ pk[4], pk[5] ^ pk[6], pk[7]);
#else
    st_shared_v4(swz(chunk_base, r, wg * 2), pk[0], pk[1], pk[2], pk[3]);
    st_shared_v4(swz(chunk_base, r, wg * 2 + 1), pk[4], pk[5], pk[6], pk[7]);
    if constexpr (kSplit) {
      st_shared_v4(swz(chunk_base + kActBytes, r, wg * 2), pl[0], pl[1], pl[2], pl[3]);
      st_shared_v4(swz(chunk_base + kActBytes, r, wg * 2 + 1), pl[4], pl[5], pl[6], pl[7]);
    }
#endif
  }
}

// kDump (mode 4): a separate instantiation, so the q_3 dump costs the RGB-projecting kernels neither a register nor a branch.
// kPix (init_q=True): see make_row / layer0_step.
// tmWlo: the fp16 residual weights (split format). tmSelP / tmSelB (select variant): the 4-D map of the fp16 P whose box is
// one pair's LR patch x 64 features, and the 2-D map of the constant Q-bias tiles. P: fp32 rows, fp16 rows with kSel.
// kLiif: LIIF's imnet (ReLU MLP 580 -> 256^4 -> 3, liif.py:26 / mlp.py) instead of the dual-interactive layers, query lists only.
// kTab: layer 0's sines come out of a per-CTA phase table (Work::canon, layer0_step).
template <int CG, int FMT, bool kDump, bool kPix, bool kSel, bool kLiif, bool kTab>
__global__ void __launch_bounds__(kThreads, 1)
stage_b_umma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo,
                    const __grid_constant__ CUtensorMap tmSelP, const __grid_constant__ CUtensorMap tmSelB,
                    const __grid_constant__ SmallParams sp,
                    const __grid_constant__ PixelSource src, const __grid_constant__ OutSpec out,
                    const void* __restrict__ P, const __grid_constant__ Work wk,
                    int* __restrict__ err_flag, long long* __restrict__ trace) {
  using C = Cfg<CG, kSel>;
  constexpr bool kSplit = FMT == 2;
  static_assert(!(kSplit && kPix), "the split format is not wired for per-pixel P (init_q=True runs on the fp32 CUDA-core path)");
  static_assert(!kSel || (CG == 2 && !kSplit && !kPix), "select variant: CTA pairs, 16-bit operand formats, LR-resolution P");
  static_assert(!kLiif || (!kSel && !kPix && !kDump), "LIIF's imnet runs on the classic kernel (query lists, fp32 P)");
  static_assert(!kTab || (CG == 2 && !kSplit && !kPix && !kLiif), "phase table: CTA pairs, 16-bit operand formats, grid decodes");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_act = smem;                     // 2 x 64 KB
  uint8_t* s_w = smem + 2 * kActBytes;       // weight stages
  Smem& sm = *reinterpret_cast<Smem*>(smem + C::kCtrlOff);
  const uint32_t sel0 = smem_u32(smem + C::kSelOff), asel0 = smem_u32(smem + C::kASelOff);  // select variant only
  // phase table (kTab): rows 0..7 at tab_lo, rows 8..15 at tab_hi (see kTabHalf)
  const uint32_t tab_tail = smem_u32(smem + C::kCtrlOff + sizeof(Smem));
  const uint32_t tab_lo = kSel ? sel0 + kTabHalf : tab_tail;
  const uint32_t tab_hi = kSel ? tab_tail : tab_tail + kTabHalf;

  // warp index through a shuffle: ptxas then knows it is warp-uniform, so everything indexed by it (feature offsets
  // into the constant-bank parameters, TMEM columns) goes through uniform registers / LDCU instead of the ADU
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = rank == 0;
  const int unit_id = blockIdx.x / CG;        // CTA pair (or CTA) index
  const int n_units = gridDim.x / CG;
  // optional timeline capture (DIINN_TRACE=1): leader CTA of unit 0, first 8 tiles, clock64 at pipeline events
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  // The timeline points are compiled in only with -DDIINN_TRACE_BUILD=1 (tools/trace_stage_b.py builds that side library):
  // even never-taken `if (tracer) clock64()` sites split the epilogue into more basic blocks and cost issue slots.
#if DIINN_TRACE_BUILD
#define DIINN_TR(tile, idx) do { if (tracing && (tile) < 8) trace[(tile) * 128 + (idx)] = clock64(); } while (0)
#else
#define DIINN_TR(tile, idx) do { } while (0)
#endif

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) atomicExch(err_flag, 1);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmW);
    if constexpr (kSplit) prefetch_tensormap(&tmWlo);
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&sm.w_full[i], 1);
      mbar_init(&sm.w_empty[i], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&sm.act_ready[0][0] + i, kEpiWarps * CG);  // every epilogue warp writes part of a chunk
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.tmem_full[i], 1);
      mbar_init(&sm.tmem_empty[i], kEpiWarps * CG);
    }
    sm.pf_tile = 0;
    mbar_init(&sm.a01_free, 1);
    mbar_init(&sm.sel_full, 1);
    mbar_init(&sm.sel_empty, 1);
    if constexpr (kSel) {
      prefetch_tensormap(&tmSelP);
      prefetch_tensormap(&tmSelB);
    }
    fence_barrier_init();
  }
  if constexpr (kSel) {
    // the B_sel stage starts all-zero: the P16 box fills rows [0, box_r*box_c) of the K-branch CTA's tile and nothing ever
    // writes the rows behind them (the bias slots and the padding up to K_sel must multiply to zero there)
    // (kTab: K_sel = 16, the tile is the lower half of the stage and the upper half holds phase-table rows)
    for (int i = threadIdx.x; i < (kTab ? kTabHalf : kSelBytes) / 16; i += kThreads) st_shared_v4(sel0 + i * 16, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if constexpr (kTab) {
    // sin(Wq0 (rel_h, rel_w, ratio) + bq0) for every phase x feature pair, with exactly the operations of layer0_step's
    // compute path (the launcher folded w_ratio * ratio + b into the bias: grid decode), rounded to fp16
    const int n = wk.s_h * wk.s_w * (kD / 2);
    for (int i = threadIdx.x; i < n; i += kThreads) {
      const int phase = i >> 7, fp = i & 127;
      const int p_h = phase / wk.s_w, p_w = phase - p_h * wk.s_w;
      const float rel_h = canon_rel(p_h, wk.s_h), rel_w = canon_rel(p_w, wk.s_w);
      const float2* w = &sp.wq0_p[fp][0];
      float2 t = __ffma2_rn(w[0], make_float2(rel_h, rel_h), w[3]);
      t = __ffma2_rn(w[1], make_float2(rel_w, rel_w), t);
      const uint32_t s16 = pack_f16x2_sat(act_sin<false>(t.x), act_sin<false>(t.y));
      const uint32_t row = (phase < 8 ? tab_lo : tab_hi) + static_cast<uint32_t>(phase & 7) * 512u;
      st_shared_b32(row + ((((fp >> 2) ^ (phase & 7))) << 4) + (fp & 3) * 4, s16);
    }
  }
  if (warp == 2) tmem_alloc<CG>(&sm.tmem_ptr, 512);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_ptr;
  const uint32_t act0 = smem_u32(s_act);
  grid_dep_launch();  // the next decode's first kernel is a normal launch, so this only matters inside CUDA graphs
  grid_dep_wait();    // P is complete once stage A (or the LR chain of modes 1 / 2) has finished (PDL, see ptx.cuh)

  if (warp < 4) {
  setmaxnreg_dec<kRegsCtrl>();
  if (warp == 0) {
    // ===================== weight producer =====================
    uint32_t it = 0, sel_it = 0;
    // one weight stage: tile s24 of the (layer-1, half, kc) sequence out of map tm
    int tr_tile = 0;  // tile counter of this producer
    auto load_w = [&](const CUtensorMap* tm, int s24) {
      const int st = it % C::kStages;
      const uint32_t ph = (it / C::kStages) & 1;
      mbar_wait(&sm.w_empty[st], ph ^ 1);
#if DIINN_TRACE_BUILD
      // the stage came free = the MMAs of the weight tile kStages earlier have RETIRED: the tensor pipe's own timeline, as
      // long as the producer was already waiting here (16 + 4 + 4 slots: 96..111, 124..127, 60..63)
      if (!kSplit && lane == 0 && (s24 & 3) == 0) DIINN_TR(tr_tile, 96 + (s24 >> 2) * 2);
      if (!kSplit && lane == 0 && s24 == 23) DIINN_TR(tr_tile, 110);
#endif
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(&sm.w_full[st], C::kStageBytes * CG);
        void* dst = s_w + st * C::kStageBytes;
        if constexpr (CG == 1)
          tma_load_2d(dst, tm, &sm.w_full[st], 0, s24 * 256);
        else
          tma_load_2d_2sm(dst, tm, &sm.w_full[st], 0, s24 * 256 + rank * 128);
      }
      __syncwarp();
      ++it;
    };
    for (int work = unit_id; work < wk.n_work; work += n_units, ++tr_tile) {
      if (lane == 0) st_shared_volatile_u32(smem_u32(&sm.pf_tile), static_cast<uint32_t>(tr_tile));
#if DIINN_TRACE_BUILD
      if (lane == 0) DIINN_TR(tr_tile, 109);
#endif
      __syncwarp();
      // the whole warp walks the ring (so the stage index and barrier addresses stay in uniform registers and the
      // TMA / mbarrier instructions are issued without a per-lane broadcast loop); one elected lane issues
      if constexpr (kSplit) {
        // per (layer, half) and K-chunk pair g: hi, lo tiles of chunks 2g, 2g+1 (for the small a_lo.w_hi and a_hi.w_lo
        // terms), then their hi tiles again (a_hi.w_hi) -- see the MMA issuer
        for (int sl = 0; sl < 72; ++sl) {
          const int lh = sl / 12, j = sl % 12, g = j / 6, jj = j % 6;  // jj: 0..3 = (kc, hi|lo) of the small terms, 4..5 = hi again
          const int kc = 2 * g + (jj < 4 ? (jj >> 1) : jj - 4);
          load_w((jj < 4 && (jj & 1)) ? &tmWlo : &tmW, lh * 4 + kc);
        }
      } else {
        int b_img = 0, ih0 = 0, iw0 = 0;
        if constexpr (kSel) {
          const int per_img = wk.tiles_y * wk.n_txp;
          b_img = work / per_img;
          const int rem = work - b_img * per_img;
          const int ty = rem / wk.n_txp, txp = rem - ty * wk.n_txp;
          pair_origin<CG>(src, wk, ty, txp, ih0, iw0);
        }
        for (int lh = 0; lh < 6; ++lh) {  // (layer-1, half)
          if constexpr (kSel) {
            // B_sel stage of this half slot: the K-branch CTA (rank 0: features [128h, 128h+128) of P's block `layer`)
            // fetches the pair's LR patch out of the fp16 P, the Q-branch CTA the constant bias tile
            const uint32_t fb_stride = static_cast<uint32_t>(wk.ksel) * 128u;
            mbar_wait(&sm.sel_empty, (sel_it & 1) ^ 1);
#if DIINN_TRACE_BUILD
            if (lane == 0) DIINN_TR(tr_tile, 97 + lh * 2);
#endif
            if (elect_one()) {
              if (leader)
                mbar_arrive_expect_tx(&sm.sel_full, 2u * 128u * static_cast<uint32_t>(wk.box_r * wk.box_c + wk.ksel));
              uint8_t* dst = smem + C::kSelOff;
              if (rank == 0) {
                const int f0 = ((lh >> 1) + 1) * kD + (lh & 1) * 128;
                tma_load_4d_2sm(dst, &tmSelP, &sm.sel_full, f0, iw0, ih0 - src.lr_row0, b_img);
                tma_load_4d_2sm(dst + fb_stride, &tmSelP, &sm.sel_full, f0 + 64, iw0, ih0 - src.lr_row0, b_img);
              } else {
                tma_load_2d_2sm(dst, &tmSelB, &sm.sel_full, 0, (lh * 2 + 0) * wk.ksel);
                tma_load_2d_2sm(dst + fb_stride, &tmSelB, &sm.sel_full, 0, (lh * 2 + 1) * wk.ksel);
              }
            }
            __syncwarp();
            ++sel_it;
          }
          for (int kc = 0; kc < 4; ++kc) load_w(&tmW, lh * 4 + kc);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // The whole warp runs the loop and waits on the barriers; only the tcgen05 instructions sit under elect_one().
    // With every operand derived from warp-uniform values ptxas builds the descriptors in uniform registers; issued
    // from a divergent single-lane region instead, each UTCHMMA costs an ELECT + 5x R2UR.BROADCAST waterfall
    // (~100 clk), which capped the tensor pipe at ~200 clk per MMA instead of 128.
    if (leader) {
      constexpr uint32_t idesc = FMT == 0 ? umma_idesc_bf16(128 * CG, 256) : umma_idesc_f16(128 * CG, 256);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t it = 0;            // weight stage counter
      uint32_t sel_it = 0;        // B_sel stage counter (select variant)
      uint32_t act_phase = 0;     // bit b: parity to wait for on act_ready[b][*]
      uint32_t slot_uses = 0;     // completed uses per TMEM slot (same for both slots at layer granularity)
      int t = 0;
      for (int work = unit_id; work < wk.n_work; work += n_units, ++t) {
        const int X = t & 1;
#pragma unroll 1
        for (int layer = 1; layer <= 3; ++layer) {
          // L1: buf X, L2: buf X^1, L3: buf X; the split format has ONE buffer (hi half | lo half), rewritten in place
          const int bin = kSplit ? 0 : ((layer == 2) ? (X ^ 1) : X);
          const uint32_t a_base = act0 + bin * kActBytes;
          const uint32_t aph = (act_phase >> bin) & 1;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            // slot h must have been drained by the epilogue of its previous use
            if constexpr (CG == 2) mbar_wait_cluster(&sm.tmem_empty[h], (slot_uses & 1) ^ 1);
            else mbar_wait(&sm.tmem_empty[h], (slot_uses & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) DIINN_TR(t, (layer - 1) * 20 + h * 10);
            const uint32_t d_tmem = tmem_u + h * 256;
            // one weight stage = four K = 16 MMAs of A chunk `a0` against it
            auto stage_mmas = [&](uint32_t a0, bool first, bool commit_full) {
              const int st = it % C::kStages;
              mbar_wait(&sm.w_full[st], (it / C::kStages) & 1);
              tc_fence_after();
              const uint32_t b0 = smem_u32(s_w + st * C::kStageBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16<CG>(d_tmem, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc,
                                (first && k == 0) ? 0u : 1u);
                umma_commit<CG>(&sm.w_empty[st]);
                if (commit_full) umma_commit<CG>(&sm.tmem_full[h]);
              }
              __syncwarp();
              ++it;
            };
            auto wait_chunk = [&](int kc) {
              if (h == 0) {  // half 1 re-reads chunks whose readiness half 0 already observed
                if constexpr (CG == 2) mbar_wait_cluster(&sm.act_ready[bin][kc], aph);
                else mbar_wait(&sm.act_ready[bin][kc], aph);
              }
            };
            if constexpr (kSel) {
              // accumulators start as  [ P16[l(row)] | bq ]  = A_sel (one-hot rows) x B_sel; K-chunk 0's barrier also covers
              // this tile's A_sel rows (every epilogue warp writes its share before it signals chunk 0)
              wait_chunk(0);
              if (lane == 0) DIINN_TR(t, 112 + (layer - 1) * 4 + h * 2);
              mbar_wait(&sm.sel_full, sel_it & 1);
              tc_fence_after();
              if (lane == 0) DIINN_TR(t, 113 + (layer - 1) * 4 + h * 2);
              const uint32_t a_sel = asel0 + (t & 1) * kASelBytes;
              constexpr uint32_t idesc_sel = umma_idesc_f16_bmn(128 * CG, 256);
              if (elect_one()) {
                umma_bf16<CG>(d_tmem, umma_desc_k_sw64(a_sel), umma_desc_mn_sw128(sel0, wk.sel_lbo, wk.sel_sbo), idesc_sel, 0u);
                if (wk.ksel == 32)
                  umma_bf16<CG>(d_tmem, umma_desc_k_sw64(a_sel + 32),
                                umma_desc_mn_sw128(sel0 + wk.sel_kstep, wk.sel_lbo, wk.sel_sbo), idesc_sel, 1u);
                umma_commit<CG>(&sm.sel_empty);
              }
              __syncwarp();
              ++sel_it;
            }
            if constexpr (!kSplit) {
#pragma unroll 1
              for (int kc = 0; kc < 4; ++kc) {
                wait_chunk(kc);
                if (lane == 0) DIINN_TR(t, (layer - 1) * 20 + h * 10 + 1 + 2 * kc);
                stage_mmas(a_base + kc * kChunkBytes, kc == 0 && !kSel, kc == 3);
                if (lane == 0) DIINN_TR(t, (layer - 1) * 20 + h * 10 + 2 + 2 * kc);
              }
            } else {
              // Split format. The tensor core adds every MMA's result to the fp32 accumulator with truncation, so each add at
              // full accumulator magnitude costs up to one ulp: the small terms (a_lo.w_hi, a_hi.w_lo: 2^-11 of the result)
              // of a K-chunk pair go in BEFORE that pair's a_hi.w_hi. Chunk pairs (0,1) then (2,3), so that chunks 0, 1 of the
              // in-place activation buffer are dead halfway through half 1 and half 0's epilogue can start overwriting them.
#pragma unroll 1
              for (int g = 0; g < 2; ++g) {
#pragma unroll 1
                for (int kc = 2 * g; kc < 2 * g + 2; ++kc) {
                  wait_chunk(kc);
                  const uint32_t a0 = a_base + kc * kChunkBytes;
                  stage_mmas(a0 + kActBytes, kc == 0, false);  // a_lo . w_hi
                  stage_mmas(a0, false, false);                // a_hi . w_lo
                }
#pragma unroll 1
                for (int kc = 2 * g; kc < 2 * g + 2; ++kc)
                  stage_mmas(a_base + kc * kChunkBytes, false, kc == 3);  // a_hi . w_hi
                if (g == 0 && h == 1) {
                  if (elect_one()) umma_commit<CG>(&sm.a01_free);
                  __syncwarp();
                }
              }
            }
            if (lane == 0) DIINN_TR(t, (layer - 1) * 20 + h * 10 + 9);
          }
          act_phase ^= 1u << bin;
          ++slot_uses;
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== L2 prefetcher =====================
    // P is produced by stage A right before this kernel and is far larger than L2, so a tile's first touch of its P rows would
    // be an HBM-latency fetch in the middle of the pipeline. This warp pulls the rows of the tile TWO work items ahead into L2
    // (cp.async.bulk.prefetch.L2). It used to be the weight producer's job at the top of its tile loop, where the ~1 900 clk the
    // prefetch instructions take (timeline: tools/trace_stage_b.py) delayed the next tile's first B_sel / weight requests and
    // left the tensor pipe idle at every tile boundary.
    auto prefetch = [&](int w_) {
      if constexpr (kPix) prefetch_tile_pixels<CG>(src, P, wk, w_, rank, lane);
      else prefetch_tile_rows<CG, kSel ? 2 : 4>(src, P, wk, w_, rank, lane);
    };
    // Pacing: the weight producer publishes the index of the tile it has started (a plain shared-memory counter, polled
    // here). Deliberately NOT an mbarrier phase: a prefetcher that falls two phases behind (query lists issue 128 prefetches per
    // tile) would wait for a parity that never comes back; with a counter it simply stops waiting until it has caught up.
    if (unit_id + n_units < wk.n_work) prefetch(unit_id + n_units);
    int t3 = 0;
    for (int work = unit_id; work < wk.n_work; work += n_units, ++t3) {
      while (ld_shared_volatile_u32(smem_u32(&sm.pf_tile)) < static_cast<uint32_t>(t3)) nanosleep_ns(200);
      if (work + 2 * n_units < wk.n_work) prefetch(work + 2 * n_units);
      __syncwarp();
    }
    __syncwarp();
  }
  } else {
    // ===================== epilogue warps =====================
    // All 16 warps drain half slot 0, then half slot 1, of every layer. A warp owns one TMEM lane quarter (32 tile
    // rows) and one 16-feature group fg of every 64-feature K-chunk, so a half slot is two steps per warp and EVERY STEP
    // COMPLETES ONE K-CHUNK of the next layer's A operand: the chunk published last -- which gates the next layer's
    // first half slot -- is one step (not a whole half-slot epilogue) behind the layer's last MMA.
    setmaxnreg_inc<kRegsEpi>();
    const int ew = warp - 4;
    const int fg = ew >> 2;             // 16-feature group inside every 64-feature chunk
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;  // tile row == TMEM lane
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool tracer = tracing && warp == 4 && lane == 0;
    uint32_t full_uses = 0;             // completed layers (each half slot is used once per layer)
    int t = 0;
    int work = unit_id;
    RowCtx rc{};

    // P slices of this warp's next two steps. Software pipeline over the whole step sequence of the kernel (layer 0
    // chunk pairs and half slots alike): the step that consumes ka / kb re-issues the load for the same step of the
    // NEXT unit, so every slice has at least one full step (>= 700 clk) plus the barrier wait to arrive from L2.
    float4 ka[4], kb[4];

    // this warp's share of K-chunks [kc0, kc0 + 2) of layer 0 for the tile described by rcx -> activation buffer
    // `bufidx`; nb = P slice base of the next unit (nullptr: none)
    // (select variant: nb counts fp16 elements of the P16 row, and asel >= 0 names the A_sel buffer this tile's one-hot
    // rows go to, written before K-chunk 0 is signalled)
    auto load_p = [&](const float* q, float4 (&v)[4]) {
      if constexpr (kSel || kPix) load16h(reinterpret_cast<const uint16_t*>(q), v);
      else load16(q, v);
    };
    constexpr bool kHalfP = kSel || kPix;     // fp16 P rows: every column offset of the float pointer halves
    constexpr int kPD = kHalfP ? 2 : 1;
    constexpr int kP64 = 64 / kPD;            // 64 P columns, in units of the float pointer that carries the row
    auto layer0_unit = [&](int bufidx, int kc0, const RowCtx& rcx, const float* nb, int asel = -1) {
      const uint32_t buf = act0 + bufidx * kActBytes;
      if constexpr (kSel) {
        if (asel >= 0) {
          // this row's A_sel: 1.0 at its LR cell's slot; this warp writes 16-byte unit fg = slots [8 fg, 8 fg + 8) of the
          // 64-byte row (64B swizzle: unit ^= (row >> 1) & 3)
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int s0 = 8 * fg + 2 * j;
            w[j] = (rcx.slot == s0 ? 0x3C00u : 0u) | (rcx.slot == s0 + 1 ? 0x3C000000u : 0u);
          }
          st_shared_v4(asel0 + asel * kASelBytes + r * 64 + ((fg ^ ((r >> 1) & 3)) << 4), w[0], w[1], w[2], w[3]);
        }
      }
      // (the P prefetch is issued AFTER the proxy fence: fence.proxy.async lowers to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC,
      // which waits for every outstanding global load of the thread -- a prefetch issued just before it exposes its
      // whole L2 latency in every step; measured)
      const uint32_t trow = (rcx.phase < 8 ? tab_lo : tab_hi) + static_cast<uint32_t>(rcx.phase & 7) * 512u;
      const int tkey = rcx.phase & 7;
      const bool canon = wk.canon != 0;
      // Both k_0 slices of the NEXT unit are fetched behind this unit's LAST proxy fence: fence.proxy.async drains the thread's
      // outstanding global loads, so a load issued between the two steps exposed its whole L2 latency at the second fence
      // (timeline: 1 750 clk per unit against 760 for the unit without loads; same-box A/B 1.680 -> 1.663 ms on c3).
      layer0_step<FMT, kPix, kSel, kLiif, kTab>(buf, kc0, fg, r, rcx, sp, ka, trow, tkey, canon);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) signal<CG>(&sm.act_ready[bufidx][kc0]);
      layer0_step<FMT, kPix, kSel, kLiif, kTab>(buf, kc0 + 1, fg, r, rcx, sp, kb, trow, tkey, canon);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) signal<CG>(&sm.act_ready[bufidx][kc0 + 1]);
      if (nb) load_p(nb, ka);
      if (nb) load_p(nb + kP64, kb);
    };

    if (work < wk.n_work) {
      rc = make_row<CG, kPix, kSel, kLiif>(src, out, P, wk, work, rank, r, fg == 0);
      const float* p0 = rc.prow + fg * (16 / kPD);
      load_p(p0, ka);
      load_p(p0 + kP64, kb);
      layer0_unit(0, 0, rc, p0 + 2 * kP64, 0);  // first tile -> buffer 0 (A_sel buffer 0)
      layer0_unit(0, 2, rc, kSel ? nullptr : p0 + kD / kPD);
    }
    for (; work < wk.n_work; work += n_units, ++t) {
      const int X = t & 1;
      const int next_work = work + n_units;
      const bool has_next = next_work < wk.n_work;
      RowCtx rc_next{};
      const float* const pw = rc.prow + fg * (16 / kPD);  // this warp's column of the tile row's P entry
      const float* pn = nullptr;                               // same for the next tile
      float rgb[3] = {0.f, 0.f, 0.f};  // this warp's share of the RGB projection (scalar FFMA: measured faster than FFMA2 here)
      // mode 4: this row's q_3 vector in the dump buffer (rows outside the image / band get no store target)
      __nv_bfloat16* q3row = nullptr;
      float* q3row_f = nullptr;  // the split format dumps fp32
      if constexpr (kDump && !kSplit) q3row = rc.valid ? out.q3 + rc.out_off * kD : nullptr;
      if constexpr (kDump && kSplit) q3row_f = rc.valid ? out.q3f + rc.out_off * kD : nullptr;
#pragma unroll 1
      for (int layer = 1; layer <= 3; ++layer) {
        // L1 writes X^1, L2 writes X, (L3 writes nothing); split format: the one buffer, in place
        const int bout = kSplit ? 0 : ((layer == 2) ? X : (X ^ 1));
        const uint32_t out_base = act0 + bout * kActBytes;
        const bool last = layer == 3;
        if (layer == 2 && has_next) {
          rc_next = make_row<CG, kPix, kSel, kLiif>(src, out, P, wk, next_work, rank, r, fg == 0);
          pn = rc_next.prow + fg * (16 / kPD);
          if constexpr (kSel) {  // only layer 0 reads P here: the next tile's first two k_0 slices, a whole layer ahead
            load_p(pn, ka);
            load_p(pn + kP64, kb);
          }
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          // Split format: this layer's epilogue (and the next tile's layer 0) overwrite the buffer the layer's MMAs read:
          // K-chunks 0, 1 once half 1 is past them (a01_free), chunks 2, 3 once half slot 1 is complete.
          if constexpr (kSplit) {
            if (h == 0) mbar_wait(&sm.a01_free, full_uses & 1);       // chunks 0, 1 (this half's output, layer 0's first pair)
            else mbar_wait(&sm.tmem_full[1], full_uses & 1);          // chunks 2, 3: every MMA of the layer has retired
          }
          // Layer 0 of the next tile goes into buffer X^1, whose last readers are layer 2's MMAs: complete, because
          // this warp has itself consumed tmem_full[1] of layer 2. Its four chunks are interleaved with layer 3's halves.
          if (last && has_next) {
            if constexpr (kSel) layer0_unit(X ^ 1, 2 * h, rc_next, h == 0 ? pn + 2 * kP64 : nullptr, h == 0 ? ((t + 1) & 1) : -1);
            else layer0_unit(kSplit ? 0 : (X ^ 1), 2 * h, rc_next, pw + (3 * kD + h * 128) / kPD);
          }
          // unit after this one: the other half / the next layer / layer 0 of the next tile / the next tile's layer 1
          const float* nb;
          if constexpr (kSel) nb = nullptr;  // the accumulators arrive with P[l] and bq in them
          else if (!last) nb = (h == 0) ? pw + (layer * kD + 128) / kPD : (layer == 1 || !has_next) ? pw + (layer + 1) * kD / kPD : pn;
          else nb = has_next ? (h == 0 ? pn + 128 / kPD : pn + kD / kPD) : (h == 0 ? pw + (3 * kD + 128) / kPD : nullptr);
          const uint32_t tslot = tlane + h * 256;
          if (tracer) DIINN_TR(t, 64 + (layer - 1) * 10 + h * 5);
          mbar_wait(&sm.tmem_full[h], full_uses & 1);
          tc_fence_after();
          if (tracer) DIINN_TR(t, 64 + (layer - 1) * 10 + h * 5 + 1);
          Raw ra, rb;
          epi_load(tslot, fg * 16, ra);
          tmem_ld_wait();
#if DIINN_TOUCH_KA
          { float tch = ka[0].x + ka[1].y + ka[2].z + ka[3].w; asm volatile("" ::"f"(tch) : "memory"); }
#endif
#if DIINN_FINE_TRACE
          if (tracer && layer == 2) DIINN_TR(t, 96 + h * 8 + 0);
#endif
          if (last) epi_math<true, FMT, kDump, kSel, kLiif, kPix>(ra, out_base, layer, h, 0, fg, r, sp, ka, rgb, q3row, q3row_f);
          else epi_math<false, FMT, false, kSel, kLiif, kPix>(ra, out_base, layer, h, 0, fg, r, sp, ka, rgb);
#if DIINN_FINE_TRACE
          if (tracer && layer == 2) DIINN_TR(t, 96 + h * 8 + 1);
#endif
          epi_load(tslot, 64 + fg * 16, rb);
          if (!last) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) signal<CG>(&sm.act_ready[bout][2 * h]);
          }
          if (nb) load_p(nb, ka);  // after the fence (see layer0_unit)
          if (tracer) DIINN_TR(t, 64 + (layer - 1) * 10 + h * 5 + 2);
          tmem_ld_wait();
#if DIINN_FINE_TRACE
          if (tracer && layer == 2) DIINN_TR(t, 96 + h * 8 + 2);
#endif
          tc_fence_before();
          __syncwarp();
          if (lane == 0) signal<CG>(&sm.tmem_empty[h]);
          if (tracer) DIINN_TR(t, 64 + (layer - 1) * 10 + h * 5 + 4);
          if (last) epi_math<true, FMT, kDump, kSel, kLiif, kPix>(rb, out_base, layer, h, 1, fg, r, sp, kb, rgb, q3row, q3row_f);
          else epi_math<false, FMT, false, kSel, kLiif, kPix>(rb, out_base, layer, h, 1, fg, r, sp, kb, rgb);
#if DIINN_FINE_TRACE
          if (tracer && layer == 2) DIINN_TR(t, 96 + h * 8 + 3);
#endif
          if (!last) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) signal<CG>(&sm.act_ready[bout][2 * h + 1]);
          }
          if (nb) load_p(nb + kP64, kb);
          if (tracer) DIINN_TR(t, 64 + (layer - 1) * 10 + h * 5 + 3);
        }
        ++full_uses;
      }
      // Combine the four partial RGB projections of a row (one per 16-feature group) in a fixed order (bit-reproducible).
      // The warps of feature group 3 reduce and store; the others drop their partials in smem.
      const int pidx = fg;
      if constexpr (kDump) {
        // mode 4: nothing to reduce or store here, csrc/mode4.cu projects q_3 through the 3x3 conv afterwards
      } else if (pidx != 3) {
        if (t > 0) named_bar_sync(2, kEpiThreads);  // the reducers have read the previous tile's partials
        float* pp = &sm.partial[pidx][r][0];
        pp[0] = rgb[0], pp[1] = rgb[1], pp[2] = rgb[2];
        named_bar_arrive(1, kEpiThreads);
      } else {
        named_bar_sync(1, kEpiThreads);
        float o[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
          o[ch] = ((sm.partial[0][r][ch] + sm.partial[1][r][ch]) + sm.partial[2][r][ch]) + rgb[ch] + sp.bl[ch];
        if (has_next) named_bar_arrive(2, kEpiThreads);
        if (src.mode == 1 && src.ensemble) {
          // lanes 4j..4j+3 hold the four neighbours of one query: out = sum_v pred_v * area_{3-v} / sum(area)
          // (liif.py:117-127; butterfly sums, every lane of the warp participates)
          const float a_sw = __shfl_xor_sync(0xffffffffu, rc.area, 3);
          float tot = rc.area + __shfl_xor_sync(0xffffffffu, rc.area, 1);
          tot += __shfl_xor_sync(0xffffffffu, tot, 2);
          const float wgt = a_sw / tot;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            float t2 = o[ch] * wgt;
            t2 += __shfl_xor_sync(0xffffffffu, t2, 1);
            t2 += __shfl_xor_sync(0xffffffffu, t2, 2);
            o[ch] = t2;
          }
        }
        if (rc.valid) {
          const int64_t cs = src.mode == 0 ? out.chan_stride : 1;
          store_out(out, rc.out_off, o[0]);
          store_out(out, rc.out_off + cs, o[1]);
          store_out(out, rc.out_off + 2 * cs, o[2]);
        }
      }
      rc = rc_next;
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<CG>(tmem_base, 512);
}

}  // namespace sb

template <int CG, int FMT, bool kDump, bool kPix, bool kSel, bool kLiif = false, bool kTab = false>
static int launch_variant(Handle* h, cudaLaunchConfig_t* cfg, const CUtensorMap& tm, const CUtensorMap& tm_lo,
                          const CUtensorMap& tm_selp, const CUtensorMap& tm_selb, const PixelSource& src, const OutSpec& out,
                          const void* P, const sb::Work& wk, int* err_flag, long long* trace) {
  using namespace sb;
  constexpr int kBytes = static_cast<int>(smem_bytes_tab<CG, kSel, kTab>());
  cfg->dynamicSmemBytes = kBytes;
  DIINN_CUDA_OK(h, cudaFuncSetAttribute(stage_b_umma_kernel<CG, FMT, kDump, kPix, kSel, kLiif, kTab>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kBytes));
  if (getenv("DIINN_DEBUG_OCC")) {
    int nc = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, stage_b_umma_kernel<CG, FMT, kDump, kPix, kSel, kLiif, kTab>, cfg);
    fprintf(stderr, "[diinn] stage B: grid %u CTAs, cluster %d, max active clusters %d (%s)\n", cfg->gridDim.x, CG, nc,
            cudaGetErrorString(e));
  }
  if (src.mode == 0 && !kPix) {
    // grid decode: `ratio` is one constant per image, so Q.0's w_ratio * ratio + b is folded into the bias here
    static thread_local SmallParams spf;
    spf = h->small;
    for (int i = 0; i < kD / 2; ++i) {
      spf.wq0_p[i][3].x = fmaf(spf.wq0_p[i][2].x, src.ratio, spf.wq0_p[i][3].x);
      spf.wq0_p[i][3].y = fmaf(spf.wq0_p[i][2].y, src.ratio, spf.wq0_p[i][3].y);
    }
    spf.q0_folded = 1;
    DIINN_CUDA_OK(h, cudaLaunchKernelEx(cfg, stage_b_umma_kernel<CG, FMT, kDump, kPix, kSel, kLiif, kTab>, tm, tm_lo, tm_selp, tm_selb, spf,
                                        src, out, P, wk, err_flag, trace));
    return DIINN_OK;
  }
  DIINN_CUDA_OK(h, cudaLaunchKernelEx(cfg, stage_b_umma_kernel<CG, FMT, kDump, kPix, kSel, kLiif, kTab>, tm, tm_lo, tm_selp, tm_selb,
                                      h->small, src, out, P, wk, err_flag, trace));
  return DIINN_OK;
}

namespace {

// host twin of axis_index() (common.cuh): one rounded fp32 multiply, then floorf
inline int host_axis_index(const AxisParams& p, int j) {
  volatile float t = (static_cast<float>(j) + 0.5f) * p.scale;
  const int i = static_cast<int>(floorf(t));
  return i < p.n_in - 1 ? i : p.n_in - 1;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Geometry of one stage-B launch: patch shape, work items, and whether the select-MMA variant applies (then P is fp16).
struct Plan {
  sb::Work wk{};
  int cta_group = 2;
  bool sel = false;
  bool tab = false;  // layer 0 from the phase table (Work::canon and room for the table)
};

Plan make_plan(const Handle* h, const PixelSource& src, int cta_group, int fmt, bool chain_mode) {
  using namespace sb;
  Plan pl;
  if (cta_group == 0) {
    static int env_cg = -1;
    if (env_cg < 0) env_cg = env_int("DIINN_CTA_GROUP", 2) == 1 ? 1 : 2;
    cta_group = env_cg;
  }
  pl.cta_group = cta_group;
  Work& wk = pl.wk;
  wk.pw_log2 = kPatchWLog2Default;
  const bool pix = src.per_pixel_p != 0;
  if (src.mode != 0) {
    const int64_t total = static_cast<int64_t>(src.B) * src.Q * (src.ensemble ? 4 : 1);
    wk.n_work = static_cast<int>((total + kTileM * cta_group - 1) / (kTileM * cta_group));
    return pl;
  }
  // Select variant: CTA pairs, 16-bit formats, LR-resolution P (modes 1 / 2: the K chain updates the fp16 P), a pair patch of <= 30 LR cells.
  // Whether it applies must NOT depend on the row range or the patch shape a launch ends up with: row tiles of one image
  // have to be bit-identical to the full decode, and the two variants differ in the last bits (fp16 P added inside the
  // tensor core vs fp32 P added by the epilogue). So the decision is taken for the DEFAULT patch shape from a bound that
  // only knows the scale factors -- n pixels of an axis touch at most floor((n - 1) * n_in / n_up) + 2 source cells -- and
  // the wave-count search below only considers shapes on the same side of it.
  static const int env_nosel = env_int("DIINN_NO_SEL", 0);
  const bool sel_candidate = cta_group == 2 && fmt != kFmtSplit && !pix && !env_nosel;
  auto box_of = [&](int l2, int max_rows, int& br, int& bc) {
    const int pw = 1 << l2, ph = kTileM >> l2;
    br = static_cast<int>(floor(static_cast<double>(ph - 1) * src.H / src.H_up)) + 2;
    bc = static_cast<int>(floor(static_cast<double>(2 * pw - 1) * src.W / src.W_up)) + 2;
    // columns always start at 0: on an integer scale factor that divides the pair's width every pair covers exactly
    // 2 pw / s_w cells (x4, 4x32 patches: 16, where the bound says 17). Rows start wherever the caller's tile does: bound kept.
    if (src.W_up % src.W == 0 && (2 * pw) % (src.W_up / src.W) == 0) bc = 2 * pw / (src.W_up / src.W);
    br = br < max_rows ? br : max_rows;
    bc = bc < src.W ? bc : src.W;
  };
  auto sel_ok = [&](int l2) {
    int br, bc;
    box_of(l2, src.H, br, bc);  // (whole-image extents: the same answer for every row tile)
    return sel_candidate && br * bc <= 32;  // every B_sel row may hold a cell (the Q bias rides in all of them)
  };
  pl.sel = sel_ok(kPatchWLog2Default);
  // patch shape: fewest waves over the launch's CTA (pair)s; ties keep the default 8x16 (DIINN_PATCH_W_LOG2 pins it)
  static const int env_pw = env_int("DIINN_PATCH_W_LOG2", -1);
  const int max_units = h->sm_count / cta_group > 0 ? h->sm_count / cta_group : 1;
  long long best_waves = -1;
  const int cand[4] = {4, 5, 3, 6};
  for (int ci = 0; ci < 4; ++ci) {
    const int l2 = cand[ci];
    if (env_pw >= 3 && env_pw <= 6 && l2 != env_pw && sel_ok(env_pw) == pl.sel) continue;
    if (pix && l2 != kPatchWLog2Default) continue;  // per-pixel P (init_q): the chunk geometry assumes 8-row patches
    if (sel_ok(l2) != pl.sel) continue;
    const int pw = 1 << l2, ph = kTileM >> l2;
    const long long n = static_cast<long long>(src.B) * ((src.row1 - src.row0 + ph - 1) / ph) *
                        (((src.W_up + pw - 1) / pw + cta_group - 1) / cta_group);
    const long long waves = (n + max_units - 1) / max_units;
    // (another shape must save at least 2 % of the waves: the default's timings are the measured ones)
    if (best_waves < 0 || waves * 50 <= best_waves * 49) best_waves = waves, wk.pw_log2 = l2;
  }
  const int pw = 1 << wk.pw_log2, ph = kTileM >> wk.pw_log2;
  const int tiles_x = (src.W_up + pw - 1) / pw;
  wk.tiles_y = (src.row1 - src.row0 + ph - 1) / ph;
  wk.n_txp = (tiles_x + cta_group - 1) / cta_group;
  wk.n_work = src.B * wk.tiles_y * wk.n_txp;
  if (pl.sel) {
    box_of(wk.pw_log2, src.lr_rows, wk.box_r, wk.box_c);  // (the P16 tensor of this launch holds lr_rows LR rows)
    // The bound above decides WHETHER the variant runs; the box itself is the exact extent of this launch's patches (x4 with
    // aligned 8 x 32 patches: 2 x 8 = 16 cells where the bound says 27), which is what picks one or two K = 16 select MMAs.
    // Unselected slots multiply zeros, so K_sel does not change a single bit of the result.
    int er = 1, ec = 1;
    for (int ty = 0; ty < wk.tiles_y; ++ty) {
      const int a = src.row0 + ty * ph, b = (a + ph - 1 < src.row1 - 1) ? a + ph - 1 : src.row1 - 1;
      const int n = host_axis_index(src.ax_h, b) - host_axis_index(src.ax_h, a) + 1;
      er = n > er ? n : er;
    }
    for (int tx = 0; tx < wk.n_txp; ++tx) {
      const int a = (tx * cta_group * pw < src.W_up - 1) ? tx * cta_group * pw : src.W_up - 1;
      const int b = (a + cta_group * pw - 1 < src.W_up - 1) ? a + cta_group * pw - 1 : src.W_up - 1;
      const int n = host_axis_index(src.ax_w, b) - host_axis_index(src.ax_w, a) + 1;
      ec = n > ec ? n : ec;
    }
    wk.box_r = er < wk.box_r ? er : wk.box_r;
    wk.box_c = ec < wk.box_c ? ec : wk.box_c;
    wk.ksel = wk.box_r * wk.box_c <= 16 ? 16 : 32;
    // B_sel tile of one CTA: two 64-feature blocks of K_sel rows x 128 B; MN-major SWIZZLE_128B atoms are 8 K-rows (1 KB)
    wk.sel_lbo = static_cast<uint32_t>(env_int("DIINN_SEL_LBO", wk.ksel * 128));
    wk.sel_sbo = static_cast<uint32_t>(env_int("DIINN_SEL_SBO", 1024));
    wk.sel_kstep = static_cast<uint32_t>(env_int("DIINN_SEL_KSTEP", 2048));
  }
  // Canonical relative coordinates + phase table (Work::canon). Decided from the image geometry and the format alone, so every
  // row tile of an image agrees; only whether the table fits (select variant: K_sel = 16) may differ between launches, and
  // the computing and the looking-up kernels are bit-identical (layer0_step). nearest-exact picks floor(j / s) exactly as long
  // as the fp32 rounding of (j + 0.5) * fl(1/s) (<= n_in * 2^-23) stays below the 1/(2s) distance to the next integer.
  static const int env_nocanon = env_int("DIINN_NO_CANON", 0), env_notab = env_int("DIINN_NO_TAB", 0);
  if (cta_group == 2 && fmt != kFmtSplit && !pix && !src.liif && !env_nocanon && src.H_up % src.H == 0 &&
      src.W_up % src.W == 0 && src.H_up <= (1 << 20) && src.W_up <= (1 << 20)) {
    const int s_h = src.H_up / src.H, s_w = src.W_up / src.W;
    if (s_h * s_w <= 16) {
      wk.canon = 1, wk.s_h = s_h, wk.s_w = s_w;
      pl.tab = !env_notab && (!pl.sel || wk.ksel == 16);
    }
  }
  return pl;
}

}  // namespace

// Host-only view of make_plan for tests (diinn_debug_plan_stage_b): which kernel a launch would take, without a device.
// out[12] = {select variant, phase table, canonical coordinates, s_h, s_w, log2 patch width, K_sel, box rows, box columns,
//            work items, tiles_y, pairs per patch row}
void plan_stage_b_probe(int sm_count, int decoder_mode, int B, int H, int W, int H_up, int W_up, int row0, int row1, int fmt,
                        int* out) {
  Handle h;
  h.sm_count = sm_count;
  h.cfg.mode = decoder_mode;
  PixelSource src{};
  src.mode = 0;
  src.ax_h = make_axis(H, H_up);
  src.ax_w = make_axis(W, W_up);
  src.B = B, src.H = H, src.W = W, src.H_up = H_up, src.W_up = W_up, src.row0 = row0, src.row1 = row1;
  src.lr_row0 = 0, src.lr_rows = H;
  const Plan pl = make_plan(&h, src, 0, fmt, decoder_mode == 1 || decoder_mode == 2);
  const int v[12] = {pl.sel, pl.tab, pl.wk.canon, pl.wk.s_h, pl.wk.s_w, pl.wk.pw_log2, pl.sel ? pl.wk.ksel : 0,
                     pl.sel ? pl.wk.box_r : 0, pl.sel ? pl.wk.box_c : 0, pl.wk.n_work, pl.wk.tiles_y, pl.wk.n_txp};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
}

// Does the stage-B launch for this source take the select-MMA variant, i.e. must stage A write P as fp16?
bool stage_b_wants_p16(const Handle* h, const PixelSource& src, int fmt) {
  return make_plan(h, src, 0, fmt, h->cfg.mode == 1 || h->cfg.mode == 2).sel;
}

// fmt: kFmtBf16 / kFmtF16 / kFmtSplit (handle.h). P: fp32 rows -- fp16 rows when stage_b_wants_p16() says so. tap: debug only.
int launch_stage_b_umma(Handle* h, const PixelSource& src, const OutSpec& out, const void* P, int cta_group, int fmt,
                        cudaStream_t s, int4* tap) {
  using namespace sb;
  if (fmt < 0 || fmt > kFmtSplit) return fail(h, DIINN_ERR_BAD_DTYPE, "stage B: unknown operand format");
  Plan pl = make_plan(h, src, cta_group, fmt, h->cfg.mode == 1 || h->cfg.mode == 2);
  cta_group = pl.cta_group;
  Work& wk = pl.wk;
  wk.tap = tap;
  int* err_flag = h->err_flag;   // per-handle device scratch, allocated by diinn_create (no allocation in decode)
  long long* trace = h->trace_dev;
  const bool pix = src.per_pixel_p != 0;

  int units = h->sm_count / cta_group;
  if (units > wk.n_work) units = wk.n_work;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(units * cta_group, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
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
  // instantiations: kDump = mode 4 (q_3 dumped instead of projected), kPix = init_q=True (per-pixel P, q_0 given)
  const bool dump = out.q3 != nullptr || out.q3f != nullptr;
  if (pix && src.mode != 0) return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "per-pixel P is implemented for the HR grid only");
  if (fmt == kFmtSplit && pix) return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "the split format is not wired for per-pixel P");
  if (dump && ((fmt == kFmtSplit) != (out.q3f != nullptr)))
    return fail(h, DIINN_ERR_BAD_ARG, "stage B: the split format dumps fp32 q_3, the 16-bit formats bf16");
  const int ci = cta_group - 1;
  const CUtensorMap& tm = h->tmapWB[fmt == kFmtBf16 ? 0 : 1][ci];
  const CUtensorMap& tm_lo = h->tmapWBlo[ci];
  CUtensorMap tm_selp = tm;  // placeholders unless the select variant runs
  const CUtensorMap& tm_selb = pl.sel ? h->tmapSelB[wk.ksel == 16 ? 0 : 1] : tm;
  if (pl.sel) {
    // fp16 P as (1024 features, W, LR rows, B); box = 64 features x the pair's LR patch
    const uint64_t dims[4] = {static_cast<uint64_t>(kPCols), static_cast<uint64_t>(src.W), static_cast<uint64_t>(src.lr_rows),
                              static_cast<uint64_t>(src.B)};
    const uint64_t strides[3] = {kPCols * 2ull, static_cast<uint64_t>(src.W) * kPCols * 2ull,
                                 static_cast<uint64_t>(src.lr_rows) * src.W * kPCols * 2ull};
    const uint32_t box[4] = {64, static_cast<uint32_t>(wk.box_c), static_cast<uint32_t>(wk.box_r), 1};
    int rc0 = make_tmap_4d_bf16(h, &tm_selp, P, dims, strides, box);  // (16-bit elements; TMA does no arithmetic)
    if (rc0) return rc0;
  }
  int rc;
#define DIINN_SB_LAUNCH(CGv, FMTv, DUMPv, PIXv, SELv) \
  launch_variant<CGv, FMTv, DUMPv, PIXv, SELv>(h, &cfg, tm, tm_lo, tm_selp, tm_selb, src, out, P, wk, err_flag, trace)
#define DIINN_SB_PICK(CGv, FMTv)                                                                                     \
  (dump ? (pix ? DIINN_SB_LAUNCH(CGv, FMTv, true, true, false) : DIINN_SB_LAUNCH(CGv, FMTv, true, false, false))       \
        : (pix ? DIINN_SB_LAUNCH(CGv, FMTv, false, true, false) : DIINN_SB_LAUNCH(CGv, FMTv, false, false, false)))
#define DIINN_SB_PICK_SEL(FMTv) (dump ? DIINN_SB_LAUNCH(2, FMTv, true, false, true) : DIINN_SB_LAUNCH(2, FMTv, false, false, true))
#define DIINN_SB_TAB(FMTv, DUMPv, SELv) \
  launch_variant<2, FMTv, DUMPv, false, SELv, false, true>(h, &cfg, tm, tm_lo, tm_selp, tm_selb, src, out, P, wk, err_flag, trace)
#define DIINN_SB_PICK_TAB(FMTv)                                                                    \
  (pl.sel ? (dump ? DIINN_SB_TAB(FMTv, true, true) : DIINN_SB_TAB(FMTv, false, true))              \
          : (dump ? DIINN_SB_TAB(FMTv, true, false) : DIINN_SB_TAB(FMTv, false, false)))
#define DIINN_SB_PICK_SPLIT(CGv) \
  (dump ? DIINN_SB_LAUNCH(CGv, 2, true, false, false) : DIINN_SB_LAUNCH(CGv, 2, false, false, false))
  if (src.liif) {
    if (src.mode != 1 || dump || pix || cta_group != 2)
      return fail(h, DIINN_ERR_UNSUPPORTED_MODE, "LIIF's imnet runs on query lists with CTA pairs only");
#define DIINN_SB_LIIF(FMTv) \
  launch_variant<2, FMTv, false, false, false, true>(h, &cfg, tm, tm_lo, tm_selp, tm_selb, src, out, P, wk, err_flag, trace)
    rc = fmt == kFmtSplit ? DIINN_SB_LIIF(2) : fmt == kFmtF16 ? DIINN_SB_LIIF(1) : DIINN_SB_LIIF(0);
#undef DIINN_SB_LIIF
  } else if (pl.tab) rc = fmt == kFmtF16 ? DIINN_SB_PICK_TAB(1) : DIINN_SB_PICK_TAB(0);
  else if (pl.sel) rc = fmt == kFmtF16 ? DIINN_SB_PICK_SEL(1) : DIINN_SB_PICK_SEL(0);
  else if (cta_group == 1) rc = fmt == kFmtSplit ? DIINN_SB_PICK_SPLIT(1) : fmt == kFmtF16 ? DIINN_SB_PICK(1, 1) : DIINN_SB_PICK(1, 0);
  else rc = fmt == kFmtSplit ? DIINN_SB_PICK_SPLIT(2) : fmt == kFmtF16 ? DIINN_SB_PICK(2, 1) : DIINN_SB_PICK(2, 0);
#undef DIINN_SB_PICK_SPLIT
#undef DIINN_SB_PICK_TAB
#undef DIINN_SB_TAB
#undef DIINN_SB_PICK_SEL
#undef DIINN_SB_PICK
#undef DIINN_SB_LAUNCH
  if (rc) return rc;
  h->launches += 1;
  return DIINN_OK;
}

}  // namespace diinn

// 3x3 SAME convolution block on the 5th-generation tensor cores (tcgen05 / TMEM), fp32 in and
// out, 3xTF32 split precision — nnlib.run_cnn / run_dcnn layers (nnlib.py:214-255, :339-402).
//
// Why 3xTF32: a single TF32 pass (10-bit mantissa) moves the attention box by ~1e-3 px and the
// sharp sigmoid(gamma*v-5) edges turn that into 1e-2..1e-1 errors in attn_box / y_out, far
// outside the 1e-3 parity bar (measured with the oracle).  Each operand is split
// v = hi + lo (hi = v rounded to the nearest tf32, lo = v - hi, exact) and
// D = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, which is ~2^-22 accurate.
//
// Implicit GEMM without im2col: a padded input tile is staged as "pixel slots"
// [(TH+2) rows x TWP = TW+2 columns] in a channel-plane layout [c/4][slot][4 floats] (16 B per
// slot per plane).  For a K-major SWIZZLE_NONE shared-memory descriptor (8 rows x 16 B core
// matrices, SBO = 128 B, LBO = plane stride) a filter tap (ky,kx) is then nothing but a START
// ADDRESS offset of (ky*TWP + kx) slots: the A operand of every tap is the same buffer.
// M = 128 consecutive output slots per MMA (2 of every TWP are halo garbage and are dropped in
// the epilogue), N = the CTA's share of Cout (padded to 16), K = 8 channels per instruction.
//
// Persistent, warp-specialised pipeline (one CTA per SM, 544 threads):
//   warp 16     TMA: one lane issues, per (tile, channel-chunk) item, one cp.async.bulk.tensor (4-D
//               tiled map over the NHWC activation, box = [4 channels][TW+2][TH+2][1]) per channel
//               plane - the box lands in shared memory already in the plane layout, image borders
//               are the TMA's out-of-bounds zero fill - plus one cp.async.bulk for the filter slice;
//               completion is counted on raw[s] (mbarrier complete_tx).
//   warps 8-15  converters: split the landed fp32 values IN PLACE into hi (tf32, nearest) and lo planes
//               (for the transposed convolutions: scatter the low-resolution box into the zero-inserted
//               tile); fence.proxy.async + mbarrier arrive (full[s]).  Layers whose channel counts are
//               not multiples of 4 (no legal tensor map) are gathered by these warps with plain loads.
//   warps 4-7   one lane each issues its share of the n_mt x 9 taps x (2|3) tcgen05.mma per item, tcgen05.commit ->
//               empty[s]; after a tile's last chunk tcgen05.commit -> tmem_full[buf].
//   warps 0-3   epilogue: tcgen05.ld (thread = TMEM lane = slot), folded BN scale/shift, ReLU,
//               2x2 max-pool (horizontal by shuffle, vertical through a small shared tile),
//               stores; arrive tmem_empty[buf].  Accumulators are double buffered in TMEM so the
//               epilogue of tile i overlaps the MMAs of tile i+1.
// Narrow layers (N <= 32) are bound by the tensor core's shared-memory reads of A, so there the
// two filter parts are stacked along N ([B_hi; B_lo], one MMA with N' = 2N reads A_hi once) and
// summed in the epilogue; small feature maps split N over CTAs to fill the 148 SMs.
#include <cuda.h>  // CUtensorMap and its enums only; the encoder is fetched with cudaGetDriverEntryPoint
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace {

// Warp roles.  The epilogue is the longest stage of the pipeline on the large layers (tools/conv_timeline.py wait
// accounting, profiles/r04f_*: 81 % of controller layer 1 with four epilogue warps, the MMA warps waiting for free
// accumulators), so it is run by TWO groups of four warps: a warp may read the TMEM lanes [32 (w % 4), +32), group A = warps
// 0-3 takes the even m-tiles of a tile, group B = warps 15-18 the odd ones.
constexpr int kEpiThreads = 128;   // one epilogue group: warps 0-3 (A), warps 15-18 (B)
constexpr int kEpiAll = 2 * kEpiThreads;
constexpr int kMmaWarp0 = 4;       // warps 4-7: one issuing lane each, m-tiles dealt round-robin
constexpr int kMmaWarps = 4;
constexpr int kProdThreads = 224;  // warps 8-14 (seven: 20 warps = 5 per SM sub-partition keep 96 registers per thread)
constexpr int kProdWarp0 = kMmaWarp0 + kMmaWarps;
constexpr int kEpiWarpB0 = kProdWarp0 + kProdThreads / 32;  // warp 15 (group B = warps 15-18: lane quarters 3, 0, 1, 2)
constexpr int kTmaWarp = kEpiWarpB0 + kEpiThreads / 32;      // warp 19
constexpr int kThreads = 32 * (kTmaWarp + 1);                // 640
constexpr int kMaxStages = 4;
constexpr int kStageUnroll = 4;
constexpr int kKsplitDefault = 8;  // partial accumulators per m-tile used for precision (see make_plan)
constexpr int kPoolLd = 20;  // floats per staged half-row of a 16-column chunk (16 + 4 padding)

struct UmmaConvParams {
  const float *x1;
  const float *x2;
  const float *wpack;  // [n_split][chunks][9][KC/4][2*NPc][4]   (rows 0..NPc-1 = hi, NPc.. = lo)
  const float *scale;
  const float *shift;
  float *y;
  int C1, C2, Cin, Cout;
  int B, Hin, Win, Hout, Wout;
  int up, pool, relu;
  int TH, TW, TWP, n_mt, KC, n_chunks;
  int NPc, n_split, merged;  // output channels per CTA (padded), CTAs along N, stacked-B mode
  int rowstack;              // the three kx taps of a filter row stacked along N (see issue_chunk_rs)
  int mt_stride;             // output slots per m-tile: 128, or 126 with rowstack (tiles overlap by two rows of D)
  int slots_alloc;           // input slots allocated per plane
  int tiles_x, tiles_y, n_items;
  int stages;
  int w_resident;   // the CTA's whole filter image stays in shared memory for all of its tiles
  int w_res_bytes;  // bytes of that resident image (0 when streaming)
  int acc_cols;     // TMEM columns of one accumulator buffer (n_mt * ksplit * cols_per_mt)
  int ksplit;       // independent accumulators per m-tile: (tap, k8) step q goes to accumulator q % ksplit
  int nbuf;         // accumulator buffers in TMEM (2: the epilogue of tile i overlaps the MMAs of tile i+1)
  int stage_bytes;  // bytes of one shared-memory stage (input hi + lo + filter slice)
  int vec4;         // both sources have channel counts divisible by 4
  int tma;          // inputs (and streamed filters) arrive by TMA; the producer warps only convert
  int RW, RH;       // TMA box: columns x rows of input pixels (low resolution when up == 2)
  int raw_plane_bytes;  // up == 2 only: bytes per channel plane of the landed low-resolution box
  int raw_off;      // up == 2 only: offset of that landing zone inside a stage
  int split_corr;   // merged mode: accumulate A_lo B_hi in the second accumulator half (RA_UMMA_JOINT_CORR=1: first)
  int four_term;    // merged mode: the A_lo instruction also spans [B_hi; B_lo] (adds the lo x lo partial product)
  int f16;          // fp16 hi / lo operand split (ra_conv3x3_umma_set_f16): kind::f16, K = 16 per instruction
  int fold;         // f16: the tensor core adds the 2^11-scaled correction half into the main half at the end of a tile
                    // (kind::tf32 MMAs with the correction columns as TMEM A operand and 2^-11 I as B): the epilogue reads
                    // N columns per m-tile instead of 2N
  int pdl;          // launched with programmatic stream serialization: griddepcontrol.wait before touching activations
  int grid;         // CTAs that serve this layer (= gridDim.x of a single-layer launch; <= gridDim.x inside a chain)
  long long *dbg;   // optional per-CTA timeline (ra_debug_conv_timeline), 8 slots per CTA
};

__device__ __forceinline__ bool getenv_split_corr(const UmmaConvParams &p) { return p.split_corr != 0; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // sm_100 descriptor version; layout_type 0 = SWIZZLE_NONE
  return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::f16 (fp16 operands, fp32 accumulate): the same shared-memory operand rows (32 bytes) hold K = 16 elements.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from TENSOR MEMORY (columns a_tmem.. of the CTA's allocation: lane = row, one 32-bit column per K element),
// B from shared memory: D[128 x N] (+)= A[128 x 8] B[8 x N].
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool F16>
__device__ __forceinline__ void umma_k(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  if (F16)
    umma_f16(d_tmem, a_desc, b_desc, idesc, accumulate);
  else
    umma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}

// One lane of a converged warp (always the same one for a full mask).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}

__device__ __forceinline__ void mbar_inval(uint32_t mbar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}

// One box of a 4-D tiled tensor map -> shared memory; completion (bytes) is counted on mbar.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3,
                                            uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Contiguous global -> shared bulk copy (filter slices).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n\t"  // suspend-time hint: fewer re-polls
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}

// hi part of the 3xTF32 split: v rounded to NEAREST tf32 (not truncated), so |lo| <= 2^-12 |v| and the
// tensor core's own truncation of lo costs 2^-23 instead of 2^-22.
__device__ __forceinline__ float tf32_hi(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// Asynchronous TMEM load of 16 columns (no wait): pair with tmem_wait16 before the values are used.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// tcgen05.wait::ld; the registers are in/out operands so that no use of them can be scheduled above the wait.
__device__ __forceinline__ void tmem_wait16(uint32_t *r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

struct Item {
  int b, y0, x0;
};

// CTA c serves output-channel split ns = c % n_split and the spatial tiles c / n_split + k * (grid / n_split).
__device__ __forceinline__ Item decode_tile(const UmmaConvParams &p, int t) {
  Item it;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  it.b = t / p.tiles_y;
  it.y0 = ty * p.TH;
  it.x0 = tx * p.TW;
  return it;
}

// Operands of one chunk's MMAs, all warp-uniform (see the MMA issuer warps).
struct IssueCtx {
  uint64_t a0_hi, a1_hi, a0_lo, a1_lo, b;  // A descriptors of the warp's two m-tiles (hi / lo parts), filter desc
  uint32_t d0, d1;                         // TMEM columns of the two m-tiles' first accumulator
  uint32_t idesc_n, idesc_2n, idesc_lo;  // idesc_lo: the A_lo instruction of the merged mode (N or 2N wide)
  uint32_t TWP, a_k8, b_k8, b_tap, b_lo, cols_mt, wrap, init_steps, lo_col;
};

// The 9 taps x K8N k8-steps of one channel chunk, run by the ELECTED lane only (the caller branches on it): the
// compiler then knows a single thread is active and feeds the MMA's uniform-register operands with plain R2UR
// moves from ordinary integer arithmetic.  Taps are a rolled loop - fully unrolled variants made the kernel 335 KB
// of code and every launch paid the instruction-cache misses; the k8 steps and the warp's one or two m-tiles (TWO)
// are unrolled.  Step q accumulates into partial accumulator q % ksplit (column offset jc, wraps at wrap).
template <int K8N, bool MERGED, bool TWO, bool F16 = false>
__device__ __forceinline__ void issue_chunk(const IssueCtx &c) {
  uint32_t jc = 0, a_row = 0, b_tap = 0;
  int q = 0;
#pragma unroll 1
  for (int ky = 0; ky < 3; ++ky, a_row += c.TWP) {
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx, b_tap += c.b_tap) {
#pragma unroll
      for (int k8 = 0; k8 < K8N; ++k8, ++q) {
        const uint64_t a_off = (uint64_t)(a_row + (uint32_t)kx + (uint32_t)k8 * c.a_k8);
        const uint64_t b_hi = c.b + (uint64_t)(b_tap + (uint32_t)k8 * c.b_k8);
        const uint32_t flag = (uint32_t)q < c.init_steps ? 0u : 1u;
        const uint32_t d0 = c.d0 + jc, d1 = c.d1 + jc;
        const uint64_t a0h = c.a0_hi + a_off, a1h = c.a1_hi + a_off, a0l = c.a0_lo + a_off, a1l = c.a1_lo + a_off;
        if (MERGED) {
          // D[:, 0:N] += A_hi B_hi and D[:, N:2N] += A_hi B_lo in ONE instruction, then D[:, 0:N] += A_lo B_hi.
          // RA_UMMA_4TERM=1 widens the A_lo instruction to [B_hi; B_lo] as well (adds lo x lo, 2^-22 of a product):
          // MEASURED to change nothing in the end-to-end error (tools/dbg_parity.py: ctrl_out 5e-6 either way) and
          // to cost 0.2 ms per forward, so the three-term form is the default.
          // The A_lo B_hi correction goes to the SECOND half of the accumulator (c.lo_col = N), next to A_hi B_lo:
          // the tensor core truncates every fp32 accumulation, a bias proportional to the accumulator's magnitude,
          // so the full-magnitude half D[:, 0:N] should see one add per step, not two (tools/dbg_parity.py:
          // controller output error vs the fp32 oracle with both corrections in one half / split).
          umma_k<F16>(d0, a0h, b_hi, c.idesc_2n, flag);
          if (TWO) umma_k<F16>(d1, a1h, b_hi, c.idesc_2n, flag);
          umma_k<F16>(d0 + c.lo_col, a0l, b_hi, c.idesc_lo, 1u);
          if (TWO) umma_k<F16>(d1 + c.lo_col, a1l, b_hi, c.idesc_lo, 1u);
        } else {
          const uint64_t b_lo = b_hi + (uint64_t)c.b_lo;
          umma_tf32(d0, a0h, b_hi, c.idesc_n, flag);
          if (TWO) umma_tf32(d1, a1h, b_hi, c.idesc_n, flag);
          umma_tf32(d0, a0h, b_lo, c.idesc_n, 1u);
          if (TWO) umma_tf32(d1, a1h, b_lo, c.idesc_n, 1u);
          umma_tf32(d0, a0l, b_hi, c.idesc_n, 1u);
          if (TWO) umma_tf32(d1, a1l, b_hi, c.idesc_n, 1u);
        }
        jc += c.cols_mt;
        if (jc == c.wrap) jc = 0;
      }
    }
  }
}

// Row-stacked taps (narrow layers, 6 * NPc <= 256).  A tcgen05.mma with M = 128, K = 8 occupies the tensor core for
// max(48, N/2) cycles whatever N <= 96 is - the A operand streams from shared memory - so a layer with 16 or 32 output
// channels pays the A read nine times per k8 step for very little math.  Here the three kx taps of a filter row share
// ONE read of A: B is stacked along N as [hi kx0 | hi kx1 | hi kx2 | lo kx0 | lo kx1 | lo kx2] (6 NPc columns) and
//   D'[r, kx * NPc + co] (+)= sum_c A[r + ky * TWP, c] * W[ky, kx, c, co]
// is accumulated for every A row r; the output of slot s is D'[s, kx=0] + D'[s+1, kx=1] + D'[s+2, kx=2], combined by
// the epilogue across TMEM lanes (m-tiles overlap by two rows so that all three live in one tile).  Two instructions
// per (ky, k8) step instead of six: A_hi x [all six blocks], then A_lo x [hi blocks] into the lo half.
template <int K8N, bool TWO>
__device__ __forceinline__ void issue_chunk_rs(const IssueCtx &c) {
  uint32_t jc = 0, a_row = 0, b_row = 0;
  int q = 0;
#pragma unroll 1
  for (int ky = 0; ky < 3; ++ky, a_row += c.TWP, b_row += c.b_tap) {
#pragma unroll
    for (int k8 = 0; k8 < K8N; ++k8, ++q) {
      const uint64_t a_off = (uint64_t)(a_row + (uint32_t)k8 * c.a_k8);
      const uint64_t b = c.b + (uint64_t)(b_row + (uint32_t)k8 * c.b_k8);
      const uint32_t flag = (uint32_t)q < c.init_steps ? 0u : 1u;
      const uint32_t d0 = c.d0 + jc, d1 = c.d1 + jc;
      umma_tf32(d0, c.a0_hi + a_off, b, c.idesc_2n, flag);
      if (TWO) umma_tf32(d1, c.a1_hi + a_off, b, c.idesc_2n, flag);
      umma_tf32(d0 + c.lo_col, c.a0_lo + a_off, b, c.idesc_n, 1u);
      if (TWO) umma_tf32(d1 + c.lo_col, c.a1_lo + a_off, b, c.idesc_n, 1u);
      jc += c.cols_mt;
      if (jc == c.wrap) jc = 0;
    }
  }
}

template <int K8N, bool F16 = false>
__device__ __forceinline__ void issue_chunk_k8(const IssueCtx &c, bool merged, bool two) {
  if constexpr (F16) {  // merged only (make_plan)
    if (two)
      issue_chunk<K8N, true, true, true>(c);
    else
      issue_chunk<K8N, true, false, true>(c);
    return;
  }
  if (merged) {
    if (two)
      issue_chunk<K8N, true, true>(c);
    else
      issue_chunk<K8N, true, false>(c);
  } else {
    if (two)
      issue_chunk<K8N, false, true>(c);
    else
      issue_chunk<K8N, false, false>(c);
  }
}

// Wait accounting of the timeline (slots 8-14: cycles the TMA warp waited for a free stage, the converters for a landed
// box, MMA warp 0 for a staged chunk / for a free accumulator, the epilogue for a complete accumulator; cycles the epilogue /
// the converters were busy): who starves whom.
#define RA_WAIT(mbar, parity, acc)            \
  do {                                        \
    if ((F16 && !CHAIN && p.dbg != nullptr)) {                   \
      const long long t0_ = clock64();        \
      mbar_wait(mbar, parity);                \
      (acc) += clock64() - t0_;               \
    } else {                                  \
      mbar_wait(mbar, parity);                \
    }                                         \
  } while (0)
#define RA_DBG_PUT(slot, v)                                                            \
  do {                                                                                 \
    if ((F16 && !CHAIN && p.dbg != nullptr)) p.dbg[(size_t)blockIdx.x * 16 + (slot)] = (long long)(v);    \
  } while (0)
#define RA_DBG(slot)                                                                                   \
  do {                                                                                                 \
    if ((!CHAIN && p.dbg != nullptr)) p.dbg[(size_t)blockIdx.x * 16 + (slot)] = clock64();                          \
  } while (0)

// hi / lo' of two activations, each packed as a half2 (first value in the low half): hi = fp16(x), lo' = fp16((x - hi) * 2^11).
// Two packing conversions and two unpacking adds per pair (F2FP.PACK_AB / HADD2.F32: ALU / FMA pipe); a formulation that
// built hi with integer rounding to save conversions cost 2.5x the instructions and made the converters slower (r04g / r04i).
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t &h, uint32_t &l) {
  const __half2 hh = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
  h = *reinterpret_cast<const uint32_t *>(&hh);
  l = *reinterpret_cast<const uint32_t *>(&ll);
}

// fp16 split of one landed stage (TMA-mode converters of the fp16 variant).  The activations land PIXEL-MAJOR: one TMA box
// per 16-channel group, 64 bytes per pixel, SWIZZLE_64B (the 16-byte unit u of pixel `off` sits at unit u ^ ((off >> 1) & 3)
// of its 64-byte row: zones are 512-byte aligned) - a box row of 64 bytes instead of the 16 bytes of the 4-channel planes,
// which the TMA unit serves at ~4 cycles per row whatever its length (the 16-byte rows bound controller layers 1 / 3 / 5 at
// ~4 bytes per cycle and SM).  The two fp32 quads of an 8-channel unit become 8 halves: hi = fp16(x) goes to plane 2k of the
// operand region, lo' = fp16((x - hi) * 2^11) to plane 2k + 1 (the scale keeps lo' out of the fp16 subnormals; its products
// are accumulated in their own TMEM columns and scaled back by the epilogue).  Slot-fastest thread order: the swizzle makes
// the reads of 8 consecutive pixels conflict-free, the writes are consecutive 16-byte units.
__device__ __forceinline__ void convert_stage_f16(uint4 *hi4, const uint4 *zone4, int zone_units, int planes, int box_slots,
                                                  int slots_alloc, int up, int TWP, int RW, int y0, int x0, int cx, int cy,
                                                  int ptid) {
  constexpr int U = 2;
  for (int k = 0; k < planes / 2; ++k) {  // 8-channel units of the chunk (no division in the slot loop)
    const uint4 *zone = zone4 + (k >> 1) * zone_units;
    uint4 *hi_k = hi4 + 2 * k * slots_alloc;
    const int h2 = (k & 1) * 2;
    for (int base = ptid; base < box_slots; base += kProdThreads * U) {
      uint4 va[U], vb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int slot = base + u * kProdThreads;
        va[u] = vb[u] = make_uint4(0u, 0u, 0u, 0u);
        if (slot < box_slots) {
          int off = slot;  // up == 1: the box is the tile with its halo, RW == TWP
          if (up != 1) {
            const int r = slot / TWP, col = slot - r * TWP;
            const int vy = y0 - 2 + r, vx = x0 - 2 + col;  // zero-inserted (virtual) pixel
            off = ((vy | vx) & 1) == 0 ? ((vy >> 1) - cy) * RW + ((vx >> 1) - cx) : -1;
          }
          if (off >= 0) {
            const uint4 *row = zone + off * 4;
            const int sw = (off >> 1) & 3;
            va[u] = row[h2 ^ sw];
            vb[u] = row[(h2 + 1) ^ sw];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int slot = base + u * kProdThreads;
        if (slot >= box_slots) continue;
        uint32_t h[4], l[4];
        split_f16x2(__uint_as_float(va[u].x), __uint_as_float(va[u].y), h[0], l[0]);
        split_f16x2(__uint_as_float(va[u].z), __uint_as_float(va[u].w), h[1], l[1]);
        split_f16x2(__uint_as_float(vb[u].x), __uint_as_float(vb[u].y), h[2], l[2]);
        split_f16x2(__uint_as_float(vb[u].z), __uint_as_float(vb[u].w), h[3], l[3]);
        hi_k[slot] = make_uint4(h[0], h[1], h[2], h[3]);
        hi_k[slot + slots_alloc] = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

// One conv layer, run by a whole CTA (all roles).  `tm1` / `tm2` point at tensor maps in kernel-parameter space (single
// launch) or in global memory (chain).  CHAIN: the layer is one of several run back to back by a persistent grid: TMEM
// is allocated by the caller (`tmem_base_in`), the mbarriers of the previous layer are invalidated before they are
// initialised again, `first` marks the layer that performs the programmatic-dependent-launch handshake.
// RS: the row-stacked-taps variant (opt-in, see make_plan) is a separate instantiation, so that the default kernel
// carries none of its registers / shared memory.
template <bool RS, bool CHAIN, bool F16 = false, bool FOLD = false>
__device__ __forceinline__ void conv_layer(const UmmaConvParams &p, const CUtensorMap &tm1, const CUtensorMap &tm2,
                                           unsigned char *smem_dyn, uint32_t tmem_base_in, bool first,
                                           const unsigned int *chain_counter = nullptr,
                                           unsigned int chain_target = 0u) {
  if (threadIdx.x == 0) RA_DBG(0);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sc_s[256], sh_s[256];  // folded-BN scale / shift of this CTA's channels
  __shared__ float xw_s[RS ? 2 * 4 * 48 : 1];  // rowstack epilogue: rows the next warp hands to lanes 30 / 31
  __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_raw[kMaxStages], bar_tfull[2],
      bar_tempty[2];
  // TMA destinations want 128-byte alignment: round the dynamic window up (the launcher adds the slack)
  // (fp16 variant: 1024 bytes - its swizzled landing zones are addressed by absolute shared-memory address bits)
  constexpr uint32_t kAlign = F16 ? 1024u : 128u;
  unsigned char *smem_raw = smem_dyn + ((kAlign - (smem_u32(smem_dyn) & (kAlign - 1u))) & (kAlign - 1u));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler, too
  const int planes = p.KC / 4;
  const uint32_t plane_bytes = (uint32_t)p.slots_alloc * 16u;
  const int in_floats = planes * p.slots_alloc * 4;  // one of hi / lo
  // offset of the filter slice inside a stage: behind the hi and lo activation planes; the fp16 split keeps hi and lo' in
  // alternating planes of ONE such region (8 halves take the 16 bytes of 4 floats)
  const uint32_t w_off = (F16 ? 1u : 2u) * (uint32_t)in_floats * 4u;
  // filter image of one chunk: [9 taps][planes][2 NPc rows][16 bytes]; fp16 mode: a plane is 8 channels, not 4
  const int w_chunk_floats = 9 * (F16 ? planes / 2 : planes) * 2 * p.NPc * 4;
  unsigned char *stage_base = smem_raw + p.w_res_bytes;  // [resident filter image][stages][pool tile]
  float *pool_s = reinterpret_cast<float *>(stage_base + (size_t)p.stages * p.stage_bytes);
  const int ns = blockIdx.x % p.n_split;
  const int tile0 = blockIdx.x / p.n_split, tile_step = p.grid / p.n_split;
  const int n_tiles = p.n_items / p.n_split;

  // Barriers are initialised by the TMA warp, which (TMA mode) fires the first `stages` loads right away - before
  // the block-wide setup barrier - so the activation fetch overlaps TMEM allocation and the filter load.
  int tma_prologue = 0;  // (tile, chunk) items already requested by the early issue (TMA warp only)
  if (warp == kTmaWarp) {
    const bool leader = elect_one();
    if (leader) {
      if (CHAIN && !first) {  // the previous layer's barriers: every wait on them has completed (CTA-wide sync)
        for (int s = 0; s < kMaxStages; ++s) {
          mbar_inval(smem_u32(&bar_full[s]));
          mbar_inval(smem_u32(&bar_empty[s]));
          mbar_inval(smem_u32(&bar_raw[s]));
        }
        for (int s = 0; s < 2; ++s) {
          mbar_inval(smem_u32(&bar_tfull[s]));
          mbar_inval(smem_u32(&bar_tempty[s]));
        }
      }
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(smem_u32(&bar_full[s]), kProdThreads);
        mbar_init(smem_u32(&bar_empty[s]), kMmaWarps);
        mbar_init(smem_u32(&bar_raw[s]), 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(smem_u32(&bar_tfull[s]), kMmaWarps);
        mbar_init(smem_u32(&bar_tempty[s]), RS ? kEpiThreads : kEpiAll);  // (rowstack: group A only)
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (p.tma) {
      if (leader) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm1)) : "memory");
        if (p.C2 > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm2)) : "memory");
      }
      // the activations are the previous kernel's output: wait for it (programmatic dependent launch) here, and
      // only here - every other access to dependent memory is ordered after these loads
      if (p.pdl && first) asm volatile("griddepcontrol.wait;" ::: "memory");
      if (CHAIN && !first) {
        // split grid barrier: every CTA ARRIVED when it finished the previous layer (chain kernel loop); only this
        // warp WAITS, and only now - the barrier set-up above, the TMEM / filter / scale-shift loads of the other
        // warps overlap the tail of the previous layer on slower SMs
        // (every lane polls and the loop exit is warp-uniform: the TMA instructions below want a converged warp)
        unsigned int seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(chain_counter) : "memory");
        } while (__any_sync(0xffffffffu, seen < chain_target));
        asm volatile("fence.proxy.async;" ::: "memory");  // other CTAs' generic-proxy stores -> our TMA reads
        __syncwarp();
      }
      const uint32_t w_bytes = p.w_resident ? 0u : (uint32_t)w_chunk_floats * 4u;
      const uint32_t tx_bytes = (uint32_t)planes * (uint32_t)(p.RW * p.RH) * 16u + w_bytes;
      // one box per 4-channel plane (16-byte rows), or - fp16 variant - per 16-channel group (64-byte rows, landing zone)
      const uint32_t dst_plane = (F16 || p.up == 2) ? (uint32_t)p.raw_plane_bytes : plane_bytes;
      constexpr int box_c = F16 ? 16 : 4;
      const int n_boxes = F16 ? p.KC / 16 : planes;
      int g = 0;
      for (int tile = tile0; tile < n_tiles && g < p.stages; tile += tile_step) {
        const Item it = decode_tile(p, tile);
        const int cx = p.up == 2 ? (it.x0 >> 1) - 1 : it.x0 - 1;
        const int cy = p.up == 2 ? ((it.y0 - 1) >> 1) : it.y0 - 1;
        for (int ch = 0; ch < p.n_chunks && g < p.stages; ++ch, ++g) {
          const uint32_t st = smem_u32(stage_base + (size_t)g * p.stage_bytes);  // first pass: stage g, all empty
          const uint32_t bar = smem_u32(&bar_raw[g]);
          if (leader) mbar_arrive_expect_tx(bar, tx_bytes);
          const uint32_t dst0 = (F16 || p.up == 2) ? st + (uint32_t)p.raw_off : st;
          for (int pl = 0; pl < n_boxes; ++pl) {
            const int cc = ch * p.KC + box_c * pl;
            const uint32_t dst = dst0 + (uint32_t)pl * dst_plane;
            if (cc < p.C1 || p.C2 == 0) {
              if (leader) tma_load_4d(dst, &tm1, cc, cx, cy, it.b, bar);
            } else {
              if (leader) tma_load_4d(dst, &tm2, cc - p.C1, cx, cy, it.b, bar);
            }
          }
          if (!p.w_resident) {
            const float *wsrc = p.wpack + ((size_t)ns * p.n_chunks + ch) * w_chunk_floats;
            if (leader) bulk_load(st + w_off, wsrc, w_bytes, bar);
          }
          __syncwarp();
          if (g == 0 && leader) RA_DBG(2);  // first stage requested
        }
      }
      tma_prologue = g;
    }
  }
  if (!CHAIN && warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (p.w_resident) {
    // the whole filter image of this CTA's channel split: loaded once, reused by every tile
    const float4 *src = reinterpret_cast<const float4 *>(p.wpack + (size_t)ns * p.n_chunks * w_chunk_floats);
    float4 *dstw = reinterpret_cast<float4 *>(smem_raw);
    const int n4 = p.n_chunks * w_chunk_floats / 4;
    for (int base = tid; base < n4; base += kThreads * kStageUnroll) {
      float4 t[kStageUnroll];
#pragma unroll
      for (int u = 0; u < kStageUnroll; ++u) {
        const int idx = base + u * kThreads;
        t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < n4) t[u] = __ldg(src + idx);
      }
#pragma unroll
      for (int u = 0; u < kStageUnroll; ++u) {
        const int idx = base + u * kThreads;
        if (idx < n4) dstw[idx] = t[u];
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  // fold (fp16 variant): B = 2^-11 I in the K-major operand layout [NPc / 4 planes][NPc rows][4 floats], behind the pool tile
  float *fold_s = pool_s + (p.pool == 2 ? p.n_mt * 64 * kPoolLd : 0);
  if (F16 && FOLD) {
    for (int i = tid; i < (p.NPc / 4) * p.NPc; i += kThreads) {
      const int pl = i / p.NPc, row = i - pl * p.NPc;
      const int hit = row - 4 * pl;  // the k index (0..3) of this plane that equals the row, if any
      reinterpret_cast<float4 *>(fold_s)[i] = make_float4(hit == 0 ? 4.8828125e-4f : 0.f, hit == 1 ? 4.8828125e-4f : 0.f,
                                                          hit == 2 ? 4.8828125e-4f : 0.f, hit == 3 ? 4.8828125e-4f : 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = CHAIN ? tmem_base_in : __shfl_sync(0xffffffffu, tmem_base_s, 0);  // uniform register
  const int slots_in = (p.TH + 2) * p.TWP + 2;  // slots that carry real (or zero-padding) data
  if (threadIdx.x == 0) RA_DBG(1);  // setup done (barriers, TMEM, resident filters)
  // programmatic dependent launch: the next kernel of the stream may start its own setup on SMs this grid has
  // left; it still waits (griddepcontrol.wait) for this grid to complete before it reads our output
  if (p.pdl && first) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == kTmaWarp) {
    // =============================== TMA issuer ===============================
    // (warp-uniform control flow, one elected lane issues: TMA operands live in uniform registers, see the MMA warps)
    if (p.tma) {
      const bool leader = elect_one();
      const uint32_t w_bytes = p.w_resident ? 0u : (uint32_t)w_chunk_floats * 4u;
      const uint32_t tx_bytes = (uint32_t)planes * (uint32_t)(p.RW * p.RH) * 16u + w_bytes;
      // one box per 4-channel plane (16-byte rows), or - fp16 variant - per 16-channel group (64-byte rows, landing zone)
      const uint32_t dst_plane = (F16 || p.up == 2) ? (uint32_t)p.raw_plane_bytes : plane_bytes;
      constexpr int box_c = F16 ? 16 : 4;
      const int n_boxes = F16 ? p.KC / 16 : planes;
      int g = 0;
      long long w_tma = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step) {
        const Item it = decode_tile(p, tile);
        // first input pixel of the box (negative / past-the-end coordinates are zero-filled = SAME padding)
        const int cx = p.up == 2 ? (it.x0 >> 1) - 1 : it.x0 - 1;
        const int cy = p.up == 2 ? ((it.y0 - 1) >> 1) : it.y0 - 1;
        for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
          if (g < tma_prologue) continue;  // requested before the setup barrier
          const int s = g % p.stages;
          RA_WAIT(smem_u32(&bar_empty[s]), (uint32_t)(((g / p.stages) & 1) ^ 1), w_tma);
          const uint32_t st = smem_u32(stage_base + (size_t)s * p.stage_bytes);
          const uint32_t bar = smem_u32(&bar_raw[s]);
          if (leader) mbar_arrive_expect_tx(bar, tx_bytes);
          const uint32_t dst0 = (F16 || p.up == 2) ? st + (uint32_t)p.raw_off : st;
          for (int pl = 0; pl < n_boxes; ++pl) {
            const int cc = ch * p.KC + box_c * pl;
            const uint32_t dst = dst0 + (uint32_t)pl * dst_plane;
            // channels past Cin (padding planes) fall outside the map and come back as zeros
            if (cc < p.C1 || p.C2 == 0) {
              if (leader) tma_load_4d(dst, &tm1, cc, cx, cy, it.b, bar);
            } else {
              if (leader) tma_load_4d(dst, &tm2, cc - p.C1, cx, cy, it.b, bar);
            }
          }
          if (!p.w_resident) {
            const float *wsrc = p.wpack + ((size_t)ns * p.n_chunks + ch) * w_chunk_floats;
            if (leader) bulk_load(st + w_off, wsrc, w_bytes, bar);
          }
          __syncwarp();
        }
      }
      if (leader) RA_DBG_PUT(8, w_tma);
      if (CHAIN) {
        // Drain: the MMA warps release every stage with an asynchronous tcgen05.commit -> mbarrier arrive.  Nobody waits
        // for the LAST release of a stage inside a layer, and the next layer re-initialises these barriers right after
        // the CTA-wide sync: an arrival still in flight would land on the fresh barrier.  Wait for all of them here.
        const int first_pending = g > p.stages ? g - p.stages : 0;
        for (int gg = first_pending; gg < g; ++gg)
          mbar_wait(smem_u32(&bar_empty[gg % p.stages]), (uint32_t)((gg / p.stages) & 1));
        __syncwarp();
      }
    }
  } else if (warp >= kProdWarp0 && warp < kEpiWarpB0 && p.tma) {
    // =============================== converters (TMA mode) ===============================
    const int ptid = tid - (kEpiThreads + 32 * kMmaWarps);
    const int box_slots = (p.TH + 2) * p.TWP;  // slots the taps of real output pixels can touch
    const int n_conv = planes * box_slots;
    int g = 0;
    long long w_conv = 0, b_conv = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      const Item it = decode_tile(p, tile);
      const int cx = (it.x0 >> 1) - 1, cy = (it.y0 - 1) >> 1;  // up == 2: origin of the landed box
      for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
        const int s = g % p.stages;
        RA_WAIT(smem_u32(&bar_raw[s]), (uint32_t)((g / p.stages) & 1), w_conv);
        const long long tb_ = (F16 && !CHAIN && p.dbg != nullptr) ? clock64() : 0;
        unsigned char *st = stage_base + (size_t)s * p.stage_bytes;
        float4 *hi4 = reinterpret_cast<float4 *>(st);
        float4 *lo4 = hi4 + in_floats / 4;
        const float4 *raw4 = reinterpret_cast<const float4 *>(st + p.raw_off);
        const int raw_plane4 = p.raw_plane_bytes >> 4;
        if constexpr (F16) {
          convert_stage_f16(reinterpret_cast<uint4 *>(st), reinterpret_cast<const uint4 *>(st + p.raw_off), raw_plane4, planes,
                            box_slots, p.slots_alloc, p.up, p.TWP, p.RW, it.y0, it.x0, cx, cy, ptid);
        } else
        for (int base = ptid; base < n_conv; base += kProdThreads * kStageUnroll) {
          float4 v[kStageUnroll];
          int dst[kStageUnroll];
#pragma unroll
          for (int u = 0; u < kStageUnroll; ++u) {
            const int idx = base + u * kProdThreads;
            dst[u] = -1;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < n_conv) {
              const int c4 = idx / box_slots, slot = idx - c4 * box_slots;  // slot fastest: conflict-free
              dst[u] = c4 * p.slots_alloc + slot;
              if (p.up == 1) {
                v[u] = hi4[dst[u]];  // the box landed in place
              } else {
                const int r = slot / p.TWP, col = slot - r * p.TWP;
                const int vy = it.y0 - 2 + r, vx = it.x0 - 2 + col;  // zero-inserted (virtual) pixel
                if (((vy | vx) & 1) == 0) v[u] = raw4[c4 * raw_plane4 + ((vy >> 1) - cy) * p.RW + ((vx >> 1) - cx)];
              }
            }
          }
#pragma unroll
          for (int u = 0; u < kStageUnroll; ++u) {
            if (dst[u] < 0) continue;
            const float4 h = make_float4(tf32_hi(v[u].x), tf32_hi(v[u].y), tf32_hi(v[u].z), tf32_hi(v[u].w));
            const float4 l = make_float4(v[u].x - h.x, v[u].y - h.y, v[u].z - h.z, v[u].w - h.w);
            hi4[dst[u]] = h;
            lo4[dst[u]] = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (tensor core)
        mbar_arrive(smem_u32(&bar_full[s]));
        if (F16 && !CHAIN && p.dbg != nullptr) b_conv += clock64() - tb_;
      }
    }
    if (ptid == 0) {
      RA_DBG(3);  // converters done
      RA_DBG_PUT(9, w_conv);
      RA_DBG_PUT(14, b_conv);
    }
  } else if (warp >= kProdWarp0 && warp < kEpiWarpB0) {
    // =============================== producers (plain-load mode) ===============================
    const int ptid = tid - (kEpiThreads + 32 * kMmaWarps);
    if (p.pdl && first) asm volatile("griddepcontrol.wait;" ::: "memory");  // the activations are the previous kernel's output
    int g = 0;  // running (tile, chunk) counter
    for (int tile = tile0; tile < n_tiles; tile += tile_step) {
      const Item it = decode_tile(p, tile);
      for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
        const int s = g % p.stages;
        mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((g / p.stages) & 1) ^ 1));
        float *in_hi = reinterpret_cast<float *>(stage_base + (size_t)s * p.stage_bytes);
        float *in_lo = in_hi + in_floats;
        float *w_s = in_lo + in_floats;
        const int c0 = ch * p.KC;
        // ---- input chunk: slot i <-> virtual pixel (y0 - up + i / TWP, x0 - up + i % TWP)
        const int n_in = p.slots_alloc * planes;
        for (int base = ptid; base < n_in; base += kProdThreads * kStageUnroll) {
          float4 v[kStageUnroll];
          int dst[kStageUnroll];
#pragma unroll
          for (int u = 0; u < kStageUnroll; ++u) {
            const int idx = base + u * kProdThreads;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            dst[u] = -1;
            if (idx < n_in) {
              const int c4 = idx % planes;  // plane fastest: the threads of one pixel read one contiguous run
              const int slot = idx / planes;
              dst[u] = (c4 * p.slots_alloc + slot) * 4;
              if (slot < slots_in) {
                const int r = slot / p.TWP, col = slot - r * p.TWP;
                const int vy = it.y0 - p.up + r, vx = it.x0 - p.up + col;
                bool on = vy >= 0 && vx >= 0 && vy < p.Hout && vx < p.Wout;
                int iy = vy, ix = vx;
                if (p.up == 2) {
                  on = on && (((vy | vx) & 1) == 0);
                  iy >>= 1;
                  ix >>= 1;
                }
                const int cc = c0 + c4 * 4;
                if (on && cc < p.Cin) {
                  const size_t pix = ((size_t)it.b * p.Hin + iy) * p.Win + ix;
                  if (p.vec4) {
                    v[u] = (cc < p.C1) ? __ldg(reinterpret_cast<const float4 *>(p.x1 + pix * p.C1 + cc))
                                       : __ldg(reinterpret_cast<const float4 *>(p.x2 + pix * p.C2 + (cc - p.C1)));
                  } else {
                    float t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const int c = cc + j;
                      t[j] = 0.f;
                      if (c < p.C1)
                        t[j] = __ldg(p.x1 + pix * p.C1 + c);
                      else if (c < p.Cin)
                        t[j] = __ldg(p.x2 + pix * p.C2 + (c - p.C1));
                    }
                    v[u] = make_float4(t[0], t[1], t[2], t[3]);
                  }
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < kStageUnroll; ++u) {
            if (dst[u] < 0) continue;
            const float4 h = make_float4(tf32_hi(v[u].x), tf32_hi(v[u].y), tf32_hi(v[u].z), tf32_hi(v[u].w));
            const float4 l = make_float4(v[u].x - h.x, v[u].y - h.y, v[u].z - h.z, v[u].w - h.w);
            *reinterpret_cast<float4 *>(in_hi + dst[u]) = h;
            *reinterpret_cast<float4 *>(in_lo + dst[u]) = l;
          }
        }
        // ---- filter slice of this (n-split, chunk): already hi/lo split, plane layout
        if (!p.w_resident) {
          const float4 *src =
              reinterpret_cast<const float4 *>(p.wpack + ((size_t)ns * p.n_chunks + ch) * w_chunk_floats);
          float4 *dstw = reinterpret_cast<float4 *>(w_s);
          const int n4 = w_chunk_floats / 4;
          for (int base = ptid; base < n4; base += kProdThreads * kStageUnroll) {
            float4 t[kStageUnroll];
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
              const int idx = base + u * kProdThreads;
              t[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (idx < n4) t[u] = __ldg(src + idx);
            }
#pragma unroll
            for (int u = 0; u < kStageUnroll; ++u) {
              const int idx = base + u * kProdThreads;
              if (idx < n4) dstw[idx] = t[u];
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (tensor core)
        mbar_arrive(smem_u32(&bar_full[s]));
        if (g == 0 && ptid == 0) RA_DBG(2);  // first stage staged
      }
    }
    if (ptid == 0) RA_DBG(3);  // producers done
  } else if (warp >= kMmaWarp0 && warp < kProdWarp0) {
    // =============================== MMA issuers ===============================
    // tcgen05.mma takes its operands from UNIFORM registers.  Issued from an `if (lane == 0)` region the compiler
    // cannot prove them warp-uniform and wraps every MMA in an ELECT / R2UR.BROADCAST loop (~185 cycles per
    // instruction); generic warp-uniform loops are no better (dependent uniform-datapath instructions cost ~5 cycles
    // each, ~450 cycles of loop control per step).  What works: branch on elect.sync - the compiler then knows ONE
    // thread is active and uses plain R2UR moves (~50-75 cycles per instruction, measured) - and let kMmaWarps warps
    // issue in parallel.  Warp mw owns m-tiles mw and mw + 4 (their own TMEM accumulator columns).
    {
      const bool leader = elect_one();
      const int mw = warp - kMmaWarp0;
      const int cols_mt = RS ? 6 * p.NPc : (p.merged ? 2 * p.NPc : p.NPc);
      IssueCtx c;
      // rowstack: idesc_n = the A_lo instruction (3 NPc wide), idesc_2n = the A_hi instruction (6 NPc wide)
      const uint32_t n_lo = RS ? 3u * (uint32_t)p.NPc : (uint32_t)p.NPc;
      const uint32_t n_hi = RS ? 6u * (uint32_t)p.NPc : 2u * (uint32_t)p.NPc;
      constexpr bool f16 = F16;
      const uint32_t fmt = f16 ? 0u : 2u;  // a / b format: F16 = 0, TF32 = 2; c format F32
      c.idesc_n = (1u << 4) | (fmt << 7) | (fmt << 10) | ((n_lo >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      c.idesc_2n = (1u << 4) | (fmt << 7) | (fmt << 10) | ((n_hi >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      c.idesc_lo = p.four_term ? c.idesc_2n : c.idesc_n;
      // column offset of the A_lo B_hi correction inside an m-tile's accumulator (0: same half as the main sum)
      c.lo_col = RS ? 3u * (uint32_t)p.NPc
                            : ((p.merged && !p.four_term && getenv_split_corr(p)) ? (uint32_t)p.NPc : 0u);
      // bytes between channel planes of the filter image
      const uint32_t w_plane = (uint32_t)((RS ? 6 : 2) * p.NPc) * 16u;
      c.TWP = (uint32_t)p.TWP;
      // A start-address step of one K step: two 4-channel planes (tf32, K = 8) or four (fp16, K = 16: two 8-channel units
      // that sit in planes 2k / 2k + 2)
      c.a_k8 = (f16 ? 4u : 2u) * (plane_bytes >> 4);
      c.b_k8 = 2u * (w_plane >> 4);                     // same for the filter image
      c.b_tap = (uint32_t)(f16 ? planes / 2 : planes) * (w_plane >> 4);  // filter image step of one tap (rowstack: of one filter row)
      c.b_lo = (uint32_t)p.NPc;                         // rows NPc..2NPc-1 of a plane hold the lo part
      c.cols_mt = (uint32_t)cols_mt;
      c.wrap = (uint32_t)(p.ksplit * cols_mt);          // TMEM columns of one m-tile (all its partial accumulators)
      const bool has0 = mw < p.n_mt, has1 = mw + kMmaWarps < p.n_mt;
      const bool merged = p.merged != 0;
      const uint32_t a_mt0 = (uint32_t)(mw * p.mt_stride), a_mt1 = (uint32_t)((mw + kMmaWarps) * p.mt_stride);
      const int k8n = f16 ? p.KC / 16 : p.KC / 8;
      int g = 0, t = 0;
      long long w_full = 0, w_tempty = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step, ++t) {
        const int buf = p.nbuf == 2 ? (t & 1) : 0;
        const uint32_t use = p.nbuf == 2 ? (uint32_t)(t >> 1) : (uint32_t)t;
        RA_WAIT(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u, w_tempty);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
        c.d0 = acc + (uint32_t)mw * c.wrap;
        c.d1 = acc + (uint32_t)(mw + kMmaWarps) * c.wrap;
        for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
          const int s = g % p.stages;
          RA_WAIT(smem_u32(&bar_full[s]), (uint32_t)((g / p.stages) & 1), w_full);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (g == 0 && mw == 0 && leader) RA_DBG(4);  // first MMA can issue
          // Descriptors differ only in the start-address field (bits 0-13, 16-byte units): a tap, a k8 step, an
          // m-tile are small additive constants.
          const uint32_t a_hi = smem_u32(stage_base + (size_t)s * p.stage_bytes);
          // fp16: the K-adjacent core matrices (8-channel units) are two fp32 planes apart; lo' sits one plane behind hi
          const uint64_t dA = make_desc(a_hi, f16 ? 2u * plane_bytes : plane_bytes, 128);
          const uint64_t lo_off = f16 ? (uint64_t)(plane_bytes >> 4) : (uint64_t)(((uint32_t)in_floats * 4u) >> 4);
          c.a0_hi = dA + (uint64_t)a_mt0;
          c.a1_hi = dA + (uint64_t)a_mt1;
          c.a0_lo = c.a0_hi + lo_off;
          c.a1_lo = c.a1_hi + lo_off;
          const uint32_t w_addr = p.w_resident ? smem_u32(smem_raw) + (uint32_t)ch * (uint32_t)w_chunk_floats * 4u
                                               : a_hi + w_off;
          c.b = make_desc(w_addr, w_plane, 128);
          c.init_steps = ch == 0 ? (uint32_t)p.ksplit : 0u;  // the first MMA into each accumulator overwrites
          if (has0 && leader) {
            if constexpr (RS) {
              if (k8n == 1) {
                if (has1) issue_chunk_rs<1, true>(c); else issue_chunk_rs<1, false>(c);
              } else if (k8n == 2) {
                if (has1) issue_chunk_rs<2, true>(c); else issue_chunk_rs<2, false>(c);
              } else {
                if (has1) issue_chunk_rs<4, true>(c); else issue_chunk_rs<4, false>(c);
              }
            } else if (k8n == 1)
              issue_chunk_k8<1, F16>(c, merged, has1);
            else if (k8n == 2)
              issue_chunk_k8<2, F16>(c, merged, has1);
            else
              issue_chunk_k8<4, F16>(c, merged, has1);
          }
          __syncwarp();
          if (leader) umma_commit(smem_u32(&bar_empty[s]));  // the stage may be refilled once these MMAs have read it
          __syncwarp();
        }
        if (F16 && FOLD && has0 && leader) {
          // main half += 2^-11 x correction half, on the tensor core (in order behind the tile's accumulations)
          const uint64_t idb = make_desc(smem_u32(fold_s), (uint32_t)p.NPc * 16u, 128);
          const uint32_t idesc_f = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)p.NPc >> 3) << 17) |
                                   ((uint32_t)(128 >> 4) << 24);
          for (int half = 0; half < (has1 ? 2 : 1); ++half) {
            const uint32_t dm = half ? c.d1 : c.d0;
            for (int ks = 0; ks < p.ksplit; ++ks) {
              const uint32_t dk = dm + (uint32_t)(ks * cols_mt);
              for (int j = 0; j < p.NPc / 8; ++j)
                umma_tf32_ts(dk, dk + (uint32_t)(p.NPc + 8 * j), idb + (uint64_t)(2 * j * p.NPc), idesc_f, 1u);
            }
          }
        }
        __syncwarp();
        if (leader) umma_commit(smem_u32(&bar_tfull[buf]));  // accumulators of this tile are complete
        __syncwarp();
      }
      if (mw == 0 && leader) {
        RA_DBG(5);  // all MMAs issued
        RA_DBG_PUT(10, w_full);
        RA_DBG_PUT(11, w_tempty);
      }
    }
  } else if (!(RS && warp >= kEpiWarpB0)) {
    // =============================== epilogue (warps 0-3 and 15-18) ===============================
    // eg: the group (it takes the m-tiles eg, eg + 2, ...; rowstack: group A alone), et: thread within the group = TMEM lane /
    // slot within an m-tile, e2: thread within the whole epilogue (loops shared by both groups)
    constexpr int kGroups = RS ? 1 : 2;
    constexpr int kEpiN = RS ? kEpiThreads : kEpiAll;
    const int eg = warp >= kEpiWarpB0 ? 1 : 0;
    const int wq = warp & 3;
    const int et = wq * 32 + lane;
    const int e2 = eg * kEpiThreads + et;
    const int cols_mt = RS ? 6 * p.NPc : (p.merged ? 2 * p.NPc : p.NPc);
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;  // warp w may touch TMEM lanes [32 (w % 4), +32)
    int xw_par = 0;  // rowstack: parity of the cross-warp exchange buffer
    const int slots_out = p.TH * p.TWP;
    const int chain_cols = p.ksplit * cols_mt;  // TMEM columns of one m-tile
    const int co_base = ns * p.NPc;
    // folded-BN scale / shift of this CTA's channels: once into shared memory (padded channels: 0)
    for (int i = e2; i < p.NPc; i += kEpiN) {
      const int co = co_base + i;
      sc_s[i] = co < p.Cout ? __ldg(p.scale + co) : 0.f;
      sh_s[i] = co < p.Cout ? __ldg(p.shift + co) : 0.f;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiN) : "memory");

    // v[16] <- columns cb..cb+15 of m-tile mt: sum of the merged halves and of the K-split partial accumulators;
    // the loads of one (hi half, lo half) pair are in flight together.  Then BN scale/shift (+ ReLU).
    // Software pipeline: TMEM reads are slow (64 bytes per cycle and SM, a round trip of ~1000 cycles per load pair while the
    // tensor core accumulates into the other buffer: profiles/r04h_timeline_f16.txt) and they were the larger half of the
    // epilogue; with one partial accumulator (ksplit == 1: every large layer) the loads of the NEXT call (nmt, ncb; nmt < 0:
    // none) are issued into r0 / r1 as soon as this call has combined them into v, and fly during BN / pooling / stores.
// (fold: the tensor core has already added the second half)
#define two_halves (p.merged && !(F16 && FOLD))
    uint32_t r0[16], r1[16];
    bool pre = false;  // r0 / r1 already hold the (in-flight) loads of the coming call
    // relu_now = false: the caller applies the ReLU itself (pooling layers: after the max, on a quarter of the values)
    auto load16 = [&](uint32_t acc, int mt, int cb, float *v, int nmt, int ncb, bool relu_now) {
      if constexpr (RS) {
        // g[j] = D'[lane, kx block, channel cb + j] (hi half + lo half + K-split partials); the output of this
        // lane's slot is g(kx=0)[lane] + g(kx=1)[lane + 1] + g(kx=2)[lane + 2].  Lanes 30 / 31 take the rows of the
        // next warp from shared memory (its lanes 0 / 1 publish them; double-buffered, one barrier per call).
        float *xw = xw_s + xw_par * (4 * 48);
        xw_par ^= 1;
        float g[16];
        auto load_group = [&](int kx) {
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = 0.f;
          for (int ks = 0; ks < p.ksplit; ++ks) {
            const uint32_t col = acc + (uint32_t)(mt * chain_cols + ks * cols_mt + kx * p.NPc + cb);
            tmem_ld16_issue(col, r0);
            tmem_ld16_issue(col + 3u * (uint32_t)p.NPc, r1);
            tmem_wait16(r0);
            tmem_wait16(r1);
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
          }
        };
        load_group(0);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = g[j];
        load_group(1);
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xw[warp * 48 + j] = g[j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = __shfl_down_sync(0xffffffffu, g[j], 1);
          if (lane != 31) v[j] += t;
        }
        load_group(2);
        if (lane < 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xw[warp * 48 + 16 + lane * 16 + j] = g[j];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = __shfl_down_sync(0xffffffffu, g[j], 2);
          if (lane < 30) v[j] += t;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        const float *nx = xw + ((warp + 1) & 3) * 48;
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += nx[j] + nx[32 + j];
        } else if (lane == 30) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += nx[16 + j];
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 sc = *reinterpret_cast<const float4 *>(sc_s + cb + j);
          const float4 sh = *reinterpret_cast<const float4 *>(sh_s + cb + j);
          v[j] = fmaf(v[j], sc.x, sh.x);
          v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
          v[j + 2] = fmaf(v[j + 2], sc.z, sh.z);
          v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
        }
        if (p.relu && relu_now) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        return;
      }
      const uint32_t base = acc + (uint32_t)(mt * chain_cols + cb);
      if (!pre) {
        tmem_ld16_issue(base, r0);
        if (two_halves) tmem_ld16_issue(base + (uint32_t)p.NPc, r1);
      }
      tmem_wait16(r0);
      // fp16 mode: the second half holds the 2^11-scaled corrections (fmaf with 1.0 is the plain sum, bit for bit)
      constexpr float lo_scale = F16 ? 4.8828125e-4f : 1.0f;
      if (two_halves) {
        tmem_wait16(r1);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r1[j]), lo_scale, __uint_as_float(r0[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r0[j]);
      }
      pre = p.ksplit == 1 && nmt >= 0;
      if (pre) {
        const uint32_t nb = acc + (uint32_t)(nmt * chain_cols + ncb);
        tmem_ld16_issue(nb, r0);
        if (two_halves) tmem_ld16_issue(nb + (uint32_t)p.NPc, r1);
      }
      for (int ks = 1; ks < p.ksplit; ++ks) {  // partial accumulators of the K split
        tmem_ld16_issue(base + (uint32_t)(ks * cols_mt), r0);
        if (two_halves) tmem_ld16_issue(base + (uint32_t)(ks * cols_mt + p.NPc), r1);
        tmem_wait16(r0);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(r0[j]);
        if (two_halves) {
          tmem_wait16(r1);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(r1[j]), lo_scale, v[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 sc = *reinterpret_cast<const float4 *>(sc_s + cb + j);
        const float4 sh = *reinterpret_cast<const float4 *>(sh_s + cb + j);
        v[j] = fmaf(v[j], sc.x, sh.x);
        v[j + 1] = fmaf(v[j + 1], sc.y, sh.y);
        v[j + 2] = fmaf(v[j + 2], sc.z, sh.z);
        v[j + 3] = fmaf(v[j + 3], sc.w, sh.w);
      }
      if (p.relu && relu_now) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      }
    };

#undef two_halves
    int t = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step, ++t) {
      const Item it = decode_tile(p, tile);
      const int buf = p.nbuf == 2 ? (t & 1) : 0;
      const uint32_t use = p.nbuf == 2 ? (uint32_t)(t >> 1) : (uint32_t)t;
      mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (t == 0 && e2 == 0) RA_DBG(6);  // first accumulator complete
      const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols) + lane_sel;
      const int cb_end = (p.Cout - co_base) < p.NPc ? (p.Cout - co_base) : p.NPc;  // real channels of this split
      if (p.pool == 2) {
        for (int cb = 0; cb < cb_end; cb += 16) {
          for (int mt = eg; mt < p.n_mt; mt += kGroups) {
            const int s = mt * p.mt_stride + et;
            float v[16];
            {
              const bool more = mt + kGroups < p.n_mt;  // then the same chunk of my next m-tile, else the next chunk
              const bool nxt = more || cb + 16 < cb_end;
              load16(acc, mt, cb, v, nxt ? (more ? mt + kGroups : eg) : -1, more ? cb : cb + 16, false);
            }
            // horizontal max with the next slot (same image row: TW, x0 are even so pairs do not straddle)
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], __shfl_down_sync(0xffffffffu, v[j], 1));
            if ((lane & 1) == 0 && s < slots_out && et < p.mt_stride) {
              float *dst = pool_s + (size_t)(s >> 1) * kPoolLd;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiN) : "memory");
          // vertical max + store of this 16-channel chunk: pooled pixel (py, px) <- staged half-rows of
          // slots (2py)*TWP+2px and +TWP
          const int ph = p.TH / 2, pw = p.TW / 2;
          const int Ho = p.Hout / 2, Wo = p.Wout / 2;
          const int pw_shift = (pw & (pw - 1)) == 0 ? 31 - __clz(pw) : -1;  // tile widths are mostly powers of two
          const float floor_v = p.relu ? 0.f : -INFINITY;  // ReLU commutes with the max
          for (int idx = e2; idx < ph * pw * 4; idx += kEpiN) {
            const int c4 = idx & 3;
            const int pix = idx >> 2;
            const int py = pw_shift >= 0 ? pix >> pw_shift : pix / pw, px = pix - py * pw;
            const int gy = (it.y0 >> 1) + py, gx = (it.x0 >> 1) + px;
            const int co = co_base + cb + c4 * 4;
            if (gy >= Ho || gx >= Wo || co >= p.Cout) continue;
            const int s0 = (2 * py) * p.TWP + 2 * px;
            const float4 va = *reinterpret_cast<const float4 *>(pool_s + (size_t)(s0 >> 1) * kPoolLd + c4 * 4);
            const float4 vc =
                *reinterpret_cast<const float4 *>(pool_s + (size_t)((s0 + p.TWP) >> 1) * kPoolLd + c4 * 4);
            const float4 m = make_float4(fmaxf(fmaxf(va.x, vc.x), floor_v), fmaxf(fmaxf(va.y, vc.y), floor_v),
                                         fmaxf(fmaxf(va.z, vc.z), floor_v), fmaxf(fmaxf(va.w, vc.w), floor_v));
            float *dst = p.y + (((size_t)it.b * Ho + gy) * Wo + gx) * p.Cout + co;
            if ((p.Cout & 3) == 0) {
              *reinterpret_cast<float4 *>(dst) = m;
            } else {
              const float tt[4] = {m.x, m.y, m.z, m.w};
              for (int j = 0; j < 4; ++j)
                if (co + j < p.Cout) dst[j] = tt[j];
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiN) : "memory");
        }
      } else {
        for (int mt = eg; mt < p.n_mt; mt += kGroups) {
          const int s = mt * p.mt_stride + et;
          const int oy_l = s / p.TWP, ox_l = s - oy_l * p.TWP;
          const int oy = it.y0 + oy_l, ox = it.x0 + ox_l;
          const bool valid = s < slots_out && et < p.mt_stride && ox_l < p.TW && oy < p.Hout && ox < p.Wout;
          float *dst_px = p.y + (((size_t)it.b * p.Hout + oy) * p.Wout + ox) * p.Cout + co_base;
          for (int cb = 0; cb < cb_end; cb += 16) {
            float v[16];
            {
              const bool more = cb + 16 < cb_end;  // then the next chunk of this m-tile, else my next m-tile
              const bool nxt = more || mt + kGroups < p.n_mt;
              load16(acc, mt, cb, v, nxt ? (more ? mt : mt + kGroups) : -1, more ? cb + 16 : 0, true);
            }  // (warp-collective: every lane takes part, valid or not)
            if (!valid) continue;
            float *dst = dst_px + cb;
            if ((p.Cout & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (co_base + cb + j < p.Cout)
                  *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (co_base + cb + j < p.Cout) dst[j] = v[j];
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bar_tempty[buf]));
    }
  }

  if (CHAIN) __threadfence();  // this layer's outputs are read by other CTAs after the grid barrier
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) RA_DBG(7);  // all roles done
  if (!CHAIN && warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// F16: the fp16 hi / lo operand split (RA_UMMA_F16 plans) - like RS a separate instantiation.
// FOLD: F16 + the correction half folded on the tensor core (mode + 4): one more instantiation.
template <bool RS, bool F16 = false, bool FOLD = false>
__global__ void __launch_bounds__(kThreads, 1)
    conv3x3_umma_kernel(const __grid_constant__ UmmaConvParams p, const __grid_constant__ CUtensorMap tm1,
                        const __grid_constant__ CUtensorMap tm2) {
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  conv_layer<RS, false, F16, FOLD>(p, tm1, tm2, smem_dyn, 0u, true);
}

// A CHAIN of conv layers in one launch: the patch network of a decode step (6 + 7 layers) or controller layers 1-7.
// Every layer of those stacks is a 15-30 us launch for 1-2 us of tensor work (pipeline fill / drain, CTA start-up, TMEM
// allocation); here a persistent grid of one CTA per SM runs the layers back to back with a grid-wide barrier between
// them: TMEM is allocated once, the next layer's barrier set-up and filter load start the moment the barrier opens.
// Layer l is served by the first layers[l].p.grid CTAs (its own tile plan); the others only take part in the barrier.
// All CTAs are co-resident (grid <= number of SMs, one CTA per SM), so the spin barrier cannot deadlock: kernels of
// other graph branches that hold an SM finish on their own.
struct ChainLayer {
  UmmaConvParams p;
  CUtensorMap tm1, tm2;
};
constexpr int kChainMax = 16;
// The whole chain travels as ONE kernel argument (kernel parameters may be up to 32 KB since CUDA 12.1): the layer
// parameters are then read through the constant bank and the tensor maps live in parameter space, exactly like in the
// single-layer kernel - descriptors in global memory made every field a load in the MMA-issue and epilogue loops.
struct ChainArgs {
  ChainLayer layers[kChainMax];
  int n_layers;
  unsigned int *counter;
};

__global__ void __launch_bounds__(kThreads, 1) conv3x3_umma_chain_kernel(const __grid_constant__ ChainArgs args) {
  const ChainLayer *layers = args.layers;
  const int n_layers = args.n_layers;
  unsigned int *counter = args.counter;
  const int pdl = 0;
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  __shared__ uint32_t chain_tmem_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&chain_tmem_s)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = chain_tmem_s;
  for (int l = 0; l < n_layers; ++l) {
    const ChainLayer &L = layers[l];
    if ((int)blockIdx.x < L.p.grid) {
      // (layers fed by plain loads instead of TMA have no single waiting warp: they would need a full barrier - the
      // prepare call refuses them)
      // (one counter per layer boundary: CTAs that sit a layer out arrive for it at once, which must not count towards
      // the barriers of other layers)
      conv_layer<false, true>(L.p, L.tm1, L.tm2, smem_dyn, tmem_base, l == 0, counter + (l > 0 ? l - 1 : 0), gridDim.x);
    } else if (l == 0 && pdl) {
      // idle in the first layer: still take part in the programmatic-dependent-launch handshake
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (l + 1 < n_layers) {
      // ARRIVE at the grid barrier of layer l (conv_layer ended with __threadfence + __syncthreads: this CTA's stores
      // are ordered before the increment); the wait happens inside the next layer, right before its first TMA load
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter + l, 1u);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Operand format of the merged-mode layers (N <= 64 output channels per CTA): 1 (default) = fp16 hi / lo split wherever
// the layer can be fed in 16-channel TMA boxes, 0 = 3xTF32 everywhere, 2 = like 1 but never an 8-channel chunk (calibration).
// Taken from RA_UMMA_F16 on first use; ra_conv3x3_umma_set_f16 overrides it (filter images packed for the old mode become
// invalid).  MEASURED (tools/f16_ab.py, profiles/r04*): same error against the fp64 convolution / the CPU oracle as 3xTF32
// (slightly lower), the KITTI eval forward 14.9 -> 13.1 ms with the 4-channel-plane feed, and see DESIGN 4.1 for the
// 16-channel-box feed.
// + 4 (modes 5 / 6): the tensor core folds the correction half into the main half at the end of a tile (UmmaConvParams::fold;
// correct, TMEM reads halved, measured no faster - the layers are bound by shared-memory bandwidth, DESIGN 4.1 - so off by default).
int g_f16_mode = -1;
int umma_f16_raw() {
  if (g_f16_mode < 0) {
    const char *e = getenv("RA_UMMA_F16");
    g_f16_mode = e == nullptr ? 1 : atoi(e);
    if (g_f16_mode < 0 || g_f16_mode > 6 || (g_f16_mode & 3) == 3 || g_f16_mode == 4) g_f16_mode = 1;
  }
  return g_f16_mode;
}
int umma_f16_mode() { return umma_f16_raw() & 3; }
bool umma_fold_mode() { return (umma_f16_raw() & 4) != 0; }

// Tile plan shared by the launcher and the weight packer (through ra_conv3x3_umma_plan).
struct Plan {
  int KC, NP, NPc, n_split, merged, TH, TW, TWP, n_mt, slots_alloc, n_chunks, stages, acc_cols, stage_bytes;
  int rowstack, mt_stride;
  int f16;  // fp16 hi / lo operand split (ra_conv3x3_umma_set_f16): merged mode, KC % 16 == 0, TMA feed in 16-channel boxes
  int zone_bytes;  // f16: the landing zone of the 64-byte-per-pixel boxes the planner counted into stage_bytes
  int ksplit, nbuf;
  int w_resident, w_res_bytes, grid;
  size_t smem_bytes;
};

// Cost model (cycles per CTA), calibrated with tools/umma_rate.cu, tools/conv_timeline.py and the ncu captures
// under profiles/: a tcgen05.mma with M=128, K=8 occupies the tensor core for max(48, N/2) cycles (the A operand
// streams from shared memory) and one issuing warp sustains an instruction every 50-75 cycles.
// fp16 split: can the layer be fed in 16-channel TMA boxes?  (16-byte global strides; a box must not straddle the two inputs
// of a channel concatenation; channels past Cin are zero-filled by the TMA unit)
bool f16_feasible(int C1, int C2) {
  return umma_f16_mode() != 0 && getenv("RA_CONV_NO_TMA") == nullptr && (C1 % 4) == 0 && (C2 % 4) == 0 &&
         (C2 == 0 || (C1 % 16) == 0);
}

int make_plan(int C1, int C2, int Cout, int Hout, int Wout, int pool, int B, Plan *pl) {
  const int Cin = C1 + C2;
  const bool f16_ok = f16_feasible(C1, C2);
  const int NP = round_up(Cout, 16);
  if (NP > 256) return RA_ERR_UNSUPPORTED;
  if ((Wout & 1) || Wout < 2) return RA_ERR_UNSUPPORTED;
  if (pool == 2 && (Hout & 1)) return RA_ERR_UNSUPPORTED;
  if (B < 1) B = 1;
  const size_t smem_cap = 200 * 1024;
  const double kProdBytesPerCycle = 40.0;  // TMA + in-place hi/lo split by the converter warps
  double best = 1e30;
  Plan bp{};
  bool found = false;
  for (int n_split = 1; n_split <= 8; n_split *= 2) {
    if (NP % (16 * n_split) != 0) break;
    const int NPc = NP / n_split;
   for (int merged = 1; merged >= 0; --merged) {
    // merged: [B_hi; B_lo] stacked along N, one MMA (N' = 2 NPc <= 256) reads A_hi once; 2 instead of 3 MMAs
    if (merged && 2 * NPc > 256) continue;
    // 0 (default): off, 1: by the cost model, 2: on every narrow layer (calibration).  MEASURED on B200
    // (tools/rowstack_ab.py, profiles/r02c_rowstack_ab.txt): results identical to 1e-6, the MMA phase shrinks 3x, but
    // the epilogue reads three times the TMEM columns and becomes the bound - controller layer 1 takes 141 us instead
    // of 82 us, the 48x48-patch layers gain nothing (they are fill / drain latency) - so the mode stays opt-in until the
    // epilogue is spread over more warps.
    static const int rs_mode = []() {
      const char *e = getenv("RA_UMMA_ROWSTACK");
      return e == nullptr ? 0 : atoi(e);
    }();
    const bool rs_ok = merged && 6 * NPc <= 256 && rs_mode != 0;
    // rowstack: the three kx taps of a filter row stacked along N (6 NPc columns), see issue_chunk_rs
   for (int rs = rs_ok ? 1 : 0; rs >= 0; --rs) {
    if (rs_mode == 2 && !rs && NP <= 32) continue;  // calibration: narrow layers run row-stacked or not at all
    const int cols_mt = rs ? 6 * NPc : (merged ? 2 * NPc : NPc);
    const int mt_stride = rs ? 126 : 128;
    const int mt_max = 512 / cols_mt;  // all 512 TMEM columns, single accumulator buffer
    if (mt_max < 1) continue;
    const double n_mma = (merged || rs) ? 2.0 : 3.0;  // instructions per step and m-tile, one dependent chain
    // tensor-core time of one M=128 x K=8 instruction: max(~48, N/2) cycles (tools/umma_rate.cu)
    const double hw_n = NPc / 2 > 48 ? NPc / 2 : 48.0, hw_2n = NPc > 48 ? (double)NPc : 48.0;
    const double hw_rs = (3.0 * NPc > 48 ? 3.0 * NPc : 48.0) + (1.5 * NPc > 48 ? 1.5 * NPc : 48.0);
    const double hw_cycles = rs ? hw_rs : (merged ? hw_2n + hw_n : 3.0 * hw_n);
    // RA_UMMA_F16 (experiment, eval only): 1 = fp16 hi / lo split where the cost model likes it, 2 = wherever it is
    // possible.  kind::f16 consumes K = 16 per instruction at the cost of a kind::tf32 one (tools/umma_kind_rate.cu):
    // half the instructions per channel chunk.
    const int f16_mode = umma_f16_mode();
    for (int KC = 8; KC <= 32; KC *= 2) {
      if (KC > 8 && KC / 2 >= Cin) continue;
      const int f16 = (f16_ok && merged && !rs && KC % 16 == 0) ? 1 : 0;
      if (f16_mode == 2 && f16_ok && merged && !rs && !f16 && Cin > 8) continue;
      const int planes = KC / 4;
      const int n_chunks = (Cin + KC - 1) / KC;
      const size_t w_chunk_bytes = (size_t)9 * (f16 ? planes / 2 : planes) * 2 * NPc * 16;
      const size_t w_total = w_chunk_bytes * n_chunks;
      for (int TW = Wout; TW >= 2; TW = (TW % 2 == 0 ? TW / 2 : 0)) {
        if (TW & 1) break;
        const int TWP = TW + 2;
        for (int TH = (pool == 2 ? 2 : 1); TH <= Hout; TH += (pool == 2 ? 2 : 1)) {
          const int n_mt = (TH * TWP + mt_stride - 1) / mt_stride;
          if (n_mt > mt_max || n_mt > 2 * kMmaWarps) break;  // each MMA warp owns at most two m-tiles
          const int slots_alloc = round_up(n_mt * 128 + 2 * TWP + 2, 8);  // planes stay 128-byte aligned (TMA)
          if (f16 && (TWP > 256 || TH + 2 > 256)) continue;  // TMA box extents
          // fp16: hi / lo' share one region; the boxes land pixel-major in a zone behind it (counted at its stride-1 size;
          // + alignment of the zone inside the stage)
          const size_t zone_bytes = f16 ? (size_t)(KC / 16) * round_up(TWP * (TH + 2) * 64, 512) + 512 : 0;
          const size_t in_bytes = (size_t)(f16 ? 1 : 2) * planes * slots_alloc * 16 + zone_bytes;
          const size_t pool_bytes = pool == 2 ? (size_t)(n_mt * 64) * kPoolLd * 4 : 0;
          const int tiles = ((Wout + TW - 1) / TW) * ((Hout + TH - 1) / TH) * B;
          int grid_t = ra::kNumSMs / n_split;  // CTAs per channel split
          if (grid_t > tiles) grid_t = tiles;
          if (grid_t < 1) continue;
          const int tiles_per_cta = (tiles + grid_t - 1) / grid_t;
          for (int nbuf = 2; nbuf >= 1; --nbuf) {
          if (n_mt * cols_mt * nbuf > 512) continue;
          // K split (partial accumulators per m-tile, summed by the epilogue): measured useless - back-to-back MMAs
          // into one accumulator already pipeline at the tensor core's rate - so the planner keeps 1; the kernel
          // path stays (RA_UMMA_KSPLIT with a forced plan) as the experiment that shows it.
          const int ksplit = 1;
          for (int resident = 1; resident >= 0; --resident) {
            const size_t stage_bytes = in_bytes + (resident ? 0 : w_chunk_bytes);
            const size_t fixed = pool_bytes + (resident ? w_total : 0);
            if (fixed + 2 * stage_bytes > smem_cap) continue;
            int st = (int)((smem_cap - fixed) / stage_bytes);
            if (st > kMaxStages) st = kMaxStages;
            const double per_tap = (double)(rs ? 3 : 9) * (f16 ? KC / 16 : KC / 8) * n_chunks;  // (tap | filter row, k) steps per tile
            const double step_tc = n_mt * hw_cycles;  // tensor-core occupancy of one step
            // issue: ~75 cycles per instruction for a warp that owns one m-tile, ~50 with two (tools/conv_timeline.py)
            const int mt_warp = (n_mt + kMmaWarps - 1) / kMmaWarps;
            const double step_issue = (mt_warp == 1 ? 75.0 : 100.0) * n_mma;
            const double step = step_tc > step_issue ? step_tc : step_issue;
            const double mma_item = per_tap * step;
            const double prod_item = (double)n_chunks * stage_bytes / kProdBytesPerCycle;
            // measured: slow (TMEM latency); rowstack reads three times the columns and exchanges rows between lanes
            const double epi_item = 500.0 + n_mt * (NPc / 16) * (rs ? 1500.0 : (merged ? 550.0 : 450.0));
            double item = mma_item > prod_item ? mma_item : prod_item;
            if (nbuf == 1)
              item += epi_item;  // single accumulator buffer: the next tile's MMAs wait for the epilogue
            else if (epi_item > item)
              item = epi_item;
            item += 500.0 * n_chunks + 1500.0;  // barrier round trips per chunk / per tile
            // pipeline fill: the first stage of every CTA and the last epilogue are exposed
            const double cost = tiles_per_cta * item + prod_item / n_chunks + epi_item +
                                (resident ? (double)w_total / 40.0 : 0.0) + 2000.0;
            if (cost < best) {
              best = cost;
              found = true;
              bp.KC = KC;
              bp.NP = NP;
              bp.NPc = NPc;
              bp.n_split = n_split;
              bp.merged = merged;
              bp.rowstack = rs;
              bp.f16 = f16;
              bp.zone_bytes = (int)zone_bytes;
              bp.mt_stride = mt_stride;
              bp.TH = TH;
              bp.TW = TW;
              bp.TWP = TWP;
              bp.n_mt = n_mt;
              bp.slots_alloc = slots_alloc;
              bp.n_chunks = n_chunks;
              bp.acc_cols = n_mt * ksplit * cols_mt;
              bp.ksplit = ksplit;
              bp.nbuf = nbuf;
              bp.stage_bytes = (int)stage_bytes;
              bp.stages = st;
              bp.w_resident = resident;
              bp.w_res_bytes = resident ? (int)w_total : 0;
              bp.grid = grid_t * n_split;
              bp.smem_bytes = fixed + (size_t)st * stage_bytes;
            }
          }
          }
        }
      }
    }
   }
   }
  }
  if (!found) return RA_ERR_UNSUPPORTED;
  // K split for PRECISION: the TMEM accumulators truncate every fp32 accumulation, a bias proportional to the
  // accumulator's magnitude and to the number of accumulating instructions.  With k partial
  // accumulators per m-tile (step q -> accumulator q % k, summed in round-to-nearest fp32 by the epilogue) each one
  // sees 1/k of the adds at 1/k of the magnitude.  Free TMEM columns are used for it, up to kKsplitMax.
  {
    static const int cap = []() {
      const char *e = getenv("RA_UMMA_KSPLIT");
      return e ? (atoi(e) < 1 ? 1 : atoi(e)) : kKsplitDefault;
    }();
    const int cols_mt = bp.rowstack ? 6 * bp.NPc : (bp.merged ? 2 * bp.NPc : bp.NPc);
    int ks = 512 / (bp.nbuf * bp.n_mt * cols_mt);
    const int steps = (bp.rowstack ? 3 : 9) * (bp.f16 ? bp.KC / 16 : bp.KC / 8) * bp.n_chunks;
    if (ks > steps) ks = steps;
    if (ks > cap) ks = cap;
    if (ks < 1) ks = 1;
    bp.ksplit = ks;
    bp.acc_cols = bp.n_mt * ks * cols_mt;
  }
  *pl = bp;
  return RA_OK;
}

// Calibration hook (tools/bench_conv_layers.py): RA_UMMA_FORCE="KC,TH,TW,n_split,resident" overrides the search.
int make_plan_forced(int C1, int C2, int Cout, int Hout, int Wout, int pool, int B, Plan *pl) {
  const int Cin = C1 + C2;
  const char *f = getenv("RA_UMMA_FORCE");
  if (f == nullptr) return make_plan(C1, C2, Cout, Hout, Wout, pool, B, pl);
  int KC = 8, TH = 2, TW = 2, n_split = 1, resident = 0;
  if (sscanf(f, "%d,%d,%d,%d,%d", &KC, &TH, &TW, &n_split, &resident) != 5) return RA_ERR_INVALID_ARG;
  const int NP = round_up(Cout, 16);
  if (NP % (16 * n_split) != 0 || TW > Wout || TH > Hout || (TW & 1) || (pool == 2 && (TH & 1))) return RA_ERR_UNSUPPORTED;
  Plan bp{};
  bp.KC = KC;
  bp.NP = NP;
  bp.NPc = NP / n_split;
  bp.n_split = n_split;
  bp.merged = bp.NPc <= 64 ? 1 : 0;
  bp.rowstack = 0;
  bp.f16 = (f16_feasible(C1, C2) && bp.merged && KC % 16 == 0 && TW + 2 <= 256 && TH + 2 <= 256) ? 1 : 0;
  bp.mt_stride = 128;
  const int cols_mt = bp.merged ? 2 * bp.NPc : bp.NPc;
  bp.TH = TH;
  bp.TW = TW;
  bp.TWP = TW + 2;
  bp.n_mt = (TH * bp.TWP + 127) / 128;
  if (bp.n_mt * cols_mt > 512 || bp.n_mt > 2 * kMmaWarps) return RA_ERR_UNSUPPORTED;
  bp.nbuf = bp.n_mt * cols_mt * 2 <= 512 ? 2 : 1;
  bp.ksplit = 512 / (bp.nbuf * bp.n_mt * cols_mt);
  if (bp.ksplit > (8 + bp.n_mt - 1) / bp.n_mt) bp.ksplit = (8 + bp.n_mt - 1) / bp.n_mt;
  if (const char *ks = getenv("RA_UMMA_KSPLIT")) bp.ksplit = atoi(ks) < 1 ? 1 : (atoi(ks) < bp.ksplit ? atoi(ks) : bp.ksplit);
  bp.slots_alloc = round_up(bp.n_mt * 128 + 2 * bp.TWP + 2, 8);
  bp.n_chunks = (Cin + KC - 1) / KC;
  bp.acc_cols = bp.n_mt * bp.ksplit * cols_mt;
  const int planes = KC / 4;
  bp.zone_bytes = bp.f16 ? (KC / 16) * round_up(bp.TWP * (TH + 2) * 64, 512) + 512 : 0;
  const size_t in_bytes = (size_t)(bp.f16 ? 1 : 2) * planes * bp.slots_alloc * 16 + bp.zone_bytes;
  const size_t w_chunk = (size_t)9 * (bp.f16 ? planes / 2 : planes) * 2 * bp.NPc * 16;
  const size_t pool_bytes = pool == 2 ? (size_t)(bp.n_mt * 64) * kPoolLd * 4 : 0;
  const size_t stage_bytes = in_bytes + (resident ? 0 : w_chunk);
  const size_t fixed = pool_bytes + (resident ? w_chunk * bp.n_chunks : 0);
  if (fixed + 2 * stage_bytes > 200 * 1024) return RA_ERR_UNSUPPORTED;
  int st = (int)((200 * 1024 - fixed) / stage_bytes);
  if (st > kMaxStages) st = kMaxStages;
  bp.stages = st;
  bp.stage_bytes = (int)stage_bytes;
  bp.w_resident = resident;
  bp.w_res_bytes = resident ? (int)(w_chunk * bp.n_chunks) : 0;
  const int tiles = ((Wout + TW - 1) / TW) * ((Hout + TH - 1) / TH) * B;
  int grid_t = ra::kNumSMs / n_split;
  if (grid_t > tiles) grid_t = tiles;
  bp.grid = grid_t * n_split;
  bp.smem_bytes = fixed + (size_t)st * stage_bytes;
  *pl = bp;
  return RA_OK;
}

long long *g_conv_dbg = nullptr;

// ---- tensor maps -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

struct MapKey {
  const void *ptr;
  int C, W, H, B, RW, RH, box_c;
  bool operator==(const MapKey &o) const {
    return ptr == o.ptr && C == o.C && W == o.W && H == o.H && B == o.B && RW == o.RW && RH == o.RH && box_c == o.box_c;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey &k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.ptr);
    const int v[7] = {k.C, k.W, k.H, k.B, k.RW, k.RH, k.box_c};
    for (int i = 0; i < 7; ++i) h = h * 0x9E3779B97F4A7C15ull + (uint64_t)v[i];
    return (size_t)h;
  }
};

// 4-D tiled map over an NHWC fp32 activation: dims (C, W, H, B) innermost first, box = 4 channels x RW x RH x 1
// pixel block = one channel plane of the shared-memory tile.  No swizzle, zero fill out of bounds.
// box_c = 16 (fp16 variant): 16 channels = 64-byte rows, pixel-major, SWIZZLE_64B (see convert_stage_f16).
bool activation_map(const float *x, int C, int W, int H, int B, int RW, int RH, int box_c, CUtensorMap *out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{x, C, W, H, B, RW, RH, box_c};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return true;
  }
  EncodeTiledFn enc = tensor_map_encoder();
  if (enc == nullptr) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)RW, (cuuint32_t)RH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, tm);
  *out = tm;
  return true;
}

}  // namespace

// fp16 hi / lo filter image (RA_UMMA_F16 plans): v [rows][KC/4][NPc][4] = the fp32 filter in the operand order of one half
// of the tf32 image (params.umma_layout; rows = n_split * n_chunks * 9 taps)  ->  out [rows][KC/8][2 NPc][8 halves]: rows
// 0..NPc-1 of a plane hold hi = fp16(w) of 8 consecutive input channels, rows NPc..2NPc-1 lo' = fp16((w - hi) * 2^11) -
// the same split convert_stage_f16 applies to the activations.
__global__ void __launch_bounds__(256) umma_pack_f16_kernel(const float4 *__restrict__ v, long long n, int planes8, int NPc,
                                                            uint4 *__restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row, 8-channel plane, filter row)
  if (i >= n) return;
  const int r = (int)(i % NPc);
  const long long rk = i / NPc;  // row * planes8 + k
  const float4 a = v[(2 * rk) * NPc + r], b = v[(2 * rk + 1) * NPc + r];
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half h0 = __float2half_rn(x[2 * j]), h1 = __float2half_rn(x[2 * j + 1]);
    const __half l0 = __float2half_rn((x[2 * j] - __half2float(h0)) * 2048.0f);
    const __half l1 = __float2half_rn((x[2 * j + 1] - __half2float(h1)) * 2048.0f);
    h[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
    l[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
  }
  out[rk * 2 * NPc + r] = make_uint4(h[0], h[1], h[2], h[3]);
  out[rk * 2 * NPc + NPc + r] = make_uint4(l[0], l[1], l[2], l[3]);
}

extern "C" int ra_conv3x3_umma_set_f16(int mode) {
  const int prev = umma_f16_raw();
  if (mode >= 0 && mode <= 6 && (mode & 3) != 3 && mode != 4) g_f16_mode = mode;
  return prev;
}

extern "C" int ra_umma_pack_f16(const float *v, long long rows, int KC, int NPc, float *out, void *stream) {
  if (!v || !out || rows < 0 || KC < 16 || (KC % 16) != 0 || NPc <= 0) return RA_ERR_INVALID_ARG;
  const long long n = rows * (KC / 8) * NPc;
  if (n == 0) return RA_OK;
  umma_pack_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ra::as_stream(stream)>>>(
      reinterpret_cast<const float4 *>(v), n, KC / 8, NPc, reinterpret_cast<uint4 *>(out));
  return ra::finish_launch("umma_pack_f16_kernel");
}

// Diagnostics: when set, every conv3x3_umma CTA writes 16 slots: 8 clock64() stamps (start, setup done, first stage
// staged, producers done, first MMA issuable, MMAs issued, first accumulator complete, all done) and the wait accounting
// of RA_WAIT (8 TMA warp waiting for a free stage, 9 converters waiting for a box, 10 / 11 MMA warp 0 waiting for a staged
// chunk / a free accumulator, 14 converters busy; cycles; fp16 variant only).
extern "C" int ra_debug_conv_timeline(long long *device_buf) {
  g_conv_dbg = device_buf;
  return RA_OK;
}

// Plan query for the host-side weight packer.
extern "C" int ra_conv3x3_umma_plan_split(int C1, int C2, int Cout, int Hout, int Wout, int pool, int B, int *KC,
                                          int *NPc, int *n_split, int *n_chunks, int *rowstack) {
  if (C1 < 1 || C2 < 0) return RA_ERR_INVALID_ARG;
  Plan pl;
  const int rc = make_plan_forced(C1, C2, Cout, Hout, Wout, pool, B, &pl);
  if (rc != RA_OK) return rc;
  if (KC) *KC = pl.KC;
  if (NPc) *NPc = pl.NPc;
  if (n_split) *n_split = pl.n_split;
  if (n_chunks) *n_chunks = pl.n_chunks;
  if (rowstack) *rowstack = pl.rowstack | (pl.f16 << 1);  // layout flags of the packed filter image: bit 0 row-stacked, bit 1 fp16
  return RA_OK;
}

extern "C" int ra_conv3x3_umma_plan(int Cin, int Cout, int Hout, int Wout, int pool, int B, int *KC, int *NPc,
                                    int *n_split, int *n_chunks, int *rowstack) {
  return ra_conv3x3_umma_plan_split(Cin, 0, Cout, Hout, Wout, pool, B, KC, NPc, n_split, n_chunks, rowstack);
}

// Full plan dump (diagnostics / DESIGN.md tables): info[19] = KC, NPc, n_split, n_chunks, TH, TW, n_mt, stages,
// merged, w_resident, grid, smem_bytes, acc_cols, stage_bytes, w_res_bytes, slots_alloc, ksplit, nbuf.
extern "C" int ra_conv3x3_umma_plan_info(int Cin, int Cout, int Hout, int Wout, int pool, int B, int *info) {
  Plan pl;
  const int rc = make_plan_forced(Cin, 0, Cout, Hout, Wout, pool, B, &pl);
  if (rc != RA_OK) return rc;
  if (!info) return RA_ERR_INVALID_ARG;
  const int v[19] = {pl.KC, pl.NPc, pl.n_split, pl.n_chunks, pl.TH, pl.TW, pl.n_mt, pl.stages, pl.merged,
                     pl.w_resident, pl.grid, (int)pl.smem_bytes, pl.acc_cols, pl.stage_bytes, pl.w_res_bytes,
                     pl.slots_alloc, pl.ksplit, pl.nbuf, pl.rowstack | (pl.f16 << 1)};
  for (int i = 0; i < 19; ++i) info[i] = v[i];
  return RA_OK;
}

namespace {

// Tile plan, kernel parameters and tensor maps of one layer (shared by the single-layer launch and the chain).
// Returns RA_OK with *empty = true for an empty batch.
int build_layer(const float *x1, int C1, const float *x2, int C2, const float *wpack, const float *scale,
                const float *shift, int B, int Hin, int Win, int Cout, int upsample, int pool, int relu, float *y,
                UmmaConvParams *pp, CUtensorMap *tm1p, CUtensorMap *tm2p, Plan *plp, size_t *smem_out, bool *empty) {
  UmmaConvParams &p = *pp;
  CUtensorMap &tm1 = *tm1p, &tm2 = *tm2p;
  Plan &pl = *plp;
  *empty = false;
  if (!x1 || !wpack || !scale || !shift || !y || C1 < 1 || C2 < 0 || (C2 > 0 && !x2) || B < 0 || Hin < 1 || Win < 1 ||
      Cout < 1)
    return RA_ERR_INVALID_ARG;
  if ((upsample != 1 && upsample != 2) || (pool != 1 && pool != 2)) return RA_ERR_UNSUPPORTED;
  p.x1 = x1;
  p.x2 = x2;
  p.wpack = wpack;
  p.scale = scale;
  p.shift = shift;
  p.y = y;
  p.C1 = C1;
  p.C2 = C2;
  p.Cin = C1 + C2;
  p.Cout = Cout;
  p.B = B;
  p.Hin = Hin;
  p.Win = Win;
  p.Hout = Hin * upsample;
  p.Wout = Win * upsample;
  p.up = upsample;
  p.pool = pool;
  p.relu = relu;
  if (B == 0) {
    *empty = true;
    return RA_OK;
  }
  const int rc = make_plan_forced(C1, C2, Cout, p.Hout, p.Wout, pool, B, &pl);
  if (rc != RA_OK) return rc;
  p.TH = pl.TH;
  p.TW = pl.TW;
  p.TWP = pl.TWP;
  p.n_mt = pl.n_mt;
  p.KC = pl.KC;
  p.n_chunks = pl.n_chunks;
  p.NPc = pl.NPc;
  p.n_split = pl.n_split;
  p.merged = pl.merged;
  p.rowstack = pl.rowstack;
  p.f16 = pl.f16;
  p.mt_stride = pl.mt_stride;
  p.slots_alloc = pl.slots_alloc;
  p.stages = pl.stages;
  p.w_resident = pl.w_resident;
  p.w_res_bytes = pl.w_res_bytes;
  p.acc_cols = pl.acc_cols;
  p.ksplit = pl.ksplit;
  p.nbuf = pl.nbuf;
  p.stage_bytes = pl.stage_bytes;
  p.tiles_x = (p.Wout + p.TW - 1) / p.TW;
  p.tiles_y = (p.Hout + p.TH - 1) / p.TH;
  const long long items = (long long)p.tiles_x * p.tiles_y * B * p.n_split;
  if (items > 0x7fffffffLL) return RA_ERR_UNSUPPORTED;
  p.n_items = (int)items;
  p.vec4 = ((C1 & 3) == 0 && (C2 & 3) == 0) ? 1 : 0;
  p.dbg = g_conv_dbg;
  // ---- TMA feed: needs channel counts that are multiples of 4 (16-byte global strides) and boxes <= 256
  const bool no_tma = getenv("RA_CONV_NO_TMA") != nullptr;  // diagnostics: plain-load producers everywhere
  // + static shared memory (2.3 KB; 3.8 KB in the rowstack variant) <= 227 KB
  const size_t kSmemMax = (size_t)(pl.rowstack ? 222 : 224) * 1024;
  p.tma = 0;
  p.RW = upsample == 2 ? p.TW / 2 + 1 : p.TWP;
  p.RH = upsample == 2 ? p.TH / 2 + 2 : p.TH + 2;
  p.raw_plane_bytes = 0;
  p.raw_off = 0;
  size_t smem_bytes = pl.smem_bytes;
  memset(&tm1, 0, sizeof(tm1));
  memset(&tm2, 0, sizeof(tm2));
  if (!no_tma && p.vec4 && p.RW <= 256 && p.RH <= 256 && (reinterpret_cast<uintptr_t>(x1) & 15) == 0 &&
      (C2 == 0 || (reinterpret_cast<uintptr_t>(x2) & 15) == 0) && (reinterpret_cast<uintptr_t>(wpack) & 15) == 0) {
    bool ok = true;
    int stages = pl.stages, stage_bytes = pl.stage_bytes;
    if (pl.f16) {
      // fp16 variant: every stage = [hi / lo' operand planes][filter slice][pad][landing zone: one 64-byte-per-pixel box per
      // 16-channel group], zones 512-byte aligned (SWIZZLE_64B pattern; the kernel aligns its window to 1024 bytes)
      ok = C2 == 0 || (C1 % 16) == 0;
      p.raw_plane_bytes = round_up(p.RW * p.RH * 64, 512);
      p.raw_off = round_up(pl.stage_bytes - pl.zone_bytes, 512);
      stage_bytes = p.raw_off + (p.KC / 16) * p.raw_plane_bytes;
      p.w_res_bytes = round_up(pl.w_res_bytes, 512);
      const size_t fixed = pl.smem_bytes - (size_t)pl.stages * pl.stage_bytes - pl.w_res_bytes + p.w_res_bytes;
      while (stages > 2 && fixed + (size_t)stages * stage_bytes > kSmemMax - 1024) --stages;
      ok = ok && fixed + (size_t)stages * stage_bytes <= kSmemMax - 1024;
      if (ok) smem_bytes = fixed + (size_t)stages * stage_bytes;
    } else if (upsample == 2) {
      // landing zone of the low-resolution box, appended to every stage; drop stages if it does not fit
      p.raw_plane_bytes = round_up(p.RW * p.RH * 16, 128);
      p.raw_off = pl.stage_bytes;
      stage_bytes = pl.stage_bytes + (p.KC / 4) * p.raw_plane_bytes;
      const size_t fixed = pl.smem_bytes - (size_t)pl.stages * pl.stage_bytes;
      while (stages > 2 && fixed + (size_t)stages * stage_bytes > kSmemMax - 128) --stages;
      ok = fixed + (size_t)stages * stage_bytes <= kSmemMax - 128;
      if (ok) smem_bytes = fixed + (size_t)stages * stage_bytes;
    }
    const int box_c = pl.f16 ? 16 : 4;
    ok = ok && activation_map(x1, C1, Win, Hin, B, p.RW, p.RH, box_c, &tm1);
    ok = ok && (C2 == 0 || activation_map(x2, C2, Win, Hin, B, p.RW, p.RH, box_c, &tm2));
    if (ok) {
      p.tma = 1;
      p.stages = stages;
      p.stage_bytes = stage_bytes;
    } else {
      p.raw_plane_bytes = 0;
      p.raw_off = 0;
      p.w_res_bytes = pl.w_res_bytes;
      smem_bytes = pl.smem_bytes;
    }
  }
  p.fold = 0;
  if (p.f16 && p.tma) {
    const size_t id_bytes = (size_t)(p.NPc / 4) * p.NPc * 16;
    if (umma_fold_mode() && smem_bytes + id_bytes <= kSmemMax - 1024) {
      p.fold = 1;
      smem_bytes += id_bytes;
    }
  }
  if (p.f16 && !p.tma) {  // the fp16 split lives in the TMA-mode converters only; the filter image is already packed for it
    ra::set_last_error("conv3x3_umma: the fp16 plan needs the TMA feed (16-byte aligned tensors; C1 % 16 == 0 for a "
                       "channel concatenation - plan it with ra_conv3x3_umma_plan_split)", cudaSuccess);
    return RA_ERR_UNSUPPORTED;
  }
  p.grid = pl.grid;
  p.pdl = 0;
  p.four_term = getenv("RA_UMMA_4TERM") != nullptr ? 1 : 0;
  p.split_corr = getenv("RA_UMMA_JOINT_CORR") == nullptr ? 1 : 0;
  if (p.f16) {  // the corrections are 2^11-scaled: they must live in their own accumulator half, and lo' x lo' is not formed
    p.four_term = 0;
    p.split_corr = 1;
  }
  *smem_out = smem_bytes + (p.f16 ? 1024 : 128);  // alignment slack (the kernel rounds its window up to 128 / 1024 bytes)
  return RA_OK;
}

bool conv_attrs() {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(conv3x3_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_umma_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               224 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_umma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(conv3x3_umma_kernel)", e);
      return false;
    }
    // diagnostics: pin the SM shared-memory carve-out to its maximum for every launch of this kernel, so that
    // consecutive layers with different dynamic sizes never trigger a carve-out reconfiguration
    if (getenv("RA_CONV_CARVEOUT") != nullptr)
      (void)cudaFuncSetAttribute(conv3x3_umma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 (int)cudaSharedmemCarveoutMaxShared);
    attr_set = true;
  }
  return true;
}

bool conv_use_pdl() {
  static const bool use_pdl = []() {
    const char *e = getenv("RA_CONV_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return use_pdl;
}

}  // namespace

extern "C" int ra_conv3x3_umma_f32(const float *x1, int C1, const float *x2, int C2, const float *wpack,
                                   const float *scale, const float *shift, int B, int Hin, int Win, int Cout,
                                   int upsample, int pool, int relu, float *y, void *stream) {
  UmmaConvParams p;
  CUtensorMap tm1, tm2;
  Plan pl;
  size_t smem_bytes = 0;
  bool empty = false;
  const int brc = build_layer(x1, C1, x2, C2, wpack, scale, shift, B, Hin, Win, Cout, upsample, pool, relu, y, &p, &tm1,
                              &tm2, &pl, &smem_bytes, &empty);
  if (brc != RA_OK || empty) return brc;
  if (!conv_attrs()) return RA_ERR_CUDA;
  const bool use_pdl = conv_use_pdl();
  p.pdl = use_pdl ? 1 : 0;
  if (use_pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = ra::as_stream(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = p.rowstack ? cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<true>, p, tm1, tm2)
                    : p.fold   ? cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<false, true, true>, p, tm1, tm2)
                    : p.f16    ? cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<false, true>, p, tm1, tm2)
                               : cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<false>, p, tm1, tm2);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaLaunchKernelEx(conv3x3_umma_kernel)", e);
      return RA_ERR_CUDA;
    }
  } else {
    if (p.rowstack)
      conv3x3_umma_kernel<true><<<pl.grid, kThreads, smem_bytes, ra::as_stream(stream)>>>(p, tm1, tm2);
    else if (p.fold)
      conv3x3_umma_kernel<false, true, true><<<pl.grid, kThreads, smem_bytes, ra::as_stream(stream)>>>(p, tm1, tm2);
    else if (p.f16)
      conv3x3_umma_kernel<false, true><<<pl.grid, kThreads, smem_bytes, ra::as_stream(stream)>>>(p, tm1, tm2);
    else
      conv3x3_umma_kernel<false><<<pl.grid, kThreads, smem_bytes, ra::as_stream(stream)>>>(p, tm1, tm2);
  }
  return ra::finish_launch("conv3x3_umma_kernel");
}

// ---- chain of layers in one launch ------------------------------------------------------------------------------
extern "C" size_t ra_conv3x3_umma_chain_desc_bytes(int n_layers) {
  return (n_layers < 1 || n_layers > kChainMax) ? 0 : sizeof(ChainArgs);
}

extern "C" int ra_conv3x3_umma_chain_prepare(const ra_conv_layer_t *layers, int n_layers, void *desc_host, int *grid_out,
                                             size_t *smem_out) {
  if (!layers || n_layers < 1 || n_layers > kChainMax || !desc_host || !grid_out || !smem_out) return RA_ERR_INVALID_ARG;
  ChainArgs *args = static_cast<ChainArgs *>(desc_host);
  memset(args, 0, sizeof(ChainArgs));
  int grid = 1;
  size_t smem = 0;
  int rc = RA_OK;
  for (int l = 0; l < n_layers && rc == RA_OK; ++l) {
    const ra_conv_layer_t &L = layers[l];
    Plan pl;
    size_t sb = 0;
    bool empty = false;
    rc = build_layer(L.x1, L.C1, L.x2, L.C2, L.wpack, L.scale, L.shift, L.B, L.Hin, L.Win, L.Cout, L.upsample, L.pool,
                     L.relu, L.y, &args->layers[l].p, &args->layers[l].tm1, &args->layers[l].tm2, &pl, &sb, &empty);
    // (the chain runs the default variant, and its split grid barrier waits in the TMA warp)
    if (rc == RA_OK && (empty || pl.rowstack || pl.f16 || (l > 0 && !args->layers[l].p.tma))) rc = RA_ERR_UNSUPPORTED;
    if (rc != RA_OK) break;
    args->layers[l].p.dbg = nullptr;
    args->layers[l].p.pdl = 0;  // the chain is launched behind a memset node, without the programmatic-launch attribute
    if (pl.grid > grid) grid = pl.grid;
    if (sb > smem) smem = sb;
  }
  if (rc == RA_OK && grid > ra::kNumSMs) rc = RA_ERR_UNSUPPORTED;  // the spin barrier needs co-resident CTAs
  args->n_layers = n_layers;
  *grid_out = grid;
  *smem_out = smem;
  return rc;
}

extern "C" int ra_conv3x3_umma_chain_run(const void *desc_host, int n_layers, int grid, size_t smem_bytes,
                                         unsigned int *counter_dev, void *stream) {
  if (!desc_host || n_layers < 1 || n_layers > kChainMax || grid < 1 || grid > ra::kNumSMs || !counter_dev)
    return RA_ERR_INVALID_ARG;
  if (!conv_attrs()) return RA_ERR_CUDA;
  cudaStream_t s = ra::as_stream(stream);
  cudaError_t e = cudaMemsetAsync(counter_dev, 0, kChainMax * sizeof(unsigned int), s);
  if (e != cudaSuccess) {
    ra::set_last_error("cudaMemsetAsync(chain counter)", e);
    return RA_ERR_CUDA;
  }
  ChainArgs args = *static_cast<const ChainArgs *>(desc_host);  // (copied into the launch / the graph node)
  args.counter = counter_dev;
  conv3x3_umma_chain_kernel<<<grid, kThreads, smem_bytes, s>>>(args);
  return ra::finish_launch("conv3x3_umma_chain_kernel");
}

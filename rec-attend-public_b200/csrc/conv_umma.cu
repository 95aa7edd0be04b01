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
// Persistent, warp-specialised pipeline (one CTA per SM, 512 threads):
//   warps 8-15  producers: stage the input chunk (with the hi/lo split) and the pre-packed filter
//               slice of (tile, channel-chunk) items into a ring of shared-memory stages;
//               fence.proxy.async + mbarrier arrive (full[s]).
//   warps 4-7   one lane each issues its share of the n_mt x 9 taps x (2|3) tcgen05.mma per item, tcgen05.commit ->
//               empty[s]; after a tile's last chunk tcgen05.commit -> tmem_full[buf].
//   warps 0-3   epilogue: tcgen05.ld (thread = TMEM lane = slot), folded BN scale/shift, ReLU,
//               2x2 max-pool (horizontal by shuffle, vertical through a small shared tile),
//               stores; arrive tmem_empty[buf].  Accumulators are double buffered in TMEM so the
//               epilogue of tile i overlaps the MMAs of tile i+1.
// Narrow layers (N <= 32) are bound by the tensor core's shared-memory reads of A, so there the
// two filter parts are stacked along N ([B_hi; B_lo], one MMA with N' = 2N reads A_hi once) and
// summed in the epilogue; small feature maps split N over CTAs to fill the 148 SMs.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kEpiThreads = 128;   // warps 0-3
constexpr int kMmaWarp0 = 4;       // warps 4-7: one issuing lane each, m-tiles dealt round-robin
constexpr int kMmaWarps = 4;
constexpr int kProdThreads = 256;  // warps 8-15
constexpr int kThreads = kEpiThreads + 32 * kMmaWarps + kProdThreads;
constexpr int kMaxStages = 4;
constexpr int kStageUnroll = 4;
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
  int slots_alloc;           // input slots allocated per plane
  int tiles_x, tiles_y, n_items;
  int stages;
  int w_resident;   // the CTA's whole filter image stays in shared memory for all of its tiles
  int w_res_bytes;  // bytes of that resident image (0 when streaming)
  int acc_cols;     // TMEM columns of one accumulator buffer (n_mt * cols_per_mt)
  int stage_bytes;  // bytes of one shared-memory stage (input hi + lo + filter slice)
  int vec4;         // both sources have channel counts divisible by 4
  long long *dbg;   // optional per-CTA timeline (ra_debug_conv_timeline), 8 slots per CTA
};

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

__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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

#define RA_DBG(slot)                                                                                   \
  do {                                                                                                 \
    if (p.dbg != nullptr) p.dbg[(size_t)blockIdx.x * 8 + (slot)] = clock64();                          \
  } while (0)

__global__ void __launch_bounds__(kThreads, 1) conv3x3_umma_kernel(UmmaConvParams p) {
  if (threadIdx.x == 0) RA_DBG(0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int planes = p.KC / 4;
  const uint32_t plane_bytes = (uint32_t)p.slots_alloc * 16u;
  const int in_floats = planes * p.slots_alloc * 4;  // one of hi / lo
  const int w_chunk_floats = 9 * planes * 2 * p.NPc * 4;
  unsigned char *stage_base = smem_raw + p.w_res_bytes;  // [resident filter image][stages][pool tile]
  float *pool_s = reinterpret_cast<float *>(stage_base + (size_t)p.stages * p.stage_bytes);
  const int ns = blockIdx.x % p.n_split;
  const int tile0 = blockIdx.x / p.n_split, tile_step = gridDim.x / p.n_split;
  const int n_tiles = p.n_items / p.n_split;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), kProdThreads);
      mbar_init(smem_u32(&bar_empty[s]), kMmaWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), kMmaWarps);
      mbar_init(smem_u32(&bar_tempty[s]), kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int slots_in = (p.TH + 2) * p.TWP + 2;  // slots that carry real (or zero-padding) data
  if (threadIdx.x == 0) RA_DBG(1);  // setup done (barriers, TMEM, resident filters)

  if (warp >= kMmaWarp0 + kMmaWarps) {
    // =============================== producers ===============================
    // (Splitting the producers into groups that stage different chunks concurrently was measured: it does not
    // help — the first stage takes twice as long — and needs stages % groups == 0 for the parity waits.)
    const int ptid = tid - (kEpiThreads + 32 * kMmaWarps);
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
  } else if (warp >= kMmaWarp0) {
    // =============================== MMA issuer ===============================
    // The issue of one tcgen05.mma costs ~130 cycles of scalar work in the issuing thread (descriptor
    // arithmetic + R2UR moves, see profiles/), 2-3x the tensor core's own 45-64 cycles per M=128 x K=8
    // instruction, so kMmaWarps threads issue in parallel, each for its own m-tiles (= its own TMEM
    // accumulators; no two threads ever accumulate into the same columns).
    if (lane == 0) {
      const int mw = warp - kMmaWarp0;
      const int cols_mt = p.merged ? 2 * p.NPc : p.NPc;
      const uint32_t idesc_n =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NPc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_2n =
          (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * p.NPc) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t w_plane = (uint32_t)(2 * p.NPc) * 16u;  // bytes between channel planes of the filter image
      const uint32_t plane16 = plane_bytes >> 4, wplane16 = w_plane >> 4;
      const int k8n = p.KC / 8;
      int g = 0, t = 0;
      for (int tile = tile0; tile < n_tiles; tile += tile_step, ++t) {
        const int buf = t & 1;
        mbar_wait(smem_u32(&bar_tempty[buf]), (uint32_t)(((t >> 1) & 1) ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols);
        for (int ch = 0; ch < p.n_chunks; ++ch, ++g) {
          const int s = g % p.stages;
          mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((g / p.stages) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (g == 0 && mw == 0) RA_DBG(4);  // first MMA can issue
          // Descriptors differ only in the start-address field (bits 0-13, 16-byte units), so the loops
          // below only add small constants.  The m-tile loop is INNERMOST: consecutive MMAs then target
          // different TMEM accumulators and pipeline, instead of forming one chain of dependent accumulations.
          const uint32_t a_hi = smem_u32(stage_base + (size_t)s * p.stage_bytes);
          const uint64_t dA_hi = make_desc(a_hi, plane_bytes, 128);
          const uint64_t dA_lo = dA_hi + (uint64_t)(((uint32_t)in_floats * 4u) >> 4);
          const uint32_t w_addr = p.w_resident ? smem_u32(smem_raw) + (uint32_t)ch * (uint32_t)w_chunk_floats * 4u
                                               : a_hi + 2u * (uint32_t)in_floats * 4u;
          const uint64_t dB = make_desc(w_addr, w_plane, 128);
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_t = (uint32_t)((tap / 3) * p.TWP + (tap % 3));
            for (int k8 = 0; k8 < k8n; ++k8) {
              const uint64_t ao_hi = dA_hi + (uint64_t)(a_t + (uint32_t)(2 * k8) * plane16);
              const uint64_t ao_lo = dA_lo + (uint64_t)(a_t + (uint32_t)(2 * k8) * plane16);
              const uint64_t db_hi = dB + (uint64_t)((uint32_t)(tap * planes + 2 * k8) * wplane16);
              const uint32_t acc_flag = (ch == 0 && tap == 0 && k8 == 0) ? 0u : 1u;
              if (p.merged) {
                // D[:, 0:N] += A_hi B_hi and D[:, N:2N] += A_hi B_lo in ONE instruction ...
                for (int mt = mw; mt < p.n_mt; mt += kMmaWarps)
                  umma_tf32(acc + (uint32_t)(mt * cols_mt), ao_hi + (uint64_t)(mt * 128), db_hi, idesc_2n, acc_flag);
                // ... then D[:, 0:N] += A_lo B_hi
                for (int mt = mw; mt < p.n_mt; mt += kMmaWarps)
                  umma_tf32(acc + (uint32_t)(mt * cols_mt), ao_lo + (uint64_t)(mt * 128), db_hi, idesc_n, 1u);
              } else {
                const uint64_t db_lo = db_hi + (uint64_t)p.NPc;  // rows NPc..2NPc-1 of the plane
                for (int mt = mw; mt < p.n_mt; mt += kMmaWarps)
                  umma_tf32(acc + (uint32_t)(mt * cols_mt), ao_hi + (uint64_t)(mt * 128), db_hi, idesc_n, acc_flag);
                for (int mt = mw; mt < p.n_mt; mt += kMmaWarps)
                  umma_tf32(acc + (uint32_t)(mt * cols_mt), ao_hi + (uint64_t)(mt * 128), db_lo, idesc_n, 1u);
                for (int mt = mw; mt < p.n_mt; mt += kMmaWarps)
                  umma_tf32(acc + (uint32_t)(mt * cols_mt), ao_lo + (uint64_t)(mt * 128), db_hi, idesc_n, 1u);
              }
            }
          }
          umma_commit(smem_u32(&bar_empty[s]));  // the stage may be refilled once these MMAs have read it
        }
        umma_commit(smem_u32(&bar_tfull[buf]));  // accumulators of this tile are complete
      }
      if (mw == 0) RA_DBG(5);  // all MMAs issued
    }
  } else {
    // =============================== epilogue (warps 0-3) ===============================
    const int cols_mt = p.merged ? 2 * p.NPc : p.NPc;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;  // warp w may touch TMEM lanes [32w, 32w+32)
    const int slots_out = p.TH * p.TWP;
    int t = 0;
    for (int tile = tile0; tile < n_tiles; tile += tile_step, ++t) {
      const Item it = decode_tile(p, tile);
      const int buf = t & 1;
      mbar_wait(smem_u32(&bar_tfull[buf]), (uint32_t)((t >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (t == 0 && tid == 0) RA_DBG(6);  // first accumulator complete
      const uint32_t acc = tmem_base + (uint32_t)(buf * p.acc_cols) + lane_sel;
      const int co_base = ns * p.NPc;
      for (int cb = 0; cb < p.NPc; cb += 16) {
        if (co_base + cb >= p.Cout) break;  // padded channel chunks
        float sc[16], sh[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = co_base + cb + j;
          sc[j] = co < p.Cout ? __ldg(p.scale + co) : 0.f;
          sh[j] = co < p.Cout ? __ldg(p.shift + co) : 0.f;
        }
        for (int mt = 0; mt < p.n_mt; ++mt) {
          const int s = mt * 128 + tid;
          float v[16];
          tmem_ld16(acc + (uint32_t)(mt * cols_mt + cb), v);
          if (p.merged) {
            float v2[16];
            tmem_ld16(acc + (uint32_t)(mt * cols_mt + p.NPc + cb), v2);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += v2[j];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = fmaf(v[j], sc[j], sh[j]);
            if (p.relu) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.pool == 2) {
            // horizontal max with the next slot (same image row: TW, x0 are even so pairs do not straddle)
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], __shfl_down_sync(0xffffffffu, v[j], 1));
            if ((lane & 1) == 0 && s < slots_out) {
              float *dst = pool_s + (size_t)(s >> 1) * kPoolLd;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            const int oy_l = s / p.TWP, ox_l = s - oy_l * p.TWP;
            const int oy = it.y0 + oy_l, ox = it.x0 + ox_l;
            if (s < slots_out && ox_l < p.TW && oy < p.Hout && ox < p.Wout) {
              float *dst = p.y + (((size_t)it.b * p.Hout + oy) * p.Wout + ox) * p.Cout + co_base + cb;
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
        if (p.pool == 2) {
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
          // vertical max + store of this 16-channel chunk: pooled pixel (py, px) <- staged half-rows of
          // slots (2py)*TWP+2px and +TWP
          const int ph = p.TH / 2, pw = p.TW / 2;
          const int Ho = p.Hout / 2, Wo = p.Wout / 2;
          for (int idx = tid; idx < ph * pw * 4; idx += kEpiThreads) {
            const int c4 = idx & 3;
            const int pix = idx >> 2;
            const int px = pix % pw, py = pix / pw;
            const int gy = (it.y0 >> 1) + py, gx = (it.x0 >> 1) + px;
            const int co = co_base + cb + c4 * 4;
            if (gy >= Ho || gx >= Wo || co >= p.Cout) continue;
            const int s0 = (2 * py) * p.TWP + 2 * px;
            const float4 va = *reinterpret_cast<const float4 *>(pool_s + (size_t)(s0 >> 1) * kPoolLd + c4 * 4);
            const float4 vc =
                *reinterpret_cast<const float4 *>(pool_s + (size_t)((s0 + p.TWP) >> 1) * kPoolLd + c4 * 4);
            const float4 m = make_float4(fmaxf(va.x, vc.x), fmaxf(va.y, vc.y), fmaxf(va.z, vc.z), fmaxf(va.w, vc.w));
            float *dst = p.y + (((size_t)it.b * Ho + gy) * Wo + gx) * p.Cout + co;
            if ((p.Cout & 3) == 0) {
              *reinterpret_cast<float4 *>(dst) = m;
            } else {
              const float tt[4] = {m.x, m.y, m.z, m.w};
              for (int j = 0; j < 4; ++j)
                if (co + j < p.Cout) dst[j] = tt[j];
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bar_tempty[buf]));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) RA_DBG(7);  // all roles done
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Tile plan shared by the launcher and the weight packer (through ra_conv3x3_umma_plan).
struct Plan {
  int KC, NP, NPc, n_split, merged, TH, TW, TWP, n_mt, slots_alloc, n_chunks, stages, acc_cols, stage_bytes;
  int w_resident, w_res_bytes, grid;
  size_t smem_bytes;
};

// Cost model (cycles per CTA), calibrated with tools/umma_rate.cu and the ncu captures under profiles/:
// a tcgen05.mma with M=128, K=8 costs max(64, N/2) cycles whatever N is (the A operand streams from shared
// memory at 64 B/cycle), and the 224 producer threads stage about 20 B/cycle (latency-bound L2 loads).
int make_plan(int Cin, int Cout, int Hout, int Wout, int pool, int B, Plan *pl) {
  const int NP = round_up(Cout, 16);
  if (NP > 256) return RA_ERR_UNSUPPORTED;
  if ((Wout & 1) || Wout < 2) return RA_ERR_UNSUPPORTED;
  if (pool == 2 && (Hout & 1)) return RA_ERR_UNSUPPORTED;
  if (B < 1) B = 1;
  const size_t smem_cap = 200 * 1024;
  const double kProdBytesPerCycle = 20.0;
  double best = 1e30;
  Plan bp{};
  bool found = false;
  for (int n_split = 1; n_split <= 8; n_split *= 2) {
    if (NP % (16 * n_split) != 0) break;
    const int NPc = NP / n_split;
   for (int merged = 1; merged >= 0; --merged) {
    // merged: [B_hi; B_lo] stacked along N, one MMA (N' = 2 NPc <= 256) reads A_hi once; 2 instead of 3 MMAs
    if (merged && 2 * NPc > 256) continue;
    const int cols_mt = merged ? 2 * NPc : NPc;
    const int mt_max = 256 / cols_mt;  // two accumulator buffers in 512 TMEM columns
    if (mt_max < 1) continue;
    // tensor-core time of one M=128 x K=8 instruction: max(~48, N/2) cycles (tools/umma_rate.cu)
    const double hw_n = NPc / 2 > 48 ? NPc / 2 : 48.0, hw_2n = NPc > 48 ? (double)NPc : 48.0;
    const double hw_cycles = merged ? hw_2n + hw_n : 3.0 * hw_n;
    const double issue_cycles = (merged ? 2.0 : 3.0) * 130.0;  // scalar cost of issuing, per issuing thread
    for (int KC = 8; KC <= 32; KC *= 2) {
      if (KC > 8 && KC / 2 >= Cin) continue;
      const int planes = KC / 4;
      const int n_chunks = (Cin + KC - 1) / KC;
      const size_t w_chunk_bytes = (size_t)9 * planes * 2 * NPc * 16;
      const size_t w_total = w_chunk_bytes * n_chunks;
      for (int TW = Wout; TW >= 2; TW = (TW % 2 == 0 ? TW / 2 : 0)) {
        if (TW & 1) break;
        const int TWP = TW + 2;
        for (int TH = (pool == 2 ? 2 : 1); TH <= Hout; TH += (pool == 2 ? 2 : 1)) {
          const int n_mt = (TH * TWP + 127) / 128;
          if (n_mt > mt_max) break;
          const int slots_alloc = n_mt * 128 + 2 * TWP + 2;
          const size_t in_bytes = (size_t)2 * planes * slots_alloc * 16;
          const size_t pool_bytes = pool == 2 ? (size_t)(n_mt * 64) * kPoolLd * 4 : 0;
          const int tiles = ((Wout + TW - 1) / TW) * ((Hout + TH - 1) / TH) * B;
          int grid_t = ra::kNumSMs / n_split;  // CTAs per channel split
          if (grid_t > tiles) grid_t = tiles;
          if (grid_t < 1) continue;
          const int tiles_per_cta = (tiles + grid_t - 1) / grid_t;
          for (int resident = 1; resident >= 0; --resident) {
            const size_t stage_bytes = in_bytes + (resident ? 0 : w_chunk_bytes);
            const size_t fixed = pool_bytes + (resident ? w_total : 0);
            if (fixed + 2 * stage_bytes > smem_cap) continue;
            int st = (int)((smem_cap - fixed) / stage_bytes);
            if (st > kMaxStages) st = kMaxStages;
            const double per_tap = (double)9 * (KC / 8) * n_chunks;
            const double mma_hw = n_mt * per_tap * hw_cycles;
            const double mma_issue = ((n_mt + kMmaWarps - 1) / kMmaWarps) * per_tap * issue_cycles;
            const double mma_item = mma_hw > mma_issue ? mma_hw : mma_issue;
            const double prod_item = (double)n_chunks * stage_bytes / kProdBytesPerCycle;
            const double epi_item = 400.0 + 60.0 * n_mt * (NPc / 16);
            double item = mma_item > prod_item ? mma_item : prod_item;
            if (epi_item > item) item = epi_item;
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
              bp.TH = TH;
              bp.TW = TW;
              bp.TWP = TWP;
              bp.n_mt = n_mt;
              bp.slots_alloc = slots_alloc;
              bp.n_chunks = n_chunks;
              bp.acc_cols = n_mt * cols_mt;
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
  if (!found) return RA_ERR_UNSUPPORTED;
  *pl = bp;
  return RA_OK;
}

// Calibration hook (tools/bench_conv_layers.py): RA_UMMA_FORCE="KC,TH,TW,n_split,resident" overrides the search.
int make_plan_forced(int Cin, int Cout, int Hout, int Wout, int pool, int B, Plan *pl) {
  const char *f = getenv("RA_UMMA_FORCE");
  if (f == nullptr) return make_plan(Cin, Cout, Hout, Wout, pool, B, pl);
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
  const int cols_mt = bp.merged ? 2 * bp.NPc : bp.NPc;
  bp.TH = TH;
  bp.TW = TW;
  bp.TWP = TW + 2;
  bp.n_mt = (TH * bp.TWP + 127) / 128;
  if (bp.n_mt * cols_mt > 256) return RA_ERR_UNSUPPORTED;
  bp.slots_alloc = bp.n_mt * 128 + 2 * bp.TWP + 2;
  bp.n_chunks = (Cin + KC - 1) / KC;
  bp.acc_cols = bp.n_mt * cols_mt;
  const int planes = KC / 4;
  const size_t in_bytes = (size_t)2 * planes * bp.slots_alloc * 16;
  const size_t w_chunk = (size_t)9 * planes * 2 * bp.NPc * 16;
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

}  // namespace

// Diagnostics: when set, every conv3x3_umma CTA writes 8 clock64() stamps (start, setup done, first stage
// staged, producers done, first MMA issuable, MMAs issued, first accumulator complete, all done).
extern "C" int ra_debug_conv_timeline(long long *device_buf) {
  g_conv_dbg = device_buf;
  return RA_OK;
}

// Plan query for the host-side weight packer.
extern "C" int ra_conv3x3_umma_plan(int Cin, int Cout, int Hout, int Wout, int pool, int B, int *KC, int *NPc,
                                    int *n_split, int *n_chunks) {
  Plan pl;
  const int rc = make_plan_forced(Cin, Cout, Hout, Wout, pool, B, &pl);
  if (rc != RA_OK) return rc;
  if (KC) *KC = pl.KC;
  if (NPc) *NPc = pl.NPc;
  if (n_split) *n_split = pl.n_split;
  if (n_chunks) *n_chunks = pl.n_chunks;
  return RA_OK;
}

// Full plan dump (diagnostics / DESIGN.md tables): info[16] = KC, NPc, n_split, n_chunks, TH, TW, n_mt, stages,
// merged, w_resident, grid, smem_bytes, acc_cols, stage_bytes, w_res_bytes, slots_alloc.
extern "C" int ra_conv3x3_umma_plan_info(int Cin, int Cout, int Hout, int Wout, int pool, int B, int *info) {
  Plan pl;
  const int rc = make_plan_forced(Cin, Cout, Hout, Wout, pool, B, &pl);
  if (rc != RA_OK) return rc;
  if (!info) return RA_ERR_INVALID_ARG;
  const int v[16] = {pl.KC, pl.NPc, pl.n_split, pl.n_chunks, pl.TH, pl.TW, pl.n_mt, pl.stages, pl.merged,
                     pl.w_resident, pl.grid, (int)pl.smem_bytes, pl.acc_cols, pl.stage_bytes, pl.w_res_bytes,
                     pl.slots_alloc};
  for (int i = 0; i < 16; ++i) info[i] = v[i];
  return RA_OK;
}

extern "C" int ra_conv3x3_umma_f32(const float *x1, int C1, const float *x2, int C2, const float *wpack,
                                   const float *scale, const float *shift, int B, int Hin, int Win, int Cout,
                                   int upsample, int pool, int relu, float *y, void *stream) {
  if (!x1 || !wpack || !scale || !shift || !y || C1 < 1 || C2 < 0 || (C2 > 0 && !x2) || B < 0 || Hin < 1 || Win < 1 ||
      Cout < 1)
    return RA_ERR_INVALID_ARG;
  if ((upsample != 1 && upsample != 2) || (pool != 1 && pool != 2)) return RA_ERR_UNSUPPORTED;
  UmmaConvParams p;
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
  if (B == 0) return RA_OK;
  Plan pl;
  const int rc = make_plan_forced(p.Cin, Cout, p.Hout, p.Wout, pool, B, &pl);
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
  p.slots_alloc = pl.slots_alloc;
  p.stages = pl.stages;
  p.w_resident = pl.w_resident;
  p.w_res_bytes = pl.w_res_bytes;
  p.acc_cols = pl.acc_cols;
  p.stage_bytes = pl.stage_bytes;
  p.tiles_x = (p.Wout + p.TW - 1) / p.TW;
  p.tiles_y = (p.Hout + p.TH - 1) / p.TH;
  const long long items = (long long)p.tiles_x * p.tiles_y * B * p.n_split;
  if (items > 0x7fffffffLL) return RA_ERR_UNSUPPORTED;
  p.n_items = (int)items;
  p.vec4 = ((C1 & 3) == 0 && (C2 & 3) == 0) ? 1 : 0;
  p.dbg = g_conv_dbg;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(conv3x3_umma_kernel)", e);
      return RA_ERR_CUDA;
    }
    attr_set = true;
  }
  conv3x3_umma_kernel<<<pl.grid, kThreads, pl.smem_bytes, ra::as_stream(stream)>>>(p);
  return ra::finish_launch("conv3x3_umma_kernel");
}

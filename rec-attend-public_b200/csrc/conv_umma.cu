// 3x3 SAME convolution block on the 5th-generation tensor cores (tcgen05 / TMEM), fp32 in and
// out, 3xTF32 split precision — nnlib.run_cnn / run_dcnn layers (nnlib.py:214-255, :339-402).
//
// Why 3xTF32: a single TF32 pass (10-bit mantissa) moves the attention box by ~1e-3 px and the
// sharp sigmoid(gamma*v-5) edges turn that into 1e-2..1e-1 errors in attn_box / y_out, far
// outside the 1e-3 parity bar (measured with the oracle).  Each operand is split
// v = hi + lo (hi = v with the low 13 mantissa bits cleared, lo = v - hi, exact) and
// D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, which is ~2^-21 accurate.
//
// Implicit GEMM without im2col: the CTA stages a padded input tile as "pixel slots"
// [(TH+2) rows x TWP = TW+2 columns], channel-plane layout [c/4][slot][4 floats] (16 B per
// slot per plane).  For a K-major SWIZZLE_NONE shared-memory descriptor (8 rows x 16 B core
// matrices, SBO = 128 B, LBO = plane stride) a filter tap (ky,kx) is then nothing but a START
// ADDRESS offset of (ky*TWP + kx) slots: the A operand of every tap is the same buffer.
// M = 128 consecutive output slots per MMA (2 of every TWP are halo garbage and are dropped in
// the epilogue), N = Cout padded to 16, K = 8 channels per instruction.
// Per 8/16-channel chunk: all 128 threads stage input (+hi/lo split) and the pre-packed filter
// slice, one thread issues n_mt x 9 taps x 3 passes tcgen05.mma, tcgen05.commit -> mbarrier.
// Epilogue: tcgen05.ld (thread t <-> TMEM lane t <-> slot), folded BN scale/shift, ReLU and
// the 2x2 max-pool (horizontal by shuffle, vertical through a shared staging tile).
#include "common.cuh"

namespace {

struct UmmaConvParams {
  const float *x1;
  const float *x2;
  const float *wpack;  // [chunks][9][2 (hi,lo)][KC/4][NP][4]
  const float *scale;
  const float *shift;
  float *y;
  int C1, C2, Cin, Cout, NP;
  int B, Hin, Win, Hout, Wout;
  int up, pool, relu;
  int TH, TW, TWP, n_mt, KC, n_chunks;
  int slots_alloc;  // input slots allocated per plane
  int tiles_x, tiles_y;
  int tmem_cols;
  int vec4;  // both sources have channel counts divisible by 4
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

constexpr int kThreads = 256;
constexpr int kStageUnroll = 4;

__global__ void __launch_bounds__(kThreads) conv3x3_umma_kernel(UmmaConvParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int planes = p.KC / 4;
  const uint32_t plane_bytes = (uint32_t)p.slots_alloc * 16u;
  // shared layout: in_hi [planes][slots][4] | in_lo [planes][slots][4] | w [9][2][planes][NP][4]
  float *in_hi = reinterpret_cast<float *>(smem_raw);
  float *in_lo = in_hi + (size_t)planes * p.slots_alloc * 4;
  float *w_s = in_lo + (size_t)planes * p.slots_alloc * 4;
  const int w_chunk_floats = 9 * 2 * planes * p.NP * 4;

  int bid = blockIdx.x;
  const int tile_x = bid % p.tiles_x;
  bid /= p.tiles_x;
  const int tile_y = bid % p.tiles_y;
  const int b = bid / p.tiles_y;
  const int y0 = tile_y * p.TH, x0 = tile_x * p.TW;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t mbar_a = smem_u32(&mbar);

  const int slots_in = (p.TH + 2) * p.TWP + 2;  // slots that carry real (or zero-padding) data
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  for (int ch = 0; ch < p.n_chunks; ++ch) {
    const int c0 = ch * p.KC;
    if (ch > 0) mbar_wait(mbar_a, (uint32_t)((ch - 1) & 1));  // previous chunk's MMAs have drained smem

    // ---- stage the input chunk: slot i <-> virtual pixel (y0 - up + i / TWP, x0 - up + i % TWP).
    // kStageUnroll independent loads are issued before any is consumed (the staging is latency-bound).
    const int n_items = p.slots_alloc * planes;
    for (int base = tid; base < n_items; base += kThreads * kStageUnroll) {
      float4 v[kStageUnroll];
      int dst[kStageUnroll];
#pragma unroll
      for (int u = 0; u < kStageUnroll; ++u) {
        const int idx = base + u * kThreads;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        dst[u] = -1;
        if (idx < n_items) {
          const int c4 = idx % planes;  // plane fastest: the threads of one pixel read one contiguous run
          const int slot = idx / planes;
          dst[u] = (c4 * p.slots_alloc + slot) * 4;
          if (slot < slots_in) {
            const int r = slot / p.TWP, col = slot - r * p.TWP;
            const int vy = y0 - p.up + r, vx = x0 - p.up + col;
            bool on = vy >= 0 && vx >= 0 && vy < p.Hout && vx < p.Wout;
            int iy = vy, ix = vx;
            if (p.up == 2) {
              on = on && (((vy | vx) & 1) == 0);
              iy >>= 1;
              ix >>= 1;
            }
            const int cc = c0 + c4 * 4;
            if (on && cc < p.Cin) {
              const size_t pix = ((size_t)b * p.Hin + iy) * p.Win + ix;
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
    // ---- stage the pre-packed filter slice of this chunk (already hi/lo split, plane layout)
    {
      const float4 *src = reinterpret_cast<const float4 *>(p.wpack + (size_t)ch * w_chunk_floats);
      float4 *dst = reinterpret_cast<float4 *>(w_s);
      for (int idx = tid; idx < w_chunk_floats / 4; idx += kThreads) dst[idx] = __ldg(src + idx);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (tensor core)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(in_hi), a_lo = smem_u32(in_lo), w_a = smem_u32(w_s);
      const uint32_t w_plane = (uint32_t)p.NP * 16u;
      for (int mt = 0; mt < p.n_mt; ++mt) {
        const uint32_t d_tmem = tmem_base + (uint32_t)(mt * p.NP);
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a_off = (uint32_t)(mt * 128 + (tap / 3) * p.TWP + (tap % 3)) * 16u;
          for (int k8 = 0; k8 < p.KC / 8; ++k8) {
            const uint32_t a_k = a_off + (uint32_t)(2 * k8) * plane_bytes;
            const uint32_t wb_hi = w_a + (uint32_t)(((tap * 2 + 0) * planes + 2 * k8)) * w_plane;
            const uint32_t wb_lo = w_a + (uint32_t)(((tap * 2 + 1) * planes + 2 * k8)) * w_plane;
            const uint64_t dah = make_desc(a_hi + a_k, plane_bytes, 128);
            const uint64_t dal = make_desc(a_lo + a_k, plane_bytes, 128);
            const uint64_t dbh = make_desc(wb_hi, w_plane, 128);
            const uint64_t dbl = make_desc(wb_lo, w_plane, 128);
            const uint32_t first = (ch == 0 && tap == 0 && k8 == 0) ? 0u : 1u;
            umma_tf32(d_tmem, dah, dbh, idesc, first);
            umma_tf32(d_tmem, dah, dbl, idesc, 1u);
            umma_tf32(d_tmem, dal, dbh, idesc, 1u);
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_a)
                   : "memory");
    }
  }
  mbar_wait(mbar_a, (uint32_t)((p.n_chunks - 1) & 1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue.  Thread t owns TMEM lane t = output slot mt*128 + t.
  const int slots_out = p.TH * p.TWP;
  float *stage = reinterpret_cast<float *>(smem_raw);  // reused: all MMAs have completed
  const int sld = p.NP + 4;                            // staging row stride (floats)
  const int quarter = warp & 3;  // a warp can only touch TMEM lanes [32*(warp%4), +32)
  const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
  for (int mt = warp >> 2; mt < p.n_mt; mt += kThreads / 128) {
    const int s = mt * 128 + quarter * 32 + lane;
    const int oy_l = s / p.TWP, ox_l = s - oy_l * p.TWP;
    const int oy = y0 + oy_l, ox = x0 + ox_l;
    const bool valid = s < slots_out && ox_l < p.TW && oy < p.Hout && ox < p.Wout;
    for (int cb = 0; cb < p.NP; cb += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(lane_base + (uint32_t)(mt * p.NP + cb)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = cb + j;
        float t = __uint_as_float(r[j]);
        if (co < p.Cout) {
          t = fmaf(t, __ldg(p.scale + co), __ldg(p.shift + co));
          if (p.relu) t = fmaxf(t, 0.f);
        } else {
          t = 0.f;
        }
        v[j] = t;
      }
      if (p.pool == 2) {
        // horizontal max with the next slot (same image row: TW, x0 are even so pairs do not straddle)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float o = __shfl_down_sync(0xffffffffu, v[j], 1);
          v[j] = fmaxf(v[j], o);
        }
        if ((lane & 1) == 0 && s < slots_out) {
          float *dst = stage + (size_t)(s >> 1) * sld + cb;
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      } else if (valid) {
        float *dst = p.y + (((size_t)b * p.Hout + oy) * p.Wout + ox) * p.Cout + cb;
        if ((p.Cout & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (cb + j < p.Cout) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cb + j < p.Cout) dst[j] = v[j];
        }
      }
    }
  }
  if (p.pool == 2) {
    __syncthreads();
    // vertical max + store: pooled pixel (py, px) <- staged half-rows of slots (2py)*TWP+2px and +TWP
    const int ph = p.TH / 2, pw = p.TW / 2;
    const int Ho = p.Hout / 2, Wo = p.Wout / 2;
    const int c4n = (p.Cout + 3) / 4;
    for (int idx = tid; idx < ph * pw * c4n; idx += kThreads) {
      const int c4 = idx % c4n;
      const int pix = idx / c4n;
      const int px = pix % pw, py = pix / pw;
      const int gy = (y0 >> 1) + py, gx = (x0 >> 1) + px;
      if (gy >= Ho || gx >= Wo) continue;
      const int s0 = (2 * py) * p.TWP + 2 * px;
      const float *a = stage + (size_t)(s0 >> 1) * sld + c4 * 4;
      const float *c = stage + (size_t)((s0 + p.TWP) >> 1) * sld + c4 * 4;
      const float4 va = *reinterpret_cast<const float4 *>(a);
      const float4 vc = *reinterpret_cast<const float4 *>(c);
      const float4 m = make_float4(fmaxf(va.x, vc.x), fmaxf(va.y, vc.y), fmaxf(va.z, vc.z), fmaxf(va.w, vc.w));
      float *dst = p.y + (((size_t)b * Ho + gy) * Wo + gx) * p.Cout + c4 * 4;
      if ((p.Cout & 3) == 0) {
        *reinterpret_cast<float4 *>(dst) = m;
      } else {
        const float t[4] = {m.x, m.y, m.z, m.w};
        for (int j = 0; j < 4; ++j)
          if (c4 * 4 + j < p.Cout) dst[j] = t[j];
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Tile plan shared by the launcher and the weight packer (through ra_conv3x3_umma_plan).
struct Plan {
  int KC, NP, TH, TW, TWP, n_mt, slots_alloc, tmem_cols, n_chunks;
  size_t smem_bytes;
};

int make_plan(int Cin, int Cout, int Hout, int Wout, int pool, Plan *pl) {
  const int NP = round_up(Cout, 16);
  if (NP > 256) return RA_ERR_UNSUPPORTED;
  if (pool == 2 && ((Hout | Wout) & 1)) return RA_ERR_UNSUPPORTED;
  const int max_cols = 256;          // two CTAs per SM can hold their accumulators
  const size_t smem_cap = 100 * 1024;  // two CTAs per SM
  double best = -1.0;
  Plan bp{};
  for (int KC = 8; KC <= 16; KC += 8) {
    if (KC == 16 && Cin <= 8) continue;
    const int planes = KC / 4;
    const size_t w_bytes = (size_t)9 * 2 * planes * NP * 16;
    for (int TW = Wout; TW >= 2; TW = (TW % 2 == 0 ? TW / 2 : 0)) {
      if (TW & 1) break;
      const int TWP = TW + 2;
      for (int TH = 2; TH <= Hout + 1; TH += 2) {
        const int th = TH > Hout ? Hout : TH;
        if (pool == 2 && (th & 1)) continue;
        const int n_mt = (th * TWP + 127) / 128;
        if (n_mt * NP > max_cols && n_mt > 1) break;
        const int slots_alloc = n_mt * 128 + 2 * TWP + 2;
        const size_t in_bytes = (size_t)2 * planes * slots_alloc * 16;
        size_t smem = in_bytes + w_bytes;
        const size_t stage = pool == 2 ? (size_t)(n_mt * 64) * (NP + 4) * 4 : 0;
        if (stage > smem) smem = stage;
        if (smem > smem_cap) break;
        const double eff = (double)(th * TW) / (double)(n_mt * 128);
        // prefer MMA efficiency; among near-equal plans prefer fewer channel chunks, then bigger tiles
        const double score = eff + (KC == 16 ? 0.08 : 0.0) + 1e-4 * n_mt;
        if (score > best) {
          best = score;
          bp.KC = KC;
          bp.NP = NP;
          bp.TH = th;
          bp.TW = TW;
          bp.TWP = TWP;
          bp.n_mt = n_mt;
          bp.slots_alloc = slots_alloc;
          bp.smem_bytes = smem;
        }
        if (TH >= Hout) break;
      }
    }
  }
  if (best < 0) return RA_ERR_UNSUPPORTED;
  int cols = 32;
  while (cols < bp.n_mt * bp.NP) cols *= 2;
  if (cols > 512) return RA_ERR_UNSUPPORTED;
  bp.tmem_cols = cols;
  bp.n_chunks = (Cin + bp.KC - 1) / bp.KC;
  *pl = bp;
  return RA_OK;
}

}  // namespace

// Plan query for the host-side weight packer: channels per chunk (KC) and padded N.
extern "C" int ra_conv3x3_umma_plan(int Cin, int Cout, int Hout, int Wout, int pool, int *KC, int *NP,
                                    int *n_chunks) {
  Plan pl;
  const int rc = make_plan(Cin, Cout, Hout, Wout, pool, &pl);
  if (rc != RA_OK) return rc;
  if (KC) *KC = pl.KC;
  if (NP) *NP = pl.NP;
  if (n_chunks) *n_chunks = pl.n_chunks;
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
  Plan pl;
  const int rc = make_plan(p.Cin, Cout, p.Hout, p.Wout, pool, &pl);
  if (rc != RA_OK) return rc;
  if (B == 0) return RA_OK;
  p.NP = pl.NP;
  p.TH = pl.TH;
  p.TW = pl.TW;
  p.TWP = pl.TWP;
  p.n_mt = pl.n_mt;
  p.KC = pl.KC;
  p.n_chunks = pl.n_chunks;
  p.slots_alloc = pl.slots_alloc;
  p.tmem_cols = pl.tmem_cols;
  p.tiles_x = (p.Wout + p.TW - 1) / p.TW;
  p.tiles_y = (p.Hout + p.TH - 1) / p.TH;
  p.vec4 = ((C1 & 3) == 0 && (C2 & 3) == 0) ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(conv3x3_umma_kernel)", e);
      return RA_ERR_CUDA;
    }
    attr_set = true;
  }
  const long long nblocks = (long long)p.tiles_x * p.tiles_y * B;
  if (nblocks > 0x7fffffffLL) return RA_ERR_UNSUPPORTED;
  conv3x3_umma_kernel<<<(unsigned)nblocks, kThreads, pl.smem_bytes, ra::as_stream(stream)>>>(p);
  return ra::finish_launch("conv3x3_umma_kernel");
}

// 3x3 SAME convolution block on CUDA cores (fp32, NHWC), fused with the folded
// bias + batch-norm(eval) scale/shift, ReLU and 2x2 max-pool epilogue.
// One kernel serves nnlib.run_cnn (nnlib.py:214-255), nnlib.run_dcnn (nnlib.py:339-402:
// transposed conv = conv over a zero-inserted input with the flipped filter, skip
// concat = second input pointer) and the per-step canvas half of the first controller layer
// (add_to = the step-invariant half, computed once per forward).
//
// Tiling: a CTA owns a TH x TW tile of output pixels and CB output channels.  Each thread
// owns a 2x2 pixel quad (= one max-pool window) x CT output channels, i.e. 4*CT fp32
// accumulators; a warp spans 32 quads with the SAME channel group so the filter reads are
// shared-memory broadcasts and the input reads are conflict-free float2's from a
// channel-planar tile [CK][TH+2][TW+4].  Per input channel a thread issues 8 LDS.64 + 18
// LDS.128 for 288 FMAs.
#include <stdlib.h>

#include "common.cuh"

namespace {

struct ConvParams {
  const float *x1;
  const float *x2;
  const float *w;
  const float *scale;
  const float *shift;
  const float *add_to;
  float *y;
  int C1, C2, Cin, Cout;
  int B, Hin, Win, Hout, Wout;
  int up, pool, relu;
  int THG, TWG, NCG, CK;  // quads per tile (rows, cols), channel groups per CTA, channels per chunk
  int tiles_x, tiles_y, cout_blocks;
};

template <int CT>
__global__ void __launch_bounds__(256) conv3x3_kernel(ConvParams p) {
  extern __shared__ __align__(16) float smem[];
  const int TH = 2 * p.THG, TW = 2 * p.TWG;
  const int TWP = TW + 4;
  const int CB = p.NCG * CT;
  const int in_elems = ((p.CK * (TH + 2) * TWP + 3) / 4) * 4;
  float *in_s = smem;
  float *w_s = smem + in_elems;

  int bid = blockIdx.x;
  const int tile_x = bid % p.tiles_x;
  bid /= p.tiles_x;
  const int tile_y = bid % p.tiles_y;
  bid /= p.tiles_y;
  const int cob = bid % p.cout_blocks;
  const int b = bid / p.cout_blocks;

  const int quads = p.THG * p.TWG;
  const int tid = threadIdx.x;
  const int pg = tid % quads;
  const int cg = tid / quads;
  const int gy = pg / p.TWG, gx = pg % p.TWG;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  const int co0 = cob * CB + cg * CT;  // first output channel of this thread
  const bool co_ok = co0 < p.Cout;

  float acc[4][CT];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[q][j] = 0.f;

  const int nthreads = blockDim.x;
  const int tile_elems = (TH + 2) * (TW + 2) * p.CK;
  const int w_elems = p.CK * 9 * CB;

  for (int c0 = 0; c0 < p.Cin; c0 += p.CK) {
    __syncthreads();
    // ---- stage the input tile (channel fastest in global, channel-planar in shared)
    for (int idx = tid; idx < tile_elems; idx += nthreads) {
      const int c = idx % p.CK;
      const int pix = idx / p.CK;
      const int col = pix % (TW + 2), r = pix / (TW + 2);
      const int vy = oy0 - p.up + r, vx = ox0 - p.up + col;  // virtual (zero-inserted) coordinates
      const int cc = c0 + c;
      float v = 0.f;
      if (cc < p.Cin && vy >= 0 && vx >= 0 && vy < p.Hout && vx < p.Wout) {
        int iy = vy, ix = vx;
        bool on = true;
        if (p.up == 2) {
          on = ((vy | vx) & 1) == 0;
          iy >>= 1;
          ix >>= 1;
        }
        if (on) {
          const size_t pixoff = ((size_t)b * p.Hin + iy) * p.Win + ix;
          v = (cc < p.C1) ? __ldg(p.x1 + pixoff * p.C1 + cc) : __ldg(p.x2 + pixoff * p.C2 + (cc - p.C1));
        }
      }
      in_s[(c * (TH + 2) + r) * TWP + col] = v;
    }
    // ---- stage the filter slice [CK][9][CB]
    for (int idx = tid; idx < w_elems; idx += nthreads) {
      const int cb = idx % CB;
      const int k = (idx / CB) % 9;
      const int c = idx / (CB * 9);
      const int cc = c0 + c, co = cob * CB + cb;
      float v = 0.f;
      if (cc < p.Cin && co < p.Cout) v = __ldg(p.w + ((size_t)k * p.Cin + cc) * p.Cout + co);
      w_s[idx] = v;
    }
    __syncthreads();

    if (cg < p.NCG) {
      const int ck = min(p.CK, p.Cin - c0);
      for (int c = 0; c < ck; ++c) {
        const float *ip = in_s + (c * (TH + 2) + 2 * gy) * TWP + 2 * gx;
        float in[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float2 a = *reinterpret_cast<const float2 *>(ip + r * TWP);
          const float2 bq = *reinterpret_cast<const float2 *>(ip + r * TWP + 2);
          in[r][0] = a.x;
          in[r][1] = a.y;
          in[r][2] = bq.x;
          in[r][3] = bq.y;
        }
        const float *wp = w_s + (c * 9) * CB + cg * CT;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            float wv[CT];
            if (CT % 4 == 0) {
#pragma unroll
              for (int j = 0; j < CT; j += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(wp + (ky * 3 + kx) * CB + j);
                wv[j] = t.x;
                wv[j + 1] = t.y;
                wv[j + 2] = t.z;
                wv[j + 3] = t.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CT; ++j) wv[j] = wp[(ky * 3 + kx) * CB + j];
            }
#pragma unroll
            for (int j = 0; j < CT; ++j) {
              acc[0][j] = fmaf(in[ky][kx], wv[j], acc[0][j]);
              acc[1][j] = fmaf(in[ky][kx + 1], wv[j], acc[1][j]);
              acc[2][j] = fmaf(in[ky + 1][kx], wv[j], acc[2][j]);
              acc[3][j] = fmaf(in[ky + 1][kx + 1], wv[j], acc[3][j]);
            }
          }
      }
    }
  }

  // ---- epilogue: (+add_to) * scale + shift, ReLU, 2x2 max-pool
  if (cg >= p.NCG || !co_ok) return;
  const int oy = oy0 + 2 * gy, ox = ox0 + 2 * gx;
  float sc[CT], sh[CT];
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const bool ok = co0 + j < p.Cout;
    sc[j] = ok ? p.scale[co0 + j] : 0.f;
    sh[j] = ok ? p.shift[co0 + j] : 0.f;
  }
  float pooled[CT];
#pragma unroll
  for (int j = 0; j < CT; ++j) pooled[j] = -INFINITY;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int py = oy + (q >> 1), px = ox + (q & 1);
    if (py >= p.Hout || px >= p.Wout) continue;
    const size_t off = (((size_t)b * p.Hout + py) * p.Wout + px) * p.Cout + co0;
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      if (co0 + j >= p.Cout) continue;
      float v = acc[q][j];
      if (p.add_to != nullptr) v += p.add_to[off + j];
      v = fmaf(v, sc[j], sh[j]);
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.pool == 2)
        pooled[j] = fmaxf(pooled[j], v);
      else
        p.y[off + j] = v;
    }
  }
  if (p.pool == 2 && oy < p.Hout && ox < p.Wout) {
    const size_t off = (((size_t)b * (p.Hout / 2) + (oy >> 1)) * (p.Wout / 2) + (ox >> 1)) * p.Cout + co0;
#pragma unroll
    for (int j = 0; j < CT; ++j)
      if (co0 + j < p.Cout) p.y[off + j] = pooled[j];
  }
}

// First controller layer, per-step half (full_model.py:640-663).  conv(concat(x, canvas, d, y)) is
// linear in its input channels and only the canvas changes between decode steps, so the
// step-invariant channels are convolved once per forward (pre [B,H,W,C0]) and every step adds
// the 1-channel canvas convolution: y = pool(relu((pre + conv(canvas)) * scale + shift)).
// HBM-bound: one read of pre, one of the canvas, one write of the pooled map.  A CTA owns one pooled output
// row segment of one example (grid = segments x pooled rows x B: no 64-bit index arithmetic); the (POOL+2)-row
// canvas strip is staged once in shared memory (zero padded = SAME), a thread owns one pooled pixel x 4
// channels: POOL^2 float4 of pre (streamed, all loads issued up front), 9x4 weights in registers.
constexpr int kCcThreads = 256;
constexpr int kCcMaxCols = kCcThreads / 2 * 2 + 2;  // widest strip: C0 = 8 -> 128 pooled pixels x POOL 2, + halo

template <int POOL>
__global__ void __launch_bounds__(kCcThreads, 4) canvas_conv_kernel(const float *__restrict__ pre,
                                                                    const float *__restrict__ canvas,
                                                                    const float *__restrict__ w,
                                                                    const float *__restrict__ scale,
                                                                    const float *__restrict__ shift, int B, int H,
                                                                    int W, int C0, int relu, float *__restrict__ y) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  __shared__ float cv_s[POOL + 2][kCcMaxCols];
  __shared__ __align__(16) float w_s[9 * 64];  // [tap][C0] (C0 <= 64)
  const int cg_n = C0 >> 2;
  const int PX = kCcThreads / cg_n;  // pooled pixels per CTA
  const int Ho = H / POOL, Wo = W / POOL;
  const int b = blockIdx.z, oy = blockIdx.y;
  const int ox0 = blockIdx.x * PX;
  const int tid = threadIdx.x;
  const int cg = tid % cg_n, oxl = tid / cg_n;
  const int ox = ox0 + oxl;
  const int c = cg * 4;
  const bool live = ox < Wo;

  // the POOL^2 streamed float4 of pre first: they are the long-latency loads
  float4 a[POOL][POOL];
  if (live) {
#pragma unroll
    for (int py = 0; py < POOL; ++py)
#pragma unroll
      for (int px = 0; px < POOL; ++px)
        a[py][px] = __ldcs(reinterpret_cast<const float4 *>(
            pre + (((size_t)b * H + oy * POOL + py) * W + ox * POOL + px) * C0 + c));
  }
  // canvas strip: rows oy*POOL-1 .. oy*POOL+POOL, columns ox0*POOL-1 .. (ox0+PX)*POOL
  const int ncol = PX * POOL + 2;
  const float *cb = canvas + (size_t)b * H * W;
#pragma unroll
  for (int r = 0; r < POOL + 2; ++r) {
    const int yy = oy * POOL - 1 + r;
    for (int q = tid; q < ncol; q += kCcThreads) {
      const int xx = ox0 * POOL - 1 + q;
      cv_s[r][q] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(cb + (size_t)yy * W + xx) : 0.f;
    }
  }
  for (int i = tid; i < 9 * C0; i += kCcThreads) w_s[i] = __ldg(w + i);
  __syncthreads();
  if (!live) return;

  float cv[POOL + 2][POOL + 2];
#pragma unroll
  for (int r = 0; r < POOL + 2; ++r)
#pragma unroll
    for (int q = 0; q < POOL + 2; ++q) cv[r][q] = cv_s[r][oxl * POOL + q];
  // tap-outer accumulation: one 16-byte weight read serves the POOL^2 positions (few live registers -> 4 CTAs / SM)
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4 ww = *reinterpret_cast<const float4 *>(w_s + (ky * 3 + kx) * C0 + c);
#pragma unroll
      for (int py = 0; py < POOL; ++py)
#pragma unroll
        for (int px = 0; px < POOL; ++px) {
          const float v = cv[py + ky][px + kx];
          a[py][px].x = fmaf(v, ww.x, a[py][px].x);
          a[py][px].y = fmaf(v, ww.y, a[py][px].y);
          a[py][px].z = fmaf(v, ww.z, a[py][px].z);
          a[py][px].w = fmaf(v, ww.w, a[py][px].w);
        }
    }
  const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
  float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int py = 0; py < POOL; ++py)
#pragma unroll
    for (int px = 0; px < POOL; ++px) {
      float4 v4 = a[py][px];
      v4.x = fmaf(v4.x, sc.x, sh.x);
      v4.y = fmaf(v4.y, sc.y, sh.y);
      v4.z = fmaf(v4.z, sc.z, sh.z);
      v4.w = fmaf(v4.w, sc.w, sh.w);
      if (relu) {
        v4.x = fmaxf(v4.x, 0.f);
        v4.y = fmaxf(v4.y, 0.f);
        v4.z = fmaxf(v4.z, 0.f);
        v4.w = fmaxf(v4.w, 0.f);
      }
      best.x = fmaxf(best.x, v4.x);
      best.y = fmaxf(best.y, v4.y);
      best.z = fmaxf(best.z, v4.z);
      best.w = fmaxf(best.w, v4.w);
    }
  *reinterpret_cast<float4 *>(y + (((size_t)b * Ho + oy) * Wo + ox) * C0 + c) = best;
}

// out[b,h,w,:] = concat(a[b,h,w,:Ca], bsrc[...,:Cb], c[...,:Cc]) — builds the step-invariant
// input stack of full_model.py:640-661 once per forward.
__global__ void concat_channels_kernel(const float *__restrict__ a, int Ca, const float *__restrict__ bsrc, int Cb,
                                       const float *__restrict__ c, int Cc, size_t npix, float *__restrict__ out) {
  const int Ct = Ca + Cb + Cc;
  const size_t total = npix * Ct;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / Ct;
    const int ch = (int)(i - pix * Ct);
    float v;
    if (ch < Ca)
      v = a[pix * Ca + ch];
    else if (ch < Ca + Cb)
      v = bsrc[pix * Cb + (ch - Ca)];
    else
      v = c[pix * Cc + (ch - Ca - Cb)];
    out[i] = v;
  }
}


// ---- the same layer as a bulk-copy (TMA) pipeline -------------------------------------------------------------
// Persistent CTAs; warp 4 streams the `pre` rows of successive (example, pooled row, segment) items into a 3-stage
// shared-memory ring with cp.async.bulk (each row segment is one contiguous run of NHWC memory) and mbarrier
// transaction counts; warps 0-3 consume: canvas strip (prefetched one item ahead with cp.async), 9 taps, folded BN,
// max-pool, ReLU, two float4 stores per thread.
constexpr int kCbStages = 3;  // x 3 CTAs/SM x 16 KB = 144 KB of loads in flight per SM
constexpr int kCbConsumers = 128;              // consumer threads (warps 0-3); warp 4 is the producer
constexpr int kCbThreads = kCbConsumers + 32;
constexpr int kCbStripW = 4 + 256 + 4;         // strip row: 3 pad + left halo | up to 256 interior columns | right halo + pad

__device__ __forceinline__ uint32_t cb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cb_mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void cb_mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void cb_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cb_mbar_wait(uint32_t mbar, uint32_t parity) {
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
__device__ __forceinline__ void cb_bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}

template <int POOL>
__global__ void __launch_bounds__(kCbThreads) canvas_conv_bulk_kernel(const float *__restrict__ pre,
                                                                      const float *__restrict__ canvas,
                                                                      const float *__restrict__ w,
                                                                      const float *__restrict__ scale,
                                                                      const float *__restrict__ shift, int B, int H,
                                                                      int W, int C0, int relu, float *__restrict__ y,
                                                                      int nseg, int n_items) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  // ncu (profiles/r01j_*): the first version of this kernel was ISSUE-bound (57 % issue-active at 37 % of the DRAM
  // peak, ~540 instructions per thread and item for 144 FMAs).  Hence: two pooled pixels per thread (the canvas
  // patch, the weight reads, the item bookkeeping and the barriers are shared by twice the arithmetic), the strip
  // fetched with 16-byte cp.async, ReLU after the max.
  constexpr int NP = 2;                    // pooled pixels per thread (horizontally adjacent)
  constexpr int PW = NP * POOL;            // input columns per thread
  extern __shared__ __align__(128) unsigned char cb_smem[];
  __shared__ __align__(16) float cv_s[2][POOL + 2][kCbStripW];  // [buffer][row][3 pad | left halo | interior | right halo]
  __shared__ __align__(16) float w_s[9 * 64];
  __shared__ __align__(8) uint64_t bar_full[kCbStages], bar_empty[kCbStages];
  const int cg_n = C0 >> 2;
  const int PXT = kCbConsumers / cg_n;  // thread columns per item
  const int PX = PXT * NP;              // pooled pixels per item
  const int PXI = PX * POOL;            // input pixels per item row
  const int Ho = H / POOL, Wo = W / POOL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t row_bytes = (uint32_t)PXI * (uint32_t)C0 * 4u;  // one full row segment in a stage
  const uint32_t stage_bytes = row_bytes * POOL;
  float *stage0 = reinterpret_cast<float *>(cb_smem + ((128u - (cb_smem_u32(cb_smem) & 127u)) & 127u));

  if (tid == 0) {
    for (int s = 0; s < kCbStages; ++s) {
      cb_mbar_init(cb_smem_u32(&bar_full[s]), 1);
      cb_mbar_init(cb_smem_u32(&bar_empty[s]), kCbConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 9 * C0; i += kCbThreads) w_s[i] = __ldg(w + i);
  __syncthreads();

  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == kCbConsumers / 32) {
    // =============================== producer warp ===============================
    if (lane == 0) {
      for (int k = 0; k < my_items; ++k) {
        const int s = k % kCbStages;
        if (k >= kCbStages) cb_mbar_wait(cb_smem_u32(&bar_empty[s]), (uint32_t)(((k / kCbStages) - 1) & 1));
        int it = (int)blockIdx.x + k * (int)gridDim.x;
        const int seg = it % nseg;
        it /= nseg;
        const int oy = it % Ho, b = it / Ho;
        const int x0 = seg * PXI;
        const int npx = min(PXI, W - x0);
        const uint32_t bytes = (uint32_t)npx * (uint32_t)C0 * 4u;
        const uint32_t bar = cb_smem_u32(&bar_full[s]);
        cb_mbar_expect_tx(bar, bytes * POOL);
        const uint32_t dst = cb_smem_u32(stage0) + (uint32_t)s * stage_bytes;
#pragma unroll
        for (int py = 0; py < POOL; ++py)
          cb_bulk_load(dst + (uint32_t)py * row_bytes, pre + (((size_t)b * H + oy * POOL + py) * W + x0) * C0, bytes, bar);
      }
    }
    return;
  }

  // =============================== consumers ===============================
  const int cg = tid % cg_n, oxp = tid / cg_n;
  const int c = cg * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale + c));
  const float4 sh = __ldg(reinterpret_cast<const float4 *>(shift + c));
  const int n16 = PXI / 4;  // 16-byte pieces of a strip row's interior

  // canvas strip of the item (seg, oy, b) -> cv_s[buf]: interior columns by 16-byte cp.async (W % 4 == 0, so a piece
  // is entirely inside or outside the image), the two halo columns by 4-byte cp.async, zeros outside (SAME padding)
  auto strip_prefetch = [&](int buf, int seg, int oy, int b, bool valid) {
    if (valid) {
      const float *cb = canvas + (size_t)b * H * W;
      const int x0 = seg * PXI;
      for (int idx = tid; idx < (POOL + 2) * n16; idx += kCbConsumers) {
        const int r = idx / n16, i4 = idx - r * n16;
        const int yy = oy * POOL - 1 + r, xx = x0 + i4 * 4;
        float *dst = &cv_s[buf][r][4 + i4 * 4];
        if (yy >= 0 && yy < H && xx < W) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cb_smem_u32(dst)), "l"(cb + (size_t)yy * W + xx)
                       : "memory");
        } else {
          *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (tid < (POOL + 2) * 2) {
        const int r = tid >> 1, right = tid & 1;
        const int yy = oy * POOL - 1 + r, xx = right ? x0 + PXI : x0 - 1;
        float *dst = &cv_s[buf][r][right ? 4 + PXI : 3];
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(cb_smem_u32(dst)), "l"(cb + (size_t)yy * W + xx)
                       : "memory");
        } else {
          *dst = 0.f;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // item k of this CTA -> (seg, oy, b); decoded once, one item ahead
  int seg_n, oy_n, b_n;
  {
    int it = (int)blockIdx.x;
    seg_n = it % nseg;
    it /= nseg;
    oy_n = it % Ho;
    b_n = it / Ho;
  }
  strip_prefetch(0, seg_n, oy_n, b_n, my_items > 0);
  for (int k = 0; k < my_items; ++k) {
    const int s = k % kCbStages;
    const int seg = seg_n, oy = oy_n, b = b_n;
    {
      int it = (int)blockIdx.x + (k + 1) * (int)gridDim.x;
      seg_n = it % nseg;
      it /= nseg;
      oy_n = it % Ho;
      b_n = it / Ho;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");            // strip k has landed (this thread's part)
    asm volatile("bar.sync 1, %0;" ::"n"(kCbConsumers) : "memory");  // ... and everybody else's; item k-1 is fully read
    strip_prefetch((k + 1) & 1, seg_n, oy_n, b_n, k + 1 < my_items);  // overlaps this item's arithmetic
    cb_mbar_wait(cb_smem_u32(&bar_full[s]), (uint32_t)((k / kCbStages) & 1));
    const float *st = stage0 + (size_t)s * (stage_bytes / 4);
    const int ox = seg * PX + oxp * NP;  // first of this thread's pooled pixels
    if (ox < Wo) {                       // (Wo is even whenever PX is, so both pixels are in or out together)
      float4 a[POOL][PW];
#pragma unroll
      for (int py = 0; py < POOL; ++py)
#pragma unroll
        for (int px = 0; px < PW; ++px)
          a[py][px] = *reinterpret_cast<const float4 *>(st + ((size_t)py * PXI + oxp * PW + px) * C0 + c);
      float cv[POOL + 2][PW + 2];
#pragma unroll
      for (int r = 0; r < POOL + 2; ++r)
#pragma unroll
        for (int q = 0; q < PW + 2; ++q) cv[r][q] = cv_s[k & 1][r][3 + oxp * PW + q];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 ww = *reinterpret_cast<const float4 *>(w_s + (ky * 3 + kx) * C0 + c);
#pragma unroll
          for (int py = 0; py < POOL; ++py)
#pragma unroll
            for (int px = 0; px < PW; ++px) {
              const float v = cv[py + ky][px + kx];
              a[py][px].x = fmaf(v, ww.x, a[py][px].x);
              a[py][px].y = fmaf(v, ww.y, a[py][px].y);
              a[py][px].z = fmaf(v, ww.z, a[py][px].z);
              a[py][px].w = fmaf(v, ww.w, a[py][px].w);
            }
        }
#pragma unroll
      for (int np = 0; np < NP; ++np) {
        if (ox + np >= Wo) break;
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int py = 0; py < POOL; ++py)
#pragma unroll
          for (int px = 0; px < POOL; ++px) {
            const float4 v4 = a[py][np * POOL + px];
            best.x = fmaxf(best.x, fmaf(v4.x, sc.x, sh.x));
            best.y = fmaxf(best.y, fmaf(v4.y, sc.y, sh.y));
            best.z = fmaxf(best.z, fmaf(v4.z, sc.z, sh.z));
            best.w = fmaxf(best.w, fmaf(v4.w, sc.w, sh.w));
          }
        if (relu) {  // max(relu(.)) == relu(max(.))
          best.x = fmaxf(best.x, 0.f);
          best.y = fmaxf(best.y, 0.f);
          best.z = fmaxf(best.z, 0.f);
          best.w = fmaxf(best.w, 0.f);
        }
        *reinterpret_cast<float4 *>(y + (((size_t)b * Ho + oy) * Wo + ox + np) * C0 + c) = best;
      }
    }
    __syncwarp();
    if (lane == 0) cb_mbar_arrive(cb_smem_u32(&bar_empty[s]));  // this warp has read stage s
  }
}

}  // namespace

extern "C" int ra_conv3x3_f32(const float *x1, int C1, const float *x2, int C2, const float *w, const float *scale,
                              const float *shift, const float *add_to, int B, int Hin, int Win, int Cout, int upsample,
                              int pool, int relu, float *y, void *stream) {
  if (!x1 || !w || !scale || !shift || !y || C1 < 1 || C2 < 0 || (C2 > 0 && !x2) || B < 0 || Hin < 1 || Win < 1 ||
      Cout < 1)
    return RA_ERR_INVALID_ARG;
  if ((upsample != 1 && upsample != 2) || (pool != 1 && pool != 2)) return RA_ERR_UNSUPPORTED;
  ConvParams p;
  p.x1 = x1;
  p.x2 = x2;
  p.w = w;
  p.scale = scale;
  p.shift = shift;
  p.add_to = add_to;
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
  if (pool == 2 && ((p.Hout | p.Wout) & 1)) return RA_ERR_UNSUPPORTED;  // even sizes only
  if (B == 0) return RA_OK;

  const bool ct8 = (Cout % 8) == 0;
  const int CT = ct8 ? 8 : 1;
  const int groups = (Cout + CT - 1) / CT;
  if (p.Hout >= 64 && p.Wout >= 64) {
    p.THG = 8;
    p.TWG = 16;
  } else if (p.Wout >= 32) {
    p.THG = 4;
    p.TWG = 16;
  } else {
    p.THG = 4;
    p.TWG = 8;
  }
  const int quads = p.THG * p.TWG;
  p.NCG = 256 / quads;
  if (p.NCG > groups) p.NCG = groups;
  p.CK = p.Cin < 8 ? p.Cin : 8;
  const int TH = 2 * p.THG, TW = 2 * p.TWG, CB = p.NCG * CT;
  p.tiles_x = (p.Wout + TW - 1) / TW;
  p.tiles_y = (p.Hout + TH - 1) / TH;
  p.cout_blocks = (Cout + CB - 1) / CB;
  const size_t in_elems = ((size_t)(p.CK * (TH + 2) * (TW + 4) + 3) / 4) * 4;
  const size_t smem = (in_elems + (size_t)p.CK * 9 * CB) * sizeof(float);
  const long long nblocks = (long long)p.tiles_x * p.tiles_y * p.cout_blocks * B;
  if (nblocks > 0x7fffffffLL) return RA_ERR_UNSUPPORTED;
  const int threads = quads * p.NCG;
  cudaStream_t s = ra::as_stream(stream);
  if (ct8)
    conv3x3_kernel<8><<<(unsigned)nblocks, threads, smem, s>>>(p);
  else
    conv3x3_kernel<1><<<(unsigned)nblocks, threads, smem, s>>>(p);
  return ra::finish_launch("conv3x3_kernel");
}

extern "C" int ra_concat_channels_f32(const float *a, int Ca, const float *b, int Cb, const float *c, int Cc,
                                      size_t npix, float *out, void *stream) {
  if (!a || Ca < 1 || Cb < 0 || Cc < 0 || (Cb > 0 && !b) || (Cc > 0 && !c) || !out) return RA_ERR_INVALID_ARG;
  if (npix == 0) return RA_OK;
  const size_t total = npix * (size_t)(Ca + Cb + Cc);
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ra::kNumSMs * 16) blocks = (size_t)ra::kNumSMs * 16;
  concat_channels_kernel<<<(unsigned)blocks, 256, 0, ra::as_stream(stream)>>>(a, Ca, b, Cb, c, Cc, npix, out);
  return ra::finish_launch("concat_channels_kernel");
}

extern "C" int ra_canvas_conv_f32(const float *pre, const float *canvas, const float *w, const float *scale,
                                  const float *shift, int B, int H, int W, int C0, int pool, int relu, float *y,
                                  void *stream) {
  if (!pre || !canvas || !w || !scale || !shift || !y || B < 0 || H < 1 || W < 1 || C0 < 1) return RA_ERR_INVALID_ARG;
  if ((C0 & 3) != 0 || (pool != 1 && pool != 2)) return RA_ERR_UNSUPPORTED;
  if (pool == 2 && ((H | W) & 1)) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  // a CTA serves kCcThreads / (C0/4) pooled pixels of one pooled row; C0/4 must divide the CTA and the strip
  // must fit the static shared tile (C0 >= 8)
  const int cg_n = C0 / 4;
  if (C0 < 8 || C0 > 64 || (kCcThreads % cg_n) != 0) return RA_ERR_UNSUPPORTED;
  const int PX = kCcThreads / cg_n;
  const int Ho = H / pool, Wo = W / pool;
  if (Ho > 65535 || B > 65535) return RA_ERR_UNSUPPORTED;
  dim3 grid((Wo + PX - 1) / PX, Ho, B);
  cudaStream_t s = ra::as_stream(stream);
  // bulk-copy pipeline: needs 16-byte aligned row segments and strip pieces (C0 % 4 == 0 and W % 4 == 0 give the
  // sizes; the bases must be aligned)
  const bool no_bulk = getenv("RA_CANVAS_NO_BULK") != nullptr;  // diagnostics / tests: the plain-load kernel
  const int PXb = (kCbConsumers / cg_n) * 2;                    // pooled pixels per item of the bulk kernel
  const int nseg = (Wo + PXb - 1) / PXb;
  const long long n_items = (long long)nseg * Ho * B;
  if (!no_bulk && (W & 3) == 0 && PXb * pool <= 256 && (reinterpret_cast<uintptr_t>(pre) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(canvas) & 15) == 0 && n_items <= 0x7fffffffLL) {
    const size_t stage_bytes = (size_t)pool * (PXb * pool) * C0 * 4;
    const size_t smem = kCbStages * stage_bytes + 128;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e1 = cudaFuncSetAttribute(canvas_conv_bulk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaError_t e2 = cudaFuncSetAttribute(canvas_conv_bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      if (e1 != cudaSuccess || e2 != cudaSuccess) {
        ra::set_last_error("cudaFuncSetAttribute(canvas_conv_bulk_kernel)", e1 != cudaSuccess ? e1 : e2);
        return RA_ERR_CUDA;
      }
      attr_set = true;
    }
    int ctas = ra::kNumSMs * 3;  // 61 KB of shared memory each: three per SM
    if ((long long)ctas > n_items) ctas = (int)n_items;
    const cudaError_t le =
        pool == 2 ? ra::launch_pdl(canvas_conv_bulk_kernel<2>, dim3(ctas), dim3(kCbThreads), smem, s, pre, canvas, w, scale,
                                   shift, B, H, W, C0, relu, y, nseg, (int)n_items)
                  : ra::launch_pdl(canvas_conv_bulk_kernel<1>, dim3(ctas), dim3(kCbThreads), smem, s, pre, canvas, w, scale,
                                   shift, B, H, W, C0, relu, y, nseg, (int)n_items);
    if (le != cudaSuccess) {
      ra::set_last_error("cudaLaunchKernelEx(canvas_conv_bulk_kernel)", le);
      return RA_ERR_CUDA;
    }
    return ra::finish_launch("canvas_conv_bulk_kernel");
  }
  const cudaError_t le =
      pool == 2 ? ra::launch_pdl(canvas_conv_kernel<2>, grid, dim3(kCcThreads), (size_t)0, s, pre, canvas, w, scale, shift, B,
                                 H, W, C0, relu, y)
                : ra::launch_pdl(canvas_conv_kernel<1>, grid, dim3(kCcThreads), (size_t)0, s, pre, canvas, w, scale, shift, B,
                                 H, W, C0, relu, y);
  if (le != cudaSuccess) {
    ra::set_last_error("cudaLaunchKernelEx(canvas_conv_kernel)", le);
    return RA_ERR_CUDA;
  }
  return ra::finish_launch("canvas_conv_kernel");
}

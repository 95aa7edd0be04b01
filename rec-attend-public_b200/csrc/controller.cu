// Controller step: soft-attention read-out x LSTM x glimpse-MLP softmax, repeated n_iter
// times, then the controller head and the box-parameter maths — full_model.py:668-725,
// box_model.py:417-470, nnlib.py:476-493 (mlp), nnlib.py:637-649 (lstm), modellib.py:752-856.
//
// One CTA per example, Hd (=256) threads, the whole inner loop in one launch (the reference
// issues ~35 TF ops per glimpse iteration).  The feature map [P,Cf] is staged once in shared
// memory and re-read 5 times; thread j owns hidden unit j: its four gate pre-activations are
// dot products over the 64+256 inputs with weight reads coalesced across j (weights live in
// L2: 1.7 MB shared by every CTA).  The read-out and the softmax are warp/CTA reductions.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

struct CtrlParams {
  const float *feat;
  const float *wx, *wh, *bg;            // [4][Cf][Hd], [4][Hd][Hd], [4][Hd]; gate order i,f,o,u
  const float *gw0, *gb0, *gw1, *gb1;   // glimpse MLP
  const float *cw, *cb;                 // controller head [Hd,9], [9]
  float *h_out, *ctrl_out, *gmap_out, *box_out;
  int P, Cf, Hd, n_iter;
  int inp_h, inp_w, filt_h, filt_w, flags;
};

// Controller head -> box record (RA_BOX_* layout): full_model.py:691-725, modellib.py:752-856.
__device__ void write_box(const CtrlParams &p, const float *out_s, float *bo) {
  // full_model.py:691-725 and modellib.py:752-856
  float cn[2] = {out_s[0], out_s[1]};
  float ls[2] = {out_s[2], out_s[3]};
  if (p.flags & RA_CTRL_SQUASH) {
    for (int d = 0; d < 2; ++d) {
      cn[d] = tanhf(cn[d]);
      // -softplus(x) = -log(1 + exp(x)), computed the numerically safe way
      const float x = ls[d];
      ls[d] = -(fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))));
    }
  }
  const float img[2] = {(float)p.inp_h, (float)p.inp_w};
  const float filt[2] = {(float)p.filt_h, (float)p.filt_w};
  for (int d = 0; d < 2; ++d) {
    const float ctr = (cn[d] + 1.0f) * (img[d] / 2.0f);
    const float size = expf(ls[d]) * img[d];
    float lgv;
    if (p.flags & RA_CTRL_FIXED_VAR)
      lgv = 0.f;
    else
      lgv = logf(size) - logf(filt[d]);
    if (p.flags & RA_CTRL_DYNAMIC_VAR) lgv = out_s[4 + d];
    bo[RA_BOX_CTR_Y + d] = ctr;
    bo[RA_BOX_SIZE_Y + d] = size;
    bo[RA_BOX_LGVAR_Y + d] = lgv;
    bo[RA_BOX_TL_Y + d] = ctr - size / 2.0f;  // modellib.py:850-852
    bo[RA_BOX_BR_Y + d] = ctr + size / 2.0f;
  }
  const bool fg = (p.flags & RA_CTRL_FIXED_GAMMA) != 0;
  bo[RA_BOX_GAMMA_ATTN] = fg ? 1.0f : expf(out_s[6]);
  bo[RA_BOX_GAMMA_BOX] = expf(out_s[7]);
  bo[RA_BOX_GAMMA_Y] = fg ? expf(2.0f) : expf(out_s[8]);
  bo[13] = bo[14] = bo[15] = 0.f;
}

__global__ void __launch_bounds__(256) controller_kernel(CtrlParams p) {
  extern __shared__ __align__(16) float smem[];
  const int P = p.P, Cf = p.Cf, Hd = p.Hd;
  float *feat_s = smem;                 // [P][Cf]
  float *map_s = feat_s + (size_t)P * Cf;  // [P]
  float *x_s = map_s + P;               // [Cf]  glimpse
  float *h_s = x_s + Cf;                // [Hd]
  float *t_s = h_s + Hd;                // [Hd]  hidden layer of the glimpse MLP
  float *part_s = t_s + Hd;             // [4][Cf] read-out partials
  float *red_s = part_s + 4 * Cf;       // [32]
  float *out_s = red_s + 32;            // [16] ctrl_out

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int nt = blockDim.x;  // == Hd

  {
    const float *src = p.feat + (size_t)b * P * Cf;
    for (int i = tid; i < P * Cf; i += nt) feat_s[i] = src[i];
    const float u = 1.0f / (float)P;  // full_model.py:676-677
    for (int i = tid; i < P; i += nt) map_s[i] = u;
  }
  float c_state = 0.f;  // full_model.py:674: the LSTM state is reset at every decode step
  h_s[tid] = 0.f;
  __syncthreads();

  for (int it = 0; it < p.n_iter; ++it) {
    // glimpse map of this iteration -> output [B, n_iter, P]
    for (int i = tid; i < P; i += nt) p.gmap_out[((size_t)b * p.n_iter + it) * P + i] = map_s[i];

    // ---- read-out: x[c] = sum_p feat[p][c] * map[p]   (full_model.py:680)
    {
      const int slices = nt / Cf;  // 4 for 256/64
      const int c = tid % Cf, sl = tid / Cf;
      if (sl < slices && sl < 4) {
        float a = 0.f;
        for (int q = sl; q < P; q += min(slices, 4)) a = fmaf(feat_s[q * Cf + c], map_s[q], a);
        part_s[sl * Cf + c] = a;
      }
      __syncthreads();
      if (tid < Cf) {
        float a = 0.f;
        for (int sl2 = 0; sl2 < min(slices, 4); ++sl2) a += part_s[sl2 * Cf + tid];
        x_s[tid] = a;
      }
      __syncthreads();
    }

    // ---- LSTM gates for hidden unit j = tid (nnlib.py:641-647)
    {
      const int j = tid;
      float g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float *wxq = p.wx + (size_t)q * Cf * Hd + j;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = 0;
        for (; k + 4 <= Cf; k += 4) {
          a0 = fmaf(x_s[k], __ldg(wxq + (size_t)k * Hd), a0);
          a1 = fmaf(x_s[k + 1], __ldg(wxq + (size_t)(k + 1) * Hd), a1);
          a2 = fmaf(x_s[k + 2], __ldg(wxq + (size_t)(k + 2) * Hd), a2);
          a3 = fmaf(x_s[k + 3], __ldg(wxq + (size_t)(k + 3) * Hd), a3);
        }
        for (; k < Cf; ++k) a0 = fmaf(x_s[k], __ldg(wxq + (size_t)k * Hd), a0);
        const float *whq = p.wh + (size_t)q * Hd * Hd + j;
        float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        for (k = 0; k + 4 <= Hd; k += 4) {
          b0 = fmaf(h_s[k], __ldg(whq + (size_t)k * Hd), b0);
          b1 = fmaf(h_s[k + 1], __ldg(whq + (size_t)(k + 1) * Hd), b1);
          b2 = fmaf(h_s[k + 2], __ldg(whq + (size_t)(k + 2) * Hd), b2);
          b3 = fmaf(h_s[k + 3], __ldg(whq + (size_t)(k + 3) * Hd), b3);
        }
        for (; k < Hd; ++k) b0 = fmaf(h_s[k], __ldg(whq + (size_t)k * Hd), b0);
        g[q] = ((a0 + a1) + (a2 + a3)) + ((b0 + b1) + (b2 + b3)) + __ldg(p.bg + q * Hd + j);
      }
      const float gi = ra::sigmoidf_acc(g[0]);
      const float gf = ra::sigmoidf_acc(g[1]);
      const float go = ra::sigmoidf_acc(g[2]);
      const float u = tanhf(g[3]);
      c_state = gf * c_state + gi * u;
      const float hn = go * tanhf(c_state);
      __syncthreads();  // everyone has finished reading the old h
      h_s[j] = hn;
      __syncthreads();
    }

    if (it == p.n_iter - 1) break;  // the 5th glimpse map is dead compute (full_model.py:686-688)

    // ---- glimpse MLP layer 0: relu(h W0 + b0)
    {
      const int j = tid;
      const float *wq = p.gw0 + j;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 4 <= Hd; k += 4) {
        a0 = fmaf(h_s[k], __ldg(wq + (size_t)k * Hd), a0);
        a1 = fmaf(h_s[k + 1], __ldg(wq + (size_t)(k + 1) * Hd), a1);
        a2 = fmaf(h_s[k + 2], __ldg(wq + (size_t)(k + 2) * Hd), a2);
        a3 = fmaf(h_s[k + 3], __ldg(wq + (size_t)(k + 3) * Hd), a3);
      }
      for (; k < Hd; ++k) a0 = fmaf(h_s[k], __ldg(wq + (size_t)k * Hd), a0);
      t_s[j] = fmaxf(((a0 + a1) + (a2 + a3)) + __ldg(p.gb0 + j), 0.f);
      __syncthreads();
    }
    // ---- glimpse MLP layer 1 + softmax over the P map positions (full_model.py:350-352)
    {
      float lmax = -INFINITY;
      for (int q = tid; q < P; q += nt) {
        const float *wq = p.gw1 + q;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = 0;
        for (; k + 4 <= Hd; k += 4) {
          a0 = fmaf(t_s[k], __ldg(wq + (size_t)k * P), a0);
          a1 = fmaf(t_s[k + 1], __ldg(wq + (size_t)(k + 1) * P), a1);
          a2 = fmaf(t_s[k + 2], __ldg(wq + (size_t)(k + 2) * P), a2);
          a3 = fmaf(t_s[k + 3], __ldg(wq + (size_t)(k + 3) * P), a3);
        }
        for (; k < Hd; ++k) a0 = fmaf(t_s[k], __ldg(wq + (size_t)k * P), a0);
        const float v = ((a0 + a1) + (a2 + a3)) + __ldg(p.gb1 + q);
        map_s[q] = v;
        lmax = fmaxf(lmax, v);
      }
      lmax = ra::block_max(lmax, red_s);
      float lsum = 0.f;
      for (int q = tid; q < P; q += nt) {
        const float e = expf(map_s[q] - lmax);
        map_s[q] = e;
        lsum += e;
      }
      lsum = ra::block_sum(lsum, red_s);
      for (int q = tid; q < P; q += nt) map_s[q] = map_s[q] / lsum;
      __syncthreads();
    }
  }

  // ---- controller head: ctrl_out = h Wc + bc  (one warp per output)
  {
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    for (int o = wid; o < 9; o += nw) {
      float a = 0.f;
      for (int k = lane; k < Hd; k += 32) a = fmaf(h_s[k], __ldg(p.cw + (size_t)k * 9 + o), a);
      a = ra::warp_sum(a);
      if (lane == 0) out_s[o] = a + __ldg(p.cb + o);
    }
    p.h_out[(size_t)b * Hd + tid] = h_s[tid];
    __syncthreads();
  }

  if (tid < 9) p.ctrl_out[(size_t)b * 9 + tid] = out_s[tid];
  if (tid == 0) write_box(p, out_s, p.box_out + (size_t)b * RA_BOX_STRIDE);
}

// ------------------------------------------------------------------------------------------
// Cluster version: 8 CTAs (one thread-block cluster) serve 8 examples.  CTA r keeps the LSTM
// gate weights of hidden units [32r, 32r+32) (160 KB) and its 32 columns of the first glimpse-MLP
// layer (32 KB) in shared memory for the whole step, so the 1.7 MB of controller weights are
// read from L2 once per cluster instead of once per example and glimpse iteration; hidden
// state, MLP activations and logits are exchanged through distributed shared memory.
// CTA r also owns example r for the per-example work (read-out, softmax, head).
// ------------------------------------------------------------------------------------------
constexpr int kCl = 8;
constexpr int kHd2 = 256;
constexpr int kCf2 = 64;
constexpr int kUnits = kHd2 / kCl;     // 32 hidden units per CTA
constexpr int kKin = kCf2 + kHd2;      // 320 LSTM inputs

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(256) controller_cluster_kernel(CtrlParams p, int B) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int e_glob = (blockIdx.x / kCl) * kCl + r;  // the example this CTA owns
  const bool own = e_glob < B;
  const int P = p.P;
  const int PS = (P + kCl - 1) / kCl;  // glimpse-map positions per CTA
  const int tid = threadIdx.x;

  float *wg_s = smem;                          // [320][128]  (k, q*32+u)
  float *w0_s = wg_s + kKin * 4 * kUnits;      // [256][32]
  float *x_s = w0_s + kHd2 * kUnits;           // [64][8]
  float *h_s = x_s + kCf2 * kCl;               // [2][256][8]
  float *t_s = h_s + 2 * kHd2 * kCl;           // [256][8]  (also the gate partial-sum buffer)
  float *lg_s = t_s + kHd2 * kCl;              // [PS*8] logits / glimpse map of my example
  float *bg_s = lg_s + PS * kCl;               // [128]
  float *b0_s = bg_s + 4 * kUnits;             // [32]
  float *part_s = b0_s + kUnits;               // [256]
  float *red_s = part_s + 256;                 // [32]
  float *out_s = red_s + 32;                   // [16]

  // ---- one-time loads: my slice of the weights (float4 along the hidden-unit index, 8 loads in flight)
  {
    constexpr int kU = 8;
    const int n4 = kKin * 4 * kUnits / 4;  // float4 items of wg_s
    for (int base = tid; base < n4; base += 256 * kU) {
      float4 v[kU];
#pragma unroll
      for (int u8 = 0; u8 < kU; ++u8) {
        const int i4 = base + u8 * 256;
        v[u8] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i4 < n4) {
          const int col = (i4 * 4) % (4 * kUnits), k = (i4 * 4) / (4 * kUnits);
          const int q = col / kUnits, u = col % kUnits;
          const int j = r * kUnits + u;
          v[u8] = (k < kCf2) ? __ldg(reinterpret_cast<const float4 *>(p.wx + ((size_t)q * kCf2 + k) * kHd2 + j))
                             : __ldg(reinterpret_cast<const float4 *>(p.wh + ((size_t)q * kHd2 + (k - kCf2)) * kHd2 + j));
        }
      }
#pragma unroll
      for (int u8 = 0; u8 < kU; ++u8) {
        const int i4 = base + u8 * 256;
        if (i4 < n4) *reinterpret_cast<float4 *>(wg_s + i4 * 4) = v[u8];
      }
    }
    const int m4 = kHd2 * kUnits / 4;  // float4 items of w0_s
    for (int base = tid; base < m4; base += 256 * kU) {
      float4 v[kU];
#pragma unroll
      for (int u8 = 0; u8 < kU; ++u8) {
        const int i4 = base + u8 * 256;
        v[u8] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i4 < m4) {
          const int u = (i4 * 4) % kUnits, k = (i4 * 4) / kUnits;
          v[u8] = __ldg(reinterpret_cast<const float4 *>(p.gw0 + (size_t)k * kHd2 + r * kUnits + u));
        }
      }
#pragma unroll
      for (int u8 = 0; u8 < kU; ++u8) {
        const int i4 = base + u8 * 256;
        if (i4 < m4) *reinterpret_cast<float4 *>(w0_s + i4 * 4) = v[u8];
      }
    }
  }
  if (tid < 4 * kUnits) bg_s[tid] = __ldg(p.bg + (tid / kUnits) * kHd2 + r * kUnits + (tid % kUnits));
  if (tid < kUnits) b0_s[tid] = __ldg(p.gb0 + r * kUnits + tid);
  for (int i = tid; i < 2 * kHd2 * kCl; i += 256) h_s[i] = 0.f;  // full_model.py:674
  for (int i = tid; i < PS * kCl; i += 256) lg_s[i] = 1.0f / (float)P;  // full_model.py:676-677
  float c_state = 0.f;  // cell state of (unit tid%32, example tid/32)
  cluster.sync();

  const float *feat = p.feat + (size_t)(own ? e_glob : 0) * P * kCf2;
  for (int it = 0; it < p.n_iter; ++it) {
    const float *h_cur = h_s + (it & 1) * kHd2 * kCl;
    float *h_nxt_local = h_s + ((it + 1) & 1) * kHd2 * kCl;

    // ---- step 1 (owner): glimpse map out, read-out x = sum_p feat[p][:] * map[p]; broadcast x
    if (own)
      for (int i = tid; i < P; i += 256) p.gmap_out[((size_t)e_glob * p.n_iter + it) * P + i] = lg_s[i];
    {
      const int c = tid % kCf2, sl = tid / kCf2;  // 4 slices over p
      float a = 0.f;
      if (own) {
        // 8 independent loads in flight (the first iteration of a launch reads feat from L2, not L1)
        float a4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a4[j] = 0.f;
        int q = sl;
        for (; q + 28 < P; q += 32) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __ldg(feat + (size_t)(q + 4 * j) * kCf2 + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) a4[j] = fmaf(f[j], lg_s[q + 4 * j], a4[j]);
        }
        for (; q < P; q += 4) a4[0] = fmaf(__ldg(feat + (size_t)q * kCf2 + c), lg_s[q], a4[0]);
        a = ((a4[0] + a4[1]) + (a4[2] + a4[3])) + ((a4[4] + a4[5]) + (a4[6] + a4[7]));
      }
      part_s[tid] = a;
      __syncthreads();
      if (tid < kCf2) {
        const float x = (part_s[tid] + part_s[tid + 64]) + (part_s[tid + 128] + part_s[tid + 192]);
#pragma unroll
        for (int d = 0; d < kCl; ++d) cluster.map_shared_rank(x_s, d)[tid * kCl + r] = x;
      }
    }
    cluster.sync();

    // ---- step 2: LSTM gates of my 32 units for the 8 examples (nnlib.py:641-647)
    {
      const int col = tid % (4 * kUnits), kh = tid / (4 * kUnits);  // 2 halves of the 320 inputs
      float acc[kCl];
#pragma unroll
      for (int e = 0; e < kCl; ++e) acc[e] = 0.f;
      const int k0 = kh * (kKin / 2);
#pragma unroll 4
      for (int k = k0; k < k0 + kKin / 2; ++k) {
        const float w = wg_s[k * (4 * kUnits) + col];
        const float *v = (k < kCf2) ? (x_s + k * kCl) : (h_cur + (k - kCf2) * kCl);
        const float4 v0 = *reinterpret_cast<const float4 *>(v);
        const float4 v1 = *reinterpret_cast<const float4 *>(v + 4);
        acc[0] = fmaf(w, v0.x, acc[0]);
        acc[1] = fmaf(w, v0.y, acc[1]);
        acc[2] = fmaf(w, v0.z, acc[2]);
        acc[3] = fmaf(w, v0.w, acc[3]);
        acc[4] = fmaf(w, v1.x, acc[4]);
        acc[5] = fmaf(w, v1.y, acc[5]);
        acc[6] = fmaf(w, v1.z, acc[6]);
        acc[7] = fmaf(w, v1.w, acc[7]);
      }
      float *gp = t_s + (size_t)kh * (4 * kUnits) * kCl + col * kCl;  // [kh][col][e]
#pragma unroll
      for (int e = 0; e < kCl; ++e) gp[e] = acc[e];
      __syncthreads();
      const int u = tid % kUnits, e = tid / kUnits;
      float g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cq = q * kUnits + u;
        g[q] = (t_s[cq * kCl + e] + t_s[(4 * kUnits + cq) * kCl + e]) + bg_s[cq];
      }
      const float gi = ra::sigmoidf_acc(g[0]);
      const float gf = ra::sigmoidf_acc(g[1]);
      const float go = ra::sigmoidf_acc(g[2]);
      const float uu = tanhf(g[3]);
      c_state = gf * c_state + gi * uu;
      const float hn = go * tanhf(c_state);
      const int off = (int)(h_nxt_local - smem) + (r * kUnits + u) * kCl + e;
#pragma unroll
      for (int d = 0; d < kCl; ++d) cluster.map_shared_rank(smem, d)[off] = hn;
    }
    cluster.sync();
    if (it == p.n_iter - 1) break;  // the last glimpse map is dead compute (full_model.py:686-688)
    const float *h_new = h_nxt_local;

    // ---- step 3: glimpse MLP layer 0, my 32 columns x 8 examples: relu(h W0 + b0)
    {
      const int u = tid % kUnits, e = tid / kUnits;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll 4
      for (int k = 0; k < kHd2; k += 2) {
        a0 = fmaf(h_new[k * kCl + e], w0_s[k * kUnits + u], a0);
        a1 = fmaf(h_new[(k + 1) * kCl + e], w0_s[(k + 1) * kUnits + u], a1);
      }
      const float tv = fmaxf((a0 + a1) + b0_s[u], 0.f);
      const int off = (int)(t_s - smem) + (r * kUnits + u) * kCl + e;
#pragma unroll
      for (int d = 0; d < kCl; ++d) cluster.map_shared_rank(smem, d)[off] = tv;
    }
    cluster.sync();

    // ---- step 4: glimpse MLP layer 1 logits for my PS positions x 8 examples -> owner CTAs
    // The weights come straight from global memory (no shared memory left for them): 8 independent loads are kept
    // in flight, and when the CTA has at most 128 outputs the 256 inputs are split over two thread halves.
    if (PS * kCl <= 128) {
      const int o = tid & 127, kh = tid >> 7;
      const int pos = o % PS, e = o / PS;
      const int gp = r * PS + pos;
      const bool live = o < PS * kCl && gp < P;
      float a = 0.f;
      if (live) {
        const float *wq = p.gw1 + gp;
        float a8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = 0.f;
        const int k0 = kh * (kHd2 / 2);
        for (int k = k0; k < k0 + kHd2 / 2; k += 8) {
          float wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = __ldg(wq + (size_t)(k + j) * P);
#pragma unroll
          for (int j = 0; j < 8; ++j) a8[j] = fmaf(t_s[(k + j) * kCl + e], wv[j], a8[j]);
        }
        a = ((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]));
      }
      part_s[tid] = a;
      __syncthreads();
      if (live && kh == 0) cluster.map_shared_rank(lg_s, e)[gp] = (part_s[o] + part_s[o + 128]) + __ldg(p.gb1 + gp);
    } else {
      for (int o = tid; o < PS * kCl; o += 256) {
        const int pos = o % PS, e = o / PS;
        const int gp = r * PS + pos;
        if (gp < P) {
          const float *wq = p.gw1 + gp;
          float a8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a8[j] = 0.f;
          for (int k = 0; k < kHd2; k += 8) {
            float wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = __ldg(wq + (size_t)(k + j) * P);
#pragma unroll
            for (int j = 0; j < 8; ++j) a8[j] = fmaf(t_s[(k + j) * kCl + e], wv[j], a8[j]);
          }
          cluster.map_shared_rank(lg_s, e)[gp] =
              (((a8[0] + a8[1]) + (a8[2] + a8[3])) + ((a8[4] + a8[5]) + (a8[6] + a8[7]))) + __ldg(p.gb1 + gp);
        }
      }
    }
    cluster.sync();

    // ---- step 5 (owner): softmax over the P positions (full_model.py:350-352)
    {
      float lmax = -INFINITY;
      for (int q = tid; q < P; q += 256) lmax = fmaxf(lmax, lg_s[q]);
      lmax = ra::block_max(lmax, red_s);
      float lsum = 0.f;
      for (int q = tid; q < P; q += 256) {
        const float ev = expf(lg_s[q] - lmax);
        lg_s[q] = ev;
        lsum += ev;
      }
      lsum = ra::block_sum(lsum, red_s);
      for (int q = tid; q < P; q += 256) lg_s[q] = lg_s[q] / lsum;
      __syncthreads();
    }
  }

  // ---- head (owner): ctrl_out = h Wc + bc, box parameters
  const float *h_fin = h_s + (p.n_iter & 1) * kHd2 * kCl;
  if (own) {
    const int lane = tid & 31, wid = tid >> 5;
    for (int o = wid; o < 9; o += 8) {
      float a = 0.f;
      for (int k = lane; k < kHd2; k += 32) a = fmaf(h_fin[k * kCl + r], __ldg(p.cw + (size_t)k * 9 + o), a);
      a = ra::warp_sum(a);
      if (lane == 0) out_s[o] = a + __ldg(p.cb + o);
    }
    p.h_out[(size_t)e_glob * kHd2 + tid] = h_fin[tid * kCl + r];
  }
  __syncthreads();
  if (own && tid < 9) p.ctrl_out[(size_t)e_glob * 9 + tid] = out_s[tid];
  if (own && tid == 0) write_box(p, out_s, p.box_out + (size_t)e_glob * RA_BOX_STRIDE);
  cluster.sync();  // no CTA may exit while its shared memory can still be written remotely
}


}  // namespace

extern "C" int ra_controller_step_f32(const float *feat, int B, int P, int Cf, int Hd, int n_iter,
                                      const float *lstm_wx, const float *lstm_wh, const float *lstm_b,
                                      const float *gmlp_w0, const float *gmlp_b0, const float *gmlp_w1,
                                      const float *gmlp_b1, const float *cmlp_w, const float *cmlp_b, int inp_height,
                                      int inp_width, int filter_height, int filter_width, int flags, float *h_out,
                                      float *ctrl_out, float *glimpse_map, float *box, void *stream) {
  if (!feat || !lstm_wx || !lstm_wh || !lstm_b || !gmlp_w0 || !gmlp_b0 || !gmlp_w1 || !gmlp_b1 || !cmlp_w ||
      !cmlp_b || !h_out || !ctrl_out || !glimpse_map || !box || B < 0 || P < 1 || n_iter < 1)
    return RA_ERR_INVALID_ARG;
  if (Hd != 256 || Cf < 1 || Cf > 256 || (256 % Cf) != 0) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  CtrlParams p;
  p.feat = feat;
  p.wx = lstm_wx;
  p.wh = lstm_wh;
  p.bg = lstm_b;
  p.gw0 = gmlp_w0;
  p.gb0 = gmlp_b0;
  p.gw1 = gmlp_w1;
  p.gb1 = gmlp_b1;
  p.cw = cmlp_w;
  p.cb = cmlp_b;
  p.h_out = h_out;
  p.ctrl_out = ctrl_out;
  p.gmap_out = glimpse_map;
  p.box_out = box;
  p.P = P;
  p.Cf = Cf;
  p.Hd = Hd;
  p.n_iter = n_iter;
  p.inp_h = inp_height;
  p.inp_w = inp_width;
  p.filt_h = filter_height;
  p.filt_w = filter_width;
  p.flags = flags;
  if (Cf == kCf2 && Hd == kHd2 && getenv("RA_CTRL_NO_CLUSTER") == nullptr) {
    const int PS = (P + kCl - 1) / kCl;
    const size_t smem2 = ((size_t)kKin * 4 * kUnits + kHd2 * kUnits + kCf2 * kCl + 3 * kHd2 * kCl + PS * kCl +
                          4 * kUnits + kUnits + 256 + 32 + 16) * sizeof(float);
    if (smem2 <= 227 * 1024) {
      static bool attr2 = false;
      if (!attr2) {
        cudaError_t e = cudaFuncSetAttribute(controller_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) {
          ra::set_last_error("cudaFuncSetAttribute(controller_cluster_kernel)", e);
          return RA_ERR_CUDA;
        }
        attr2 = true;
      }
      const int clusters = (B + kCl - 1) / kCl;
      const cudaError_t le = ra::launch_pdl(controller_cluster_kernel, dim3(clusters * kCl), dim3(256), (size_t)smem2,
                                            ra::as_stream(stream), p, B);
      if (le != cudaSuccess) {
        ra::set_last_error("cudaLaunchKernelEx(controller_cluster_kernel)", le);
        return RA_ERR_CUDA;
      }
      return ra::finish_launch("controller_cluster_kernel");
    }
  }
  const size_t smem = ((size_t)P * Cf + P + Cf + 2 * Hd + 4 * Cf + 32 + 16) * sizeof(float);
  if (smem > 200 * 1024) return RA_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(controller_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(controller_kernel)", e);
      return RA_ERR_CUDA;
    }
    attr_set = true;
  }
  controller_kernel<<<B, Hd, smem, ra::as_stream(stream)>>>(p);
  return ra::finish_launch("controller_kernel");
}

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
#ifdef RA_CTRL_PROF
__device__ unsigned long long g_ctrl_prof[64];
#define CTRL_PROF(i)                                                                   \
  do {                                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                         \
      unsigned long long t_;                                                           \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                           \
      g_ctrl_prof[i] = t_;                                                             \
    }                                                                                  \
  } while (0)
#else
#define CTRL_PROF(i) do {} while (0)
#endif
constexpr int kCl = 8;
constexpr int kHd2 = 256;
constexpr int kCf2 = 64;
constexpr int kUnits = kHd2 / kCl;     // 32 hidden units per CTA
constexpr int kKin = kCf2 + kHd2;      // 320 LSTM inputs

// Cluster exchange primitives.  The CTAs of a cluster hand each other x, h, the MLP activations and the logits through
// distributed shared memory.  A cluster-wide barrier per hand-over costs ~2 us here (barrier.cluster with release /
// acquire semantics compiles to MEMBAR.ALL.GPU + CCTL.IVALL around the hardware barrier), four times per glimpse
// iteration; instead every receiving buffer has an mbarrier and the senders deliver data and completion together
// (st.async / cp.async.bulk shared::cta -> shared::cluster with mbarrier::complete_tx::bytes), so a CTA waits exactly
// for the bytes it is going to read and nothing else.
__device__ __forceinline__ uint32_t cl_smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ uint32_t cl_mapa(uint32_t addr, int rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(rank));
  return out;
}
__device__ __forceinline__ void cl_st_async(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr),
               "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cl_bulk_push(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   remote_dst),
               "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cl_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_arm(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cl_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "CL_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CL_DONE_%=;\n\t"
      "bra CL_WAIT_%=;\n\t"
      "CL_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// Hand-over protocol of one buffer and one phase: local stores + remote sends, then `cl_handover`: the block barrier
// publishes the local part, one thread arms the mbarrier with the bytes the peers deliver (the transaction count may
// run negative until then, the phase cannot complete before the arming arrival), everyone waits for the phase.
__device__ __forceinline__ void cl_handover(uint32_t bar, uint32_t remote_bytes, uint32_t parity) {
  __syncthreads();
  if (threadIdx.x == 0) cl_mbar_arm(bar, remote_bytes);
  cl_mbar_wait(bar, parity);
}

__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(256) controller_cluster_kernel(CtrlParams p, int B) {
  CTRL_PROF(0);
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int r = (int)cluster.block_rank();
  const int e_glob = (blockIdx.x / kCl) * kCl + r;  // the example this CTA owns
  const bool own = e_glob < B;
  const int P = p.P;
  const int PS = (P + kCl - 1) / kCl;  // glimpse-map positions per CTA
  const int tid = threadIdx.x;

  float *wg_s = smem;                          // [320][32][4]  (k, u, gate q): one float4 per (input, unit)
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
  __shared__ __align__(8) unsigned long long bars[4];  // x, h, t (MLP activations), lg (logits): one phase per iteration
  const uint32_t bar_x = cl_smem_u32(&bars[0]), bar_h = cl_smem_u32(&bars[1]), bar_t = cl_smem_u32(&bars[2]),
                 bar_lg = cl_smem_u32(&bars[3]);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) cl_mbar_init(cl_smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- one-time loads: my slice of the weights, as asynchronous copies (cp.async) so that all 168 loads of a thread
  // are in flight at once - the staging is bound by L2 latency, not bandwidth.  The gate weights are scattered
  // element-wise into the [k][unit][gate] order (global reads stay coalesced along the unit index).
  // The weights are not produced by the kernel in front of this one (the last controller convolution), so under
  // programmatic dependent launch they are staged BEFORE griddepcontrol.wait, while that kernel drains.
  {
    const uint32_t wg_base = cl_smem_u32(wg_s), w0_base = cl_smem_u32(w0_s);
    for (int i = tid; i < kKin * 4 * kUnits; i += 256) {
      const int u = i % kUnits, q = (i / kUnits) % 4, k = i / (4 * kUnits);
      const int j = r * kUnits + u;
      const float *src = (k < kCf2) ? p.wx + ((size_t)q * kCf2 + k) * kHd2 + j
                                    : p.wh + ((size_t)q * kHd2 + (k - kCf2)) * kHd2 + j;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wg_base + (uint32_t)(((k * kUnits + u) * 4 + q) * 4)),
                   "l"(src)
                   : "memory");
    }
    for (int i4 = tid; i4 < kHd2 * kUnits / 4; i4 += 256) {
      const int u = (i4 * 4) % kUnits, k = (i4 * 4) / kUnits;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(w0_base + (uint32_t)(i4 * 16)),
                   "l"(p.gw0 + (size_t)k * kHd2 + r * kUnits + u)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (tid < 4 * kUnits) bg_s[tid] = __ldg(p.bg + (tid / kUnits) * kHd2 + r * kUnits + (tid % kUnits));
  if (tid < kUnits) b0_s[tid] = __ldg(p.gb0 + r * kUnits + tid);
  for (int i = tid; i < 2 * kHd2 * kCl; i += 256) h_s[i] = 0.f;  // full_model.py:674
  for (int i = tid; i < PS * kCl; i += 256) lg_s[i] = 1.0f / (float)P;  // full_model.py:676-677
  float c_state = 0.f;  // cell state of (unit tid%32, example tid/32)
  CTRL_PROF(1);
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results (feat) are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  cluster.sync();
  CTRL_PROF(2);

  const float *feat = p.feat + (size_t)(own ? e_glob : 0) * P * kCf2;
  for (int it = 0; it < p.n_iter; ++it) {
    const float *h_cur = h_s + (it & 1) * kHd2 * kCl;
    float *h_nxt_local = h_s + ((it + 1) & 1) * kHd2 * kCl;
    const uint32_t ph = (uint32_t)(it & 1);

    // ---- step 1 (owner): glimpse map out, read-out x = sum_p feat[p][:] * map[p]; broadcast x
    if (own)
      for (int i = tid; i < P; i += 256) p.gmap_out[((size_t)e_glob * p.n_iter + it) * P + i] = lg_s[i];
    {
      // 16 slices over p x 16 channel quads: a warp reads two whole feature rows (512 contiguous bytes) per load and
      // every thread keeps 8 independent 16-byte loads in flight - the 128 KB of features do not fit in what is left of
      // L1 beside the weights, so every glimpse iteration streams them from L2 and the loop is bound by load latency.
      // The partial sums go through the hidden-state buffer this iteration will overwrite in step 2.
      const int c4 = tid % (kCf2 / 4), sl = tid / (kCf2 / 4);
      constexpr int kSl = 256 / (kCf2 / 4);  // 16
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (own) {
        float4 a4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *f4 = reinterpret_cast<const float4 *>(feat) + c4;
        int q = sl;
        for (; q + 7 * kSl < P; q += 8 * kSl) {
          float4 f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __ldg(f4 + (size_t)(q + kSl * j) * (kCf2 / 4));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float m = lg_s[q + kSl * j];
            a4[j].x = fmaf(f[j].x, m, a4[j].x);
            a4[j].y = fmaf(f[j].y, m, a4[j].y);
            a4[j].z = fmaf(f[j].z, m, a4[j].z);
            a4[j].w = fmaf(f[j].w, m, a4[j].w);
          }
        }
        for (; q < P; q += kSl) {
          const float4 f = __ldg(f4 + (size_t)q * (kCf2 / 4));
          const float m = lg_s[q];
          a4[0].x = fmaf(f.x, m, a4[0].x);
          a4[0].y = fmaf(f.y, m, a4[0].y);
          a4[0].z = fmaf(f.z, m, a4[0].z);
          a4[0].w = fmaf(f.w, m, a4[0].w);
        }
        a.x = ((a4[0].x + a4[1].x) + (a4[2].x + a4[3].x)) + ((a4[4].x + a4[5].x) + (a4[6].x + a4[7].x));
        a.y = ((a4[0].y + a4[1].y) + (a4[2].y + a4[3].y)) + ((a4[4].y + a4[5].y) + (a4[6].y + a4[7].y));
        a.z = ((a4[0].z + a4[1].z) + (a4[2].z + a4[3].z)) + ((a4[4].z + a4[5].z) + (a4[6].z + a4[7].z));
        a.w = ((a4[0].w + a4[1].w) + (a4[2].w + a4[3].w)) + ((a4[4].w + a4[5].w) + (a4[6].w + a4[7].w));
      }
      *reinterpret_cast<float4 *>(h_nxt_local + sl * kCf2 + c4 * 4) = a;  // [16 slices][64 channels]
      __syncthreads();
      if (tid < kCf2) {
        float x = 0.f;
#pragma unroll
        for (int j = 0; j < kSl; j += 4)
          x += (h_nxt_local[j * kCf2 + tid] + h_nxt_local[(j + 1) * kCf2 + tid]) +
               (h_nxt_local[(j + 2) * kCf2 + tid] + h_nxt_local[(j + 3) * kCf2 + tid]);
        x_s[tid * kCl + r] = x;
        const uint32_t dst = cl_smem_u32(x_s + tid * kCl + r);
#pragma unroll
        for (int d = 1; d < kCl; ++d) {
          const int peer = (r + d) % kCl;
          cl_st_async(cl_mapa(dst, peer), x, cl_mapa(bar_x, peer));
        }
      }
    }
    cl_handover(bar_x, (kCl - 1) * kCf2 * 4, ph);
    CTRL_PROF(3 + it * 5);

    // ---- step 2: LSTM gates of my 32 units for the 8 examples (nnlib.py:641-647)
    // [8 x 320] x [320 x 128] per CTA.  A lane owns ONE hidden unit: its four gate weights of input k are one float4 and
    // are applied to all 8 examples (32 accumulators), so every weight is read from shared memory exactly once.  Warp w
    // covers units [8 (w & 3), +8) and the k-half (w >> 2); its four quarter-warps interleave k (k = 4 i + quarter), so
    // the 8-example vectors of four consecutive inputs are one conflict-free 128-byte read.  The quarters are combined
    // by a transposing shuffle reduction (each lane ends with one gate x 8 examples), the halves through t_s.
    {
      const int lane = tid & 31, wid = tid >> 5;
      const int ug = wid & 3, kh = wid >> 2, qq = lane >> 3, ul = lane & 7;
      const int ucol = ug * 8 + ul;
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      const float *hv = h_cur - kCf2 * kCl;  // so that input k >= 64 indexes hv[k * 8]
      const int kbase = kh * (kKin / 2) + qq;
#pragma unroll 2
      for (int i = 0; i < kKin / 8; ++i) {
        const int k = kbase + 4 * i;
        const float4 w = *reinterpret_cast<const float4 *>(wg_s + ((size_t)k * kUnits + ucol) * 4);
        const float *v = ((k < kCf2) ? x_s : hv) + k * kCl;
        const float4 v0 = *reinterpret_cast<const float4 *>(v);
        const float4 v1 = *reinterpret_cast<const float4 *>(v + 4);
        const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[q * 8 + 0] = fmaf(ws[q], v0.x, acc[q * 8 + 0]);
          acc[q * 8 + 1] = fmaf(ws[q], v0.y, acc[q * 8 + 1]);
          acc[q * 8 + 2] = fmaf(ws[q], v0.z, acc[q * 8 + 2]);
          acc[q * 8 + 3] = fmaf(ws[q], v0.w, acc[q * 8 + 3]);
          acc[q * 8 + 4] = fmaf(ws[q], v1.x, acc[q * 8 + 4]);
          acc[q * 8 + 5] = fmaf(ws[q], v1.y, acc[q * 8 + 5]);
          acc[q * 8 + 6] = fmaf(ws[q], v1.z, acc[q * 8 + 6]);
          acc[q * 8 + 7] = fmaf(ws[q], v1.w, acc[q * 8 + 7]);
        }
      }
      if (it == 1) CTRL_PROF(42);
      // quarters 0/1 keep gates 0,1 and quarters 2/3 gates 2,3; then even quarters the lower gate, odd the upper one
      float r1[16], r2[8];
      const bool up1 = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float send = up1 ? acc[i] : acc[16 + i];
        const float keep = up1 ? acc[16 + i] : acc[i];
        r1[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
      const bool up2 = (lane & 8) != 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float send = up2 ? r1[i] : r1[8 + i];
        const float keep = up2 ? r1[8 + i] : r1[i];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      float *gp = t_s + ((size_t)(kh * 4 + qq) * kCl) * kUnits + ucol;  // [k-half][gate][example][unit]
#pragma unroll
      for (int e = 0; e < kCl; ++e) gp[e * kUnits] = r2[e];
      if (it == 1) CTRL_PROF(43);
      __syncthreads();
      if (it == 1) CTRL_PROF(44);
      const int u = tid % kUnits, e = tid / kUnits;
      float g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        g[q] = (t_s[((size_t)q * kCl + e) * kUnits + u] + t_s[((size_t)(4 + q) * kCl + e) * kUnits + u]) +
               bg_s[q * kUnits + u];
      const float gi = ra::sigmoidf_acc(g[0]);
      const float gf = ra::sigmoidf_acc(g[1]);
      const float go = ra::sigmoidf_acc(g[2]);
      const float uu = tanhf(g[3]);
      c_state = gf * c_state + gi * uu;
      const float hn = go * tanhf(c_state);
      if (it == 1) CTRL_PROF(45);
      // my block [32 units][8 examples] is 1 KB contiguous in every CTA's h buffer: one bulk push per peer
      float *blk = h_nxt_local + (size_t)r * kUnits * kCl;
      blk[u * kCl + e] = hn;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid >= 1 && tid < kCl) {
        const int peer = (r + tid) % kCl;
        cl_bulk_push(cl_mapa(cl_smem_u32(blk), peer), cl_smem_u32(blk), kUnits * kCl * 4, cl_mapa(bar_h, peer));
      }
      if (it == 1) CTRL_PROF(46);
    }
    cl_handover(bar_h, (kCl - 1) * kUnits * kCl * 4, ph);
    CTRL_PROF(4 + it * 5);
    if (it == p.n_iter - 1) break;  // the last glimpse map is dead compute (full_model.py:686-688)
    const float *h_new = h_nxt_local;

    // ---- step 3: glimpse MLP layer 0, my 32 columns x 8 examples: relu(h W0 + b0)
    // lane = column, warp = a 32-input slice, every weight applied to all 8 examples; the slices are combined through
    // the hidden-state buffer that died in step 2.
    {
      const int lane = tid & 31, wid = tid >> 5;
      float acc[kCl];
#pragma unroll
      for (int e = 0; e < kCl; ++e) acc[e] = 0.f;
#pragma unroll 8
      for (int i = 0; i < kHd2 / 8; ++i) {
        const int k = wid * (kHd2 / 8) + i;
        const float w = w0_s[k * kUnits + lane];
        const float4 v0 = *reinterpret_cast<const float4 *>(h_new + k * kCl);
        const float4 v1 = *reinterpret_cast<const float4 *>(h_new + k * kCl + 4);
        acc[0] = fmaf(v0.x, w, acc[0]);
        acc[1] = fmaf(v0.y, w, acc[1]);
        acc[2] = fmaf(v0.z, w, acc[2]);
        acc[3] = fmaf(v0.w, w, acc[3]);
        acc[4] = fmaf(v1.x, w, acc[4]);
        acc[5] = fmaf(v1.y, w, acc[5]);
        acc[6] = fmaf(v1.z, w, acc[6]);
        acc[7] = fmaf(v1.w, w, acc[7]);
      }
      float *scr = h_s + (it & 1) * kHd2 * kCl;  // [8 slices][8 examples][32 columns]
#pragma unroll
      for (int e = 0; e < kCl; ++e) scr[((size_t)wid * kCl + e) * kUnits + lane] = acc[e];
      __syncthreads();
      const int u = tid % kUnits, e = tid / kUnits;
      float a = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; w8 += 2)
        a += scr[((size_t)w8 * kCl + e) * kUnits + u] + scr[((size_t)(w8 + 1) * kCl + e) * kUnits + u];
      const float tv = fmaxf(a + b0_s[u], 0.f);
      __syncthreads();  // step 4 reuses scr after the cluster barrier, but keep the block-level hazard explicit
      float *blk = t_s + (size_t)r * kUnits * kCl;
      blk[u * kCl + e] = tv;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid >= 1 && tid < kCl) {
        const int peer = (r + tid) % kCl;
        cl_bulk_push(cl_mapa(cl_smem_u32(blk), peer), cl_smem_u32(blk), kUnits * kCl * 4, cl_mapa(bar_t, peer));
      }
    }
    cl_handover(bar_t, (kCl - 1) * kUnits * kCl * 4, ph);
    CTRL_PROF(5 + it * 5);

    // ---- step 4: glimpse MLP layer 1 logits for my PS positions x 8 examples -> owner CTAs
    // The weights come straight from L2 (no shared memory left for them), each element exactly once per CTA: a thread
    // owns one position and one slice of the 256 inputs, keeps 8 independent loads in flight and applies every weight
    // to all 8 examples; the slices are combined through the dead hidden-state buffer of this iteration.
    {
      const int KS = (PS <= 64) ? 4 : (PS <= 128) ? 2 : 1;  // input slices; PS * KS <= 256 when PS <= 256
      const int klen = kHd2 / KS;
      float *scr = h_s + (it & 1) * kHd2 * kCl;             // h_cur: dead after step 2, [KS][PS][8] <= 2048 floats
      for (int pbase = 0; pbase < PS; pbase += 256) {       // PS > 256 (very large maps): several passes, KS == 1
        const int o = tid, np = min(PS - pbase, 256);
        const int pos = pbase + o % np, ks = o / np;
        const int gp = r * PS + pos;
        const bool live = ks < KS && gp < P;
        float acc[kCl];
#pragma unroll
        for (int e = 0; e < kCl; ++e) acc[e] = 0.f;
        if (live) {
          const float *wq = p.gw1 + gp;
          const int k0 = ks * klen;
          for (int k = k0; k < k0 + klen; k += 8) {
            float wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = __ldg(wq + (size_t)(k + j) * P);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t0 = *reinterpret_cast<const float4 *>(t_s + (k + j) * kCl);
              const float4 t1 = *reinterpret_cast<const float4 *>(t_s + (k + j) * kCl + 4);
              acc[0] = fmaf(t0.x, wv[j], acc[0]);
              acc[1] = fmaf(t0.y, wv[j], acc[1]);
              acc[2] = fmaf(t0.z, wv[j], acc[2]);
              acc[3] = fmaf(t0.w, wv[j], acc[3]);
              acc[4] = fmaf(t1.x, wv[j], acc[4]);
              acc[5] = fmaf(t1.y, wv[j], acc[5]);
              acc[6] = fmaf(t1.z, wv[j], acc[6]);
              acc[7] = fmaf(t1.w, wv[j], acc[7]);
            }
          }
        }
        if (KS == 1) {
          if (live) {
            const float bias = __ldg(p.gb1 + gp);
            const uint32_t dst = cl_smem_u32(lg_s + gp);
#pragma unroll
            for (int e = 0; e < kCl; ++e) {
              if (e == r)
                lg_s[gp] = acc[e] + bias;
              else
                cl_st_async(cl_mapa(dst, e), acc[e] + bias, cl_mapa(bar_lg, e));
            }
          }
        } else {
          if (ks < KS) {
            float *dst = scr + ((size_t)ks * np + (pos - pbase)) * kCl;
            *reinterpret_cast<float4 *>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
          }
          __syncthreads();
          for (int o2 = tid; o2 < np * kCl; o2 += 256) {
            const int e = o2 / np, ps = o2 % np;
            const int gq = r * PS + pbase + ps;
            if (gq < P) {
              float v = scr[(size_t)ps * kCl + e];
              for (int s2 = 1; s2 < KS; ++s2) v += scr[((size_t)s2 * np + ps) * kCl + e];
              v += __ldg(p.gb1 + gq);
              if (e == r)
                lg_s[gq] = v;
              else
                cl_st_async(cl_mapa(cl_smem_u32(lg_s + gq), e), v, cl_mapa(bar_lg, e));
            }
          }
        }
      }
    }
    {
      const int n_own = max(0, min(PS, P - r * PS));  // my own positions are plain local stores
      cl_handover(bar_lg, (uint32_t)(P - n_own) * 4u, ph);
    }
    CTRL_PROF(6 + it * 5);

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
    CTRL_PROF(7 + it * 5);
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
  CTRL_PROF(40);
  cluster.sync();  // no CTA may exit while its shared memory can still be written remotely
  CTRL_PROF(41);
}


}  // namespace

#ifdef RA_CTRL_PROF
extern "C" int ra_debug_ctrl_prof(unsigned long long *host_out) {
  return cudaMemcpyFromSymbol(host_out, g_ctrl_prof, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

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
    constexpr size_t kDynMax = 227 * 1024 - 64;  // the kernel also holds 32 bytes of static shared memory (mbarriers)
    if (smem2 <= kDynMax) {
      static bool attr2 = false;
      if (!attr2) {
        cudaError_t e = cudaFuncSetAttribute(controller_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kDynMax);
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

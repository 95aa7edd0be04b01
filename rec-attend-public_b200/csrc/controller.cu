// Controller step: soft-attention read-out x LSTM x glimpse-MLP softmax, repeated n_iter
// times, then the controller head and the box-parameter maths — full_model.py:668-725,
// box_model.py:417-470, nnlib.py:476-493 (mlp), nnlib.py:637-649 (lstm), modellib.py:752-856.
//
// One CTA per example, Hd (=256) threads, the whole inner loop in one launch (the reference
// issues ~35 TF ops per glimpse iteration).  The feature map [P,Cf] is staged once in shared
// memory and re-read 5 times; thread j owns hidden unit j: its four gate pre-activations are
// dot products over the 64+256 inputs with weight reads coalesced across j (weights live in
// L2: 1.7 MB shared by every CTA).  The read-out and the softmax are warp/CTA reductions.
#include "common.cuh"

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
  if (tid == 0) {
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
    float *bo = p.box_out + (size_t)b * RA_BOX_STRIDE;
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

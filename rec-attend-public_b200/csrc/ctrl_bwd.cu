// Backward of the controller (full_model.py:668-725: soft read-out, LSTM, glimpse-MLP softmax x n_iter, linear head,
// box-parameter maths) - fourth building block of the backward pass.  Correctness-first fp32 kernels:
//
//   ra_controller_tape_f32   re-runs the controller forward for one decode step and records what the backward needs
//                            (one CTA per example, thread = hidden unit): per glimpse iteration k a record
//                            [ map_k (P) | glimpse_k (Cf) | h_{k-1} (Hd) | c_{k-1} (Hd) | gates i,f,o,u (4 Hd) |
//                              c_k (Hd) | h_k (Hd) | a1_k (Hd) | map_{k+1} (P) ]   (RA_CTRL_TAPE_* offsets)
//                            plus h_out and ctrl_out (bit-for-bit the same maths as the oracle's loop, so the test can
//                            also compare it with the production kernel of controller.cu).
//   ra_controller_head_bwd_f32  d_box (d_ctr, d_size, d_lg_var) and the three dL/dgamma -> d_ctrl_out [B,9]
//                            through ctr = (n+1) S/2, size = exp(l) S, lg_var (fixed / size-derived / dynamic),
//                            gamma = exp(lg_gamma), with the squash / fixed_var / dynamic_var / fixed_gamma flags.
//   ra_controller_bwd_f32    BPTT over the n_iter glimpse iterations: d_feat [B,P,Cf] and the pre-activation deltas
//                            dG [B,n_iter,4,Hd], dA1 [B,n_iter,Hd], dLog [B,n_iter,P] the weight gradients are built from.
//   ra_outer_sum_f32         dW[in,out] (+ db[out]) = sum over rows r of A[r,:]^T D[r,:] - the weight gradients of every
//                            dense layer (rows = (example, iteration)), summed in a fixed order.
#include "common.cuh"

namespace {

constexpr int kMaxP = 1024;  // glimpse positions (512 at Cityscapes 512x1024)
constexpr int kMaxHd = 256;
constexpr int kMaxCf = 256;

struct TapeLayout {
  int P, Cf, Hd;
  __host__ __device__ int o_map() const { return 0; }
  __host__ __device__ int o_glimpse() const { return P; }
  __host__ __device__ int o_hprev() const { return P + Cf; }
  __host__ __device__ int o_cprev() const { return P + Cf + Hd; }
  __host__ __device__ int o_gates() const { return P + Cf + 2 * Hd; }
  __host__ __device__ int o_c() const { return P + Cf + 6 * Hd; }
  __host__ __device__ int o_h() const { return P + Cf + 7 * Hd; }
  __host__ __device__ int o_a1() const { return P + Cf + 8 * Hd; }
  __host__ __device__ int o_mapn() const { return P + Cf + 9 * Hd; }
  __host__ __device__ int rec() const { return 2 * P + Cf + 9 * Hd; }
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates order i, f, o, u (the layout of lstm_wx [4,Cf,Hd], lstm_wh [4,Hd,Hd], lstm_b [4,Hd])
__global__ void controller_tape_kernel(const float *__restrict__ feat, int P, int Cf, int Hd, int n_iter,
                                       const float *__restrict__ wx, const float *__restrict__ wh,
                                       const float *__restrict__ bg, const float *__restrict__ w0,
                                       const float *__restrict__ b0, const float *__restrict__ w1,
                                       const float *__restrict__ b1, const float *__restrict__ cw,
                                       const float *__restrict__ cb, float *__restrict__ tape,
                                       float *__restrict__ h_out, float *__restrict__ ctrl_out) {
  __shared__ float map_s[kMaxP], g_s[kMaxCf], h_s[kMaxHd], c_s[kMaxHd], a1_s[kMaxHd], red[32];
  const int b = blockIdx.x, j = threadIdx.x, nt = blockDim.x;
  const TapeLayout L{P, Cf, Hd};
  const float *fb = feat + (size_t)b * P * Cf;
  for (int p = j; p < P; p += nt) map_s[p] = 1.0f / (float)P;
  for (int m = j; m < Hd; m += nt) h_s[m] = c_s[m] = 0.f;
  __syncthreads();
  for (int k = 0; k < n_iter; ++k) {
    float *rec = tape + ((size_t)b * n_iter + k) * L.rec();
    for (int p = j; p < P; p += nt) rec[L.o_map() + p] = map_s[p];
    for (int m = j; m < Hd; m += nt) {
      rec[L.o_hprev() + m] = h_s[m];
      rec[L.o_cprev() + m] = c_s[m];
    }
    for (int c = j; c < Cf; c += nt) {  // glimpse = sum_p feat[p,:] * map[p]  (full_model.py:680)
      float s = 0.f;
      for (int p = 0; p < P; ++p) s = fmaf(fb[(size_t)p * Cf + c], map_s[p], s);
      g_s[c] = s;
      rec[L.o_glimpse() + c] = s;
    }
    __syncthreads();
    float cn = 0.f, hn = 0.f;
    if (j < Hd) {  // nnlib.py:641-647
      float pre[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float s = bg[g * Hd + j];
        for (int c = 0; c < Cf; ++c) s = fmaf(g_s[c], wx[((size_t)g * Cf + c) * Hd + j], s);
        for (int m = 0; m < Hd; ++m) s = fmaf(h_s[m], wh[((size_t)g * Hd + m) * Hd + j], s);
        pre[g] = s;
      }
      const float gi = sigm(pre[0]), gf = sigm(pre[1]), go = sigm(pre[2]), gu = tanhf(pre[3]);
      cn = gf * c_s[j] + gi * gu;
      hn = go * tanhf(cn);
      rec[L.o_gates() + 0 * Hd + j] = gi;
      rec[L.o_gates() + 1 * Hd + j] = gf;
      rec[L.o_gates() + 2 * Hd + j] = go;
      rec[L.o_gates() + 3 * Hd + j] = gu;
      rec[L.o_c() + j] = cn;
      rec[L.o_h() + j] = hn;
    }
    __syncthreads();
    if (j < Hd) {
      c_s[j] = cn;
      h_s[j] = hn;
    }
    __syncthreads();
    if (j < Hd) {  // glimpse MLP layer 0: relu
      float s = b0[j];
      for (int m = 0; m < Hd; ++m) s = fmaf(h_s[m], w0[(size_t)m * Hd + j], s);
      s = fmaxf(s, 0.f);
      a1_s[j] = s;
      rec[L.o_a1() + j] = s;
    }
    __syncthreads();
    float mx = -INFINITY;  // layer 1 + softmax over the P positions
    for (int p = j; p < P; p += nt) {
      float s = b1[p];
      for (int m = 0; m < Hd; ++m) s = fmaf(a1_s[m], w1[(size_t)m * P + p], s);
      map_s[p] = s;
      mx = fmaxf(mx, s);
    }
    mx = ra::block_max(mx, red);
    float sum = 0.f;
    for (int p = j; p < P; p += nt) {
      const float e = expf(map_s[p] - mx);
      map_s[p] = e;
      sum += e;
    }
    sum = ra::block_sum(sum, red);
    for (int p = j; p < P; p += nt) {
      map_s[p] = map_s[p] / sum;
      rec[L.o_mapn() + p] = map_s[p];
    }
    __syncthreads();
  }
  for (int m = j; m < Hd; m += nt) h_out[(size_t)b * Hd + m] = h_s[m];
  if (j < 9) {
    float s = cb[j];
    for (int m = 0; m < Hd; ++m) s = fmaf(h_s[m], cw[(size_t)m * 9 + j], s);
    ctrl_out[(size_t)b * 9 + j] = s;
  }
}

// full_model.py:691-725 / modellib.py:752-856 backward; one thread per example
__global__ void controller_head_bwd_kernel(const float *__restrict__ ctrl_out, const float *__restrict__ box,
                                           const float *__restrict__ d_box, const float *__restrict__ d_gamma3, int B,
                                           float img_h, float img_w, int flags, float *__restrict__ d_ctrl_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float *co = ctrl_out + (size_t)b * 9, *bo = box + (size_t)b * RA_BOX_STRIDE, *db = d_box + (size_t)b * 6;
  float *d = d_ctrl_out + (size_t)b * 9;
  const float S[2] = {img_h, img_w};
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    float d_ctr = db[0 + a], d_size = db[2 + a], d_lgv = db[4 + a];
    const float size = bo[RA_BOX_SIZE_Y + a];
    float d_lgv_direct = 0.f;
    if (flags & RA_CTRL_DYNAMIC_VAR) {
      d_lgv_direct = d_lgv;  // lg_var = ctrl_out[4:6]
    } else if (!(flags & RA_CTRL_FIXED_VAR)) {
      d_size += d_lgv / size;  // lg_var = log(size) - log(F)
    }
    float d_n = d_ctr * S[a] / 2.0f;  // ctr = (n + 1) * S / 2
    float d_l = d_size * size;        // size = exp(l) * S
    if (flags & RA_CTRL_SQUASH) {     // n = tanh(raw), l = -softplus(raw)
      const float t = tanhf(co[0 + a]);
      d_n *= (1.0f - t * t);
      d_l *= -sigm(co[2 + a]);
    }
    d[0 + a] = d_n;
    d[2 + a] = d_l;
    d[4 + a] = d_lgv_direct;
  }
  // gains: gamma = exp(lg_gamma) -> dL/d lg_gamma = dL/dgamma * gamma; fixed_gamma pins the attention and mask gains
  const float *dg = d_gamma3 + (size_t)b * 3;
  const bool fixed = (flags & RA_CTRL_FIXED_GAMMA) != 0;
  d[6] = fixed ? 0.f : dg[0] * bo[RA_BOX_GAMMA_ATTN];
  d[7] = dg[1] * bo[RA_BOX_GAMMA_BOX];
  d[8] = fixed ? 0.f : dg[2] * bo[RA_BOX_GAMMA_Y];
}

// lane-strided partial dot product of a global row with a shared-memory vector; warp_row_dot adds the warp reduction
__device__ __forceinline__ float warp_row_partial(const float *__restrict__ row, const float *v_s, int n, int lane) {
  float s0 = 0.f, s1 = 0.f;
  int q = lane;
  for (; q + 32 < n; q += 64) {
    s0 = fmaf(__ldg(row + q), v_s[q], s0);
    s1 = fmaf(__ldg(row + q + 32), v_s[q + 32], s1);
  }
  if (q < n) s0 = fmaf(__ldg(row + q), v_s[q], s0);
  return s0 + s1;
}
__device__ __forceinline__ float warp_row_dot(const float *__restrict__ row, const float *v_s, int n, int lane) {
  return ra::warp_sum(warp_row_partial(row, v_s, n, lane));
}

__global__ void controller_bwd_kernel(const float *__restrict__ feat, int P, int Cf, int Hd, int n_iter,
                                      const float *__restrict__ wx, const float *__restrict__ wh,
                                      const float *__restrict__ w0, const float *__restrict__ w1,
                                      const float *__restrict__ cw, const float *__restrict__ tape,
                                      const float *__restrict__ d_h_ext, const float *__restrict__ d_ctrl_out,
                                      float *__restrict__ d_feat, float *__restrict__ dG, float *__restrict__ dA1,
                                      float *__restrict__ dLog) {
  __shared__ float dh_s[kMaxHd], dc_s[kMaxHd], dpre_s[4 * kMaxHd], dmap_s[kMaxP], dlog_s[kMaxP], da1_s[kMaxHd],
      dgl_s[kMaxCf], red[32];
  const int b = blockIdx.x, j = threadIdx.x, nt = blockDim.x;
  const int lane = j & 31, warp = j >> 5, nw = nt >> 5;
  const TapeLayout L{P, Cf, Hd};
  const float *fb = feat + (size_t)b * P * Cf;
  float *dfb = d_feat + (size_t)b * P * Cf;
  for (int i = j; i < P * Cf; i += nt) dfb[i] = 0.f;
  for (int m = j; m < Hd; m += nt) {  // head: ctrl_out = h W + b
    float s = d_h_ext ? d_h_ext[(size_t)b * Hd + m] : 0.f;
    for (int q = 0; q < 9; ++q) s = fmaf(cw[(size_t)m * 9 + q], d_ctrl_out[(size_t)b * 9 + q], s);
    dh_s[m] = s;
    dc_s[m] = 0.f;
  }
  __syncthreads();
  for (int k = n_iter - 1; k >= 0; --k) {
    const float *rec = tape + ((size_t)b * n_iter + k) * L.rec();
    float *dG_k = dG + ((size_t)b * n_iter + k) * 4 * Hd;
    float *dA1_k = dA1 + ((size_t)b * n_iter + k) * Hd;
    float *dLog_k = dLog + ((size_t)b * n_iter + k) * P;
    if (k < n_iter - 1) {
      // map_{k+1} = softmax(logits): dlog = map * (dmap - <map, dmap>)   (dmap_s was filled by iteration k+1)
      float dot = 0.f;
      for (int p = j; p < P; p += nt) dot = fmaf(rec[L.o_mapn() + p], dmap_s[p], dot);
      dot = ra::block_sum(dot, red);
      for (int p = j; p < P; p += nt) {
        const float v = rec[L.o_mapn() + p] * (dmap_s[p] - dot);
        dlog_s[p] = v;
        dLog_k[p] = v;
      }
      __syncthreads();
      // Every matrix-vector product below walks ROWS of a row-major weight matrix: one warp per row, lanes along the
      // row (coalesced 128-byte reads), warp reduction - a thread per row would touch 32 different sectors per load.
      for (int m = warp; m < Hd; m += nw) {  // layer 1 (linear) then relu of layer 0
        float s = warp_row_dot(w1 + (size_t)m * P, dlog_s, P, lane);
        s = (rec[L.o_a1() + m] > 0.f) ? s : 0.f;
        if (lane == 0) {
          da1_s[m] = s;
          dA1_k[m] = s;
        }
      }
      __syncthreads();
      for (int m = warp; m < Hd; m += nw) {  // a1 = relu(h_k W0 + b0): dh_k += W0 da1pre
        const float s = warp_row_dot(w0 + (size_t)m * Hd, da1_s, Hd, lane);
        if (lane == 0) dh_s[m] += s;
      }
      __syncthreads();
    } else {  // the glimpse MLP of the last iteration is dead code (full_model.py:686-688)
      for (int p = j; p < P; p += nt) dLog_k[p] = 0.f;
      for (int m = j; m < Hd; m += nt) dA1_k[m] = 0.f;
    }
    for (int m = j; m < Hd; m += nt) {  // LSTM cell backward (nnlib.py:641-647)
      const float gi = rec[L.o_gates() + m], gf = rec[L.o_gates() + Hd + m], go = rec[L.o_gates() + 2 * Hd + m],
                  gu = rec[L.o_gates() + 3 * Hd + m];
      const float tc = tanhf(rec[L.o_c() + m]);
      const float dh = dh_s[m];
      const float dc = dc_s[m] + dh * go * (1.0f - tc * tc);
      const float p_i = dc * gu * gi * (1.0f - gi);
      const float p_f = dc * rec[L.o_cprev() + m] * gf * (1.0f - gf);
      const float p_o = dh * tc * go * (1.0f - go);
      const float p_u = dc * gi * (1.0f - gu * gu);
      dpre_s[m] = p_i;
      dpre_s[Hd + m] = p_f;
      dpre_s[2 * Hd + m] = p_o;
      dpre_s[3 * Hd + m] = p_u;
      dG_k[m] = p_i;
      dG_k[Hd + m] = p_f;
      dG_k[2 * Hd + m] = p_o;
      dG_k[3 * Hd + m] = p_u;
      dc_s[m] = dc * gf;  // dc_{k-1}
    }
    __syncthreads();
    for (int m = warp; m < Hd; m += nw) {  // dh_{k-1} = sum_g Wh[g] dpre_g
      float s = 0.f;
      for (int g = 0; g < 4; ++g) s += warp_row_partial(wh + ((size_t)g * Hd + m) * Hd, dpre_s + g * Hd, Hd, lane);
      s = ra::warp_sum(s);
      if (lane == 0) dh_s[m] = s;
    }
    for (int c = warp; c < Cf; c += nw) {  // dglimpse_k = sum_g Wx[g] dpre_g
      float s = 0.f;
      for (int g = 0; g < 4; ++g) s += warp_row_partial(wx + ((size_t)g * Cf + c) * Hd, dpre_s + g * Hd, Hd, lane);
      s = ra::warp_sum(s);
      if (lane == 0) dgl_s[c] = s;
    }
    __syncthreads();
    for (int i = j; i < P * Cf; i += nt) {  // glimpse = sum_p feat[p,:] map_k[p]
      const int p = i / Cf, c = i - p * Cf;
      dfb[i] = fmaf(rec[L.o_map() + p], dgl_s[c], dfb[i]);
    }
    for (int p = warp; p < P; p += nw) {  // gradient of map_k (used by iteration k-1; map_0 is a constant)
      const float s = warp_row_dot(fb + (size_t)p * Cf, dgl_s, Cf, lane);
      if (lane == 0) dmap_s[p] = s;
    }
    __syncthreads();
  }
}

// dW[i][o] = sum_r A[r*a_stride + i] * D[r*d_stride + o]; db[o] = sum_r D[r*d_stride + o]  (fixed order over r)
// dW [n_in, n_out] = sum_r A_r^T D_r as a tiled product: a CTA owns a 64 x 64 tile of dW and walks the R rows in
// chunks of 16 staged in shared memory; every thread keeps a 4 x 4 register tile.  One pass over the rows in a fixed
// order: bit-identical from run to run.  db = column sums of D (CTAs of the first tile row).
constexpr int kOsT = 64, kOsR = 16;
// grid.z > 1: row chunks of `rows_per_chunk` rows; chunk z writes its partial products to dW + z * n_in * n_out (and
// db + z * n_out) - the caller passes the workspace there and outer_sum_finalize_kernel adds the chunks in order.
__global__ void __launch_bounds__(256) outer_sum_kernel(const float *__restrict__ A, size_t a_stride, int n_in,
                                                        const float *__restrict__ D, size_t d_stride, int n_out, int R,
                                                        int rows_per_chunk, float *__restrict__ dW,
                                                        float *__restrict__ db) {
  {
    const int z = blockIdx.z;
    const int rbeg = z * rows_per_chunk;
    A += (size_t)rbeg * a_stride;
    D += (size_t)rbeg * d_stride;
    R = max(0, min(R - rbeg, rows_per_chunk));
    dW += (size_t)z * n_in * n_out;
    if (db != nullptr) db += (size_t)z * n_out;
  }
  __shared__ float a_s[kOsR][kOsT + 4];
  __shared__ float d_s[kOsR][kOsT + 4];
  const int i0 = blockIdx.y * kOsT, o0 = blockIdx.x * kOsT;
  const int ti = threadIdx.x / 16, to = threadIdx.x % 16;  // rows 4*ti.., columns 4*to..
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float dbv[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_db = db != nullptr && blockIdx.y == 0 && ti == 0;
  for (int r0 = 0; r0 < R; r0 += kOsR) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kOsR * kOsT; idx += 256) {
      const int rr = idx / kOsT, c = idx - rr * kOsT;
      const int r = r0 + rr;
      a_s[rr][c] = (r < R && i0 + c < n_in) ? A[(size_t)r * a_stride + i0 + c] : 0.f;
      d_s[rr][c] = (r < R && o0 + c < n_out) ? D[(size_t)r * d_stride + o0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kOsR; ++rr) {
      const float4 av = *reinterpret_cast<const float4 *>(&a_s[rr][4 * ti]);
      const float4 dv = *reinterpret_cast<const float4 *>(&d_s[rr][4 * to]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(a4[a], d4[b], acc[a][b]);
      if (do_db) {
#pragma unroll
        for (int b = 0; b < 4; ++b) dbv[b] += d4[b];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + 4 * ti + a, o = o0 + 4 * to + b;
      if (i < n_in && o < n_out) dW[(size_t)i * n_out + o] = acc[a][b];
    }
  if (do_db)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (o0 + 4 * to + b < n_out) db[o0 + 4 * to + b] = dbv[b];
}

// out[i] = sum over the chunks (fixed order)
__global__ void outer_sum_finalize_kernel(const float *__restrict__ part, int chunks, size_t n, float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = part[i];
  for (int k = 1; k < chunks; ++k) s += part[(size_t)k * n + i];
  out[i] = s;
}

// row chunks of the outer-sum launch: enough CTAs to cover the machine twice, at most 32 chunks, at least 64 rows each
int outer_sum_chunks(int n_in, int n_out, int R) {
  const int tiles = ((n_out + kOsT - 1) / kOsT) * ((n_in + kOsT - 1) / kOsT);
  int z = (2 * ra::kNumSMs + tiles - 1) / tiles;
  if (z > 32) z = 32;
  const int zmax = R / 64;
  if (z > zmax) z = zmax;
  return z < 1 ? 1 : z;
}

bool dims_ok(int P, int Cf, int Hd, int n_iter) {
  return P >= 1 && P <= kMaxP && Cf >= 1 && Cf <= kMaxCf && Hd >= 9 && Hd <= kMaxHd && n_iter >= 1;
}

}  // namespace

extern "C" size_t ra_controller_tape_floats(int P, int Cf, int Hd, int n_iter) {
  if (!dims_ok(P, Cf, Hd, n_iter)) return 0;
  const TapeLayout L{P, Cf, Hd};
  return (size_t)n_iter * L.rec();  // per example
}

extern "C" int ra_controller_tape_layout(int P, int Cf, int Hd, int *offsets /* [10]: 9 field offsets + record size */) {
  if (!offsets || !dims_ok(P, Cf, Hd, 1)) return RA_ERR_INVALID_ARG;
  const TapeLayout L{P, Cf, Hd};
  const int o[10] = {L.o_map(), L.o_glimpse(), L.o_hprev(), L.o_cprev(), L.o_gates(),
                     L.o_c(),   L.o_h(),       L.o_a1(),    L.o_mapn(),  L.rec()};
  for (int i = 0; i < 10; ++i) offsets[i] = o[i];
  return RA_OK;
}

extern "C" int ra_controller_tape_f32(const float *feat, int B, int P, int Cf, int Hd, int n_iter, const float *lstm_wx,
                                      const float *lstm_wh, const float *lstm_b, const float *gmlp_w0,
                                      const float *gmlp_b0, const float *gmlp_w1, const float *gmlp_b1,
                                      const float *cmlp_w, const float *cmlp_b, float *tape, float *h_out,
                                      float *ctrl_out, void *stream) {
  if (B < 0 || !dims_ok(P, Cf, Hd, n_iter)) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!feat || !lstm_wx || !lstm_wh || !lstm_b || !gmlp_w0 || !gmlp_b0 || !gmlp_w1 || !gmlp_b1 || !cmlp_w || !cmlp_b ||
      !tape || !h_out || !ctrl_out)
    return RA_ERR_INVALID_ARG;
  controller_tape_kernel<<<B, kMaxHd, 0, ra::as_stream(stream)>>>(feat, P, Cf, Hd, n_iter, lstm_wx, lstm_wh, lstm_b,
                                                                   gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1, cmlp_w, cmlp_b,
                                                                   tape, h_out, ctrl_out);
  return ra::finish_launch("controller_tape_kernel");
}

extern "C" int ra_controller_head_bwd_f32(const float *ctrl_out, const float *box, const float *d_box,
                                          const float *d_gamma3, int B, int inp_height, int inp_width, int flags,
                                          float *d_ctrl_out, void *stream) {
  if (B < 0 || inp_height < 1 || inp_width < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!ctrl_out || !box || !d_box || !d_gamma3 || !d_ctrl_out) return RA_ERR_INVALID_ARG;
  controller_head_bwd_kernel<<<(B + 63) / 64, 64, 0, ra::as_stream(stream)>>>(ctrl_out, box, d_box, d_gamma3, B,
                                                                               (float)inp_height, (float)inp_width,
                                                                               flags, d_ctrl_out);
  return ra::finish_launch("controller_head_bwd_kernel");
}

extern "C" int ra_controller_bwd_f32(const float *feat, int B, int P, int Cf, int Hd, int n_iter, const float *lstm_wx,
                                     const float *lstm_wh, const float *gmlp_w0, const float *gmlp_w1,
                                     const float *cmlp_w, const float *tape, const float *d_h, const float *d_ctrl_out,
                                     float *d_feat, float *dG, float *dA1, float *dLog, void *stream) {
  if (B < 0 || !dims_ok(P, Cf, Hd, n_iter)) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!feat || !lstm_wx || !lstm_wh || !gmlp_w0 || !gmlp_w1 || !cmlp_w || !tape || !d_ctrl_out || !d_feat || !dG ||
      !dA1 || !dLog)
    return RA_ERR_INVALID_ARG;
  controller_bwd_kernel<<<B, kMaxHd, 0, ra::as_stream(stream)>>>(feat, P, Cf, Hd, n_iter, lstm_wx, lstm_wh, gmlp_w0,
                                                                  gmlp_w1, cmlp_w, tape, d_h, d_ctrl_out, d_feat, dG, dA1,
                                                                  dLog);
  return ra::finish_launch("controller_bwd_kernel");
}

extern "C" size_t ra_outer_sum_workspace(int n_in, int n_out, int R) {
  if (n_in < 1 || n_out < 1 || R < 1) return 0;
  const int z = outer_sum_chunks(n_in, n_out, R);
  return z > 1 ? (size_t)z * ((size_t)n_in * n_out + n_out) * sizeof(float) : 0;
}

extern "C" int ra_outer_sum_ex_f32(const float *A, size_t a_stride, int n_in, const float *D, size_t d_stride,
                                   int n_out, int R, void *ws, float *dW, float *db, void *stream) {
  if (n_in < 1 || n_out < 1 || R < 0 || !dW) return RA_ERR_INVALID_ARG;
  if (R > 0 && (!A || !D)) return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  const int z = ws ? outer_sum_chunks(n_in, n_out, R) : 1;
  const dim3 grid((n_out + kOsT - 1) / kOsT, (n_in + kOsT - 1) / kOsT, z);
  if (z == 1) {
    outer_sum_kernel<<<grid, 256, 0, s>>>(A, a_stride, n_in, D, d_stride, n_out, R, R, dW, db);
    return ra::finish_launch("outer_sum_kernel");
  }
  int rows = (R + z - 1) / z;
  rows = (rows + kOsR - 1) / kOsR * kOsR;
  float *pw = reinterpret_cast<float *>(ws);
  float *pb = pw + (size_t)z * n_in * n_out;
  outer_sum_kernel<<<grid, 256, 0, s>>>(A, a_stride, n_in, D, d_stride, n_out, R, rows, pw, db ? pb : nullptr);
  int rc = ra::finish_launch("outer_sum_kernel");
  if (rc != RA_OK) return rc;
  const size_t n = (size_t)n_in * n_out;
  outer_sum_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pw, z, n, dW);
  rc = ra::finish_launch("outer_sum_finalize_kernel");
  if (rc != RA_OK || !db) return rc;
  outer_sum_finalize_kernel<<<(n_out + 255) / 256, 256, 0, s>>>(pb, z, (size_t)n_out, db);
  return ra::finish_launch("outer_sum_finalize_kernel(db)");
}

extern "C" int ra_outer_sum_f32(const float *A, size_t a_stride, int n_in, const float *D, size_t d_stride, int n_out,
                                int R, float *dW, float *db, void *stream) {
  return ra_outer_sum_ex_f32(A, a_stride, n_in, D, d_stride, n_out, R, nullptr, dW, db, stream);
}

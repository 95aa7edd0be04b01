// Backward of one TRAINING-MODE conv block  y = pool(relu(bn_batch(conv3x3(x [, x2]) + b)))  (nnlib.py:229-253 / :372-400
// with phase_train = True) - the first building block of the backward pass (the reference gets these
// gradients from TensorFlow's autodiff, full_model.py:1049).  Plain fp32 CUDA-core kernels: correctness first, they are
// the precision reference for a later tensor-core version, like conv.cu is for conv_umma.cu.
//
//   ra_bn_train_block_bwd_f32   dy [B,H/p,W/p,C] -> d_raw [B,H,W,C], dgamma [C], dbeta [C]
//       through max-pool (gradient to the FIRST maximum of the window in row-major order, like TF / torch), ReLU and
//       tf.nn.batch_normalization on batch moments (the moments are differentiated like tf.nn.moments):
//       d_raw = gamma * rstd / N * (N * d_bn - dbeta - xhat * dgamma),  xhat = (raw - mean) * rstd.
//       Nothing but `raw` (the conv output incl. bias) and the batch statistics is kept from the forward pass; the
//       normalised / pooled activations are recomputed.
//   ra_conv3x3_bwd_weight_f32   x [, x2], d_raw -> dW [3,3,C1+C2,Cout] (conv-form HWIO, i.e. the filter
//       ra_conv3x3_f32 consumes), db [Cout]; upsample = 2 is the transposed-conv layer (zero-inserted input).
//   ra_subsample2_f32           S [B,2H,2W,C] -> dx [B,H,W,C] = S[:, off::2, off::2]
//   The data gradient itself is a convolution of d_raw with the flipped, transposed filter, ra_conv3x3_f32 (ops.py);
//   for the transposed-conv layer the input gradient sits at the ODD positions of that SAME convolution (the forward
//   tap offset is -2, so dZ[z] = conv_same(d_raw, flip(w))[z + 1] and dx[y] = dZ[2y]).
#include "common.cuh"

namespace {

constexpr int kT = 256;

// d_bn of one pooled output element: returns the gradient and the window position (py, px) it goes to.
// z = raw * inv + shift (BN output); pooled value = max over the window; ReLU after (max and ReLU commute).
template <int POOL>
__device__ __forceinline__ void pool_relu_route(const float *__restrict__ raw, int W, int C, size_t base, float inv,
                                                float shift, int relu, float g, float *g_out, int *pos_out,
                                                float *raw_at) {
  float best = -INFINITY, rbest = 0.f;
  int pos = 0;
#pragma unroll
  for (int py = 0; py < POOL; ++py)
#pragma unroll
    for (int px = 0; px < POOL; ++px) {
      const float r = raw[base + ((size_t)py * W + px) * C];
      const float z = fmaf(r, inv, shift);
      if (z > best) {  // strict: the first maximum wins
        best = z;
        pos = py * POOL + px;
        rbest = r;
      }
    }
  *g_out = (relu && !(best > 0.f)) ? 0.f : g;  // relu'(0) = 0
  *pos_out = pos;
  *raw_at = rbest;
}

// pass A: per-CTA partial sums of dbeta[c] = sum d_bn, dgamma[c] = sum d_bn * xhat  -> partial[cta][2][C] (double)
template <int POOL>
__global__ void __launch_bounds__(kT) bn_bwd_reduce_kernel(const float *__restrict__ raw, const float *__restrict__ dy,
                                                           const float *__restrict__ gamma,
                                                           const float *__restrict__ beta,
                                                           const float *__restrict__ mean, const float *__restrict__ var,
                                                           int B, int H, int W, int C, int relu, float eps,
                                                           double *__restrict__ partial) {
  // group g = blockIdx.y: its own B examples and its own (layer, step) BN copy
  {
    const size_t g = blockIdx.y;
    raw += g * (size_t)B * H * W * C;
    dy += g * (size_t)B * (H / POOL) * (W / POOL) * C;
    gamma += g * C; beta += g * C; mean += g * C; var += g * C;
    partial += g * (size_t)gridDim.x * 2 * C;
  }
  extern __shared__ double acc_s[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += kT) acc_s[i] = 0.0;
  __syncthreads();
  const int Ho = H / POOL, Wo = W / POOL;
  const size_t total = (size_t)B * Ho * Wo * C;
  for (size_t idx = (size_t)blockIdx.x * kT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kT) {
    const int c = (int)(idx % C);
    size_t pix = idx / C;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const float rstd = rsqrtf(var[c] + eps);
    const float inv = gamma[c] * rstd, shift = beta[c] - mean[c] * inv;
    const size_t base = (((size_t)b * H + (size_t)oy * POOL) * W + (size_t)ox * POOL) * C + c;
    float g, r;
    int pos;
    pool_relu_route<POOL>(raw, W, C, base, inv, shift, relu, dy[idx], &g, &pos, &r);
    if (g != 0.f) {
      const float xhat = (r - mean[c]) * rstd;
      atomicAdd(&acc_s[c], (double)g);               // shared-memory double atomics: order-dependent only in the last
      atomicAdd(&acc_s[C + c], (double)(g * xhat));  // bits of a double, invisible after the cast to float
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kT) partial[(size_t)blockIdx.x * 2 * C + i] = acc_s[i];
}

// ---- float4 versions (C % 4 == 0, 16-byte aligned tensors, C / 4 divides 256): a thread owns one channel quad for the
// whole launch (its BN parameters live in registers), walks pooled pixels with a fixed stride and keeps its partial sums
// in registers (double); the CTA adds them in a fixed order - no atomics, no per-element index arithmetic.
template <int POOL>
__device__ __forceinline__ void pool_relu_route4(const float *__restrict__ raw, int W, int C, size_t base,
                                                 const float inv[4], const float shift[4], int relu, const float4 dy4,
                                                 float g[4], int pos[4], float rbest[4], float4 win[POOL * POOL]) {
  float best[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    best[e] = -INFINITY;
    rbest[e] = 0.f;
    pos[e] = 0;
  }
#pragma unroll
  for (int py = 0; py < POOL; ++py)
#pragma unroll
    for (int px = 0; px < POOL; ++px)
      win[py * POOL + px] = __ldg(reinterpret_cast<const float4 *>(raw + base + ((size_t)py * W + px) * C));
#pragma unroll
  for (int q = 0; q < POOL * POOL; ++q) {
    const float r4[4] = {win[q].x, win[q].y, win[q].z, win[q].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float z = fmaf(r4[e], inv[e], shift[e]);
      if (z > best[e]) {  // strict: the first maximum wins
        best[e] = z;
        pos[e] = q;
        rbest[e] = r4[e];
      }
    }
  }
  const float d4[4] = {dy4.x, dy4.y, dy4.z, dy4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) g[e] = (relu && !(best[e] > 0.f)) ? 0.f : d4[e];
}

template <int POOL>
__global__ void __launch_bounds__(kT) bn_bwd_reduce4_kernel(const float *__restrict__ raw, const float *__restrict__ dy,
                                                            const float *__restrict__ gamma,
                                                            const float *__restrict__ beta,
                                                            const float *__restrict__ mean,
                                                            const float *__restrict__ var, int B, int H, int W, int C,
                                                            int relu, float eps, double *__restrict__ partial) {
  {
    const size_t g = blockIdx.y;
    raw += g * (size_t)B * H * W * C;
    dy += g * (size_t)B * (H / POOL) * (W / POOL) * C;
    gamma += g * C; beta += g * C; mean += g * C; var += g * C;
    partial += g * (size_t)gridDim.x * 2 * C;
  }
  __shared__ double red_s[kT * 8];
  const int cg_n = C >> 2, lanes = kT / cg_n;
  const int cg = threadIdx.x % cg_n, pl = threadIdx.x / cg_n;
  const int Ho = H / POOL, Wo = W / POOL;
  const size_t npix = (size_t)B * Ho * Wo;
  float inv[4], shift[4], mu[4], rstd[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cg * 4 + e;
    mu[e] = mean[c];
    rstd[e] = rsqrtf(var[c] + eps);
    inv[e] = gamma[c] * rstd[e];
    shift[e] = beta[c] - mu[e] * inv[e];
  }
  double sb[4] = {0.0, 0.0, 0.0, 0.0}, sg[4] = {0.0, 0.0, 0.0, 0.0};
  for (size_t p = (size_t)blockIdx.x * lanes + pl; p < npix; p += (size_t)gridDim.x * lanes) {
    const int ox = (int)(p % Wo);
    const size_t q = p / Wo;
    const int oy = (int)(q % Ho);
    const int b = (int)(q / Ho);
    const size_t base = (((size_t)b * H + (size_t)oy * POOL) * W + (size_t)ox * POOL) * C + cg * 4;
    const float4 dy4 = __ldg(reinterpret_cast<const float4 *>(dy + p * C + cg * 4));
    float g[4], rb[4];
    int pos[4];
    float4 win[POOL * POOL];
    pool_relu_route4<POOL>(raw, W, C, base, inv, shift, relu, dy4, g, pos, rb, win);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xhat = (rb[e] - mu[e]) * rstd[e];
      sb[e] += (double)g[e];
      sg[e] += (double)(g[e] * xhat);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    red_s[threadIdx.x * 8 + e] = sb[e];
    red_s[threadIdx.x * 8 + 4 + e] = sg[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kT) {  // i = which * C + c
    const int which = i / C, c = i - which * C;
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += red_s[(l * cg_n + (c >> 2)) * 8 + which * 4 + (c & 3)];
    partial[(size_t)blockIdx.x * 2 * C + i] = t;
  }
}

template <int POOL>
__global__ void __launch_bounds__(kT) bn_bwd_apply4_kernel(const float *__restrict__ raw, const float *__restrict__ dy,
                                                           const float *__restrict__ gamma,
                                                           const float *__restrict__ beta,
                                                           const float *__restrict__ mean, const float *__restrict__ var,
                                                           const float *__restrict__ dgamma,
                                                           const float *__restrict__ dbeta, int B, int H, int W, int C,
                                                           int relu, float eps, float *__restrict__ d_raw) {
  {
    const size_t g = blockIdx.y;
    raw += g * (size_t)B * H * W * C;
    d_raw += g * (size_t)B * H * W * C;
    dy += g * (size_t)B * (H / POOL) * (W / POOL) * C;
    gamma += g * C; beta += g * C; mean += g * C; var += g * C; dgamma += g * C; dbeta += g * C;
  }
  const int cg_n = C >> 2, lanes = kT / cg_n;
  const int cg = threadIdx.x % cg_n, pl = threadIdx.x / cg_n;
  const int Ho = H / POOL, Wo = W / POOL;
  const size_t npix = (size_t)B * Ho * Wo;
  const float n = (float)((size_t)B * H * W);
  float inv[4], shift[4], mu[4], rstd[4], db[4], dg[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cg * 4 + e;
    mu[e] = mean[c];
    rstd[e] = rsqrtf(var[c] + eps);
    inv[e] = gamma[c] * rstd[e];
    shift[e] = beta[c] - mu[e] * inv[e];
    db[e] = dbeta[c];
    dg[e] = dgamma[c];
  }
  for (size_t p = (size_t)blockIdx.x * lanes + pl; p < npix; p += (size_t)gridDim.x * lanes) {
    const int ox = (int)(p % Wo);
    const size_t q = p / Wo;
    const int oy = (int)(q % Ho);
    const int b = (int)(q / Ho);
    const size_t base = (((size_t)b * H + (size_t)oy * POOL) * W + (size_t)ox * POOL) * C + cg * 4;
    const float4 dy4 = __ldg(reinterpret_cast<const float4 *>(dy + p * C + cg * 4));
    float g[4], rb[4];
    int pos[4];
    float4 win[POOL * POOL];
    pool_relu_route4<POOL>(raw, W, C, base, inv, shift, relu, dy4, g, pos, rb, win);
#pragma unroll
    for (int py = 0; py < POOL; ++py)
#pragma unroll
      for (int px = 0; px < POOL; ++px) {
        const int w = py * POOL + px;
        const float r4[4] = {win[w].x, win[w].y, win[w].z, win[w].w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xhat = (r4[e] - mu[e]) * rstd[e];
          const float dbn = (w == pos[e]) ? g[e] : 0.f;
          o[e] = inv[e] / n * (n * dbn - db[e] - xhat * dg[e]);
        }
        *reinterpret_cast<float4 *>(d_raw + base + ((size_t)py * W + px) * C) = make_float4(o[0], o[1], o[2], o[3]);
      }
  }
}

// one CTA per (channel, group): the per-CTA partials are added by 128 threads (strided, then a fixed tree) in double
__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const double *__restrict__ partial, int ctas, int C,
                                                              float *__restrict__ dgamma, float *__restrict__ dbeta) {
  __shared__ double sb_s[128], sg_s[128];
  const int c = blockIdx.x;
  partial += (size_t)blockIdx.y * ctas * 2 * C;
  dgamma += (size_t)blockIdx.y * C;
  dbeta += (size_t)blockIdx.y * C;
  double sb = 0.0, sg = 0.0;
  for (int k = threadIdx.x; k < ctas; k += 128) {
    sb += partial[(size_t)k * 2 * C + c];
    sg += partial[(size_t)k * 2 * C + C + c];
  }
  sb_s[threadIdx.x] = sb;
  sg_s[threadIdx.x] = sg;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sb_s[threadIdx.x] += sb_s[threadIdx.x + o];
      sg_s[threadIdx.x] += sg_s[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dbeta[c] = (float)sb_s[0];
    dgamma[c] = (float)sg_s[0];
  }
}

// pass B: d_raw for every element of the window of every pooled output
template <int POOL>
__global__ void __launch_bounds__(kT) bn_bwd_apply_kernel(const float *__restrict__ raw, const float *__restrict__ dy,
                                                          const float *__restrict__ gamma, const float *__restrict__ beta,
                                                          const float *__restrict__ mean, const float *__restrict__ var,
                                                          const float *__restrict__ dgamma,
                                                          const float *__restrict__ dbeta, int B, int H, int W, int C,
                                                          int relu, float eps, float *__restrict__ d_raw) {
  {
    const size_t g = blockIdx.y;
    raw += g * (size_t)B * H * W * C;
    d_raw += g * (size_t)B * H * W * C;
    dy += g * (size_t)B * (H / POOL) * (W / POOL) * C;
    gamma += g * C; beta += g * C; mean += g * C; var += g * C; dgamma += g * C; dbeta += g * C;
  }
  const int Ho = H / POOL, Wo = W / POOL;
  const size_t total = (size_t)B * Ho * Wo * C;
  const float n = (float)((size_t)B * H * W);
  for (size_t idx = (size_t)blockIdx.x * kT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kT) {
    const int c = (int)(idx % C);
    size_t pix = idx / C;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int b = (int)(pix / Ho);
    const float m = mean[c], rstd = rsqrtf(var[c] + eps);
    const float inv = gamma[c] * rstd, shift = beta[c] - m * inv;
    const size_t base = (((size_t)b * H + (size_t)oy * POOL) * W + (size_t)ox * POOL) * C + c;
    float g, r;
    int pos;
    pool_relu_route<POOL>(raw, W, C, base, inv, shift, relu, dy[idx], &g, &pos, &r);
    const float db = dbeta[c], dg = dgamma[c];
#pragma unroll
    for (int py = 0; py < POOL; ++py)
#pragma unroll
      for (int px = 0; px < POOL; ++px) {
        const size_t at = base + ((size_t)py * W + px) * C;
        const float xhat = (raw[at] - m) * rstd;
        const float dbn = (py * POOL + px == pos) ? g : 0.f;
        d_raw[at] = inv / n * (n * dbn - db - xhat * dg);
      }
  }
}

// ---------------------------------------------------------------------------------------------- weight gradient
// dW[t][ci][co] = sum over output pixels (b, Y, X) of Z[b, Y+ky-up, X+kx-up, ci] * G[b, Y, X, co], Z = the zero-padded
// input, zero-inserted for up = 2 (Z[2y,2x] = x[y,x]; the tap offset is -up, the forward kernel's convention), t = ky*3+kx.
//
// A CTA walks spatial tiles of TH x kWgTW output pixels (persistent over tiles, grid.x), for one block of CI_B x CO_B
// channels (grid.y).  Per tile it stages the (TH+2) x (TW+2) input window [slot][CI_B] and the gradient tile
// [pixel][CO_B] in shared memory; every thread keeps an RI x RJ register tile for ALL NINE taps (9*RI*RJ accumulators):
// per pixel one vector load of G and nine of the shifted inputs feed 9*RI*RJ FMAs.  Lanes of a warp differ in the
// output-channel index first, so the input loads are broadcasts and the gradient loads are conflict-free.  Channel blocks
// narrower than 256 / ((CI_B/RI) * (CO_B/RJ)) threads are replicated over PG pixel groups (each takes every PG-th pixel)
// and summed through shared memory at the end.  Partials per CTA are added in a fixed order by the finalize kernel:
// bit-identical from run to run.
constexpr int kWgTWMax = 64;

struct WgParams {
  const float *x1, *x2, *g;
  int C1, C2, x1_bmod, N, Hin, Win, Cout, up;
  int CI_B, CO_B, TI, TJ, PG;  // channel block, threads along ci / co, pixel groups (TI*TJ*PG == 256)
  int stages;                  // shared-memory stages of the tile pipeline (2..4)
  int TH, TW;                  // tile rows x columns (TW = 64 / 32 / 16 by the map width; TH * TW = 256 or 128 pixels)
  int vec_x, vec_g;            // 16-byte cp.async staging of the input window / the gradient tile (channel counts % 4
                               // == 0, aligned pointers); both stages are double buffered either way
  int tiles_x, tiles_y, n_tiles, n_ci_blk;
  float *partial;     // [grid.x][9][Cin][Cout]
  float *db_partial;  // [grid.x][Cout] or nullptr
};

__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? 4 : 0;  // src-size 0: the destination word is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

// NKY = filter rows per CTA: 3 (all nine taps, grid.z = 1) or 1 (three taps, grid.z = 3: a third of the accumulators,
// so two CTAs fit an SM and twice the warps hide the shared-memory latency; the tiles are staged once per filter row)
template <int RI, int RJ, int NKY>
__global__ void __launch_bounds__(kT, (NKY == 1 || RI * RJ <= 8) ? 2 : 1) conv_bwd_weight_kernel(const WgParams p) {
  constexpr int NT = NKY * 3;
  const int ky0 = (int)blockIdx.z * NKY;
  extern __shared__ __align__(16) float wg_smem[];
  const int CI_B = p.CI_B, CO_B = p.CO_B;
  const int kWgTW = p.TW, kWgTWP = p.TW + 2;  // (runtime; the names are kept from the fixed-tile version)
  const int n_slots = (p.TH + 2) * kWgTWP, n_pix = p.TH * kWgTW;
  const int stage_floats = n_slots * CI_B + n_pix * CO_B;  // [slots][CI_B] then [pixels][CO_B]
  const int Cin = p.C1 + p.C2;
  const int Ho = p.Hin * p.up, Wo = p.Win * p.up;
  const int ci_blk = blockIdx.y % p.n_ci_blk, co_blk = blockIdx.y / p.n_ci_blk;
  const int ci0 = ci_blk * CI_B, co0 = co_blk * CO_B;
  const int nci = min(CI_B, Cin - ci0), nco = min(CO_B, p.Cout - co0);
  const int tj = threadIdx.x % p.TJ, ti = (threadIdx.x / p.TJ) % p.TI, pg = threadIdx.x / (p.TJ * p.TI);
  float acc[NT][RI][RJ];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < RJ; ++j) acc[t][i][j] = 0.f;
  float dbv[RJ];
#pragma unroll
  for (int j = 0; j < RJ; ++j) dbv[j] = 0.f;
  const bool do_db = (ci_blk == 0 && blockIdx.z == 0 && p.db_partial != nullptr && ti == 0);

  // Stage one tile: input window, slot (r, c) <-> zero-inserted pixel (y0 - up + r, x0 - up + c), and gradient tile.
  auto stage = [&](int tile, float *xs, float *gs) {
    int t = tile;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    const int b = t / p.tiles_y;
    const int y0 = ty * p.TH, x0 = tx * kWgTW;
    const int b1 = p.x1_bmod > 0 ? b % p.x1_bmod : b;
    if (p.vec_x) {
      const int c4n = CI_B >> 2;
      for (int u = threadIdx.x; u < n_slots * c4n; u += kT) {
        const int slot = u / c4n, c = (u - slot * c4n) << 2;
        const int r = slot / kWgTWP, col = slot - r * kWgTWP;
        const int zy = y0 - p.up + r, zx = x0 - p.up + col;
        bool ok = c < nci && zy >= 0 && zy < Ho && zx >= 0 && zx < Wo && (p.up == 1 || (((zy | zx) & 1) == 0));
        const int iy = p.up == 1 ? zy : zy >> 1, ix = p.up == 1 ? zx : zx >> 1;
        const int ci = ci0 + c;
        const float *src = p.x1;
        if (ok) {
          if (ci < p.C1)
            src = p.x1 + (((size_t)b1 * p.Hin + iy) * p.Win + ix) * p.C1 + ci;
          else
            src = p.x2 + (((size_t)b * p.Hin + iy) * p.Win + ix) * p.C2 + (ci - p.C1);
        }
        cp_async16(xs + slot * CI_B + c, src, ok);
      }
    } else {
      for (int idx = threadIdx.x; idx < n_slots * CI_B; idx += kT) {
        const int slot = idx / CI_B, c = idx - slot * CI_B;
        const int r = slot / kWgTWP, col = slot - r * kWgTWP;
        const int zy = y0 - p.up + r, zx = x0 - p.up + col;
        // 4-byte cp.async: asynchronous like the vector path, so the deeper pipeline really overlaps
        const bool ok = c < nci && zy >= 0 && zy < Ho && zx >= 0 && zx < Wo && (p.up == 1 || (((zy | zx) & 1) == 0));
        const float *src = p.x1;
        if (ok) {
          const int iy = p.up == 1 ? zy : zy >> 1, ix = p.up == 1 ? zx : zx >> 1;
          const int ci = ci0 + c;
          if (ci < p.C1)
            src = p.x1 + (((size_t)b1 * p.Hin + iy) * p.Win + ix) * p.C1 + ci;
          else
            src = p.x2 + (((size_t)b * p.Hin + iy) * p.Win + ix) * p.C2 + (ci - p.C1);
        }
        cp_async4(xs + idx, src, ok);
      }
    }
    if (p.vec_g) {
      const int o4n = CO_B >> 2;
      for (int u = threadIdx.x; u < n_pix * o4n; u += kT) {
        const int pix = u / o4n, c = (u - pix * o4n) << 2;
        const int py = pix / kWgTW, px = pix - py * kWgTW;
        const int Y = y0 + py, X = x0 + px;
        const bool ok = c < nco && Y < Ho && X < Wo;
        const float *src = ok ? p.g + (((size_t)b * Ho + Y) * Wo + X) * p.Cout + co0 + c : p.g;
        cp_async16(gs + pix * CO_B + c, src, ok);
      }
    } else {
      for (int idx = threadIdx.x; idx < n_pix * CO_B; idx += kT) {
        const int pix = idx / CO_B, c = idx - pix * CO_B;
        const int py = pix / kWgTW, px = pix - py * kWgTW;
        const int Y = y0 + py, X = x0 + px;
        float v = 0.f;
        if (c < nco && Y < Ho && X < Wo) v = __ldg(p.g + (((size_t)b * Ho + Y) * Wo + X) * p.Cout + co0 + c);
        gs[idx] = v;
      }
    }
  };

  auto compute = [&](const float *xs, const float *gs) {
    const float *xb = xs + ti * RI;
    const float *gb = gs + tj * RJ;
    // pixel pg, pg + PG, ...: PG and the tile width are powers of two, so the walk is a row-major double loop without
    // a division per pixel (PG <= TW: every row, columns pg, pg + PG, ..; PG > TW: one column, every (PG / TW)-th row)
    const int px_step = p.PG <= kWgTW ? p.PG : kWgTW, py_step = p.PG <= kWgTW ? 1 : p.PG / kWgTW;
    for (int py = pg / kWgTW; py < p.TH; py += py_step)
    for (int px = pg % kWgTW; px < kWgTW; px += px_step) {
      const int pix = py * kWgTW + px;
      float gv[RJ];
      if constexpr (RJ == 2) {  // 8-byte aligned: CO_B and tj * RJ are even
        const float2 t2 = *reinterpret_cast<const float2 *>(gb + pix * CO_B);
        gv[0] = t2.x;
        gv[1] = t2.y;
      } else if constexpr (RJ == 4) {
        const float4 t4 = *reinterpret_cast<const float4 *>(gb + pix * CO_B);
        gv[0] = t4.x;
        gv[1] = t4.y;
        gv[2] = t4.z;
        gv[3] = t4.w;
      } else {
#pragma unroll
        for (int j = 0; j < RJ; ++j) gv[j] = gb[pix * CO_B + j];
      }
      if (do_db) {
#pragma unroll
        for (int j = 0; j < RJ; ++j) dbv[j] += gv[j];
      }
      const float *xr = xb + (size_t)(py * kWgTWP + px) * CI_B;
#pragma unroll
      for (int ky = 0; ky < NKY; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float xv[RI];
          if constexpr (RI == 2) {
            const float2 t2 = *reinterpret_cast<const float2 *>(xr + ((ky0 + ky) * kWgTWP + kx) * CI_B);
            xv[0] = t2.x;
            xv[1] = t2.y;
          } else if constexpr (RI == 4) {  // 16-byte aligned: CI_B and ti * RI are multiples of 4
            const float4 t4 = *reinterpret_cast<const float4 *>(xr + ((ky0 + ky) * kWgTWP + kx) * CI_B);
            xv[0] = t4.x;
            xv[1] = t4.y;
            xv[2] = t4.z;
            xv[3] = t4.w;
          } else {
#pragma unroll
            for (int i = 0; i < RI; ++i) xv[i] = xr[((ky0 + ky) * kWgTWP + kx) * CI_B + i];
          }
#pragma unroll
          for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < RJ; ++j) acc[ky * 3 + kx][i][j] = fmaf(xv[i], gv[j], acc[ky * 3 + kx][i][j]);
        }
    }
  };

  {
    // p.stages shared-memory stages (2..4): the copies of tiles k+1 .. k+stages-1 (cp.async where the layout allows it)
    // are in flight while tile k is multiplied.  Two are enough for the register-tiled channel blocks (a tile is ~3 us
    // of FMAs); the 1-channel canvas layer is bandwidth-bound and needs the deeper queue to keep HBM busy.
    const int S = p.stages;
    auto tile_of = [&](int it) { return (int)blockIdx.x + it * (int)gridDim.x; };
    for (int st = 0; st < S - 1; ++st) {
      float *nb = wg_smem + st * stage_floats;
      if (tile_of(st) < p.n_tiles) stage(tile_of(st), nb, nb + n_slots * CI_B);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; tile_of(it) < p.n_tiles; ++it) {
      const int nxt = it + S - 1;
      float *nb = wg_smem + (nxt % S) * stage_floats;
      if (tile_of(nxt) < p.n_tiles) stage(tile_of(nxt), nb, nb + n_slots * CI_B);
      asm volatile("cp.async.commit_group;" ::: "memory");
      // everything but the newest S-1 groups has landed
      if (S == 2)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      else if (S == 3)
        asm volatile("cp.async.wait_group 2;" ::: "memory");
      else
        asm volatile("cp.async.wait_group 3;" ::: "memory");
      __syncthreads();
      const float *cb = wg_smem + (it % S) * stage_floats;
      compute(cb, cb + n_slots * CI_B);
      __syncthreads();  // this stage may be overwritten by the copies issued in the next iteration
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  // ---- sum the pixel groups through shared memory (fixed order), write this CTA's partial
  __syncthreads();
  float *red = wg_smem;  // [PG][NT][CI_B][CO_B]  (+ [PG][CO_B] for db)
  const int blk = NT * CI_B * CO_B;
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < RJ; ++j) red[(size_t)pg * blk + (t * CI_B + ti * RI + i) * CO_B + tj * RJ + j] = acc[t][i][j];
  float *dbred = red + (size_t)p.PG * blk;
  if (do_db) {
#pragma unroll
    for (int j = 0; j < RJ; ++j) dbred[pg * CO_B + tj * RJ + j] = dbv[j];
  }
  __syncthreads();
  float *out = p.partial + (size_t)blockIdx.x * 9 * Cin * p.Cout;
  for (int idx = threadIdx.x; idx < blk; idx += kT) {
    const int co = idx % CO_B, r = idx / CO_B, ci = r % CI_B, t = r / CI_B;
    if (ci >= nci || co >= nco) continue;
    float s = 0.f;
    for (int g = 0; g < p.PG; ++g) s += red[(size_t)g * blk + idx];
    out[((size_t)(ky0 * 3 + t) * Cin + ci0 + ci) * p.Cout + co0 + co] = s;
  }
  if (ci_blk == 0 && blockIdx.z == 0 && p.db_partial != nullptr)
    for (int co = threadIdx.x; co < nco; co += kT) {
      float s = 0.f;
      for (int g = 0; g < p.PG; ++g) s += dbred[g * CO_B + co];
      p.db_partial[(size_t)blockIdx.x * p.Cout + co0 + co] = s;
    }
}

// ---- single input channel (the canvas channel of the first controller layer: [N,H,W,1] x [N,H,W,Cout], N = T*B).
// 36 FMAs per 16 + 4 bytes: bandwidth-bound, so no shared-memory tiles: a thread owns one output-channel quad and walks
// runs of 8 consecutive pixels of an image row - eight independent 16-byte gradient loads in flight, the 3 x 10 input
// window of the run in registers - and keeps its 9 x 4 sums in registers; the CTA adds its threads in a fixed order and
// writes one partial, summed by conv_bwd_weight_finalize_kernel like the tiled kernel's.  Needs Cout % 4 == 0, Cout <= 64.
constexpr int kC1Run = 8;
__global__ void __launch_bounds__(kT) conv_bwd_weight_c1_kernel(const float *__restrict__ x, int x_bmod,
                                                                const float *__restrict__ g, int N, int H, int W, int Cout,
                                                                float *__restrict__ partial, float *__restrict__ db_partial) {
  extern __shared__ float c1_sm[];  // [lanes][Cout / 4][40]
  const int q_n = Cout >> 2, lanes = kT / q_n;
  const int q = threadIdx.x % q_n, pl = threadIdx.x / q_n;
  float acc[9][4], dbv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
  const int runs_per_row = (W + kC1Run - 1) / kC1Run;
  const long long n_runs = (long long)N * H * runs_per_row;
  for (long long r = (long long)blockIdx.x * lanes + pl; pl < lanes && r < n_runs; r += (long long)gridDim.x * lanes) {
    const int xr = (int)(r % runs_per_row) * kC1Run;
    const long long row = r / runs_per_row;  // n * H + y
    const int y = (int)(row % H);
    const long long n = row / H;
    const long long nx = x_bmod > 0 ? n % x_bmod : n;
    float4 gv[kC1Run];
    const float *gp = g + ((size_t)row * W + xr) * Cout + q * 4;
#pragma unroll
    for (int i = 0; i < kC1Run; ++i)
      gv[i] = (xr + i < W) ? __ldg(reinterpret_cast<const float4 *>(gp + (size_t)i * Cout)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float win[3][kC1Run + 2];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      const float *xp = x + ((size_t)nx * H + (yy < 0 ? 0 : (yy >= H ? H - 1 : yy))) * W;
      const bool row_ok = yy >= 0 && yy < H;
#pragma unroll
      for (int i = 0; i < kC1Run + 2; ++i) {
        const int xx = xr + i - 1;
        win[ky][i] = (row_ok && xx >= 0 && xx < W) ? __ldg(xp + xx) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < kC1Run; ++i) {
      const float g4[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) dbv[e] += g4[e];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[ky * 3 + kx][e] = fmaf(win[ky][i + kx], g4[e], acc[ky * 3 + kx][e]);
    }
  }
  float *mine = c1_sm + ((size_t)pl * q_n + q) * 40;
  if (pl < lanes) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) mine[t * 4 + e] = acc[t][e];
#pragma unroll
    for (int e = 0; e < 4; ++e) mine[36 + e] = dbv[e];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 10 * Cout; o += kT) {  // o = t * Cout + co, t == 9: the bias gradient
    const int t = o / Cout, co = o - t * Cout;
    float sum = 0.f;
    for (int l = 0; l < lanes; ++l) sum += c1_sm[((size_t)l * q_n + (co >> 2)) * 40 + t * 4 + (co & 3)];
    if (t < 9)
      partial[(size_t)blockIdx.x * 9 * Cout + o] = sum;
    else if (db_partial != nullptr)
      db_partial[(size_t)blockIdx.x * Cout + co] = sum;
  }
}

// channel blocking of the weight-gradient kernel for a layer shape
struct WgPlan {
  int RI, RJ, CI_B, CO_B, TI, TJ, PG, n_ci_blk, n_co_blk, ctas, TH, TW, stages, nky;
  size_t smem;
};

WgPlan wg_plan(int N, int Ho, int Wo, int Cin, int Cout) {
  WgPlan w;
  if (Cin == 1) {  // the canvas channel of the first controller layer: 1 x 4 threads (4 channels each), 64 pixel groups
    w.RI = 1;
    w.RJ = 4;
    w.CI_B = 1;
    w.CO_B = 16;
  } else {
    // 4 x 4 register tiles: 144 FMAs per 10 16-byte shared-memory reads and one pixel's index arithmetic - the 2 x 2
    // version (RA_WGRAD_TILE=2) spends as many issue slots on loads and addresses as on FMAs
    static const int tile = getenv("RA_WGRAD_TILE") ? atoi(getenv("RA_WGRAD_TILE")) : 4;
    w.RI = (tile == 2) ? 2 : 4;
    w.RJ = (tile == 2 || tile == 42) ? 2 : 4;  // 42: 4 x 2 tiles, half the accumulators, two CTAs per SM (experiment)
    w.CI_B = Cin <= 16 ? 16 : 32;
    w.CO_B = Cout <= 16 ? 16 : 32;
  }
  {
    // MEASURED (KITTI B=32): three taps per CTA at two CTAs per SM is slower - 33.5 ms against 22.0 ms for all the
    // weight gradients: the tiles are staged three times and the FMA : load ratio falls to 12 : 1.  Opt-in only.
    static const int nky_env = getenv("RA_WGRAD_NKY") ? atoi(getenv("RA_WGRAD_NKY")) : 3;
    w.nky = (w.RI == 4 && w.RJ == 4 && nky_env == 1) ? 1 : 3;
  }
  w.TI = w.CI_B / w.RI;
  w.TJ = w.CO_B / w.RJ;
  w.PG = kT / (w.TI * w.TJ);
  w.n_ci_blk = (Cin + w.CI_B - 1) / w.CI_B;
  w.n_co_blk = (Cout + w.CO_B - 1) / w.CO_B;
  // tile: as wide as the map allows (64 / 32 / 16 columns), 256 pixels - 128 for the 32-wide channel blocks, whose two
  // stages must fit twice per SM
  w.TW = Wo > 32 ? 64 : (Wo > 16 ? 32 : 16);
  // 256 pixels per tile; 128 for the 32-wide channel blocks when two CTAs (two stages each) must fit an SM
  const int pix = ((w.RI != 4 || w.RJ != 4 || w.nky == 1) && (w.CI_B > 16 || w.CO_B > 16)) ? 128 : 256;
  w.TH = pix / w.TW;
  const size_t stage = ((size_t)(w.TH + 2) * (w.TW + 2) * w.CI_B + (size_t)w.TH * w.TW * w.CO_B) * sizeof(float);
  const size_t red = ((size_t)w.PG * w.nky * 3 * w.CI_B * w.CO_B + (size_t)w.PG * w.CO_B) * sizeof(float);
  w.stages = (w.CI_B == 1) ? 4 : 2;
  w.smem = w.stages * stage > red ? w.stages * stage : red;
  const size_t n_tiles = (size_t)N * ((Ho + w.TH - 1) / w.TH) * ((Wo + w.TW - 1) / w.TW);
  // resident CTAs per SM: 2 (register / shared-memory bound), 1 for the 4 x 4 tiles (~180 registers per thread)
  size_t cap = (size_t)ra::kNumSMs * ((w.RI == 4 && w.RJ == 4 && w.nky == 3) ? 1 : 2) /
               ((size_t)w.n_ci_blk * w.n_co_blk * (size_t)(3 / w.nky));
  if (cap < 16) cap = 16;
  w.ctas = (int)(n_tiles < cap ? (n_tiles < 1 ? 1 : n_tiles) : cap);
  return w;
}

// dw[i] = sum over chunks (fixed order, double); same for db
__global__ void conv_bwd_weight_finalize_kernel(const float *__restrict__ partial, int chunks, size_t n,
                                                float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < chunks; ++k) s += (double)partial[(size_t)k * n + i];
  out[i] = (float)s;
}

__global__ void subsample2_kernel(const float *__restrict__ src, int B, int H, int W, int C, int off,
                                  float *__restrict__ dst) {
  const size_t total = (size_t)B * H * W * C;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    size_t pix = idx / C;
    const int x = (int)(pix % W);
    pix /= W;
    const int y = (int)(pix % H);
    const int b = (int)(pix / H);
    dst[idx] = src[(((size_t)b * 2 * H + 2 * y + off) * 2 * W + 2 * x + off) * C + c];
  }
}

// out[ky][kx][co][ci] = w[2-ky][2-kx][ci][co]: the filter of the data-gradient convolution (and the map between the
// conv-form filter of a transposed-conv layer and TensorFlow's [kh,kw,Cout,Cin] layout, nnlib.py:320-325)
__global__ void filter_flip_transpose_kernel(const float *__restrict__ w, int Ci, int Co, float *__restrict__ out) {
  const int n = 9 * Ci * Co;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int ci = idx % Ci;
    int r = idx / Ci;
    const int co = r % Co;
    const int t = r / Co;
    out[idx] = w[((size_t)(8 - t) * Ci + ci) * Co + co];
  }
}

int bwd_ctas(size_t total) {
  size_t want = (total + kT - 1) / kT;
  const size_t cap = (size_t)ra::kNumSMs * 4;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" size_t ra_bn_train_block_bwd_workspace(int B, int H, int W, int C, int pool) {
  if (B < 1 || H < 1 || W < 1 || C < 1 || (pool != 1 && pool != 2)) return 0;
  return (size_t)bwd_ctas((size_t)B * (H / pool) * (W / pool) * C) * 2 * C * sizeof(double);
}

extern "C" size_t ra_bn_train_block_bwd_grouped_workspace(int G, int B, int H, int W, int C, int pool) {
  return (size_t)(G < 1 ? 0 : G) * ra_bn_train_block_bwd_workspace(B, H, W, C, pool);
}

extern "C" int ra_bn_train_block_bwd_f32(const float *raw, const float *dy, const float *gamma, const float *beta,
                                         const float *mean, const float *var, int B, int H, int W, int C, int pool,
                                         int relu, float eps, void *ws, float *d_raw, float *dgamma, float *dbeta,
                                         void *stream) {
  return ra_bn_train_block_bwd_grouped_f32(raw, dy, gamma, beta, mean, var, 1, B, H, W, C, pool, relu, eps, ws, d_raw,
                                           dgamma, dbeta, stream);
}

extern "C" int ra_bn_train_block_bwd_grouped_f32(const float *raw, const float *dy, const float *gamma,
                                                 const float *beta, const float *mean, const float *var, int G, int B,
                                                 int H, int W, int C, int pool, int relu, float eps, void *ws,
                                                 float *d_raw, float *dgamma, float *dbeta, void *stream) {
  if (B < 0 || H < 1 || W < 1 || C < 1 || G < 1 || G > 65535) return RA_ERR_INVALID_ARG;
  if (pool != 1 && pool != 2) return RA_ERR_UNSUPPORTED;
  if (pool == 2 && ((H | W) & 1)) return RA_ERR_UNSUPPORTED;
  if ((size_t)2 * C * sizeof(double) > 48 * 1024) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  if (!raw || !dy || !gamma || !beta || !mean || !var || !ws || !d_raw || !dgamma || !dbeta) return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  const size_t total = (size_t)B * (H / pool) * (W / pool) * C;
  const int ctas = bwd_ctas(total);
  double *partial = reinterpret_cast<double *>(ws);
  const size_t smem = (size_t)2 * C * sizeof(double);
  // float4 path: whole channel quads per thread (every conv layer of the model but the 1-channel mask layer)
  const bool vec4 = (C & 3) == 0 && (kT % (C >> 2)) == 0 &&
                    ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(dy) |
                      reinterpret_cast<uintptr_t>(d_raw)) & 15) == 0 && getenv("RA_BN_BWD_SCALAR") == nullptr;
  if (vec4) {
    if (pool == 2)
      bn_bwd_reduce4_kernel<2><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, B, H, W, C, relu, eps, partial);
    else
      bn_bwd_reduce4_kernel<1><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, B, H, W, C, relu, eps, partial);
  } else if (pool == 2) {
    bn_bwd_reduce_kernel<2><<<dim3(ctas, G), kT, smem, s>>>(raw, dy, gamma, beta, mean, var, B, H, W, C, relu, eps,
                                                            partial);
  } else {
    bn_bwd_reduce_kernel<1><<<dim3(ctas, G), kT, smem, s>>>(raw, dy, gamma, beta, mean, var, B, H, W, C, relu, eps,
                                                            partial);
  }
  int rc = ra::finish_launch("bn_bwd_reduce_kernel");
  if (rc != RA_OK) return rc;
  bn_bwd_finalize_kernel<<<dim3(C, G), 128, 0, s>>>(partial, ctas, C, dgamma, dbeta);
  rc = ra::finish_launch("bn_bwd_finalize_kernel");
  if (rc != RA_OK) return rc;
  if (vec4) {
    if (pool == 2)
      bn_bwd_apply4_kernel<2><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, dgamma, dbeta, B, H, W, C,
                                                           relu, eps, d_raw);
    else
      bn_bwd_apply4_kernel<1><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, dgamma, dbeta, B, H, W, C,
                                                           relu, eps, d_raw);
  } else if (pool == 2) {
    bn_bwd_apply_kernel<2><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, dgamma, dbeta, B, H, W, C, relu,
                                                        eps, d_raw);
  } else {
    bn_bwd_apply_kernel<1><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, dgamma, dbeta, B, H, W, C, relu,
                                                        eps, d_raw);
  }
  return ra::finish_launch("bn_bwd_apply_kernel");
}

extern "C" size_t ra_conv3x3_bwd_weight_workspace(int B, int Hin, int Win, int Cin, int Cout, int upsample) {
  if (B < 1 || Hin < 1 || Win < 1 || Cin < 1 || Cout < 1 || (upsample != 1 && upsample != 2)) return 0;
  const int chunks = wg_plan(B, Hin * upsample, Win * upsample, Cin, Cout).ctas;
  return ((size_t)chunks * 9 * Cin * Cout + (size_t)chunks * Cout) * sizeof(float);
}

extern "C" int ra_conv3x3_bwd_weight_f32(const float *x1, int C1, const float *x2, int C2, const float *d_out, int B,
                                         int Hin, int Win, int Cout, int upsample, void *ws, float *dw, float *db,
                                         void *stream) {
  return ra_conv3x3_bwd_weight_ex_f32(x1, C1, 0, x2, C2, d_out, B, Hin, Win, Cout, upsample, ws, dw, db, stream);
}

extern "C" int ra_conv3x3_bwd_weight_ex_f32(const float *x1, int C1, int x1_bmod, const float *x2, int C2,
                                            const float *d_out, int B, int Hin, int Win, int Cout, int upsample,
                                            void *ws, float *dw, float *db, void *stream) {
  if (x1_bmod < 0) return RA_ERR_INVALID_ARG;
  if (C1 < 1 || C2 < 0 || B < 0 || Hin < 1 || Win < 1 || Cout < 1 || !dw) return RA_ERR_INVALID_ARG;
  if (upsample != 1 && upsample != 2) return RA_ERR_UNSUPPORTED;
  cudaStream_t s = ra::as_stream(stream);
  const int Cin = C1 + C2;
  if (B == 0) {  // an empty batch contributes nothing: zero gradients (empty tensors may carry null pointers)
    cudaMemsetAsync(dw, 0, (size_t)9 * Cin * Cout * sizeof(float), s);
    if (db) cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), s);
    return ra::finish_launch("cudaMemsetAsync(dw)");
  }
  if (!x1 || (C2 > 0 && !x2) || !d_out || !ws) return RA_ERR_INVALID_ARG;
  const int Ho = Hin * upsample, Wo = Win * upsample;
  const WgPlan w = wg_plan(B, Ho, Wo, Cin, Cout);
  const int chunks = w.ctas;
  if ((size_t)w.n_ci_blk * w.n_co_blk > 65535) return RA_ERR_UNSUPPORTED;
  float *partial = reinterpret_cast<float *>(ws);
  float *db_partial = partial + (size_t)chunks * 9 * Cin * Cout;
  if (Cin == 1 && upsample == 1 && (Cout & 3) == 0 && Cout <= 64 && (kT % (Cout >> 2)) == 0 &&
      (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && getenv("RA_WGRAD_C1_TILED") == nullptr) {
    const size_t c1_smem = (size_t)kT * 40 * sizeof(float);
    static bool c1_attr = false;
    if (!c1_attr) {
      cudaFuncSetAttribute(conv_bwd_weight_c1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c1_smem);
      c1_attr = true;
    }
    conv_bwd_weight_c1_kernel<<<chunks, kT, c1_smem, s>>>(x1, x1_bmod, d_out, B, Hin, Win, Cout, partial,
                                                          db ? db_partial : nullptr);
    int rc1 = ra::finish_launch("conv_bwd_weight_c1_kernel");
    if (rc1 != RA_OK) return rc1;
    const size_t n1 = (size_t)9 * Cout;
    conv_bwd_weight_finalize_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, s>>>(partial, chunks, n1, dw);
    rc1 = ra::finish_launch("conv_bwd_weight_finalize_kernel");
    if (rc1 != RA_OK || !db) return rc1;
    conv_bwd_weight_finalize_kernel<<<(Cout + 255) / 256, 256, 0, s>>>(db_partial, chunks, (size_t)Cout, db);
    return ra::finish_launch("conv_bwd_weight_finalize_kernel(db)");
  }
  WgParams p;
  p.x1 = x1; p.x2 = x2; p.g = d_out;
  p.C1 = C1; p.C2 = C2; p.x1_bmod = x1_bmod; p.N = B; p.Hin = Hin; p.Win = Win; p.Cout = Cout; p.up = upsample;
  p.CI_B = w.CI_B; p.CO_B = w.CO_B; p.TI = w.TI; p.TJ = w.TJ; p.PG = w.PG;
  p.tiles_x = (Wo + w.TW - 1) / w.TW;
  p.TW = w.TW;
  p.TH = w.TH;
  p.stages = w.stages;
  p.tiles_y = (Ho + w.TH - 1) / w.TH;
  p.vec_x = ((C1 & 3) == 0 && (C2 & 3) == 0 && (w.CI_B & 3) == 0 &&
             ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2)) & 15) == 0) ? 1 : 0;
  p.vec_g = ((Cout & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) ? 1 : 0;
  p.n_tiles = B * p.tiles_x * p.tiles_y;
  p.n_ci_blk = w.n_ci_blk;
  p.partial = partial;
  p.db_partial = db ? db_partial : nullptr;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(conv_bwd_weight_kernel<2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    cudaFuncSetAttribute(conv_bwd_weight_kernel<4, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv_bwd_weight_kernel<4, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    cudaFuncSetAttribute(conv_bwd_weight_kernel<4, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    cudaFuncSetAttribute(conv_bwd_weight_kernel<1, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    attr_done = true;
  }
  const dim3 grid(chunks, w.n_ci_blk * w.n_co_blk, 3 / w.nky);
  if (w.RI == 2)
    conv_bwd_weight_kernel<2, 2, 3><<<grid, kT, w.smem, s>>>(p);
  else if (w.RI == 4 && w.RJ == 2)
    conv_bwd_weight_kernel<4, 2, 3><<<grid, kT, w.smem, s>>>(p);
  else if (w.RI == 4 && w.nky == 1)
    conv_bwd_weight_kernel<4, 4, 1><<<grid, kT, w.smem, s>>>(p);
  else if (w.RI == 4)
    conv_bwd_weight_kernel<4, 4, 3><<<grid, kT, w.smem, s>>>(p);
  else
    conv_bwd_weight_kernel<1, 4, 3><<<grid, kT, w.smem, s>>>(p);
  int rc = ra::finish_launch("conv_bwd_weight_kernel");
  if (rc != RA_OK) return rc;
  const size_t n = (size_t)9 * Cin * Cout;
  conv_bwd_weight_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(partial, chunks, n, dw);
  rc = ra::finish_launch("conv_bwd_weight_finalize_kernel");
  if (rc != RA_OK || !db) return rc;
  conv_bwd_weight_finalize_kernel<<<(Cout + 255) / 256, 256, 0, s>>>(db_partial, chunks, (size_t)Cout, db);
  return ra::finish_launch("conv_bwd_weight_finalize_kernel(db)");
}

extern "C" int ra_subsample2_f32(const float *src, int B, int H, int W, int C, int off, float *dst, void *stream) {
  if (B < 0 || H < 1 || W < 1 || C < 1 || (off != 0 && off != 1)) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!src || !dst) return RA_ERR_INVALID_ARG;
  const size_t total = (size_t)B * H * W * C;
  subsample2_kernel<<<bwd_ctas(total), kT, 0, ra::as_stream(stream)>>>(src, B, H, W, C, off, dst);
  return ra::finish_launch("subsample2_kernel");
}

extern "C" int ra_filter_flip_transpose_f32(const float *w, int Ci, int Co, float *out, void *stream) {
  if (!w || !out || Ci < 1 || Co < 1) return RA_ERR_INVALID_ARG;
  if ((size_t)9 * Ci * Co > 0x7fffffffULL) return RA_ERR_UNSUPPORTED;
  filter_flip_transpose_kernel<<<bwd_ctas((size_t)9 * Ci * Co), kT, 0, ra::as_stream(stream)>>>(w, Ci, Co, out);
  return ra::finish_launch("filter_flip_transpose_kernel");
}

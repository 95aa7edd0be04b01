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

__global__ void bn_bwd_finalize_kernel(const double *__restrict__ partial, int ctas, int C, float *__restrict__ dgamma,
                                       float *__restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  partial += (size_t)blockIdx.y * ctas * 2 * C;
  dgamma += (size_t)blockIdx.y * C;
  dbeta += (size_t)blockIdx.y * C;
  double sb = 0.0, sg = 0.0;
  for (int k = 0; k < ctas; ++k) {
    sb += partial[(size_t)k * 2 * C + c];
    sg += partial[(size_t)k * 2 * C + C + c];
  }
  dbeta[c] = (float)sb;
  dgamma[c] = (float)sg;
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
// input, zero-inserted for up = 2 (Z[2y,2x] = x[y,x]; the tap offset is -up, the forward kernel's convention), t = ky*3+kx.  Grid (pixel chunk, tap, channel block); a CTA stages 32 output pixels of G and of
// the shifted input in shared memory and every thread accumulates an (8 ci) x (6 co) register tile:
// thread (ti, tj) of the 16 x 16 CTA owns ci = ti + 16*i, co = tj + 16*j inside the channel block of 128 x 96.
constexpr int kWgPix = 32;
constexpr int kWgCi = 128, kWgCo = 96;
constexpr int kWgI = kWgCi / 16, kWgJ = kWgCo / 16;

__global__ void __launch_bounds__(kT) conv_bwd_weight_kernel(const float *__restrict__ x1, int C1, int x1_bmod,
                                                             const float *__restrict__ x2, int C2,
                                                             const float *__restrict__ g, int B, int Hin, int Win,
                                                             int Cout, int up, int n_ci_blk,
                                                             float *__restrict__ partial /* [chunks][9][Cin][Cout] */,
                                                             float *__restrict__ db_partial /* [chunks][Cout] */) {
  __shared__ float xs[kWgPix][kWgCi + 1];
  __shared__ float gs[kWgPix][kWgCo + 1];
  const int Cin = C1 + C2;
  const int Ho = Hin * up, Wo = Win * up;
  const int t = blockIdx.y, ky = t / 3, kx = t - ky * 3;
  const int cb = blockIdx.z, ci_blk = cb % n_ci_blk, co_blk = cb / n_ci_blk;
  const int ci0 = ci_blk * kWgCi, co0 = co_blk * kWgCo;
  const int nci = min(kWgCi, Cin - ci0), nco = min(kWgCo, Cout - co0);
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  float acc[kWgI][kWgJ];
#pragma unroll
  for (int i = 0; i < kWgI; ++i)
#pragma unroll
    for (int j = 0; j < kWgJ; ++j) acc[i][j] = 0.f;
  float db_acc = 0.f;  // threads < nco of the (tap 0, ci block 0) CTAs: column sums of G for the bias gradient
  const bool do_db = (t == 0 && ci_blk == 0 && db_partial != nullptr);
  const size_t npix = (size_t)B * Ho * Wo;
  const size_t per = (npix + gridDim.x - 1) / gridDim.x;
  const size_t p_begin = (size_t)blockIdx.x * per, p_end = min(npix, p_begin + per);
  for (size_t p0 = p_begin; p0 < p_end; p0 += kWgPix) {
    const int np = (int)min((size_t)kWgPix, p_end - p0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kWgPix * nco; idx += kT) {
      const int p = idx / nco, c = idx - p * nco;
      gs[p][c] = (p < np) ? g[(p0 + p) * Cout + co0 + c] : 0.f;
    }
    for (int idx = threadIdx.x; idx < kWgPix * nci; idx += kT) {
      const int p = idx / nci, c = idx - p * nci;
      float v = 0.f;
      if (p < np) {
        size_t q = p0 + p;
        const int X = (int)(q % Wo);
        q /= Wo;
        const int Y = (int)(q % Ho);
        const int b = (int)(q / Ho);
        const int zy = Y + ky - up, zx = X + kx - up;  // coordinate in the zero-inserted input (conv.cu: oy - up + r)
        if (zy >= 0 && zy < Ho && zx >= 0 && zx < Wo && (zy % up) == 0 && (zx % up) == 0) {
          const size_t src = ((size_t)b * Hin + zy / up) * Win + zx / up;
          const int ci = ci0 + c;
          if (ci < C1) {
            const size_t src1 = ((size_t)(x1_bmod > 0 ? b % x1_bmod : b) * Hin + zy / up) * Win + zx / up;
            v = x1[src1 * C1 + ci];
          } else {
            v = x2[src * C2 + (ci - C1)];
          }
        }
      }
      xs[p][c] = v;
    }
    __syncthreads();
    for (int p = 0; p < np; ++p) {
      float xv[kWgI], gv[kWgJ];
#pragma unroll
      for (int i = 0; i < kWgI; ++i) xv[i] = (ti + 16 * i < nci) ? xs[p][ti + 16 * i] : 0.f;
#pragma unroll
      for (int j = 0; j < kWgJ; ++j) gv[j] = (tj + 16 * j < nco) ? gs[p][tj + 16 * j] : 0.f;
#pragma unroll
      for (int i = 0; i < kWgI; ++i)
#pragma unroll
        for (int j = 0; j < kWgJ; ++j) acc[i][j] = fmaf(xv[i], gv[j], acc[i][j]);
    }
    if (do_db && threadIdx.x < nco)
      for (int p = 0; p < np; ++p) db_acc += gs[p][threadIdx.x];
  }
  float *out = partial + ((size_t)blockIdx.x * 9 + t) * Cin * Cout;
#pragma unroll
  for (int i = 0; i < kWgI; ++i)
#pragma unroll
    for (int j = 0; j < kWgJ; ++j) {
      const int ci = ti + 16 * i, co = tj + 16 * j;
      if (ci < nci && co < nco) out[(size_t)(ci0 + ci) * Cout + co0 + co] = acc[i][j];
    }
  if (do_db && threadIdx.x < nco) db_partial[(size_t)blockIdx.x * Cout + co0 + threadIdx.x] = db_acc;
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

int wg_chunks(size_t npix) {
  size_t want = (npix + 8 * kWgPix - 1) / (8 * kWgPix);  // at least 8 staged tiles per CTA
  const size_t cap = (size_t)ra::kNumSMs;
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
  if (pool == 2)
    bn_bwd_reduce_kernel<2><<<dim3(ctas, G), kT, smem, s>>>(raw, dy, gamma, beta, mean, var, B, H, W, C, relu, eps,
                                                            partial);
  else
    bn_bwd_reduce_kernel<1><<<dim3(ctas, G), kT, smem, s>>>(raw, dy, gamma, beta, mean, var, B, H, W, C, relu, eps,
                                                            partial);
  int rc = ra::finish_launch("bn_bwd_reduce_kernel");
  if (rc != RA_OK) return rc;
  bn_bwd_finalize_kernel<<<dim3((C + 127) / 128, G), 128, 0, s>>>(partial, ctas, C, dgamma, dbeta);
  rc = ra::finish_launch("bn_bwd_finalize_kernel");
  if (rc != RA_OK) return rc;
  if (pool == 2)
    bn_bwd_apply_kernel<2><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, dgamma, dbeta, B, H, W, C, relu,
                                                        eps, d_raw);
  else
    bn_bwd_apply_kernel<1><<<dim3(ctas, G), kT, 0, s>>>(raw, dy, gamma, beta, mean, var, dgamma, dbeta, B, H, W, C, relu,
                                                        eps, d_raw);
  return ra::finish_launch("bn_bwd_apply_kernel");
}

extern "C" size_t ra_conv3x3_bwd_weight_workspace(int B, int Hin, int Win, int Cin, int Cout, int upsample) {
  if (B < 1 || Hin < 1 || Win < 1 || Cin < 1 || Cout < 1 || (upsample != 1 && upsample != 2)) return 0;
  const int chunks = wg_chunks((size_t)B * Hin * upsample * Win * upsample);
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
  const size_t npix = (size_t)B * Hin * upsample * Win * upsample;
  const int chunks = wg_chunks(npix);
  const int n_ci_blk = (Cin + kWgCi - 1) / kWgCi, n_co_blk = (Cout + kWgCo - 1) / kWgCo;
  if ((size_t)n_ci_blk * n_co_blk > 65535) return RA_ERR_UNSUPPORTED;
  float *partial = reinterpret_cast<float *>(ws);
  float *db_partial = partial + (size_t)chunks * 9 * Cin * Cout;
  conv_bwd_weight_kernel<<<dim3(chunks, 9, n_ci_blk * n_co_blk), kT, 0, s>>>(x1, C1, x1_bmod, x2, C2, d_out, B, Hin, Win,
                                                                            Cout, upsample, n_ci_blk, partial,
                                                                            db ? db_partial : nullptr);
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

// Backward of the paste-back and of the Gaussian filters (third building block of the backward pass):
//
//   forward (full_model.py:738-741 with P = ones, :810-814 with the mask-head patch; modellib.py:581-641):
//       V = Fy P Fx^T  [H,W],   out = sigmoid(gamma * V - 5)
//   backward, given d_out:
//       dZ = d_out * out * (1 - out);   d_gamma = sum dZ * V;   dV = gamma * dZ
//       d_P  = Fy^T dV Fx                      [F,F]
//       d_Fy[y,i] = sum_x dV[y,x] R[i,x],  R = P Fx^T     [F,W]
//       d_Fx[x,j] = sum_i A[i,x] P[i,j],   A = Fy^T dV    [F,W]
//   and of filt[l,f] = exp(-(l - mu_f)^2 / (2 s2)) / sqrt(2 pi s2),  mu_f = ctr + (size + 1) / F * (f - (F-1)/2),
//   s2 = exp(lg_var)  (modellib.get_gaussian_filter):
//       d_mu_f = sum_l d_filt * filt * (l - mu_f) / s2;   d_ctr = sum_f d_mu_f;   d_size = sum_f d_mu_f (f - (F-1)/2) / F
//       d_lg_var = sum_{l,f} d_filt * filt * ((l - mu_f)^2 / (2 s2) - 1/2)
//   Filters are tap-major on the device: fy [B,F,H], fx [B,F,W] (ra_gaussian_filters_f32).  Entries the forward
//   pass stored as exact zeros (below exp(-30) of the peak) contribute nothing here either.
//   Plain fp32 kernels, one launch per stage; V is recomputed from R instead of inverting the sigmoid.
//   `disable_overwrite` (y_out gated by 1 - canvas) is not handled: it is off in every shipped config.
#include "common.cuh"

namespace {

constexpr int kMaxF = 64;
constexpr int kT = 256;

// R[b,i,x] = sum_j P[b,i,j] fx[b,j,x]   (P == nullptr: ones)
__global__ void __launch_bounds__(kT) pb_R_kernel(const float *__restrict__ patch, const float *__restrict__ fx, int W,
                                                  int F, float *__restrict__ R) {
  extern __shared__ float P_s[];  // [F][F]
  const int b = blockIdx.y;
  if (patch != nullptr)
    for (int i = threadIdx.x; i < F * F; i += kT) P_s[i] = patch[(size_t)b * F * F + i];
  __syncthreads();
  const int x = blockIdx.x * kT + threadIdx.x;
  if (x >= W) return;
  const float *fxb = fx + (size_t)b * F * W + x;
  float *Rb = R + (size_t)b * F * W + x;
  if (patch == nullptr) {
    float s = 0.f;
    for (int j = 0; j < F; ++j) s += fxb[(size_t)j * W];
    for (int i = 0; i < F; ++i) Rb[(size_t)i * W] = s;
    return;
  }
  for (int i = 0; i < F; ++i) {
    float s = 0.f;
    for (int j = 0; j < F; ++j) s = fmaf(P_s[i * F + j], fxb[(size_t)j * W], s);
    Rb[(size_t)i * W] = s;
  }
}

// one CTA per image row y: V, dZ, dV (stored), d_fy[:, y], the row's share of d_gamma
__global__ void __launch_bounds__(kT) pb_row_kernel(const float *__restrict__ d_out, const float *__restrict__ out,
                                                    size_t out_bstride, const float *__restrict__ fy,
                                                    const float *__restrict__ R, const float *__restrict__ gamma,
                                                    int gamma_stride, int H, int W, int F, int accumulate,
                                                    float *__restrict__ dV, float *__restrict__ d_fy,
                                                    float *__restrict__ dgamma_rows) {
  __shared__ float fy_s[kMaxF];
  __shared__ float red[32];
  __shared__ float dfy_s[kMaxF];
  const int y = blockIdx.x, b = blockIdx.y;
  for (int i = threadIdx.x; i < F; i += kT) fy_s[i] = fy[((size_t)b * F + i) * H + y];
  __syncthreads();
  const float g = gamma[(size_t)b * gamma_stride];
  const float *Rb = R + (size_t)b * F * W;
  const size_t row = (size_t)b * out_bstride + (size_t)y * W;
  float dg = 0.f;
  for (int x = threadIdx.x; x < W; x += kT) {
    float v = 0.f;
    for (int i = 0; i < F; ++i) v = fmaf(fy_s[i], Rb[(size_t)i * W + x], v);
    const float o = out[row + x];
    const float dz = d_out[row + x] * o * (1.0f - o);
    dg = fmaf(dz, v, dg);
    dV[((size_t)b * H + y) * W + x] = g * dz;
  }
  dg = ra::block_sum(dg, red);
  if (threadIdx.x == 0) dgamma_rows[(size_t)b * H + y] = dg;
  __syncthreads();  // this row of dV is complete (written by this CTA only)
  for (int i = 0; i < F; ++i) {
    float s = 0.f;
    for (int x = threadIdx.x; x < W; x += kT) s = fmaf(dV[((size_t)b * H + y) * W + x], Rb[(size_t)i * W + x], s);
    s = ra::block_sum(s, red);
    if (threadIdx.x == 0) dfy_s[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += kT) {
    float *dst = d_fy + ((size_t)b * F + i) * H + y;
    *dst = accumulate ? (*dst + dfy_s[i]) : dfy_s[i];
  }
}

// thread per column x: A[:, x] = Fy^T dV[:, x] in registers, stored; d_fx[:, x] = P^T A[:, x]
__global__ void __launch_bounds__(kT) pb_col_kernel(const float *__restrict__ dV, const float *__restrict__ fy,
                                                    const float *__restrict__ patch, int H, int W, int F,
                                                    int accumulate, float *__restrict__ A, float *__restrict__ d_fx) {
  extern __shared__ float P_s[];  // [F][F]
  const int b = blockIdx.y;
  if (patch != nullptr)
    for (int i = threadIdx.x; i < F * F; i += kT) P_s[i] = patch[(size_t)b * F * F + i];
  __syncthreads();
  const int x = blockIdx.x * kT + threadIdx.x;
  if (x >= W) return;
  float acc[kMaxF];
#pragma unroll
  for (int i = 0; i < kMaxF; ++i) acc[i] = 0.f;
  const float *fyb = fy + (size_t)b * F * H;
  for (int y = 0; y < H; ++y) {
    const float v = dV[((size_t)b * H + y) * W + x];
    if (v == 0.f) continue;
#pragma unroll
    for (int i = 0; i < kMaxF; ++i)
      if (i < F) acc[i] = fmaf(fyb[(size_t)i * H + y], v, acc[i]);
  }
#pragma unroll
  for (int i = 0; i < kMaxF; ++i)
    if (i < F) A[((size_t)b * F + i) * W + x] = acc[i];
  for (int j = 0; j < F; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxF; ++i)
      if (i < F) s = (patch != nullptr) ? fmaf(P_s[i * F + j], acc[i], s) : (s + acc[i]);
    float *dst = d_fx + ((size_t)b * F + j) * W + x;
    *dst = accumulate ? (*dst + s) : s;
  }
}

// d_P[b,i,j] = sum_x A[b,i,x] fx[b,j,x]: one warp per (i, j); d_gamma[b] = sum of the row shares (CTA 0)
__global__ void __launch_bounds__(kT) pb_dP_kernel(const float *__restrict__ A, const float *__restrict__ fx,
                                                   const float *__restrict__ dgamma_rows, int H, int W, int F,
                                                   float *__restrict__ d_patch, float *__restrict__ d_gamma) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  if (blockIdx.x == 0) {
    float s = 0.f;
    for (int y = threadIdx.x; y < H; y += kT) s += dgamma_rows[(size_t)b * H + y];
    s = ra::block_sum(s, red);
    if (threadIdx.x == 0) d_gamma[b] = s;
  }
  if (d_patch == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ij = blockIdx.x * (kT / 32) + warp;
  if (ij >= F * F) return;
  const int i = ij / F, j = ij - i * F;
  const float *Ai = A + ((size_t)b * F + i) * W, *fj = fx + ((size_t)b * F + j) * W;
  float s = 0.f;
  for (int x = lane; x < W; x += 32) s = fmaf(Ai[x], fj[x], s);
  s = ra::warp_sum(s);
  if (lane == 0) d_patch[(size_t)b * F * F + ij] = s;
}

// grid (axis, b): d_ctr, d_size, d_lg_var of one axis
__global__ void __launch_bounds__(kT) filters_bwd_kernel(const float *__restrict__ box, const float *__restrict__ fy,
                                                         const float *__restrict__ fx, const float *__restrict__ d_fy,
                                                         const float *__restrict__ d_fx, int H, int W, int F,
                                                         float *__restrict__ d_box) {
  __shared__ float red[32];
  const int axis = blockIdx.x, b = blockIdx.y;
  const int L = axis == 0 ? H : W;
  const float *f = (axis == 0 ? fy : fx) + (size_t)b * F * L;
  const float *df = (axis == 0 ? d_fy : d_fx) + (size_t)b * F * L;
  const float *bo = box + (size_t)b * RA_BOX_STRIDE;
  const float ctr = bo[RA_BOX_CTR_Y + axis], size = bo[RA_BOX_SIZE_Y + axis];
  const float s2 = expf(bo[RA_BOX_LGVAR_Y + axis]);
  const float step = (size + 1.0f) / (float)F, half = (float)(F - 1) / 2.0f;
  float d_ctr = 0.f, d_size = 0.f, d_lgv = 0.f;
  for (int idx = threadIdx.x; idx < F * L; idx += kT) {
    const int tap = idx / L, l = idx - tap * L;
    const float t = df[idx] * f[idx];
    if (t == 0.f) continue;
    const float mu = ctr + step * ((float)tap - half);
    const float d = (float)l - mu;
    const float dmu = t * d / s2;
    d_ctr += dmu;
    d_size = fmaf(dmu, ((float)tap - half) / (float)F, d_size);
    d_lgv = fmaf(t, d * d / (2.0f * s2) - 0.5f, d_lgv);
  }
  d_ctr = ra::block_sum(d_ctr, red);
  d_size = ra::block_sum(d_size, red);
  d_lgv = ra::block_sum(d_lgv, red);
  if (threadIdx.x == 0) {
    d_box[(size_t)b * 6 + 0 + axis] = d_ctr;
    d_box[(size_t)b * 6 + 2 + axis] = d_size;
    d_box[(size_t)b * 6 + 4 + axis] = d_lgv;
  }
}

}  // namespace

extern "C" size_t ra_paste_back_bwd_workspace(int B, int H, int W, int F) {
  if (B < 1 || H < 1 || W < 1 || F < 1) return 0;
  return ((size_t)B * H * W + (size_t)2 * B * F * W + (size_t)B * H) * sizeof(float);  // dV, R, A, d_gamma row shares
}

extern "C" int ra_paste_back_bwd_f32(const float *d_out, const float *out, size_t out_bstride, const float *patch,
                                     const float *fy, const float *fx, const float *gamma, int gamma_stride, int B,
                                     int H, int W, int F, int accumulate, void *ws, float *d_patch, float *d_fy,
                                     float *d_fx, float *d_gamma, void *stream) {
  if (B < 0 || H < 1 || W < 1 || F < 1 || F > kMaxF || gamma_stride < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!d_out || !out || !fy || !fx || !gamma || !ws || !d_fy || !d_fx || !d_gamma) return RA_ERR_INVALID_ARG;
  if ((patch == nullptr) != (d_patch == nullptr) || out_bstride < (size_t)H * W) return RA_ERR_INVALID_ARG;
  if (B > 65535 || H > 65535) return RA_ERR_UNSUPPORTED;
  cudaStream_t s = ra::as_stream(stream);
  float *dV = reinterpret_cast<float *>(ws);
  float *R = dV + (size_t)B * H * W, *A = R + (size_t)B * F * W, *rows = A + (size_t)B * F * W;
  const size_t psm = (size_t)F * F * sizeof(float);
  const int bx = (W + kT - 1) / kT;
  pb_R_kernel<<<dim3(bx, B), kT, psm, s>>>(patch, fx, W, F, R);
  int rc = ra::finish_launch("pb_R_kernel");
  if (rc != RA_OK) return rc;
  pb_row_kernel<<<dim3(H, B), kT, 0, s>>>(d_out, out, out_bstride, fy, R, gamma, gamma_stride, H, W, F, accumulate, dV,
                                          d_fy, rows);
  rc = ra::finish_launch("pb_row_kernel");
  if (rc != RA_OK) return rc;
  pb_col_kernel<<<dim3(bx, B), kT, psm, s>>>(dV, fy, patch, H, W, F, accumulate, A, d_fx);
  rc = ra::finish_launch("pb_col_kernel");
  if (rc != RA_OK) return rc;
  const int pairs = d_patch ? F * F : 1;
  pb_dP_kernel<<<dim3((pairs + kT / 32 - 1) / (kT / 32), B), kT, 0, s>>>(A, fx, rows, H, W, F, d_patch, d_gamma);
  return ra::finish_launch("pb_dP_kernel");
}

extern "C" int ra_gaussian_filters_bwd_f32(const float *box, const float *fy, const float *fx, const float *d_fy,
                                           const float *d_fx, int B, int H, int W, int F, float *d_box, void *stream) {
  if (B < 0 || H < 1 || W < 1 || F < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!box || !fy || !fx || !d_fy || !d_fx || !d_box) return RA_ERR_INVALID_ARG;
  if (B > 65535) return RA_ERR_UNSUPPORTED;
  filters_bwd_kernel<<<dim3(2, B), kT, 0, ra::as_stream(stream)>>>(box, fy, fx, d_fy, d_fx, H, W, F, d_box);
  return ra::finish_launch("filters_bwd_kernel");
}

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
__global__ void __launch_bounds__(kT) pb_R_kernel(const float *__restrict__ patch, const float *__restrict__ fx,
                                                  const float *__restrict__ box, int box_stride, int W, int F,
                                                  float *__restrict__ R) {
  extern __shared__ float P_s[];  // [F][F]
  const int b = blockIdx.y;
  int xlo = 0, xhi = W - 1;
  if (box != nullptr) ra::band_union(box + (size_t)b * box_stride, 1, F, W, &xlo, &xhi);
  {
    const int c0 = blockIdx.x * kT;
    if (c0 > xhi || c0 + kT - 1 < xlo) return;  // whole CTA outside the box columns: R is never read there
  }
  if (patch != nullptr)
    for (int i = threadIdx.x; i < F * F; i += kT) P_s[i] = patch[(size_t)b * F * F + i];
  __syncthreads();
  const int x = blockIdx.x * kT + threadIdx.x;
  if (x >= W || x < xlo || x > xhi) return;
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
                                                    size_t out_bstride, int n_inner, size_t out_ostride,
                                                    const float *__restrict__ fy,
                                                    const float *__restrict__ R, const float *__restrict__ gamma,
                                                    int gamma_stride, const float *__restrict__ box, int box_stride,
                                                    int H, int W, int F, int accumulate,
                                                    float *__restrict__ dV, float *__restrict__ d_fy,
                                                    float *__restrict__ dgamma_rows) {
  __shared__ float fy_s[kMaxF];
  __shared__ float red[32];
  __shared__ float dfy_s[kMaxF];
  const int y = blockIdx.x, b = blockIdx.y;
  // Only the box region matters: V = Fy P Fx^T vanishes outside the union bands, and every product below carries a
  // filter entry that is an exact zero there.
  int xlo = 0, xhi = W - 1, ylo = 0, yhi = H - 1;
  if (box != nullptr) {
    ra::band_union(box + (size_t)b * box_stride, 0, F, H, &ylo, &yhi);
    ra::band_union(box + (size_t)b * box_stride, 1, F, W, &xlo, &xhi);
  }
  if (y < ylo || y > yhi) {  // no tap reaches this row: no contribution to d_gamma, d_fy[:, y] = 0
    if (threadIdx.x == 0) dgamma_rows[(size_t)b * H + y] = 0.f;
    if (!accumulate)
      for (int i = threadIdx.x; i < F; i += kT) d_fy[((size_t)b * F + i) * H + y] = 0.f;
    return;
  }
  for (int i = threadIdx.x; i < F; i += kT) fy_s[i] = fy[((size_t)b * F + i) * H + y];
  __syncthreads();
  const float g = gamma[(size_t)b * gamma_stride];
  const float *Rb = R + (size_t)b * F * W;
  const size_t row = (size_t)(b % n_inner) * out_bstride + (size_t)(b / n_inner) * out_ostride + (size_t)y * W;
  float dg = 0.f;
  for (int x = xlo + threadIdx.x; x <= xhi; x += kT) {
    float v = 0.f;
    for (int i = 0; i < F; ++i)
      if (fy_s[i] != 0.f) v = fmaf(fy_s[i], Rb[(size_t)i * W + x], v);
    const float o = out[row + x];
    const float dz = d_out[row + x] * o * (1.0f - o);
    dg = fmaf(dz, v, dg);
    dV[((size_t)b * H + y) * W + x] = g * dz;
  }
  dg = ra::block_sum(dg, red);
  if (threadIdx.x == 0) dgamma_rows[(size_t)b * H + y] = dg;
  __syncthreads();  // this row of dV is complete (written by this CTA only)
  // d_fy[i, y] = sum_x dV[y, x] R[i, x]: one warp per in-band tap (lanes along x, coalesced rows of R), no block barrier
  // per tap.  Outside a tap's band d_fy[i, y] only ever multiplies the zero filter entry.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *dvr = dV + ((size_t)b * H + y) * W;
  for (int i = warp; i < F; i += kT / 32) {
    float s = 0.f;
    if (fy_s[i] != 0.f) {  // warp-uniform
      float s0 = 0.f, s1 = 0.f;
      int x = xlo + lane;
      for (; x + 32 <= xhi; x += 64) {
        s0 = fmaf(dvr[x], Rb[(size_t)i * W + x], s0);
        s1 = fmaf(dvr[x + 32], Rb[(size_t)i * W + x + 32], s1);
      }
      if (x <= xhi) s0 = fmaf(dvr[x], Rb[(size_t)i * W + x], s0);
      s = ra::warp_sum(s0 + s1);
    }
    if (lane == 0) dfy_s[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < F; i += kT) {
    float *dst = d_fy + ((size_t)b * F + i) * H + y;
    *dst = accumulate ? (*dst + dfy_s[i]) : dfy_s[i];
  }
}

// Column pass as two small tiled GEMMs per (example, 128-column tile):
//   A[i,x]    = sum_{y in union y-band} fy[i,y] dV[y,x]     (F x 128 tile, rows staged 32 at a time in shared memory)
//   d_fx[j,x] = sum_i P[i,j] A[i,x]                          (A tile kept in shared memory)
// 256 threads = 8 tap groups x 32 column quads; a thread owns kPcTI taps x 4 columns (F <= 8 * kPcTI = 64).
constexpr int kPcTX = 128, kPcTY = 32, kPcTI = 8;
constexpr int kPcFS = kMaxF + 4;  // row stride of fy_s: 16-byte aligned rows, 4-way instead of 32-way conflicts on the fill
__global__ void __launch_bounds__(kT) pb_col_kernel(const float *__restrict__ dV, const float *__restrict__ fy,
                                                    const float *__restrict__ patch, const float *__restrict__ box,
                                                    int box_stride, int H, int W, int F,
                                                    int accumulate, float *__restrict__ A, float *__restrict__ d_fx) {
  extern __shared__ __align__(16) float pc_sm[];
  float *P_s = pc_sm;                      // [F][F]  (patch != nullptr)
  float *fy_s = P_s + kMaxF * kMaxF;       // [kPcTY][kPcFS]   fy_s[yy][i]
  float *dv_s = fy_s + kPcTY * kPcFS;      // [kPcTY][kPcTX]
  float *A_s = dv_s + kPcTY * kPcTX;       // [kMaxF][kPcTX]
  const int b = blockIdx.y;
  const int x0 = blockIdx.x * kPcTX;
  int xlo = 0, xhi = W - 1, ylo = 0, yhi = H - 1;
  if (box != nullptr) {
    ra::band_union(box + (size_t)b * box_stride, 0, F, H, &ylo, &yhi);
    ra::band_union(box + (size_t)b * box_stride, 1, F, W, &xlo, &xhi);
  }
  if (x0 > xhi || x0 + kPcTX - 1 < xlo) {  // no tap reaches these columns: d_fx[:, x] = 0, A is never read there
    if (!accumulate)
      for (int k = threadIdx.x; k < F * kPcTX; k += kT) {
        const int j = k / kPcTX, x = x0 + k % kPcTX;
        if (x < W) d_fx[((size_t)b * F + j) * W + x] = 0.f;
      }
    return;
  }
  if (patch != nullptr)
    for (int i = threadIdx.x; i < F * F; i += kT) P_s[i] = patch[(size_t)b * F * F + i];
  const int tg = threadIdx.x / 32, xq = threadIdx.x % 32;  // taps [tg * kPcTI, +kPcTI), columns x0 + 4 xq ..
  float acc[kPcTI][4];
#pragma unroll
  for (int t = 0; t < kPcTI; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
  const float *fyb = fy + (size_t)b * F * H;
  const float *dvb = dV + (size_t)b * H * W;
  for (int y0 = ylo; y0 <= yhi; y0 += kPcTY) {
    const int ny = min(kPcTY, yhi - y0 + 1);
    __syncthreads();
    for (int k = threadIdx.x; k < kPcTY * kMaxF; k += kT) {
      const int i = k / kPcTY, yy = k % kPcTY;  // consecutive threads walk y: coalesced rows of fy
      fy_s[yy * kPcFS + i] = (i < F && yy < ny) ? fyb[(size_t)i * H + y0 + yy] : 0.f;
    }
    for (int k = threadIdx.x; k < kPcTY * kPcTX; k += kT) {
      const int yy = k / kPcTX, xx = k % kPcTX;
      const int x = x0 + xx;
      dv_s[k] = (yy < ny && x >= xlo && x <= xhi) ? dvb[(size_t)(y0 + yy) * W + x] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int yy = 0; yy < kPcTY; ++yy) {
      const float4 v = *reinterpret_cast<const float4 *>(dv_s + yy * kPcTX + 4 * xq);
      const float4 f0 = *reinterpret_cast<const float4 *>(fy_s + yy * kPcFS + tg * kPcTI);
      const float4 f1 = *reinterpret_cast<const float4 *>(fy_s + yy * kPcFS + tg * kPcTI + 4);
      const float fv[kPcTI] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int t = 0; t < kPcTI; ++t) {
        acc[t][0] = fmaf(fv[t], v.x, acc[t][0]);
        acc[t][1] = fmaf(fv[t], v.y, acc[t][1]);
        acc[t][2] = fmaf(fv[t], v.z, acc[t][2]);
        acc[t][3] = fmaf(fv[t], v.w, acc[t][3]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kPcTI; ++t) {
    const int i = tg * kPcTI + t;
    *reinterpret_cast<float4 *>(A_s + i * kPcTX + 4 * xq) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    if (i < F) {
      float *dst = A + ((size_t)b * F + i) * W + x0 + 4 * xq;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (x0 + 4 * xq + c < W) dst[c] = acc[t][c];
    }
  }
  __syncthreads();
  // d_fx tile: thread owns taps j in [tg * kPcTI, +kPcTI) x the same 4 columns
#pragma unroll
  for (int t = 0; t < kPcTI; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
  for (int i = 0; i < F; ++i) {
    const float4 av = *reinterpret_cast<const float4 *>(A_s + i * kPcTX + 4 * xq);
#pragma unroll
    for (int t = 0; t < kPcTI; ++t) {
      const int j = tg * kPcTI + t;
      const float pv = (patch != nullptr) ? ((j < F) ? P_s[i * F + j] : 0.f) : 1.0f;
      acc[t][0] = fmaf(pv, av.x, acc[t][0]);
      acc[t][1] = fmaf(pv, av.y, acc[t][1]);
      acc[t][2] = fmaf(pv, av.z, acc[t][2]);
      acc[t][3] = fmaf(pv, av.w, acc[t][3]);
    }
  }
#pragma unroll
  for (int t = 0; t < kPcTI; ++t) {
    const int j = tg * kPcTI + t;
    if (j >= F) continue;
    float *dst = d_fx + ((size_t)b * F + j) * W + x0 + 4 * xq;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int x = x0 + 4 * xq + c;
      if (x >= W) continue;
      const float v = (x >= xlo && x <= xhi) ? acc[t][c] : 0.f;
      dst[c] = accumulate ? (dst[c] + v) : v;
    }
  }
}

// d_P[b,i,j] = sum_x A[b,i,x] fx[b,j,x]: one warp per (i, j); d_gamma[b] = sum of the row shares (CTA 0)
__global__ void __launch_bounds__(kT) pb_dP_kernel(const float *__restrict__ A, const float *__restrict__ fx,
                                                   const float *__restrict__ dgamma_rows,
                                                   const float *__restrict__ box, int box_stride, int H, int W, int F,
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
  int xlo = 0, xhi = W - 1;
  if (box != nullptr) ra::tap_band(box + (size_t)b * box_stride, 1, j, F, W, &xlo, &xhi);
  float s = 0.f;
  for (int x = xlo + lane; x <= xhi; x += 32) s = fmaf(Ai[x], fj[x], s);
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


// ---------------------------------------------------------------------------------------------- glimpse backward
// forward (full_model.py:788-789, modellib.py:615-641): x_patch[i,j,d] = gamma * sum_{y,x} fy[i,y] X[y,x,d] fx[j,x],
// X = the step's input stack given as xs [B,H,W,Cs] + canvas [B,H,W]; source channel c sits at patch channel
// chan_map[c].  With G = d x_patch and T[i,x,c] = sum_y fy[i,y] X[y,x,c] (the forward's row pass, recomputed):
//   d_fx[j,x] = gamma * sum_{i,c} G[i,j,c] T[i,x,c]
//   d_fy[i,y] = sum_{x,c} S[i,x,c] X[y,x,c],   S[i,x,c] = gamma * sum_j G[i,j,c] fx[j,x]
//   d_gamma   = sum G * x_patch / gamma
// No gradient flows to X: the image is data and the canvas is behind tf.stop_gradient (full_model.py:846-848).
// Band-limited: fy[i, :] is non-zero only on tap i's y-band, fx[j, :] on tap j's x-band (ra::tap_band), so
//   T[i,x,c]  is needed for x in the union x-band and sums over y in band_i,
//   d_fx[j,x] is needed for x in band_j only (outside it multiplies the zero filter entry in the filter backward),
//   S[i,x,c]  sums over the taps j whose band holds x,  d_fy[i,y] is needed for y in band_i only.
// box == nullptr: dense filters of unknown origin - every range is the full axis.
__device__ __forceinline__ void axis_union(const float *box, int b, int axis, int F, int L, int *lo, int *hi) {
  *lo = 0;
  *hi = L - 1;
  if (box != nullptr) ra::band_union(box + (size_t)b * RA_BOX_STRIDE, axis, F, L, lo, hi);
}
__device__ __forceinline__ void axis_tap(const float *box, int b, int axis, int t, int F, int L, int *lo, int *hi) {
  *lo = 0;
  *hi = L - 1;
  if (box != nullptr) ra::tap_band(box + (size_t)b * RA_BOX_STRIDE, axis, t, F, L, lo, hi);
}

// grid ((x,c) chunks, tap groups of kExIB, b): T[b,i,x,c] = sum_{y in band_i} fy[i,y] X[y,x,c], x in the union x-band
constexpr int kExIB = 4;
__global__ void __launch_bounds__(kT) ex_T_kernel(const float *__restrict__ xs, int Cs, int xs_bmod,
                                                  const float *__restrict__ canvas, const float *__restrict__ fy,
                                                  const float *__restrict__ box, int H, int W, int F, int D,
                                                  float *__restrict__ T) {
  const int b = blockIdx.z;
  const int bx = xs_bmod > 0 ? b % xs_bmod : b;
  const int i0 = blockIdx.y * kExIB;
  int xlo, xhi;
  axis_union(box, b, 1, F, W, &xlo, &xhi);
  const int idx = blockIdx.x * kT + threadIdx.x;  // (x - xlo, c)
  const int nxc = (xhi - xlo + 1) * D;
  if (blockIdx.x * kT >= nxc) return;
  int ylo = H, yhi = -1;
#pragma unroll
  for (int k = 0; k < kExIB; ++k)
    if (i0 + k < F) {
      int lo, hi;
      axis_tap(box, b, 0, i0 + k, F, H, &lo, &hi);
      if (hi >= lo) {
        ylo = min(ylo, lo);
        yhi = max(yhi, hi);
      }
    }
  // the CTA's filter rows over its union band in shared memory (every thread needs all of them), then four independent
  // image loads in flight per round - the loop is bound by load latency, not arithmetic
  extern __shared__ float ext_w[];  // [kExIB][H]
  const int ny = max(yhi - ylo + 1, 0);
  for (int u = threadIdx.x; u < kExIB * ny; u += kT) {
    const int k = u / ny, yy = u - k * ny;
    ext_w[k * H + yy] = (i0 + k < F) ? __ldg(fy + ((size_t)b * F + i0 + k) * H + ylo + yy) : 0.f;
  }
  __syncthreads();
  if (idx >= nxc) return;
  const int x = xlo + idx / D, c = idx % D;
  float acc[kExIB];
#pragma unroll
  for (int k = 0; k < kExIB; ++k) acc[k] = 0.f;
  const bool from_xs = c < Cs;
  const float *src = from_xs ? xs + (((size_t)bx * H + ylo) * W + x) * Cs + c : canvas + ((size_t)b * H + ylo) * W + x;
  const size_t ystride = from_xs ? (size_t)W * Cs : (size_t)W;
  int yy = 0;
  for (; yy + 4 <= ny; yy += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (size_t)(yy + u) * ystride);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < kExIB; ++k) acc[k] = fmaf(ext_w[k * H + yy + u], v[u], acc[k]);
  }
  for (; yy < ny; ++yy) {
    const float v = __ldg(src + (size_t)yy * ystride);
#pragma unroll
    for (int k = 0; k < kExIB; ++k) acc[k] = fmaf(ext_w[k * H + yy], v, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < kExIB; ++k)
    if (i0 + k < F) T[(((size_t)b * F + i0 + k) * W + x) * D + c] = acc[k];
}

// grid (j, b): d_fx[b,j,x] (+)= gamma * sum_{i,c} G[b,i,j,ch(c)] T[b,i,x,c] for x in band_j (zero elsewhere); one warp
// per column x, lanes over the (i, c) pairs
__global__ void __launch_bounds__(kT) ex_dfx_kernel(const float *__restrict__ G, int cstride,
                                                    const int *__restrict__ chan_map, const float *__restrict__ T,
                                                    const float *__restrict__ gamma, int gamma_stride,
                                                    const float *__restrict__ box, int W, int F, int D, int accumulate,
                                                    float *__restrict__ d_fx) {
  extern __shared__ float G_s[];  // [F (i)][D] of tap j, source-channel order
  const int j = blockIdx.x, b = blockIdx.y;
  for (int k = threadIdx.x; k < F * D; k += kT) {
    const int i = k / D, c = k - i * D;
    G_s[k] = G[(((size_t)b * F + i) * F + j) * cstride + chan_map[c]];
  }
  __syncthreads();
  int xlo, xhi;
  axis_tap(box, b, 1, j, F, W, &xlo, &xhi);
  const float g = gamma[(size_t)b * gamma_stride];
  float *dst = d_fx + ((size_t)b * F + j) * W;
  if (!accumulate)
    for (int x = threadIdx.x; x < W; x += kT)
      if (x < xlo || x > xhi) dst[x] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int x = xlo + warp; x <= xhi; x += kT / 32) {
    float s = 0.f;
    for (int k = lane; k < F * D; k += 32) {
      const int i = k / D, c = k - i * D;
      s = fmaf(G_s[k], T[(((size_t)b * F + i) * W + x) * D + c], s);
    }
    s = ra::warp_sum(s);
    if (lane == 0) dst[x] = accumulate ? fmaf(g, s, dst[x]) : g * s;
  }
}

// grid (i, b): S[x,c] = gamma * sum_{j: x in band_j} G[b,i,j,ch(c)] fx[b,j,x] in shared memory (x in the union x-band,
// processed in chunks), then d_fy[b,i,y] (+)= sum_{x,c} S[x,c] X[b,y,x,c] for y in band_i (zero elsewhere): one warp
// per row y, partial sums over the chunks accumulated in registers of the owning warp.
constexpr int kExChunk = 128;  // columns per shared-memory chunk
__global__ void __launch_bounds__(kT) ex_dfy_kernel(const float *__restrict__ xs, int Cs, int xs_bmod,
                                                    const float *__restrict__ canvas, const float *__restrict__ G,
                                                    int cstride, const int *__restrict__ chan_map,
                                                    const float *__restrict__ fx, const float *__restrict__ gamma,
                                                    int gamma_stride, const float *__restrict__ box, int H, int W, int F,
                                                    int D, int accumulate, float *__restrict__ d_fy) {
  extern __shared__ float sm[];  // G_s [F (j)][D] | S_s [kExChunk][D] | band_s [F][2] (ints)
  float *G_s = sm, *S_s = sm + F * D;
  int *band_s = reinterpret_cast<int *>(S_s + kExChunk * D);
  const int i = blockIdx.x, b = blockIdx.y;
  const int bx = xs_bmod > 0 ? b % xs_bmod : b;
  for (int k = threadIdx.x; k < F * D; k += kT) {
    const int j = k / D, c = k - j * D;
    G_s[k] = G[(((size_t)b * F + i) * F + j) * cstride + chan_map[c]];
  }
  for (int j = threadIdx.x; j < F; j += kT) axis_tap(box, b, 1, j, F, W, &band_s[2 * j], &band_s[2 * j + 1]);
  int xlo, xhi, ylo, yhi;
  axis_union(box, b, 1, F, W, &xlo, &xhi);
  axis_tap(box, b, 0, i, F, H, &ylo, &yhi);
  float *dst = d_fy + ((size_t)b * F + i) * H;
  if (!accumulate)
    for (int y = threadIdx.x; y < H; y += kT)
      if (y < ylo || y > yhi) dst[y] = 0.f;
  const float g = gamma[(size_t)b * gamma_stride];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kRowsPerWarp = 8;  // rows y = ylo + warp + 8 * r, r < kRowsPerWarp  (bands are narrower than 64 rows;
  float acc[kRowsPerWarp];         //  wider ones take more passes below)
  for (int ybase = ylo; ybase <= yhi; ybase += kRowsPerWarp * (kT / 32)) {
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) acc[r] = 0.f;
    for (int x0 = xlo; x0 <= xhi; x0 += kExChunk) {
      const int nx = min(kExChunk, xhi - x0 + 1);
      __syncthreads();  // G_s / band_s ready; the previous chunk has been consumed
      for (int k = threadIdx.x; k < nx * D; k += kT) {
        const int x = x0 + k / D, c = k % D;
        float s = 0.f;
        for (int j = 0; j < F; ++j)
          if (x >= band_s[2 * j] && x <= band_s[2 * j + 1]) s = fmaf(G_s[j * D + c], __ldg(fx + ((size_t)b * F + j) * W + x), s);
        S_s[k] = g * s;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) {
        const int y = ybase + warp + (kT / 32) * r;
        if (y > yhi) continue;
        float s = 0.f;
        for (int k = lane; k < nx * D; k += 32) {
          const int x = x0 + k / D, c = k % D;
          const float xv = (c < Cs) ? __ldg(xs + (((size_t)bx * H + y) * W + x) * Cs + c)
                                    : __ldg(canvas + ((size_t)b * H + y) * W + x);
          s = fmaf(S_s[k], xv, s);
        }
        acc[r] += s;
      }
    }
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
      const int y = ybase + warp + (kT / 32) * r;
      const float v = ra::warp_sum(acc[r]);
      if (y <= yhi && lane == 0) dst[y] = accumulate ? (dst[y] + v) : v;
    }
  }
}

// ---- d_fy by the other contraction order (used when H <= W, so that U fits the second half of the workspace):
//   U[b,y,j,c] = sum_{x in band_j} X[b,y,x,c] fx[b,j,x]        (the forward's column pass applied to the image rows)
//   d_fy[b,i,y] = gamma * sum_{j,c} G[b,i,j,ch(c)] U[b,y,j,c]   for y in band_i
// U costs |rows| * sum_j |band_j| * D MACs per example and is shared by all F taps i, where the S formulation rebuilds
// an [x,c] plane of F-term sums for every (example, tap).
// grid (H, B): one CTA per image row inside the union y-band; slab [x][c] of the row in shared memory.
__global__ void __launch_bounds__(kT) ex_U_kernel(const float *__restrict__ xs, int Cs, int xs_bmod,
                                                  const float *__restrict__ canvas, const float *__restrict__ fx,
                                                  const float *__restrict__ box, int H, int W, int F, int D,
                                                  float *__restrict__ U) {
  extern __shared__ float sm[];  // slab [nx][D] | band_s [F][2]
  const int y = blockIdx.x, b = blockIdx.y;
  int ylo, yhi, xlo, xhi;
  axis_union(box, b, 0, F, H, &ylo, &yhi);
  if (y < ylo || y > yhi) return;  // no tap reads this row
  axis_union(box, b, 1, F, W, &xlo, &xhi);
  const int nx = xhi - xlo + 1;
  if (nx <= 0) return;
  float *slab = sm;
  int *band_s = reinterpret_cast<int *>(sm + (size_t)W * D);
  for (int j = threadIdx.x; j < F; j += kT) axis_tap(box, b, 1, j, F, W, &band_s[2 * j], &band_s[2 * j + 1]);
  const int bx = xs_bmod > 0 ? b % xs_bmod : b;
  if (Cs > 0) {
    const float *row = xs + (((size_t)bx * H + y) * W + xlo) * Cs;
    for (int k = threadIdx.x; k < nx * Cs; k += kT) {
      const int x = k / Cs, c = k - x * Cs;
      slab[x * D + c] = __ldg(row + k);
    }
  }
  if (D > Cs) {
    const float *row = canvas + ((size_t)b * H + y) * W + xlo;
    for (int x = threadIdx.x; x < nx; x += kT) slab[x * D + Cs] = __ldg(row + x);
  }
  __syncthreads();
  const float *fxb = fx + (size_t)b * F * W;
  float *dst = U + ((size_t)b * H + y) * F * D;
  for (int idx = threadIdx.x; idx < F * D; idx += kT) {
    const int j = idx / D, c = idx - j * D;
    const int lo = max(band_s[2 * j], xlo), hi = min(band_s[2 * j + 1], xhi);
    const float *fj = fxb + (size_t)j * W;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int x = lo;
    for (; x + 3 <= hi; x += 4) {
      a0 = fmaf(slab[(x - xlo) * D + c], __ldg(fj + x), a0);
      a1 = fmaf(slab[(x + 1 - xlo) * D + c], __ldg(fj + x + 1), a1);
      a2 = fmaf(slab[(x + 2 - xlo) * D + c], __ldg(fj + x + 2), a2);
      a3 = fmaf(slab[(x + 3 - xlo) * D + c], __ldg(fj + x + 3), a3);
    }
    for (; x <= hi; ++x) a0 = fmaf(slab[(x - xlo) * D + c], __ldg(fj + x), a0);
    dst[idx] = (a0 + a1) + (a2 + a3);
  }
}

// grid (i, b): d_fy[b,i,y] (+)= gamma * <G[b,i,:,:], U[b,y,:,:]> for y in band_i (zero elsewhere); one warp per row.
__global__ void __launch_bounds__(kT) ex_dfy_from_U_kernel(const float *__restrict__ G, int cstride,
                                                           const int *__restrict__ chan_map, const float *__restrict__ U,
                                                           const float *__restrict__ gamma, int gamma_stride,
                                                           const float *__restrict__ box, int H, int F, int D,
                                                           int accumulate, float *__restrict__ d_fy) {
  extern __shared__ float G_s[];  // [F (j)][D], source-channel order
  const int i = blockIdx.x, b = blockIdx.y;
  for (int k = threadIdx.x; k < F * D; k += kT) {
    const int j = k / D, c = k - j * D;
    G_s[k] = G[(((size_t)b * F + i) * F + j) * cstride + chan_map[c]];
  }
  int ylo, yhi, ulo, uhi;
  axis_tap(box, b, 0, i, F, H, &ylo, &yhi);
  axis_union(box, b, 0, F, H, &ulo, &uhi);  // rows outside the union were not produced (band_i lies inside it)
  ylo = max(ylo, ulo);
  yhi = min(yhi, uhi);
  float *dst = d_fy + ((size_t)b * F + i) * H;
  if (!accumulate)
    for (int y = threadIdx.x; y < H; y += kT)
      if (y < ylo || y > yhi) dst[y] = 0.f;
  __syncthreads();
  const float g = gamma[(size_t)b * gamma_stride];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = F * D;
  for (int y = ylo + warp; y <= yhi; y += kT / 32) {
    const float *u = U + ((size_t)b * H + y) * n;
    float s0 = 0.f, s1 = 0.f;
    int k = lane;
    for (; k + 32 < n; k += 64) {
      s0 = fmaf(G_s[k], __ldg(u + k), s0);
      s1 = fmaf(G_s[k + 32], __ldg(u + k + 32), s1);
    }
    for (; k < n; k += 32) s0 = fmaf(G_s[k], __ldg(u + k), s0);
    const float v = g * ra::warp_sum(s0 + s1);
    if (lane == 0) dst[y] = accumulate ? (dst[y] + v) : v;
  }
}

// d_gamma[b] = sum G * x_patch / gamma over the D real channels
__global__ void __launch_bounds__(kT) ex_dgamma_kernel(const float *__restrict__ G, const float *__restrict__ x_patch,
                                                       int cstride, int D, int F, const float *__restrict__ gamma,
                                                       int gamma_stride, float *__restrict__ d_gamma) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float s = 0.f;
  for (int k = threadIdx.x; k < F * F * cstride; k += kT)
    if (k % cstride < D) s = fmaf(G[(size_t)b * F * F * cstride + k], x_patch[(size_t)b * F * F * cstride + k], s);
  s = ra::block_sum(s, red);
  if (threadIdx.x == 0) d_gamma[b] = s / gamma[(size_t)b * gamma_stride];
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
  return ra_paste_back_bwd_ex_f32(d_out, out, out_bstride, B > 0 ? B : 1, 0, patch, fy, fx, gamma, gamma_stride, nullptr,
                                  B, H, W, F, accumulate, ws, d_patch, d_fy, d_fx, d_gamma, stream);
}

extern "C" int ra_paste_back_bwd_ex_f32(const float *d_out, const float *out, size_t out_bstride, int n_inner,
                                        size_t out_ostride, const float *patch, const float *fy, const float *fx,
                                        const float *gamma, int gamma_stride, const float *box, int B, int H, int W,
                                        int F, int accumulate, void *ws, float *d_patch, float *d_fy, float *d_fx,
                                        float *d_gamma, void *stream) {
  if (n_inner < 1) return RA_ERR_INVALID_ARG;
  const int box_stride = RA_BOX_STRIDE;
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
  pb_R_kernel<<<dim3(bx, B), kT, psm, s>>>(patch, fx, box, box_stride, W, F, R);
  int rc = ra::finish_launch("pb_R_kernel");
  if (rc != RA_OK) return rc;
  pb_row_kernel<<<dim3(H, B), kT, 0, s>>>(d_out, out, out_bstride, n_inner, out_ostride, fy, R, gamma, gamma_stride, box,
                                          box_stride, H, W, F, accumulate, dV, d_fy, rows);
  rc = ra::finish_launch("pb_row_kernel");
  if (rc != RA_OK) return rc;
  {
    const size_t csm = ((size_t)kMaxF * kMaxF + (size_t)kPcTY * kPcFS + (size_t)kPcTY * kPcTX + (size_t)kMaxF * kPcTX) *
                       sizeof(float);  // 74 KB
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(pb_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm);
      attr_done = true;
    }
    pb_col_kernel<<<dim3((W + kPcTX - 1) / kPcTX, B), kT, csm, s>>>(dV, fy, patch, box, box_stride, H, W, F, accumulate,
                                                                     A, d_fx);
  }
  rc = ra::finish_launch("pb_col_kernel");
  if (rc != RA_OK) return rc;
  const int pairs = d_patch ? F * F : 1;
  pb_dP_kernel<<<dim3((pairs + kT / 32 - 1) / (kT / 32), B), kT, 0, s>>>(A, fx, rows, box, box_stride, H, W, F, d_patch,
                                                                         d_gamma);
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

extern "C" size_t ra_gaussian_extract_bwd_workspace(int B, int W, int F, int D) {
  if (B < 1 || W < 1 || F < 1 || D < 1) return 0;
  return (size_t)2 * B * F * W * D * sizeof(float);  // T and S
}

extern "C" int ra_gaussian_extract_bwd_f32(const float *xs, int Cs, const float *canvas, const int32_t *chan_map,
                                           const float *fy, const float *fx, const float *gamma, int gamma_stride,
                                           const float *d_patch, const float *x_patch, int patch_cstride, int B, int H,
                                           int W, int F, int accumulate, void *ws, float *d_fy, float *d_fx,
                                           float *d_gamma, void *stream) {
  return ra_gaussian_extract_bwd_ex_f32(xs, Cs, 0, canvas, chan_map, fy, fx, gamma, gamma_stride, nullptr, d_patch,
                                        x_patch, patch_cstride, B, H, W, F, accumulate, ws, d_fy, d_fx, d_gamma, stream);
}

extern "C" int ra_gaussian_extract_bwd_ex_f32(const float *xs, int Cs, int xs_bmod, const float *canvas,
                                              const int32_t *chan_map, const float *fy, const float *fx,
                                              const float *gamma, int gamma_stride, const float *box,
                                              const float *d_patch, const float *x_patch, int patch_cstride, int B,
                                              int H, int W, int F, int accumulate, void *ws, float *d_fy, float *d_fx,
                                              float *d_gamma, void *stream) {
  if (xs_bmod < 0) return RA_ERR_INVALID_ARG;
  const int D = Cs + (canvas ? 1 : 0);
  if (B < 0 || H < 1 || W < 1 || F < 1 || F > kMaxF || Cs < 0 || D < 1 || patch_cstride < D || gamma_stride < 1)
    return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if ((Cs > 0 && !xs) || !chan_map || !fy || !fx || !gamma || !d_patch || !x_patch || !ws || !d_fy || !d_fx || !d_gamma)
    return RA_ERR_INVALID_ARG;
  if (B > 65535 || ((size_t)F * D + (size_t)kExChunk * D + 2 * F) * sizeof(float) > 48 * 1024) return RA_ERR_UNSUPPORTED;
  cudaStream_t s = ra::as_stream(stream);
  float *T = reinterpret_cast<float *>(ws);
  const int bxc = (W * D + kT - 1) / kT;
  const size_t gsm = (size_t)F * D * sizeof(float);
  ex_T_kernel<<<dim3(bxc, (F + kExIB - 1) / kExIB, B), kT, (size_t)kExIB * H * sizeof(float), s>>>(xs, Cs, xs_bmod, canvas,
                                                                                                   fy, box, H, W, F, D, T);
  int rc = ra::finish_launch("ex_T_kernel");
  if (rc != RA_OK) return rc;
  ex_dfx_kernel<<<dim3(F, B), kT, gsm, s>>>(d_patch, patch_cstride, chan_map, T, gamma, gamma_stride, box, W, F, D,
                                            accumulate, d_fx);
  rc = ra::finish_launch("ex_dfx_kernel");
  if (rc != RA_OK) return rc;
  const size_t usm = (size_t)W * D * sizeof(float) + (size_t)2 * F * sizeof(int);
  if (H <= W && usm <= 160 * 1024 && getenv("RA_EX_DFY_S") == nullptr) {
    if (usm > 48 * 1024) {
      static bool attr_done = false;
      if (!attr_done) {
        cudaFuncSetAttribute(ex_U_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_done = true;
      }
    }
    // U [B,H,F,D] lives in the second half of the workspace (B*F*W*D floats >= B*H*F*D when H <= W)
    float *U = T + (size_t)B * F * W * D;
    ex_U_kernel<<<dim3(H, B), kT, usm, s>>>(xs, Cs, xs_bmod, canvas, fx, box, H, W, F, D, U);
    rc = ra::finish_launch("ex_U_kernel");
    if (rc != RA_OK) return rc;
    ex_dfy_from_U_kernel<<<dim3(F, B), kT, gsm, s>>>(d_patch, patch_cstride, chan_map, U, gamma, gamma_stride, box, H, F,
                                                     D, accumulate, d_fy);
    rc = ra::finish_launch("ex_dfy_from_U_kernel");
    if (rc != RA_OK) return rc;
  } else {
    const size_t ysm = gsm + (size_t)kExChunk * D * sizeof(float) + (size_t)2 * F * sizeof(int);
    ex_dfy_kernel<<<dim3(F, B), kT, ysm, s>>>(xs, Cs, xs_bmod, canvas, d_patch, patch_cstride, chan_map, fx, gamma,
                                              gamma_stride, box, H, W, F, D, accumulate, d_fy);
    rc = ra::finish_launch("ex_dfy_kernel");
    if (rc != RA_OK) return rc;
  }
  ex_dgamma_kernel<<<B, kT, 0, s>>>(d_patch, x_patch, patch_cstride, D, F, gamma, gamma_stride, d_gamma);
  return ra::finish_launch("ex_dgamma_kernel");
}

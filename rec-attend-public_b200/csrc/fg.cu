// Output head and loss block of the foreground / orientation FCN (fg_model.py:174-236).
// The FCN's convolution stack runs on the conv kernels of conv_umma.cu / conv.cu; what is left is one HBM-bound pass
// over the logits: sigmoid or softmax over the semantic channels, softmax over the orientation channels, the
// thresholded / arg-max "hard" output and - when the ground truth is given - every sum the losses need
// (f_iou_all soft and hard, f_bce / f_ce, the masked orientation cross-entropy and accuracy), accumulated in
// double precision per CTA and finished by one small CTA in a fixed order (deterministic, no atomics).
#include "common.cuh"

namespace {

constexpr int kFgBlocks = ra::kNumSMs * 4;  // persistent grid-stride CTAs
constexpr int kFgThreads = 256;
constexpr int kFgAcc = 9;
constexpr int kFgMaxC = 32;  // semantic + orientation channels (9 for KITTI, 17 for Cityscapes)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// accumulators: 0 sum(y_out*y_gt)  1 sum(y_out)  2 sum(y_gt)  3 sum(y_hard*y_gt)  4 sum(y_hard)   [IoU channels]
//               5 sum of f_bce (nsc == 1) or f_ce (nsc > 1) over every semantic channel
//               6 sum(f_ce(d_out, d_gt) * mask)  7 sum(mask)  8 sum(correct * mask)
__global__ void __launch_bounds__(kFgThreads) fg_head_kernel(const float *__restrict__ logits, size_t npix, int nsc,
                                                             int nori, const float *__restrict__ y_gt,
                                                             const float *__restrict__ d_gt, float *__restrict__ y_out,
                                                             float *__restrict__ d_out, float *__restrict__ y_hard,
                                                             double *__restrict__ partial) {
  const int C = nsc + nori;
  double acc[kFgAcc];
#pragma unroll
  for (int a = 0; a < kFgAcc; ++a) acc[a] = 0.0;
  const int c0 = (nsc == 1) ? 0 : 1;  // IoU over the foreground channel, or over classes 1.. (fg_model.py:205-214)
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    const float *lg = logits + p * C;
    float ys[kFgMaxC];
    float mask = 0.f;
    if (nsc == 1) {
      ys[0] = ra::sigmoidf_acc(lg[0]);
    } else {  // tf.nn.softmax: exp(x - max) / sum
      float mx = -INFINITY, sum = 0.f;
      for (int c = 0; c < nsc; ++c) mx = fmaxf(mx, lg[c]);
      for (int c = 0; c < nsc; ++c) {
        ys[c] = expf(lg[c] - mx);
        sum += ys[c];
      }
      for (int c = 0; c < nsc; ++c) ys[c] = ys[c] / sum;
    }
    float ymax = -INFINITY;
    for (int c = 0; c < nsc; ++c) ymax = fmaxf(ymax, ys[c]);
    for (int c = 0; c < nsc; ++c) {
      const float v = ys[c];
      const float hard = (nsc == 1) ? (v > 0.5f ? 1.f : 0.f) : (v == ymax ? 1.f : 0.f);  // :201-204
      y_out[p * nsc + c] = v;
      if (y_hard != nullptr) y_hard[p * nsc + c] = hard;
      if (y_gt != nullptr) {
        const float g = y_gt[p * nsc + c];
        if (c >= c0) {
          acc[0] += (double)(v * g);
          acc[1] += (double)v;
          acc[2] += (double)g;
          acc[3] += (double)(hard * g);
          acc[4] += (double)hard;
          if (nsc > 1) mask = fmaxf(mask, g);  // y_gt_mask = max over classes 1.. (:194-198)
        }
        if (nsc == 1) {
          acc[5] += (double)(-g * logf(v + 1e-5f) - (1.f - g) * logf(1.f - v + 1e-5f));  // modellib.f_bce
          mask = g;
        } else {
          acc[5] += (double)(-g * logf(v + 1e-5f));  // modellib.f_ce
        }
      }
    }
    if (nori > 0) {
      const float *lo = lg + nsc;
      float ds[kFgMaxC];
      float mx = -INFINITY, sum = 0.f;
      for (int c = 0; c < nori; ++c) mx = fmaxf(mx, lo[c]);
      for (int c = 0; c < nori; ++c) {
        ds[c] = expf(lo[c] - mx);
        sum += ds[c];
      }
      int arg_o = 0, arg_g = 0;
      float best_o = -INFINITY, best_g = -INFINITY, ce = 0.f;
      for (int c = 0; c < nori; ++c) {
        const float v = ds[c] / sum;
        d_out[p * nori + c] = v;
        if (v > best_o) {  // tf.argmax: first maximum
          best_o = v;
          arg_o = c;
        }
        if (y_gt != nullptr && d_gt != nullptr) {
          const float g = d_gt[p * nori + c];
          ce += -g * logf(v + 1e-5f);
          if (g > best_g) {
            best_g = g;
            arg_g = c;
          }
        }
      }
      if (y_gt != nullptr && d_gt != nullptr) {
        acc[6] += (double)(ce * mask);
        acc[8] += (double)((arg_o == arg_g ? 1.f : 0.f) * mask);
      }
    }
    if (y_gt != nullptr) acc[7] += (double)mask;
  }
  if (partial == nullptr) return;
  __shared__ double red[kFgAcc][kFgThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < kFgAcc; ++a) {
    const double s = warp_sum_d(acc[a]);
    if (lane == 0) red[a][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < kFgAcc) {
    double s = 0.0;
    for (int w = 0; w < kFgThreads / 32; ++w) s += red[threadIdx.x][w];
    partial[(size_t)blockIdx.x * kFgAcc + threadIdx.x] = s;
  }
}

// out: RA_FG_* slots (see the header)
__global__ void fg_finalize_kernel(const double *__restrict__ partial, int nblocks, double num_pixel, int has_ori,
                                   int loss_is_bce, float *__restrict__ out) {
  __shared__ double tot[kFgAcc];
  if (threadIdx.x < kFgAcc) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * kFgAcc + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float inter_s = (float)tot[0], sum_o = (float)tot[1], sum_g = (float)tot[2];
    const float inter_h = (float)tot[3], sum_h = (float)tot[4];
    const float iou_soft = inter_s / (sum_o + sum_g - inter_s + 1e-5f);  // modellib.f_iou_all
    const float iou_hard = inter_h / (sum_h + sum_g - inter_h + 1e-5f);
    const float segloss = (float)(tot[5] / num_pixel);
    const float fg_loss = loss_is_bce ? segloss : -iou_soft;  // fg_model.py:223-226
    float loss = fg_loss, ori_ce = 0.f, ori_acc = 0.f;
    if (has_ori) {
      ori_ce = (float)tot[6] / (float)tot[7];   // :231-233 (0/0 = NaN without any foreground pixel, like the graph)
      ori_acc = (float)tot[8] / (float)tot[7];  // :236-239
      loss += ori_ce;
    }
    out[RA_FG_IOU_SOFT] = iou_soft;
    out[RA_FG_IOU_HARD] = iou_hard;
    out[RA_FG_SEGLOSS] = segloss;
    out[RA_FG_FOREGROUND_LOSS] = fg_loss;
    out[RA_FG_ORIENTATION_CE] = ori_ce;
    out[RA_FG_ORIENTATION_ACC] = ori_acc;
    out[RA_FG_LOSS] = loss;
    out[7] = 0.f;
  }
}

// Gradient of loss = foreground_loss [+ orientation_ce] (fg_model.py:223-240) at the logits, one pass over the pixels.
// The global sums the IoU loss and the masked cross-entropy divide by are re-read from the per-CTA partials the forward
// head (ra_fg_head_f32 on the same inputs) left in its workspace - every CTA adds them in the same fixed order.
//   nsc == 1:  y = sigmoid(z);  'iou': dL/dy = -(g U - I (1 - g)) / U^2,  U = sum y + sum g - I + 1e-5 (f_iou_all);
//              'bce': dL/dy = (-g / (y + 1e-5) + (1 - g) / (1 - y + 1e-5)) / npix;            dz = dL/dy y (1 - y)
//   nsc  > 1:  y = softmax(z);  'iou' over classes 1..: as above per class (0 for class 0);  else f_ce: -g / (y + 1e-5) / npix;
//              dz_c = y_c (dL/dy_c - sum_k y_k dL/dy_k)
//   orientation: d = softmax(zo);  dL/dd_c = -mask dg_c / (d_c + 1e-5) / sum(mask);  softmax backward.
__global__ void __launch_bounds__(kFgThreads) fg_head_bwd_kernel(const float *__restrict__ y_out,
                                                                 const float *__restrict__ d_out, size_t npix, int nsc,
                                                                 int nori, const float *__restrict__ y_gt,
                                                                 const float *__restrict__ d_gt, int loss_is_bce,
                                                                 const double *__restrict__ partial, int nblocks,
                                                                 float *__restrict__ d_logits) {
  __shared__ double tot_s[kFgAcc];
  if (threadIdx.x < kFgAcc) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * kFgAcc + threadIdx.x];
    tot_s[threadIdx.x] = s;
  }
  __syncthreads();
  const float I = (float)tot_s[0], So = (float)tot_s[1], Sg = (float)tot_s[2], M = (float)tot_s[7];
  const float U = So + Sg - I + 1e-5f;
  const float inv_npix = 1.0f / (float)npix;
  const int C = nsc + nori;
  const int c0 = (nsc == 1) ? 0 : 1;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    float *dz = d_logits + p * C;
    float mask = 0.f;
    if (nsc == 1) {
      const float y = y_out[p], g = y_gt[p];
      mask = g;
      const float dy = loss_is_bce ? (-g / (y + 1e-5f) + (1.f - g) / (1.f - y + 1e-5f)) * inv_npix
                                   : -(g * U - I * (1.f - g)) / (U * U);
      dz[0] = dy * y * (1.f - y);
    } else {
      float dy[kFgMaxC], dot = 0.f;
      for (int c = 0; c < nsc; ++c) {
        const float y = y_out[p * nsc + c], g = y_gt[p * nsc + c];
        if (c >= c0) mask = fmaxf(mask, g);
        dy[c] = loss_is_bce ? -g / (y + 1e-5f) * inv_npix : (c >= c0 ? -(g * U - I * (1.f - g)) / (U * U) : 0.f);
        dot = fmaf(y, dy[c], dot);
      }
      for (int c = 0; c < nsc; ++c) dz[c] = y_out[p * nsc + c] * (dy[c] - dot);
    }
    if (nori > 0) {
      float dd[kFgMaxC], dot = 0.f;
      for (int c = 0; c < nori; ++c) {
        const float d = d_out[p * nori + c];
        dd[c] = -mask * d_gt[p * nori + c] / (d + 1e-5f) / M;
        dot = fmaf(d, dd[c], dot);
      }
      for (int c = 0; c < nori; ++c) dz[nsc + c] = d_out[p * nori + c] * (dd[c] - dot);
    }
  }
}

}  // namespace

extern "C" size_t ra_fg_head_workspace(void) { return (size_t)kFgBlocks * kFgAcc * sizeof(double); }

extern "C" int ra_fg_head_f32(const float *logits, size_t npix, int nsc, int nori, const float *y_gt, const float *d_gt,
                              int loss_is_bce, float *y_out, float *d_out, float *y_hard, float *out, void *ws,
                              void *stream) {
  if (nsc < 1 || nori < 0 || nsc > kFgMaxC || nori > kFgMaxC) return RA_ERR_INVALID_ARG;
  if (!logits || !y_out || (nori > 0 && !d_out)) return RA_ERR_INVALID_ARG;
  if (y_gt != nullptr && (!out || !ws)) return RA_ERR_INVALID_ARG;
  if (y_gt != nullptr && nori > 0 && !d_gt) return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  if (npix == 0) return RA_OK;
  size_t want = (npix + kFgThreads - 1) / kFgThreads;
  const int blocks = (int)(want < (size_t)kFgBlocks ? want : (size_t)kFgBlocks);
  double *partial = y_gt != nullptr ? reinterpret_cast<double *>(ws) : nullptr;
  fg_head_kernel<<<blocks, kFgThreads, 0, s>>>(logits, npix, nsc, nori, y_gt, d_gt, y_out, d_out, y_hard, partial);
  int rc = ra::finish_launch("fg_head_kernel");
  if (rc != RA_OK || y_gt == nullptr) return rc;
  fg_finalize_kernel<<<1, 32, 0, s>>>(partial, blocks, (double)npix, nori > 0 ? 1 : 0, loss_is_bce, out);
  return ra::finish_launch("fg_finalize_kernel");
}

extern "C" int ra_fg_head_bwd_f32(const float *y_out, const float *d_out, size_t npix, int nsc, int nori,
                                  const float *y_gt, const float *d_gt, int loss_is_bce, const void *ws, float *d_logits,
                                  void *stream) {
  if (nsc < 1 || nori < 0 || nsc > kFgMaxC || nori > kFgMaxC) return RA_ERR_INVALID_ARG;
  if (!y_out || !y_gt || !ws || !d_logits || (nori > 0 && (!d_out || !d_gt))) return RA_ERR_INVALID_ARG;
  if (npix == 0) return RA_OK;
  size_t want = (npix + kFgThreads - 1) / kFgThreads;
  const int blocks = (int)(want < (size_t)kFgBlocks ? want : (size_t)kFgBlocks);  // the forward head's grid
  fg_head_bwd_kernel<<<blocks, kFgThreads, 0, ra::as_stream(stream)>>>(y_out, d_out, npix, nsc, nori, y_gt, d_gt,
                                                                       loss_is_bce,
                                                                       reinterpret_cast<const double *>(ws), blocks,
                                                                       d_logits);
  return ra::finish_launch("fg_head_bwd_kernel");
}

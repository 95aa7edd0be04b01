// Glue kernels of the TRAINING STEP (full_model.py:1039-1057, box_model.py:635-652; runner.py:98-105
// `sess.run([loss, train_step])`): what TensorFlow's graph does between the per-block gradients and the optimiser.
//
//   ra_param_gather_f32     flat parameter bucket -> every device-side weight image of the model in ONE launch
//                           (plain copies, permuted / padded filters, and the hi / lo tf32 images of the tcgen05
//                           convolution): after apply_gradients (full_model.py:1056) the next forward sees the new
//                           weights without a host round trip.
//   ra_param_scatter_f32    per-tensor gradients in their device layouts -> the flat gradient bucket (the buffer the
//                           NCCL all-reduce and ra_adam_step_f32 run on).
//   ra_bn_fold_f32          eval-mode scale / shift rows of every (layer, step) BN copy from gamma, beta, the EMA
//                           shadows and the conv bias (nnlib.py:113-119).
//   ra_weight_decay_f32     sum_i wd[i] * p[i]^2 / 2 (nnlib.py:59-61) of the CURRENT parameters, for the loss value.
//   ra_sum_groups_f32, ra_add_f32, ra_split_channels_f32: the accumulations autodiff inserts where a tensor has
//                           several consumers (skip connections, the shared first controller layer).
//   ra_score_bwd_f32        backward of s = sigmoid([h, core] w + b) (full_model.py:821-822, box_model.py:508-511).
//   ra_knob_box_bwd_f32     scheduled sampling: ctr = kb * ctr_gt + (1 - kb) * ctr_ctrl (full_model.py:760-776).
//   ra_iou_box_coord_bwd_f32  gradient of the box loss through modellib.f_iou_box (modellib.py:206-238) when
//                           opt['use_iou_box'] feeds the per-step coordinate IoUs to the loss (full_model.py:750-754,
//                           :926-929).
#include "common.cuh"

namespace {

constexpr int kT = 256;

inline int grid_for(size_t n, int per_sm = 8) {
  size_t b = (n + kT - 1) / kT;
  const size_t cap = (size_t)ra::kNumSMs * per_sm;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// segment of element i: largest s with seg_start[s] <= i
__device__ __forceinline__ int find_seg(const long long *__restrict__ seg_start, int nseg, long long i) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (seg_start[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// code = 0: constant zero; else ((flat index + 1) << 2) | kind, kind 0 = value, 1 = hi (nearest tf32), 2 = lo = v - hi
__global__ void __launch_bounds__(kT) param_gather_kernel(const float *__restrict__ flat, const int *__restrict__ codes,
                                                          const long long *__restrict__ seg_start,
                                                          float *const *__restrict__ seg_dst, int nseg,
                                                          long long total) {
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int s = find_seg(seg_start, nseg, i);
    const int code = codes[i];
    float v = 0.f;
    if (code != 0) {
      const float p = flat[(code >> 2) - 1];
      const int kind = code & 3;
      if (kind == 0) {
        v = p;
      } else {
        const float hi = __uint_as_float((__float_as_uint(p) + 0x1000u) & 0xFFFFE000u);
        v = (kind == 1) ? hi : (p - hi);
      }
    }
    seg_dst[s][i - seg_start[s]] = v;
  }
}

__global__ void __launch_bounds__(kT) param_scatter_kernel(float *__restrict__ flat, const int *__restrict__ codes,
                                                           const long long *__restrict__ seg_start,
                                                           const float *const *__restrict__ seg_src, int nseg,
                                                           long long total) {
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const int code = codes[i];
    if (code == 0) continue;
    const int s = find_seg(seg_start, nseg, i);
    flat[(code >> 2) - 1] = seg_src[s][i - seg_start[s]];
  }
}

__global__ void __launch_bounds__(kT) bn_fold_kernel(const float *__restrict__ gamma, const float *__restrict__ beta,
                                                     const float *__restrict__ mean, const float *__restrict__ var,
                                                     const float *__restrict__ bias, int T, int C, float eps,
                                                     float *__restrict__ scale, float *__restrict__ shift) {
  const int i = blockIdx.x * kT + threadIdx.x;
  if (i >= T * C) return;
  const int c = i % C;
  // the host fold of load_weights: inv = (1 / sqrt(var + eps)) * gamma; shift = beta - mean * inv + bias * inv
  const float inv = __fmul_rn(__fdiv_rn(1.0f, sqrtf(__fadd_rn(var[i], eps))), gamma[i]);
  scale[i] = inv;
  shift[i] = __fadd_rn(__fsub_rn(beta[i], __fmul_rn(mean[i], inv)), __fmul_rn(bias[c], inv));
}

__global__ void __launch_bounds__(kT) weight_decay_kernel(const float *__restrict__ p, const float *__restrict__ wd,
                                                          size_t n, double *__restrict__ partial) {
  __shared__ double red[kT];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) {
    const float w = wd[i];
    if (w != 0.f) s += 0.5 * (double)w * (double)p[i] * (double)p[i];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = kT / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

__global__ void weight_decay_finalize_kernel(const double *__restrict__ partial, int n, float *__restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partial[i];
    *out = (float)s;
  }
}

// dst[i] = sum_g src[g*n + i] (fixed order)
__global__ void __launch_bounds__(kT) sum_groups_kernel(const float *__restrict__ src, int G, size_t n, size_t n4,
                                                        float *__restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n4; i += (size_t)gridDim.x * kT) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < G; ++g) {
      const float4 v = ra::ldg_stream4(src + (size_t)g * n + i * 4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4 *>(dst)[i] = a;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += kT) {
      float a = 0.f;
      for (int g = 0; g < G; ++g) a += src[(size_t)g * n + i];
      dst[i] = a;
    }
}

__global__ void __launch_bounds__(kT) add_kernel(float *__restrict__ dst, const float *__restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < n; i += (size_t)gridDim.x * kT) dst[i] += src[i];
}

__global__ void __launch_bounds__(kT) split_channels_kernel(const float *__restrict__ src, size_t npix, int C1, int C2,
                                                            float *__restrict__ dst1, float *__restrict__ dst2,
                                                            int accumulate2) {
  const int C = C1 + C2;
  const size_t total = npix * C;
  for (size_t i = (size_t)blockIdx.x * kT + threadIdx.x; i < total; i += (size_t)gridDim.x * kT) {
    const size_t p = i / C;
    const int c = (int)(i - p * C);
    const float v = src[i];
    if (c < C1) {
      if (dst1) dst1[p * C1 + c] = v;
    } else if (dst2) {
      float *d = dst2 + p * C2 + (c - C1);
      *d = accumulate2 ? (*d + v) : v;
    }
  }
}

// one CTA per row n = t*B + b: dpre = d_s * s (1 - s); d_h = dpre * w[:Hd]; d_core (+)= dpre * w[Hd:]
__global__ void __launch_bounds__(kT) score_bwd_kernel(const float *__restrict__ s_out, const float *__restrict__ d_s,
                                                       int B, int T, const float *__restrict__ w, int Hd, int Cd,
                                                       float *__restrict__ dpre, float *__restrict__ d_h,
                                                       float *__restrict__ d_core, int accumulate_core) {
  const int n = blockIdx.x, t = n / B, b = n - t * B;
  const float s = s_out[(size_t)b * T + t];
  const float dp = d_s[(size_t)b * T + t] * s * (1.0f - s);
  if (threadIdx.x == 0) dpre[n] = dp;
  for (int i = threadIdx.x; i < Hd; i += kT) d_h[(size_t)n * Hd + i] = dp * w[i];
  if (d_core != nullptr)
    for (int i = threadIdx.x; i < Cd; i += kT) {
      float *d = d_core + (size_t)n * Cd + i;
      const float v = dp * w[Hd + i];
      *d = accumulate_core ? (*d + v) : v;
    }
}

// d_box[n] = (keep * d_mixed[ctr, size], d_mixed[lg_var]) + d_pre, keep = 1 - knob_box[b, t], n = t*B + b
__global__ void knob_box_bwd_kernel(const float *__restrict__ d_mixed, const float *__restrict__ d_pre,
                                    const float *__restrict__ knob_box, int B, int T, float *__restrict__ d_box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T * 6) return;
  const int n = i / 6, k = i - n * 6, t = n / B, b = n - t * B;
  float v = d_mixed[i];
  if (k < 4) v *= 1.0f - knob_box[(size_t)b * T + t];
  if (d_pre != nullptr) v += d_pre[i];
  d_box[i] = v;
}

// thread per n = t*B + b: d(ctr, size) += gradient of sum_m wgt[m] * f_iou_box(box_n, gt_m),
// wgt[m] = -match_box[b,t,m] / (B * max(1, sum match_box[b]))  (the box loss of full_model.py:942-973 on the
// coordinate IoU, modellib.py:206-238: no eps, strict overlap test).
__global__ void iou_box_coord_bwd_kernel(const float *__restrict__ box, const float *__restrict__ tl_gt,
                                         const float *__restrict__ br_gt, const float *__restrict__ match_box, int B,
                                         int T, float scale, float *__restrict__ d_box) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= B * T) return;
  const int t = n / B, b = n - t * B;
  const float *bo = box + (size_t)n * RA_BOX_STRIDE;
  const float *mb = match_box + (size_t)b * T * T;
  float cnt = 0.f;
  for (int i = 0; i < T * T; ++i) cnt += mb[i];
  cnt = fmaxf(cnt, 1.0f);
  float ctr[2] = {bo[RA_BOX_CTR_Y], bo[RA_BOX_CTR_X]}, size[2] = {bo[RA_BOX_SIZE_Y], bo[RA_BOX_SIZE_X]};
  float tl[2], br[2];
  for (int a = 0; a < 2; ++a) {
    tl[a] = ctr[a] - size[a] / 2.0f;
    br[a] = ctr[a] + size[a] / 2.0f;
  }
  const float area_a = (br[0] - tl[0]) * (br[1] - tl[1]);
  float d_tl[2] = {0.f, 0.f}, d_br[2] = {0.f, 0.f};
  for (int m = 0; m < T; ++m) {
    const float wgt = -scale * mb[(size_t)t * T + m] / ((float)B * cnt);
    if (wgt == 0.f) continue;
    const float *tg = tl_gt + ((size_t)b * T + m) * 2, *bg = br_gt + ((size_t)b * T + m) * 2;
    float lo[2], hi[2], ext[2];
    for (int a = 0; a < 2; ++a) {
      lo[a] = fmaxf(tl[a], tg[a]);
      hi[a] = fminf(br[a], bg[a]);
      ext[a] = hi[a] - lo[a];
    }
    const bool flag = ext[0] > 0.f && ext[1] > 0.f;
    const float inter = flag ? ext[0] * ext[1] : 0.f;
    const float area_b = (bg[0] - tg[0]) * (bg[1] - tg[1]);
    const float uni = area_a + area_b - inter;
    const float d_inter = wgt * (1.0f / uni + inter / (uni * uni));
    const float d_area = -wgt * inter / (uni * uni);
    for (int a = 0; a < 2; ++a) {
      const int o = 1 - a;
      const float d_ext = flag ? d_inter * ext[o] : 0.f;
      if (br[a] <= bg[a]) d_br[a] += d_ext;
      if (tl[a] >= tg[a]) d_tl[a] -= d_ext;
      const float side = br[o] - tl[o];
      d_br[a] += d_area * side;
      d_tl[a] -= d_area * side;
    }
  }
  for (int a = 0; a < 2; ++a) {
    d_box[(size_t)n * 6 + a] += d_tl[a] + d_br[a];
    d_box[(size_t)n * 6 + 2 + a] += (d_br[a] - d_tl[a]) / 2.0f;
  }
}

// modellib.f_weighted_coverage (modellib.py:268-302): cov = (1/B) sum_b sum_m max_n iou[b,n,m] * wt[b,m],
// wt = area_m / (sum_m area_m + [area_m == 0]).  Also the gradient coefficients coeff[b,n,m] = wt[b,m] at the FIRST
// arg-max n (zero elsewhere): handed to ra_iou_loss_bwd_f32 in place of the matching (their sum per example is <= 1, so
// its 1 / max(1, sum) normalisation is the identity).  area from `area` [B,M], or - rect != NULL - the pixel count of
// the filled GT rectangle (get_filled_box_idx, modellib.py:704-749: idx >= tl and idx <= br on the pixel grid).
__global__ void __launch_bounds__(256) wt_cov_kernel(const float *__restrict__ iou, const float *__restrict__ area,
                                                     const float *__restrict__ rect, int H, int W, int B, int N, int M,
                                                     float *__restrict__ cov, float *__restrict__ coeff) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float tot = 0.f;
    for (int pass = 0; pass < 2; ++pass)
      for (int m = 0; m < M; ++m) {
        float ar;
        if (rect != nullptr) {
          const float *r = rect + ((size_t)b * M + m) * 4;
          const float y0 = fmaxf(ceilf(r[0]), 0.f), x0 = fmaxf(ceilf(r[1]), 0.f);
          const float y1 = fminf(floorf(r[2]), (float)(H - 1)), x1 = fminf(floorf(r[3]), (float)(W - 1));
          ar = fmaxf(y1 - y0 + 1.f, 0.f) * fmaxf(x1 - x0 + 1.f, 0.f);
        } else {
          ar = area[(size_t)b * M + m];
        }
        if (pass == 0) {
          tot += ar;
          continue;
        }
        const float wt = ar / (tot + (ar == 0.f ? 1.f : 0.f));
        float best = -INFINITY;
        int arg = 0;
        for (int n = 0; n < N; ++n) {
          const float v = iou[((size_t)b * N + n) * M + m];
          if (v > best) {
            best = v;
            arg = n;
          }
          if (coeff != nullptr) coeff[((size_t)b * N + n) * M + m] = 0.f;
        }
        if (coeff != nullptr) coeff[((size_t)b * N + arg) * M + m] = wt;
        acc += best * wt;
      }
  }
  acc = ra::block_sum(acc, red);
  if (threadIdx.x == 0) *cov = acc / (float)B;
}

// segm_loss_fn / box_loss_fn == 'wt_cov' (full_model.py:967,1013-1014): replace the IoU losses of the loss block by
// the negative weighted coverages and rebuild the total (the weight-decay term is added afterwards by the caller).
__global__ void loss_select_kernel(float *__restrict__ scal, const float *__restrict__ box_cov,
                                   const float *__restrict__ segm_cov, float mix, float segm_coeff) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (box_cov != nullptr) scal[RA_LOSS_BOX] = -*box_cov;
  if (segm_cov != nullptr) scal[RA_LOSS_SEGM] = -*segm_cov;
  scal[RA_LOSS_TOTAL] = scal[RA_LOSS_BOX] + segm_coeff * scal[RA_LOSS_SEGM] + mix * scal[RA_LOSS_CONF];
}

}  // namespace

extern "C" int ra_wt_cov_f32(const float *iou, const float *area, const float *rect, int H, int W, int B, int N, int M,
                             float *cov, float *coeff, void *stream) {
  if (B < 1 || N < 1 || M < 1 || !iou || !cov || (!area && !rect)) return RA_ERR_INVALID_ARG;
  wt_cov_kernel<<<1, 256, 0, ra::as_stream(stream)>>>(iou, area, rect, H, W, B, N, M, cov, coeff);
  return ra::finish_launch("wt_cov_kernel");
}

extern "C" int ra_loss_select_f32(float *scal, const float *box_cov, const float *segm_cov, float mix, float segm_coeff,
                                  void *stream) {
  if (!scal) return RA_ERR_INVALID_ARG;
  loss_select_kernel<<<1, 32, 0, ra::as_stream(stream)>>>(scal, box_cov, segm_cov, mix, segm_coeff);
  return ra::finish_launch("loss_select_kernel");
}

extern "C" int ra_param_gather_f32(const float *flat, const int32_t *codes, const long long *seg_start,
                                   float *const *seg_dst, int nseg, long long total, void *stream) {
  if (nseg < 0 || total < 0) return RA_ERR_INVALID_ARG;
  if (nseg == 0 || total == 0) return RA_OK;
  if (!flat || !codes || !seg_start || !seg_dst) return RA_ERR_INVALID_ARG;
  param_gather_kernel<<<grid_for((size_t)total), kT, 0, ra::as_stream(stream)>>>(flat, codes, seg_start, seg_dst, nseg,
                                                                               total);
  return ra::finish_launch("param_gather_kernel");
}

extern "C" int ra_param_scatter_f32(float *flat, const int32_t *codes, const long long *seg_start,
                                    const float *const *seg_src, int nseg, long long total, void *stream) {
  if (nseg < 0 || total < 0) return RA_ERR_INVALID_ARG;
  if (nseg == 0 || total == 0) return RA_OK;
  if (!flat || !codes || !seg_start || !seg_src) return RA_ERR_INVALID_ARG;
  param_scatter_kernel<<<grid_for((size_t)total), kT, 0, ra::as_stream(stream)>>>(flat, codes, seg_start, seg_src, nseg,
                                                                                total);
  return ra::finish_launch("param_scatter_kernel");
}

extern "C" int ra_bn_fold_f32(const float *gamma, const float *beta, const float *ema_mean, const float *ema_var,
                              const float *bias, int T, int C, float eps, float *scale, float *shift, void *stream) {
  if (T < 1 || C < 1 || !gamma || !beta || !ema_mean || !ema_var || !bias || !scale || !shift) return RA_ERR_INVALID_ARG;
  bn_fold_kernel<<<(T * C + kT - 1) / kT, kT, 0, ra::as_stream(stream)>>>(gamma, beta, ema_mean, ema_var, bias, T, C, eps,
                                                                         scale, shift);
  return ra::finish_launch("bn_fold_kernel");
}

extern "C" size_t ra_weight_decay_workspace(void) { return (size_t)ra::kNumSMs * 2 * sizeof(double); }

extern "C" int ra_weight_decay_f32(const float *param, const float *wd, size_t n, void *ws, float *out, void *stream) {
  if (!param || !wd || !ws || !out) return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  const int ctas = ra::kNumSMs * 2;
  weight_decay_kernel<<<ctas, kT, 0, s>>>(param, wd, n, reinterpret_cast<double *>(ws));
  int rc = ra::finish_launch("weight_decay_kernel");
  if (rc != RA_OK) return rc;
  weight_decay_finalize_kernel<<<1, 32, 0, s>>>(reinterpret_cast<double *>(ws), ctas, out);
  return ra::finish_launch("weight_decay_finalize_kernel");
}

extern "C" int ra_sum_groups_f32(const float *src, int G, size_t n, float *dst, void *stream) {
  if (G < 1 || !src || !dst) return RA_ERR_INVALID_ARG;
  if (n == 0) return RA_OK;
  // float4 path needs 16-byte aligned rows (n % 4 == 0 keeps every group aligned); otherwise the scalar tail does it all
  const bool vec = !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) && !(n & 3);
  sum_groups_kernel<<<vec ? grid_for(n / 4) : 1, kT, 0, ra::as_stream(stream)>>>(src, G, n, vec ? n / 4 : 0, dst);
  return ra::finish_launch("sum_groups_kernel");
}

extern "C" int ra_add_f32(float *dst, const float *src, size_t n, void *stream) {
  if (n == 0) return RA_OK;
  if (!dst || !src) return RA_ERR_INVALID_ARG;
  add_kernel<<<grid_for(n), kT, 0, ra::as_stream(stream)>>>(dst, src, n);
  return ra::finish_launch("add_kernel");
}

extern "C" int ra_split_channels_f32(const float *src, size_t npix, int C1, int C2, float *dst1, float *dst2,
                                     int accumulate2, void *stream) {
  if (C1 < 0 || C2 < 0 || C1 + C2 < 1) return RA_ERR_INVALID_ARG;
  if (npix == 0) return RA_OK;
  if (!src) return RA_ERR_INVALID_ARG;
  split_channels_kernel<<<grid_for(npix * (C1 + C2)), kT, 0, ra::as_stream(stream)>>>(src, npix, C1, C2, dst1, dst2,
                                                                                     accumulate2);
  return ra::finish_launch("split_channels_kernel");
}

extern "C" int ra_score_bwd_f32(const float *s_out, const float *d_s, int B, int T, const float *w, int Hd, int Cd,
                                float *dpre, float *d_h, float *d_core, int accumulate_core, void *stream) {
  if (B < 0 || T < 1 || Hd < 1 || Cd < 0) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!s_out || !d_s || !w || !dpre || !d_h || (Cd > 0 && !d_core)) return RA_ERR_INVALID_ARG;
  score_bwd_kernel<<<B * T, kT, 0, ra::as_stream(stream)>>>(s_out, d_s, B, T, w, Hd, Cd, dpre, d_h,
                                                            Cd > 0 ? d_core : nullptr, accumulate_core);
  return ra::finish_launch("score_bwd_kernel");
}

extern "C" int ra_knob_box_bwd_f32(const float *d_mixed, const float *d_pre, const float *knob_box, int B, int T,
                                   float *d_box, void *stream) {
  if (B < 0 || T < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!d_mixed || !knob_box || !d_box) return RA_ERR_INVALID_ARG;
  knob_box_bwd_kernel<<<(B * T * 6 + 127) / 128, 128, 0, ra::as_stream(stream)>>>(d_mixed, d_pre, knob_box, B, T, d_box);
  return ra::finish_launch("knob_box_bwd_kernel");
}

extern "C" int ra_iou_box_coord_bwd_f32(const float *box, const float *tl_gt, const float *br_gt,
                                        const float *match_box, int B, int T, float scale, float *d_box,
                                        void *stream) {
  if (B < 0 || T < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!box || !tl_gt || !br_gt || !match_box || !d_box) return RA_ERR_INVALID_ARG;
  iou_box_coord_bwd_kernel<<<(B * T + 63) / 64, 64, 0, ra::as_stream(stream)>>>(box, tl_gt, br_gt, match_box, B, T, scale,
                                                                               d_box);
  return ra::finish_launch("iou_box_coord_bwd_kernel");
}

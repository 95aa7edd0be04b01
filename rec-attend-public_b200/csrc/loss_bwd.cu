// Backward of the matching loss block (full_model.py:942-1034) down to the model outputs: the gradients TensorFlow's
// autodiff hands to y_out, attn_box and s_out.  The matchings are constants (ops.NoGradient("Hungarian"),
// modellib.py:11).  Second building block of the backward pass.
//
//   L_iou = -(scale / B) * sum_b (1 / cnt_b) * sum_{n,m} match[b,n,m] * I / U,   cnt_b = max(1, sum match[b])
//       I = sum_p a[n,p] g[m,p],  U = sum_p a[n,p] + sum_p g[m,p] - I + H*W*1e-5     (modellib.py:104-155)
//   dL/da[b,n,p] = sum_m ( c1[b,n,m] * g[b,m,p] + c0[b,n,m] ),
//       w = scale * match / (B * cnt_b),  c1 = -w (U + I) / U^2,  c0 = w I / U^2.
//   g is a stack of ground-truth masks (segmentation loss, :983-1012) or of filled rectangles (box loss, :931-973).
//   Two launches: one CTA per (b, n) reduces I, sum a, sum g for its matched m (normally one) and stores the
//   coefficients; an HBM-bound pass writes the gradient (reads the matched g rows, writes da).
//
//   L_conf = (scale / (B T)) * sum_{b,t} [ -ms log(cummin_t(s) + 1e-5) - (1 - ms) log(1 - cummax^rev_t(s) + 1e-5) ]
//   (modellib.py:316-339, 430-437): the gradient flows to the arg-min of each prefix and the arg-max of each suffix.
#include "common.cuh"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ float g_value(const float *__restrict__ gm, const float *__restrict__ rc, int W, int p) {
  if (gm != nullptr) return gm[p];
  const int yy = p / W;
  const float fy = (float)yy, fx = (float)(p - yy * W);
  return (fy >= rc[0] && fx >= rc[1] && fy <= rc[2] && fx <= rc[3]) ? 1.f : 0.f;  // modellib.py:745-747
}

__global__ void __launch_bounds__(kT) iou_bwd_coeff_kernel(const float *__restrict__ a, size_t a_bstride,
                                                           const float *__restrict__ b_masks,
                                                           const float *__restrict__ b_rect,
                                                           const float *__restrict__ match, int B, int N, int M, int H,
                                                           int W, float scale, float *__restrict__ c1,
                                                           float *__restrict__ c0) {
  __shared__ float red[32];
  const int n = blockIdx.x, b = blockIdx.y;
  const int HW = H * W;
  const float *mt = match + (size_t)b * N * M;
  float cnt = 0.f;
  for (int i = threadIdx.x; i < N * M; i += kT) cnt += mt[i];
  cnt = fmaxf(ra::block_sum(cnt, red), 1.0f);
  const float *an = a + (size_t)b * a_bstride + (size_t)n * HW;
  float sa = 0.f;
  bool any = false;
  for (int m = 0; m < M; ++m) any |= (mt[n * M + m] != 0.f);
  if (any) {
    for (int p = threadIdx.x; p < HW; p += kT) sa += an[p];
    sa = ra::block_sum(sa, red);
  }
  for (int m = 0; m < M; ++m) {
    const float w_nm = mt[n * M + m];  // block-uniform
    float k1 = 0.f, k0 = 0.f;
    if (w_nm != 0.f) {
      const float *gm = b_masks ? b_masks + ((size_t)b * M + m) * HW : nullptr;
      const float *rc = b_rect ? b_rect + ((size_t)b * M + m) * 4 : nullptr;
      float inter = 0.f, sg = 0.f;
      for (int p = threadIdx.x; p < HW; p += kT) {
        const float g = g_value(gm, rc, W, p);
        inter = fmaf(an[p], g, inter);
        sg += g;
      }
      inter = ra::block_sum(inter, red);
      sg = ra::block_sum(sg, red);
      const float U = sa + sg - inter + (float)HW * 1e-5f;
      const float w = scale * w_nm / ((float)B * cnt);
      k1 = -w * (U + inter) / (U * U);
      k0 = w * inter / (U * U);
    }
    if (threadIdx.x == 0) {
      c1[((size_t)b * N + n) * M + m] = k1;
      c0[((size_t)b * N + n) * M + m] = k0;
    }
  }
}

__global__ void __launch_bounds__(kT) iou_bwd_apply_kernel(const float *__restrict__ b_masks,
                                                           const float *__restrict__ b_rect,
                                                           const float *__restrict__ c1, const float *__restrict__ c0,
                                                           int N, int M, int H, int W, float *__restrict__ da,
                                                           size_t a_bstride) {
  __shared__ float k1_s[64];
  __shared__ int m_s[64];
  __shared__ int cnt_s;
  __shared__ float k0_s;
  const int n = blockIdx.y, b = blockIdx.z;
  const int HW = H * W;
  if (threadIdx.x == 0) {  // the matched columns of this row (a matching has at most one)
    int k = 0;
    float k0 = 0.f;
    for (int m = 0; m < M; ++m) {
      const float v1 = c1[((size_t)b * N + n) * M + m], v0 = c0[((size_t)b * N + n) * M + m];
      if (v1 != 0.f || v0 != 0.f) {
        k1_s[k] = v1;
        m_s[k] = m;
        ++k;
        k0 += v0;
      }
    }
    cnt_s = k;
    k0_s = k0;
  }
  __syncthreads();
  const int k = cnt_s;
  const float k0 = k0_s;
  float *dn = da + (size_t)b * a_bstride + (size_t)n * HW;
  for (int p = blockIdx.x * kT + threadIdx.x; p < HW; p += gridDim.x * kT) {
    float v = k0;
    for (int j = 0; j < k; ++j) {
      const int m = m_s[j];
      const float *gm = b_masks ? b_masks + ((size_t)b * M + m) * HW : nullptr;
      const float *rc = b_rect ? b_rect + ((size_t)b * M + m) * 4 : nullptr;
      v = fmaf(k1_s[j], g_value(gm, rc, W, p), v);
    }
    dn[p] = v;
  }
}

__global__ void conf_bwd_kernel(const float *__restrict__ s_out, const float *__restrict__ match, int B, int T, int M,
                                float scale, float *__restrict__ ds) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float *s = s_out + (size_t)b * T;
  float *d = ds + (size_t)b * T;
  for (int t = 0; t < T; ++t) d[t] = 0.f;
  const float k = scale / ((float)B * (float)T);
  // prefix minimum: gradient of -ms[t] * log(min_{i<=t} s[i] + eps) goes to the arg-min (torch.cummin: the later
  // index wins a tie; ties have measure zero for a sigmoid output)
  float run = INFINITY;
  int arg = 0;
  for (int t = 0; t < T; ++t) {
    if (s[t] <= run) {
      run = s[t];
      arg = t;
    }
    float ms = 0.f;
    for (int m = 0; m < M; ++m) ms += match[((size_t)b * T + t) * M + m];
    d[arg] += k * (-ms / (run + 1e-5f));
  }
  // suffix maximum (modellib.f_cum_max runs from the end): -(1 - ms[t]) * log(1 - max_{i>=t} s[i] + eps)
  run = -INFINITY;
  arg = T - 1;
  for (int t = T - 1; t >= 0; --t) {
    if (s[t] >= run) {
      run = s[t];
      arg = t;
    }
    float ms = 0.f;
    for (int m = 0; m < M; ++m) ms += match[((size_t)b * T + t) * M + m];
    d[arg] += k * ((1.f - ms) / (1.f - run + 1e-5f));
  }
}

}  // namespace

extern "C" size_t ra_iou_loss_bwd_workspace(int B, int N, int M) {
  if (B < 1 || N < 1 || M < 1) return 0;
  return (size_t)2 * B * N * M * sizeof(float);
}

extern "C" int ra_iou_loss_bwd_f32(const float *a, size_t a_bstride, const float *b_masks, const float *b_rect,
                                   const float *match, int B, int N, int M, int H, int W, float scale, void *ws,
                                   float *da, void *stream) {
  if (B < 0 || N < 1 || M < 1 || M > 64 || H < 1 || W < 1) return RA_ERR_INVALID_ARG;
  if ((b_masks == nullptr) == (b_rect == nullptr)) return RA_ERR_INVALID_ARG;  // exactly one form of g
  if (B == 0) return RA_OK;
  if (!a || !match || !ws || !da || a_bstride < (size_t)N * H * W) return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  float *c1 = reinterpret_cast<float *>(ws), *c0 = c1 + (size_t)B * N * M;
  iou_bwd_coeff_kernel<<<dim3(N, B), kT, 0, s>>>(a, a_bstride, b_masks, b_rect, match, B, N, M, H, W, scale, c1, c0);
  int rc = ra::finish_launch("iou_bwd_coeff_kernel");
  if (rc != RA_OK) return rc;
  int bx = (H * W + kT * 4 - 1) / (kT * 4);
  if (bx > 32) bx = 32;
  iou_bwd_apply_kernel<<<dim3(bx, N, B), kT, 0, s>>>(b_masks, b_rect, c1, c0, N, M, H, W, da, a_bstride);
  return ra::finish_launch("iou_bwd_apply_kernel");
}

extern "C" int ra_conf_loss_bwd_f32(const float *s_out, const float *match, int B, int T, int M, float scale, float *ds,
                                    void *stream) {
  if (B < 0 || T < 1 || M < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!s_out || !match || !ds) return RA_ERR_INVALID_ARG;
  conf_bwd_kernel<<<(B + 63) / 64, 64, 0, ra::as_stream(stream)>>>(s_out, match, B, T, M, scale, ds);
  return ra::finish_launch("conf_bwd_kernel");
}

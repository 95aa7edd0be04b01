// Instance-label post-processing of the decoder outputs — utils/postprocess.py as chained by
// full_model_eval.py:112-125: apply_confidence (:17-31) -> apply_one_label (:34-55) -> apply_threshold (:6-14)
// [-> mask_foreground (:146-155) -> remove_tiny (:106-143)].
//
// The reference walks [B,T,H,W] float arrays five times on the host (numpy).  Here one pass over y_out does the
// confidence weighting, the first-max argmax over the T instances, the threshold and the foreground mask per pixel
// and emits an int32 label map (0 = background, t+1 = instance t) plus the per-instance areas; a second, small pass
// removes tiny instances and (optionally) expands the label map into the reference's dense [B,T,H,W] float form.
// HBM-bound: B*T*H*W*4 bytes read once.
#include "common.cuh"

namespace {

constexpr int kPpThreads = 256;
constexpr int kPpMaxT = 64;

// pass 1: grid (chunks of 4*kPpThreads pixels, B)
__global__ void __launch_bounds__(kPpThreads) pp_label_kernel(const float *__restrict__ y_out,
                                                              const float *__restrict__ s_out,
                                                              const float *__restrict__ fg, int T, int HW, double thresh,
                                                              int *__restrict__ label, double *__restrict__ area_ws) {
  __shared__ float s_s[kPpMaxT];
  __shared__ double area_s[kPpMaxT];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < T) {
    s_s[tid] = s_out[(size_t)b * T + tid];
    area_s[tid] = 0.0;
  }
  __syncthreads();
  const int p0 = (blockIdx.x * kPpThreads + tid) * 4;
  if (p0 < HW) {
    const float *yb = y_out + (size_t)b * T * HW + p0;
    float best[4] = {0.f, 0.f, 0.f, 0.f};
    int arg[4] = {0, 0, 0, 0};
    const bool vec = p0 + 4 <= HW;  // HW % 4 == 0 is checked by the launcher, so always true; kept for safety
    for (int t = 0; t < T; ++t) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) {
        const float4 q = ra::ldg_stream4(yb + (size_t)t * HW);
        v[0] = q.x;
        v[1] = q.y;
        v[2] = q.z;
        v[3] = q.w;
      }
      const float s = s_s[t];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float w = __fmul_rn(v[e], s);  // apply_confidence: y_out * s_mask (fp32)
        // np.argmax: the FIRST maximum wins (strict >); t == 0 initialises
        if (t == 0 || w > best[e]) {
          best[e] = w;
          arg[e] = t;
        }
      }
    }
    int lab[4];
    float4 f4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (fg != nullptr) f4 = *reinterpret_cast<const float4 *>(fg + (size_t)b * HW + p0);
    const float fv[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // apply_one_label keeps only the arg-max instance, apply_threshold compares it with a Python float (double)
      const bool on = (double)best[e] > thresh;
      // mask_foreground multiplies the {0,1} mask by fg; remove_tiny sums that product
      const float contrib = on ? fv[e] : 0.f;
      lab[e] = (on && fv[e] != 0.f) ? arg[e] + 1 : 0;
      if (contrib != 0.f) atomicAdd(&area_s[arg[e]], (double)contrib);
    }
    *reinterpret_cast<int4 *>(label + (size_t)b * HW + p0) = make_int4(lab[0], lab[1], lab[2], lab[3]);
  }
  __syncthreads();
  if (tid < T && area_s[tid] != 0.0) atomicAdd(&area_ws[(size_t)b * T + tid], area_s[tid]);
}

// pass 2: tiny-instance removal, confidences, optional dense expansion.  grid (chunks, B)
__global__ void __launch_bounds__(kPpThreads) pp_finalize_kernel(const float *__restrict__ s_out,
                                                                 const float *__restrict__ fg,
                                                                 const double *__restrict__ area_ws, int T, int HW,
                                                                 float tiny_thr, int *__restrict__ label,
                                                                 float *__restrict__ y_hard, float *__restrict__ conf,
                                                                 float *__restrict__ area) {
  __shared__ int keep_s[kPpMaxT];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid < T) {
    const float a = (float)area_ws[(size_t)b * T + tid];
    // remove_tiny_single: is_not_tiny = size > threshold (off when threshold == 0, postprocess.py:113-114)
    const int keep = (tiny_thr == 0.f || a > tiny_thr) ? 1 : 0;
    keep_s[tid] = keep;
    if (blockIdx.x == 0) {
      const float s_hard = s_out[(size_t)b * T + tid] > 0.5f ? 1.f : 0.f;  // apply_confidence: s_out > 0.5
      conf[(size_t)b * T + tid] = s_hard * (float)keep;
      if (area != nullptr) area[(size_t)b * T + tid] = a;
    }
  }
  __syncthreads();
  if (tiny_thr == 0.f && y_hard == nullptr) return;  // nothing to remove, no dense output: the label map stands
  const int p0 = (blockIdx.x * kPpThreads + tid) * 4;
  if (p0 >= HW) return;
  int *lp = label + (size_t)b * HW + p0;
  int4 l4 = *reinterpret_cast<const int4 *>(lp);
  int lab[4] = {l4.x, l4.y, l4.z, l4.w};
  bool changed = false;
#pragma unroll
  for (int e = 0; e < 4; ++e)
    if (lab[e] > 0 && !keep_s[lab[e] - 1]) {
      lab[e] = 0;
      changed = true;
    }
  if (changed) *reinterpret_cast<int4 *>(lp) = make_int4(lab[0], lab[1], lab[2], lab[3]);
  if (y_hard != nullptr) {
    float4 f4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (fg != nullptr) f4 = *reinterpret_cast<const float4 *>(fg + (size_t)b * HW + p0);
    for (int t = 0; t < T; ++t) {
      const float4 o = make_float4(lab[0] == t + 1 ? f4.x : 0.f, lab[1] == t + 1 ? f4.y : 0.f,
                                   lab[2] == t + 1 ? f4.z : 0.f, lab[3] == t + 1 ? f4.w : 0.f);
      __stcs(reinterpret_cast<float4 *>(y_hard + ((size_t)b * T + t) * HW + p0), o);
    }
  }
}

}  // namespace

extern "C" size_t ra_postprocess_workspace(int B, int T) { return (size_t)B * T * sizeof(double); }

extern "C" int ra_postprocess_f32(const float *y_out, const float *s_out, const float *fg, int B, int T, int H, int W,
                                  double thresh, float remove_tiny, void *workspace, int32_t *label, float *y_hard,
                                  float *conf, float *area, void *stream) {
  if (B < 0 || T < 1 || H < 1 || W < 1 || remove_tiny < 0.f) return RA_ERR_INVALID_ARG;
  if (T > kPpMaxT || ((size_t)H * W) % 4 != 0) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;  // empty batch: nothing to read or write (pointers may be NULL)
  if (!y_out || !s_out || !workspace || !label || !conf) return RA_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(y_out) & 15) != 0 || (reinterpret_cast<uintptr_t>(label) & 15) != 0 ||
      (fg != nullptr && (reinterpret_cast<uintptr_t>(fg) & 15) != 0) ||
      (y_hard != nullptr && (reinterpret_cast<uintptr_t>(y_hard) & 15) != 0))
    return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  const int HW = H * W;
  cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)B * T * sizeof(double), s);
  if (e != cudaSuccess) {
    ra::set_last_error("cudaMemsetAsync(postprocess workspace)", e);
    return RA_ERR_CUDA;
  }
  dim3 grid((HW / 4 + kPpThreads - 1) / kPpThreads, B);
  pp_label_kernel<<<grid, kPpThreads, 0, s>>>(y_out, s_out, fg, T, HW, thresh, label,
                                              reinterpret_cast<double *>(workspace));
  int rc = ra::finish_launch("pp_label_kernel");
  if (rc != RA_OK) return rc;
  pp_finalize_kernel<<<grid, kPpThreads, 0, s>>>(s_out, fg, reinterpret_cast<const double *>(workspace), T, HW,
                                                 remove_tiny, label, y_hard, conf, area);
  return ra::finish_launch("pp_finalize_kernel");
}

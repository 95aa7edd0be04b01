// Optimiser block of the reference on ONE flat fp32 bucket — full_model.py:1039-1057 / box_model.py:635-652:
// gradients of total_loss (data term + wd * l2_loss(w), nnlib.py:59-61), tf.clip_by_value(g, -1, 1) per element,
// tf.train.AdamOptimizer(lr, epsilon=1e-7).apply_gradients.  The reference builds ~60 small per-variable op chains;
// here every trainable tensor lives in one contiguous bucket (the same buffer the NCCL all-reduce of the gradients
// runs on), so the whole block is a single HBM-bound launch: 5 reads + 3 writes of 4 bytes per element.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) adam_step_kernel(float *__restrict__ param, const float *__restrict__ grad,
                                                        float *__restrict__ m, float *__restrict__ v,
                                                        const float *__restrict__ wd, size_t n, float grad_scale,
                                                        float lr_t, float beta1, float beta2, float eps, float clip) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float p = param[i];
    // average over ranks first (grad_scale = 1/world), then the weight-decay gradient wd*w of this rank's (identical)
    // weights, then the clip: the reference clips the gradient of the TOTAL loss of the global batch
    float g = __fmul_rn(grad[i], grad_scale);
    if (wd != nullptr) g = __fadd_rn(g, __fmul_rn(wd[i], p));
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
    // TF-0.12 ApplyAdam: m += (g - m)(1 - b1); v += (g*g - v)(1 - b2); var -= lr_t * m / (sqrt(v) + eps)
    const float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(g, m[i]), 1.0f - beta1));
    const float vi = __fadd_rn(v[i], __fmul_rn(__fsub_rn(__fmul_rn(g, g), v[i]), 1.0f - beta2));
    m[i] = mi;
    v[i] = vi;
    param[i] = __fsub_rn(p, __fdiv_rn(__fmul_rn(mi, lr_t), __fadd_rn(sqrtf(vi), eps)));
  }
}

}  // namespace

extern "C" int ra_adam_step_f32(float *param, const float *grad, float *m, float *v, const float *wd, size_t n,
                                float grad_scale, float lr, float beta1, float beta2, float eps, float clip, int step_t,
                                void *stream) {
  if (!param || !grad || !m || !v || step_t < 1 || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f))
    return RA_ERR_INVALID_ARG;
  if (n == 0) return RA_OK;
  // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), computed like TF does: in fp32 from the running beta powers
  float b1p = 1.f, b2p = 1.f;
  for (int i = 0; i < step_t; ++i) {
    b1p *= beta1;
    b2p *= beta2;
  }
  const float lr_t = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)ra::kNumSMs * 8) blocks = (size_t)ra::kNumSMs * 8;
  adam_step_kernel<<<(unsigned)blocks, 256, 0, ra::as_stream(stream)>>>(param, grad, m, v, wd, n, grad_scale, lr_t, beta1,
                                                                       beta2, eps, clip);
  return ra::finish_launch("adam_step_kernel");
}

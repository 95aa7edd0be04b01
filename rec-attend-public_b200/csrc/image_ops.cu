// In-graph augmentation of the reference — image_ops.random_transformation (image_ops.py:9-113): zero-pad by `padding`,
// crop at a random offset, optional vertical / horizontal flip and H<->W transpose.  TensorFlow's random streams cannot
// be reproduced, so the draws (offset, flips, transpose) are explicit arguments; with
// phase_train = False the reference returns the centre slice, i.e. the identity (offset = padding, no flips).
// One gather pass, HBM-bound: every output element is read from one input element or is a padding zero.
#include "common.cuh"

namespace {

// src / dst viewed as [N][H][W][C] (images NHWC: N = B; mask stacks [B,T,H,W]: N = B*T, C = 1)
__global__ void __launch_bounds__(256) random_transformation_kernel(const float *__restrict__ src, size_t N, int H, int W,
                                                                    int C, int padding, int off_y, int off_x, int vflip,
                                                                    int hflip, int transpose, float *__restrict__ dst) {
  const size_t total = N * (size_t)H * W * C;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % C);
    size_t r = idx / C;
    const int j = (int)(r % W);
    r /= W;
    const int i = (int)(r % H);
    const size_t n = r / H;
    // tf.transpose after the flips (image_ops.py:96-100): output (i, j) <- flipped (j, i); needs H == W
    int u = transpose ? j : i, v = transpose ? i : j;
    // tf.reverse (image_ops.py:90-94): flipped (u, v) <- cropped (H-1-u, W-1-v)
    if (vflip) u = H - 1 - u;
    if (hflip) v = W - 1 - v;
    // tf.slice of the padded image (image_ops.py:52-56): cropped (u, v) <- source (u + off_y - padding, ...)
    const int sy = u + off_y - padding, sx = v + off_x - padding;
    float val = 0.f;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) val = __ldg(src + ((n * H + sy) * W + sx) * C + c);
    dst[idx] = val;
  }
}

// dst[i] = (float)src[i]: the {0,1} ground-truth masks arrive as bytes (the datasets hold PNG masks,
// data_api/ins_seg_dataset.py:169-172) and are expanded on the device - a quarter of the host->device traffic.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t *__restrict__ src, size_t n, size_t n16,
                                                        float *__restrict__ dst) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float4 *d = reinterpret_cast<float4 *>(dst) + i * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      d[k] = make_float4((float)(w[k] & 0xffu), (float)((w[k] >> 8) & 0xffu), (float)((w[k] >> 16) & 0xffu),
                         (float)(w[k] >> 24));
  }
  if (blockIdx.x == 0)
    for (size_t i = n16 * 16 + threadIdx.x; i < n; i += blockDim.x) dst[i] = (float)src[i];
}

}  // namespace

extern "C" int ra_random_transformation_f32(const float *src, size_t N, int H, int W, int C, int padding, int off_y,
                                            int off_x, int vflip, int hflip, int transpose, float *dst, void *stream) {
  if (H < 1 || W < 1 || C < 1 || padding < 0 || off_y < 0 || off_x < 0 || off_y > 2 * padding || off_x > 2 * padding)
    return RA_ERR_INVALID_ARG;
  if (transpose && H != W) return RA_ERR_UNSUPPORTED;  // the reference's static shapes only allow it for square inputs
  if (N == 0) return RA_OK;
  if (!src || !dst || src == dst) return RA_ERR_INVALID_ARG;
  const size_t total = N * (size_t)H * W * C;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ra::kNumSMs * 16) blocks = (size_t)ra::kNumSMs * 16;
  random_transformation_kernel<<<(unsigned)blocks, 256, 0, ra::as_stream(stream)>>>(src, N, H, W, C, padding, off_y, off_x,
                                                                                   vflip, hflip, transpose, dst);
  return ra::finish_launch("random_transformation_kernel");
}

extern "C" int ra_u8_to_f32(const uint8_t *src, size_t n, float *dst, void *stream) {
  if (n == 0) return RA_OK;
  if (!src || !dst) return RA_ERR_INVALID_ARG;
  const bool vec = !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15);
  const size_t n16 = vec ? n / 16 : 0;
  size_t blocks = (n16 + 255) / 256;
  if (blocks > (size_t)ra::kNumSMs * 8) blocks = (size_t)ra::kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  u8_to_f32_kernel<<<(unsigned)blocks, 256, 0, ra::as_stream(stream)>>>(src, n, n16, dst);
  return ra::finish_launch("u8_to_f32_kernel");
}

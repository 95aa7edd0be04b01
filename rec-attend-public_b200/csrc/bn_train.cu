// Training-mode batch normalisation of a conv block — nnlib.batch_norm with phase_train = True (nnlib.py:65-128):
//   batch_mean, batch_var = tf.nn.moments(x, [0, 1, 2])            (biased variance, two passes like TF)
//   ema_x -= (1 - decay) * (ema_x - batch_x)                        (ExponentialMovingAverage, decay 0.9, :103-108)
//   y = pool(relu(x * inv + (beta - mean * inv))),  inv = gamma * rsqrt(var + 1e-3)   (tf.nn.batch_normalization, :119)
// x [B,H,W,C] is the raw convolution output (bias included), NHWC (float4 path when C % 4 == 0).  Three launches: per-CTA channel sums,
// per-CTA centred squared sums (every CTA first reduces the partial sums to the mean, in double), then the fused
// normalise + ReLU + max-pool pass whose CTA 0 also publishes the statistics and updates the EMA shadows.
// HBM-bound: x is read three times (the second and third pass hit L2 for the patch-network layers).
#include "common.cuh"

namespace {

constexpr int kBnThreads = 256;
constexpr int kBnMaxC = 256;

// Sum the per-CTA partials of every channel with the whole CTA, in double, in a fixed order; out_s[c] = sum / npix.
// Every CTA of the next pass runs this before it can start streaming, so it is built for latency:
//  * fast path (C % 4 == 0 and C / 4 divides 32): the [ctas][C] array is read as float4, four independent loads per
//    thread and round - a thread always meets the same channel quad -, then lanes holding the same quad are combined
//    with shuffles and the 8 warps through shared memory: one L2 round trip per 4096 partial quads instead of a
//    dependent chain of ctas / groups loads;
//  * general path: thread t takes channel t % C and every (kBnThreads / C)-th partial.  Needs C <= kBnThreads.
__device__ __forceinline__ double shfl_xor_double(double v, int mask) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, mask);
  hi = __shfl_xor_sync(0xffffffffu, hi, mask);
  return __hiloint2double(hi, lo);
}

// COHERENT: the partials were written by other CTAs of the SAME launch (fused kernel below): read them through L2
// (ld.global.cg), not through the non-coherent path.
template <bool COHERENT = false>
__device__ __forceinline__ void reduce_partials(const float *__restrict__ partial, int ctas, int C, double inv_npix,
                                                float *out_s, double *scratch /* [kBnThreads * 4] */) {
  const int tid = threadIdx.x;
  const int q_n = C >> 2;
  if ((C & 3) == 0 && q_n <= 32 && (32 % q_n) == 0 && (reinterpret_cast<uintptr_t>(partial) & 15) == 0) {
    const int lane = tid & 31, warp = tid >> 5;
    const float4 *p4 = reinterpret_cast<const float4 *>(partial);
    const int n4 = ctas * q_n;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < n4; i += kBnThreads * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i + u * kBnThreads;
        v[u] = idx < n4 ? (COHERENT ? __ldcg(p4 + idx) : __ldg(p4 + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[0] += (double)v[u].x;
        a[1] += (double)v[u].y;
        a[2] += (double)v[u].z;
        a[3] += (double)v[u].w;
      }
    }
    // lanes l, l + q_n, l + 2 q_n, .. hold the same channel quad (256 and 32 are multiples of q_n)
    for (int m = 16; m >= q_n; m >>= 1) {
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] += shfl_xor_double(a[e], m);
    }
    __syncthreads();  // scratch may still be read by a previous call
    if (lane < q_n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) scratch[warp * C + lane * 4 + e] = a[e];  // 8 * C <= 8 * 128 doubles?  see below
    }
    __syncthreads();
    if (tid < C) {
      double t = 0.0;
      for (int w = 0; w < kBnThreads / 32; ++w) t += scratch[w * C + tid];
      out_s[tid] = (float)(t * inv_npix);
    }
    __syncthreads();
    return;
  }
  const int groups = kBnThreads / C;
  const int c = tid % C, j = tid / C;
  // four independent accumulators: the loads of a thread's partials overlap instead of forming one dependent chain
  // (every CTA of the next pass runs this before it can start streaming)
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (j < groups) {
    int k = j;
    for (; k + 3 * groups < ctas; k += 4 * groups) {
      const float p0 = __ldcg(partial + (size_t)k * C + c), p1 = __ldcg(partial + (size_t)(k + groups) * C + c);
      const float p2 = __ldcg(partial + (size_t)(k + 2 * groups) * C + c);
      const float p3 = __ldcg(partial + (size_t)(k + 3 * groups) * C + c);
      s0 += (double)p0;
      s1 += (double)p1;
      s2 += (double)p2;
      s3 += (double)p3;
    }
    for (; k < ctas; k += groups) s0 += (double)__ldcg(partial + (size_t)k * C + c);
  }
  const double s = (s0 + s1) + (s2 + s3);
  scratch[tid] = s;
  __syncthreads();
  if (tid < C) {
    double t = 0.0;
    for (int g = 0; g < groups; ++g) t += scratch[g * C + tid];
    out_s[tid] = (float)(t * inv_npix);
  }
  __syncthreads();
}

// partial[cta][C] <- sum over this CTA's pixels of (x - center[c])^pow, pow in {1, 2}
// V = 4: float4 path (C % 4 == 0); V = 1: any channel count (the 1-channel mask layer of the deconv head)
template <int POW, int V>
__global__ void __launch_bounds__(kBnThreads) bn_partial_kernel(const float *__restrict__ x, size_t npix, int C,
                                                                const float *__restrict__ prev_partial, int prev_ctas,
                                                                float *__restrict__ partial) {
  __shared__ float center_s[kBnMaxC];
  __shared__ float red_s[kBnThreads * 4];
  __shared__ double dbl_s[kBnThreads * 4];
  const int tid = threadIdx.x;
  const int cg_n = C / V;
  if (POW == 2) {  // the mean, from the first pass's partial sums
    reduce_partials(prev_partial, prev_ctas, C, 1.0 / (double)npix, center_s, dbl_s);
  } else {
    for (int c = tid; c < C; c += kBnThreads) center_s[c] = 0.f;
    __syncthreads();
  }
  // thread -> channel group cg = tid % cg_n, pixel lane pl = tid / cg_n; threads beyond lanes * cg_n idle (C = 96)
  const int cg = tid % cg_n, lanes = kBnThreads / cg_n, pl = tid / cg_n;
  float ctr[V], acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    ctr[e] = center_s[cg * V + e];
    acc[e] = 0.f;
  }
  // eight independent loads in flight per thread (the big first-layer tensors are streamed from HBM)
  constexpr int U = 8;
  const size_t stride = (size_t)gridDim.x * lanes;
  for (size_t p0 = (size_t)blockIdx.x * lanes + pl; pl < lanes && p0 < npix; p0 += U * stride) {
    float v[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t p = p0 + u * stride;
#pragma unroll
      for (int e = 0; e < V; ++e) v[u][e] = (POW == 1) ? 0.f : ctr[e];  // neutral element for pixels past the end
      if (p < npix) {
        if (V == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4 *>(x + p * C + cg * 4));
          v[u][0] = q.x;
          v[u][V > 1 ? 1 : 0] = q.y;
          v[u][V > 2 ? 2 : 0] = q.z;
          v[u][V > 3 ? 3 : 0] = q.w;
        } else {
          v[u][0] = __ldg(x + p * C + cg);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < V; ++e) {
        if (POW == 1) {
          acc[e] += v[u][e];
        } else {
          const float d = v[u][e] - ctr[e];
          acc[e] = fmaf(d, d, acc[e]);
        }
      }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) red_s[tid * V + e] = acc[e];
  __syncthreads();
  for (int c = tid; c < C; c += kBnThreads) {
    const int g = c / V, e = c % V;
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red_s[(l * cg_n + g) * V + e];
    partial[(size_t)blockIdx.x * C + c] = s;
  }
}

template <int POOL, int V>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float *__restrict__ x, int B, int H, int W, int C,
                                                              const float *__restrict__ sum_partial,
                                                              const float *__restrict__ sq_partial, int ctas,
                                                              const float *__restrict__ gamma,
                                                              const float *__restrict__ beta, float eps, float decay,
                                                              int relu, float *__restrict__ ema_mean,
                                                              float *__restrict__ ema_var, float *__restrict__ batch_mean,
                                                              float *__restrict__ batch_var, float *__restrict__ y) {
  __shared__ float inv_s[kBnMaxC], sh_s[kBnMaxC], mean_s[kBnMaxC], var_s[kBnMaxC];
  __shared__ double dbl_s[kBnThreads * 4];
  const int tid = threadIdx.x;
  const size_t npix = (size_t)B * H * W;
  reduce_partials(sum_partial, ctas, C, 1.0 / (double)npix, mean_s, dbl_s);
  reduce_partials(sq_partial, ctas, C, 1.0 / (double)npix, var_s, dbl_s);
  for (int c = tid; c < C; c += kBnThreads) {
    const float mean = mean_s[c], var = var_s[c];
    const float inv = gamma[c] * rsqrtf(var + eps);
    inv_s[c] = inv;
    sh_s[c] = beta[c] - mean * inv;
    if (blockIdx.x == 0) {
      if (batch_mean != nullptr) batch_mean[c] = mean;
      if (batch_var != nullptr) batch_var[c] = var;
      // tf.train.ExponentialMovingAverage.apply: shadow -= (1 - decay) * (shadow - value)
      if (ema_mean != nullptr) ema_mean[c] = ema_mean[c] - (1.0f - decay) * (ema_mean[c] - mean);
      if (ema_var != nullptr) ema_var[c] = ema_var[c] - (1.0f - decay) * (ema_var[c] - var);
    }
  }
  __syncthreads();
  const int cg_n = C / V;
  const int Ho = H / POOL, Wo = W / POOL;
  const unsigned total = (unsigned)((size_t)B * Ho * Wo * cg_n);  // < 2^32, checked by the launcher: 32-bit index math
  for (unsigned idx = blockIdx.x * kBnThreads + tid; idx < total; idx += gridDim.x * kBnThreads) {
    const int cg = (int)(idx % (unsigned)cg_n);
    unsigned pix = idx / (unsigned)cg_n;
    const int ox = (int)(pix % (unsigned)Wo);
    pix /= (unsigned)Wo;
    const int oy = (int)(pix % (unsigned)Ho);
    const int b = (int)(pix / (unsigned)Ho);
    float best[V];
#pragma unroll
    for (int e = 0; e < V; ++e) best[e] = -INFINITY;
#pragma unroll
    for (int py = 0; py < POOL; ++py)
#pragma unroll
      for (int px = 0; px < POOL; ++px) {
        const float *src = x + (((size_t)b * H + oy * POOL + py) * W + ox * POOL + px) * C + cg * V;
        float v[V];
        if (V == 4) {
          const float4 q = __ldg(reinterpret_cast<const float4 *>(src));
          v[0] = q.x;
          v[V > 1 ? 1 : 0] = q.y;
          v[V > 2 ? 2 : 0] = q.z;
          v[V > 3 ? 3 : 0] = q.w;
        } else {
          v[0] = __ldg(src);
        }
#pragma unroll
        for (int e = 0; e < V; ++e) best[e] = fmaxf(best[e], fmaf(v[e], inv_s[cg * V + e], sh_s[cg * V + e]));
      }
    float *dst = y + (((size_t)b * Ho + oy) * Wo + ox) * C + cg * V;
#pragma unroll
    for (int e = 0; e < V; ++e) dst[e] = relu ? fmaxf(best[e], 0.f) : best[e];  // max(relu(.)) == relu(max(.))
  }
}

// ---- small maps (the patch network, the last controller layers): ONE launch.  A CTA owns whole output rows; their
// input rows are one contiguous chunk that is copied to shared memory once (cp.async) and serves all three passes -
// channel sums, centred squared sums, normalise + ReLU + pool - with a grid barrier after each of the two reductions
// (all CTAs are co-resident: grid <= number of SMs, one CTA per SM).  One read of x instead of three, one launch instead
// of three.  Needs C % 4 == 0, C / 4 dividing 256, 16-byte aligned x, the slice in <= kBnFusedSmem bytes.
constexpr size_t kBnFusedSmem = 180 * 1024;

__device__ __forceinline__ void bn_grid_barrier(unsigned int *counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    unsigned long long polls = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      // all CTAs are co-resident by construction; a bounded wait turns a violated assumption (e.g. two such kernels
      // sharing the SMs from concurrent streams) into an error instead of a hung device
      if (++polls > (1ull << 31)) __trap();
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

template <int POOL>
__global__ void __launch_bounds__(kBnThreads, 1) bn_fused_kernel(const float *__restrict__ x, int B, int H, int W, int C,
                                                                 const float *__restrict__ gamma,
                                                                 const float *__restrict__ beta, float eps, float decay,
                                                                 int relu, float *__restrict__ sum_p,
                                                                 float *__restrict__ sq_p, unsigned int *counters,
                                                                 float *__restrict__ ema_mean, float *__restrict__ ema_var,
                                                                 float *__restrict__ batch_mean,
                                                                 float *__restrict__ batch_var, float *__restrict__ y) {
  extern __shared__ __align__(16) float xs_s[];  // my input rows [rows_in][W][C]
  __shared__ float mean_s[kBnMaxC], var_s[kBnMaxC], inv_s[kBnMaxC], sh_s[kBnMaxC];
  __shared__ float red_s[kBnThreads * 4];
  __shared__ double dbl_s[kBnThreads * 4];
  const int tid = threadIdx.x;
  const int Ho = H / POOL, Wo = W / POOL;
  const int rows_out_total = B * Ho;
  const int rows_per = (rows_out_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int r0 = min((int)blockIdx.x * rows_per, rows_out_total), r1 = min(r0 + rows_per, rows_out_total);
  const size_t row_floats = (size_t)W * C;
  const int rows_in = (r1 - r0) * POOL;
  const float *src = x + (size_t)r0 * POOL * row_floats;
  const int n4 = (int)((size_t)rows_in * row_floats / 4);
  {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(xs_s);
    for (int i = tid; i < n4; i += kBnThreads)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (uint32_t)i * 16u), "l"(src + (size_t)i * 4)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const int cg_n = C >> 2, lanes = kBnThreads / cg_n;
  const int cg = tid % cg_n, pl = tid / cg_n;
  const int npix_mine = rows_in * W;
  const double inv_npix = 1.0 / (double)((size_t)B * H * W);
  const float4 *x4 = reinterpret_cast<const float4 *>(xs_s);
  // ---- pass 1 / pass 2: per-thread sums over its pixels (fixed stride), CTA sum in a fixed order, per-CTA partial
  for (int pass = 0; pass < 2; ++pass) {
    float4 ctr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pass == 1) ctr = make_float4(mean_s[cg * 4], mean_s[cg * 4 + 1], mean_s[cg * 4 + 2], mean_s[cg * 4 + 3]);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = pl; p < npix_mine; p += lanes) {
      const float4 v = x4[p * cg_n + cg];
      if (pass == 0) {
        a.x += v.x;
        a.y += v.y;
        a.z += v.z;
        a.w += v.w;
      } else {
        const float dx = v.x - ctr.x, dy = v.y - ctr.y, dz = v.z - ctr.z, dw = v.w - ctr.w;
        a.x = fmaf(dx, dx, a.x);
        a.y = fmaf(dy, dy, a.y);
        a.z = fmaf(dz, dz, a.z);
        a.w = fmaf(dw, dw, a.w);
      }
    }
    *reinterpret_cast<float4 *>(red_s + tid * 4) = a;
    __syncthreads();
    float *part = (pass == 0 ? sum_p : sq_p) + (size_t)blockIdx.x * C;
    for (int c = tid; c < C; c += kBnThreads) {
      const int g = c >> 2, e = c & 3;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red_s[(l * cg_n + g) * 4 + e];
      part[c] = s;
    }
    bn_grid_barrier(counters + pass, gridDim.x);
    reduce_partials<true>(pass == 0 ? sum_p : sq_p, (int)gridDim.x, C, inv_npix, pass == 0 ? mean_s : var_s, dbl_s);
  }
  for (int c = tid; c < C; c += kBnThreads) {
    const float mean = mean_s[c], var = var_s[c];
    const float inv = gamma[c] * rsqrtf(var + eps);
    inv_s[c] = inv;
    sh_s[c] = beta[c] - mean * inv;
    if (blockIdx.x == 0) {
      if (batch_mean != nullptr) batch_mean[c] = mean;
      if (batch_var != nullptr) batch_var[c] = var;
      if (ema_mean != nullptr) ema_mean[c] = ema_mean[c] - (1.0f - decay) * (ema_mean[c] - mean);
      if (ema_var != nullptr) ema_var[c] = ema_var[c] - (1.0f - decay) * (ema_var[c] - var);
    }
  }
  __syncthreads();
  // ---- normalise + ReLU + pool out of shared memory
  const int n_out = (r1 - r0) * Wo * cg_n;
  float4 *y4 = reinterpret_cast<float4 *>(y + (size_t)r0 * Wo * C);
  for (int idx = tid; idx < n_out; idx += kBnThreads) {
    const int g = idx % cg_n;
    const int opix = idx / cg_n;
    const int ox = opix % Wo, orow = opix / Wo;
    const float4 inv = *reinterpret_cast<const float4 *>(inv_s + g * 4), sh = *reinterpret_cast<const float4 *>(sh_s + g * 4);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int py = 0; py < POOL; ++py)
#pragma unroll
      for (int px = 0; px < POOL; ++px) {
        const float4 v = x4[((orow * POOL + py) * W + ox * POOL + px) * cg_n + g];
        best.x = fmaxf(best.x, fmaf(v.x, inv.x, sh.x));
        best.y = fmaxf(best.y, fmaf(v.y, inv.y, sh.y));
        best.z = fmaxf(best.z, fmaf(v.z, inv.z, sh.z));
        best.w = fmaxf(best.w, fmaf(v.w, inv.w, sh.w));
      }
    if (relu) best = make_float4(fmaxf(best.x, 0.f), fmaxf(best.y, 0.f), fmaxf(best.z, 0.f), fmaxf(best.w, 0.f));
    y4[idx] = best;
  }
}

// grid of the fused kernel, or 0 when the shape does not qualify
int bn_fused_ctas(const float *x, int B, int H, int W, int C, int pool) {
  if ((C & 3) != 0 || (kBnThreads % (C >> 2)) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 0;
  // MEASURED (KITTI B=32 training step, in-graph): 11-12 us per block against ~14 us for the three launches - the two
  // grid barriers (atomics + polling over 148 CTAs + the partial reduction) cost what the saved passes gain: 80.4 ->
  // 80.1 ms per step.  Not worth a spin-waiting kernel on the default path: opt-in (RA_BN_FUSED=1).
  if (getenv("RA_BN_FUSED") == nullptr) return 0;
  const int rows_out = B * (H / pool);
  const int ctas = rows_out < ra::kNumSMs ? rows_out : ra::kNumSMs;
  if (ctas < 1) return 0;
  const int rows_per = (rows_out + ctas - 1) / ctas;
  const size_t bytes = (size_t)rows_per * pool * W * C * sizeof(float);
  return bytes <= kBnFusedSmem ? ctas : 0;
}

int bn_ctas(size_t npix, int C) {
  const int lanes = kBnThreads / ((C & 3) ? C : C / 4);
  // at least 16 pixels per thread (two rounds of 8 loads in flight): the small patch-network maps get a few dozen
  // CTAs, whose partials the next pass re-reduces in no time
  size_t want = (npix + (size_t)lanes * 16 - 1) / ((size_t)lanes * 16);
  // every CTA of the next pass re-reduces these partials (ctas x C floats from L2), so their count is bounded by the
  // channel count: 4 CTAs per SM for the wide 16-channel maps (which need the loads in flight to stream from HBM), one
  // per SM from 64 channels on
  size_t per_sm = (size_t)64 / (size_t)(C < 16 ? 16 : C);
  if (per_sm < 1) per_sm = 1;
  const size_t cap = (size_t)ra::kNumSMs * per_sm;
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" size_t ra_bn_train_workspace(int B, int H, int W, int C) {
  if (B < 1 || H < 1 || W < 1 || C < 1 || C > kBnMaxC) return 0;
  // floats: per-CTA sums, then per-CTA centred squared sums (three-pass kernels: bn_ctas CTAs; fused kernel: <= one per
  // SM), then the two barrier counters of the fused kernel
  size_t ctas = (size_t)bn_ctas((size_t)B * H * W, C);
  if (ctas < (size_t)ra::kNumSMs) ctas = (size_t)ra::kNumSMs;
  return (size_t)2 * ctas * C + 4;
}

extern "C" int ra_bn_train_block_f32(const float *x, int B, int H, int W, int C, const float *gamma, const float *beta,
                                     float eps, float decay, int pool, int relu, float *workspace, float *ema_mean,
                                     float *ema_var, float *batch_mean, float *batch_var, float *y, void *stream) {
  if (B < 0 || H < 1 || W < 1 || C < 1) return RA_ERR_INVALID_ARG;
  if (C > kBnMaxC || (pool != 1 && pool != 2)) return RA_ERR_UNSUPPORTED;
  if (pool == 2 && ((H | W) & 1)) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  if (!x || !gamma || !beta || !workspace || !y) return RA_ERR_INVALID_ARG;
  cudaStream_t s = ra::as_stream(stream);
  const size_t npix = (size_t)B * H * W;
  const int fctas = bn_fused_ctas(x, B, H, W, C, pool);
  if (fctas > 0) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(bn_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnFusedSmem);
      cudaFuncSetAttribute(bn_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBnFusedSmem);
      attr_done = true;
    }
    size_t wctas = (size_t)bn_ctas(npix, C);
    if (wctas < (size_t)ra::kNumSMs) wctas = (size_t)ra::kNumSMs;
    float *f_sum = workspace, *f_sq = workspace + (size_t)fctas * C;
    unsigned int *counters = reinterpret_cast<unsigned int *>(workspace + (size_t)2 * wctas * C);
    cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), s);
    const int rows_out = B * (H / pool);
    const int rows_per = (rows_out + fctas - 1) / fctas;
    const size_t smem = (size_t)rows_per * pool * W * C * sizeof(float);
    if (pool == 2)
      bn_fused_kernel<2><<<fctas, kBnThreads, smem, s>>>(x, B, H, W, C, gamma, beta, eps, decay, relu, f_sum, f_sq, counters,
                                                        ema_mean, ema_var, batch_mean, batch_var, y);
    else
      bn_fused_kernel<1><<<fctas, kBnThreads, smem, s>>>(x, B, H, W, C, gamma, beta, eps, decay, relu, f_sum, f_sq, counters,
                                                        ema_mean, ema_var, batch_mean, batch_var, y);
    return ra::finish_launch("bn_fused_kernel");
  }
  const int ctas = bn_ctas(npix, C);
  float *sum_p = workspace, *sq_p = workspace + (size_t)ctas * C;
  const bool vec = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec)
    bn_partial_kernel<1, 4><<<ctas, kBnThreads, 0, s>>>(x, npix, C, nullptr, 0, sum_p);
  else
    bn_partial_kernel<1, 1><<<ctas, kBnThreads, 0, s>>>(x, npix, C, nullptr, 0, sum_p);
  int rc = ra::finish_launch("bn_partial_kernel<1>");
  if (rc != RA_OK) return rc;
  if (vec)
    bn_partial_kernel<2, 4><<<ctas, kBnThreads, 0, s>>>(x, npix, C, sum_p, ctas, sq_p);
  else
    bn_partial_kernel<2, 1><<<ctas, kBnThreads, 0, s>>>(x, npix, C, sum_p, ctas, sq_p);
  rc = ra::finish_launch("bn_partial_kernel<2>");
  if (rc != RA_OK) return rc;
  const size_t total = (size_t)B * (H / pool) * (W / pool) * (vec ? C / 4 : C);
  if (total >= 0xffffffffULL) return RA_ERR_UNSUPPORTED;
  size_t blocks = (total + kBnThreads - 1) / kBnThreads;
  if (blocks > (size_t)ra::kNumSMs * 8) blocks = (size_t)ra::kNumSMs * 8;
#define RA_BN_APPLY(P, V)                                                                                            \
  bn_apply_kernel<P, V><<<(unsigned)blocks, kBnThreads, 0, s>>>(x, B, H, W, C, sum_p, sq_p, ctas, gamma, beta, eps, decay, \
                                                                relu, ema_mean, ema_var, batch_mean, batch_var, y)
  if (pool == 2) {
    if (vec) RA_BN_APPLY(2, 4); else RA_BN_APPLY(2, 1);
  } else {
    if (vec) RA_BN_APPLY(1, 4); else RA_BN_APPLY(1, 1);
  }
#undef RA_BN_APPLY
  return ra::finish_launch("bn_apply_kernel");
}

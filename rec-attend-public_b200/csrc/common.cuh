// Shared helpers for the sm_100a kernels of librecattend_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "rec_attend_b200.h"

namespace ra {

// thread-local last-error text, surfaced through ra_last_error()
void set_last_error(const char *what, cudaError_t e);
int finish_launch(const char *what);  // cudaGetLastError -> RA_OK / RA_ERR_CUDA

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions through a caller-provided shared scratch of >= 32 floats.
// All threads of the block must call; the result is returned to every thread.
__device__ __forceinline__ float block_sum(float v, float *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}
__device__ __forceinline__ float block_min(float v, float *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_min(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : INFINITY;
  r = warp_min(r);
  return r;
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// The decode loop is a serial chain of ~30 launches per step.  Kernels launched through launch_pdl carry
// cudaLaunchAttributeProgrammaticStreamSerialization: once every CTA of the PREVIOUS kernel of the stream has called
// pdl_trigger (or exited), the next grid is scheduled and waits in pdl_wait until the previous grid has completed and
// its memory is visible.  Every PDL kernel starts with pdl_wait(); pdl_trigger(); so the only thing that overlaps is
// the launch / CTA scheduling latency - never a data access.  (Both are no-ops for a kernel launched normally.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();  // RA_PDL=0 switches the attribute off (capi.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- support bands of the Gaussian attention filters (csrc/gauss.cu build_filters_kernel) ----------------------
// Filter entries whose exponent is below -kBandCut are stored as exact zeros, so tap t of an axis is non-zero only
// inside [mu_t - R, mu_t + R], R = sqrt(2 var kBandCut) + 1.  Consumers recompute (supersets of) the bands from the
// box record instead of reading them: no memory traffic, no dependence on a band buffer of another step.
constexpr float kBandCut = 30.0f;

// Superset of the union of the F taps' bands along one axis (taps are monotone in t); two extra pixels absorb the
// fast-math error.  Empty: lo > hi.  NaN boxes keep the full range.
__device__ __forceinline__ void band_union(const float *__restrict__ bo, int axis, int F, int L, int *lo_out,
                                           int *hi_out) {
  const float ctr = bo[RA_BOX_CTR_Y + axis];
  const float size = bo[RA_BOX_SIZE_Y + axis];
  const float var = __expf(bo[RA_BOX_LGVAR_Y + axis]);
  const float step = (size + 1.0f) / (float)F;
  const float half = (float)(F - 1) / 2.0f;
  const float R = __fsqrt_rn(2.0f * var * kBandCut) * 1.0001f + 1.0f;
  const float m0 = ctr - step * half, m1 = ctr + step * half;
  float lo = floorf(fminf(m0, m1) - R) - 2.0f, hi = ceilf(fmaxf(m0, m1) + R) + 2.0f;
  int ilo = 0, ihi = L - 1;
  if (lo == lo && hi == hi) {
    lo = fminf(fmaxf(lo, 0.f), (float)L);
    hi = fmaxf(fminf(hi, (float)(L - 1)), -1.f);
    ilo = (int)lo;
    ihi = (int)hi;
  }
  *lo_out = ilo;
  *hi_out = ihi;
}

// Superset (by two pixels) of the band of ONE tap.
__device__ __forceinline__ void tap_band(const float *__restrict__ bo, int axis, int t, int F, int L, int *lo_out,
                                         int *hi_out) {
  const float ctr = bo[RA_BOX_CTR_Y + axis];
  const float size = bo[RA_BOX_SIZE_Y + axis];
  const float var = __expf(bo[RA_BOX_LGVAR_Y + axis]);
  const float step = (size + 1.0f) / (float)F;
  const float mu = ctr + step * ((float)t - (float)(F - 1) / 2.0f);
  const float R = __fsqrt_rn(2.0f * var * kBandCut) * 1.0001f + 1.0f;
  float lo = floorf(mu - R) - 2.0f, hi = ceilf(mu + R) + 2.0f;
  int ilo = 0, ihi = L - 1;
  if (lo == lo && hi == hi) {
    lo = fminf(fmaxf(lo, 0.f), (float)L);
    hi = fmaxf(fminf(hi, (float)(L - 1)), -1.f);
    ilo = (int)lo;
    ihi = (int)hi;
  }
  *lo_out = ilo;
  *hi_out = ihi;
}

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  return __ldcs(reinterpret_cast<const float4 *>(p));
}

}  // namespace ra

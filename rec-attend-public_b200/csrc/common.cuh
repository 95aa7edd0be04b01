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

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  return __ldcs(reinterpret_cast<const float4 *>(p));
}

}  // namespace ra

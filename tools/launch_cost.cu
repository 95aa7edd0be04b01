// Where does the fixed per-launch cost of a big-shared-memory persistent kernel come from?  Builds CUDA graphs of
// N dependent launches of trivial kernels that differ in ONE property each and prints us per launch.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/launch_cost tools/launch_cost.cu && tools/launch_cost
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

struct Big { unsigned char b[256]; };  // stands in for two CUtensorMap kernel parameters

__global__ void k_plain(float *out) {
  extern __shared__ float sm[];
  if (threadIdx.x == 0) sm[0] = 1.f;
  __syncthreads();
  if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0) out[0] = sm[0];
}

__global__ void k_bigparam(float *out, const __grid_constant__ Big p) {
  extern __shared__ float sm[];
  if (threadIdx.x == 0) sm[0] = (float)p.b[0];
  __syncthreads();
  if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0) out[0] = sm[0];
}

__global__ void k_tmem(float *out) {
  extern __shared__ float sm[];
  __shared__ uint32_t base;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&base)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  if (threadIdx.x == 0) sm[0] = 1.f;
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u));
  if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0) out[0] = sm[0];
}

struct Launch {
  int kind;  // 0 plain, 1 bigparam, 2 tmem
  int grid, threads;
  size_t smem;
};

static float run(const char *name, const std::vector<Launch> &seq, int reps, float *out) {
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  Big bp{};
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int r = 0; r < reps; ++r)
    for (const Launch &l : seq) {
      if (l.kind == 0) k_plain<<<l.grid, l.threads, l.smem, s>>>(out);
      if (l.kind == 1) k_bigparam<<<l.grid, l.threads, l.smem, s>>>(out, bp);
      if (l.kind == 2) k_tmem<<<l.grid, l.threads, l.smem, s>>>(out);
    }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const float us = ms * 1e3f / (5.f * reps * seq.size());
  printf("%-72s %7.2f us / launch   (%s)\n", name, us, cudaGetErrorString(cudaGetLastError()));
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaStreamDestroy(s);
  return us;
}

int main() {
  float *out;
  cudaMalloc(&out, 4);
  const size_t big = 190 * 1024, mid = 150 * 1024;
  cudaFuncSetAttribute(k_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  cudaFuncSetAttribute(k_bigparam, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
  const int R = 100;
  run("plain, 148 x 256 thr, 1 KB smem", {{0, 148, 256, 1024}}, R, out);
  run("plain, 148 x 544 thr, 1 KB smem", {{0, 148, 544, 1024}}, R, out);
  run("plain, 148 x 544 thr, 190 KB smem", {{0, 148, 544, big}}, R, out);
  run("plain, 148 x 544 thr, alternating 190 / 150 KB smem", {{0, 148, 544, big}, {0, 148, 544, mid}}, R, out);
  run("plain, 148 x 544 thr, alternating 190 KB / 1 KB smem", {{0, 148, 544, big}, {0, 148, 256, 1024}}, R, out);
  run("plain, 148 x 544 thr, alternating 190 / 100 KB smem", {{0, 148, 544, big}, {0, 148, 544, 100 * 1024}}, R, out);
  run("256-byte grid_constant param, 148 x 544 thr, 190 KB smem", {{1, 148, 544, big}}, R, out);
  run("TMEM alloc 512 + dealloc, 148 x 544 thr, 190 KB smem", {{2, 148, 544, big}}, R, out);
  run("TMEM alloc 512 + dealloc, 148 x 544 thr, 1 KB smem", {{2, 148, 544, 1024}}, R, out);
  run("TMEM alloc, 64 CTAs x 544 thr, 190 KB smem", {{2, 64, 544, big}}, R, out);
  cudaFuncSetAttribute(k_plain, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  run("carveout=max: plain, alternating 190 / 150 KB smem", {{0, 148, 544, big}, {0, 148, 544, mid}}, R, out);
  run("carveout=max: plain, alternating 190 KB / 1 KB smem (same kernel)", {{0, 148, 544, big}, {0, 148, 256, 1024}}, R, out);
  return 0;
}

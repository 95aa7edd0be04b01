// Probe for the tcgen05 building blocks used by csrc/conv_umma.cu: K-major SWIZZLE_NONE shared
// memory descriptors over a "channel-plane" layout, kind::tf32 MMA into TMEM, commit ->
// mbarrier, tcgen05.ld epilogue, and the 3xTF32 split (hi*hi + hi*lo + lo*hi).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int tmem_cols(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

template <int M_ROWS, int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const float *A, const float *B, float *D, int split3) {
  // planes: [K/4][rows][4] for hi and lo parts
  extern __shared__ __align__(128) unsigned char smem[];
  float *a_hi = reinterpret_cast<float *>(smem);
  float *a_lo = a_hi + M_ROWS * K;
  float *b_hi = a_lo + M_ROWS * K;
  float *b_lo = b_hi + N * K;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < M_ROWS * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const float v = A[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const int o = ((k >> 2) * M_ROWS + r) * 4 + (k & 3);
    a_hi[o] = hi;
    a_lo[o] = v - hi;
  }
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int r = i / K, k = i % K;
    const float v = B[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const int o = ((k >> 2) * N + r) * 4 + (k & 3);
    b_hi[o] = hi;
    b_lo[o] = v - hi;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(tmem_cols(N)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M_ROWS >> 4) << 24);
    const uint32_t a_plane = M_ROWS * 16, b_plane = N * 16;
    uint32_t acc = 0;
    for (int pass = 0; pass < (split3 ? 3 : 1); ++pass) {
      const float *ap = (pass == 2) ? a_lo : a_hi;
      const float *bp = (pass == 1) ? b_lo : b_hi;
      for (int k8 = 0; k8 < K / 8; ++k8) {
        const uint64_t ad = make_desc(smem_u32(ap) + (2 * k8) * a_plane, a_plane, 128);
        const uint64_t bd = make_desc(smem_u32(bp) + (2 * k8) * b_plane, b_plane, 128);
        umma_tf32(tmem_base, ad, bd, idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                 : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done)
          : "r"(smem_u32(&mbar)), "r"(0u)
          : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: thread t <-> accumulator row t (TMEM lane), N columns
  if (tid < M_ROWS) {
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr + c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(tmem_cols(N)));
  }
}

template <int M_ROWS, int N, int K>
int run(int split3) {
  std::vector<float> A(M_ROWS * K), B(N * K), D(M_ROWS * N, -1.f);
  srand(1);
  for (auto &v : A) v = (float)rand() / RAND_MAX * 2 - 1;
  for (auto &v : B) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)(2 * M_ROWS * K + 2 * N * K) * 4;
  cudaFuncSetAttribute(probe_kernel<M_ROWS, N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<M_ROWS, N, K><<<1, 128, smem>>>(dA, dB, dD, split3);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("M=%d N=%d K=%d split3=%d: CUDA error %s\n", M_ROWS, N, K, split3, cudaGetErrorString(e));
    return 1;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double max_err = 0, max_err_tf32 = 0, max_ref = 0;
  for (int m = 0; m < M_ROWS; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0, ref_t = 0;
      for (int k = 0; k < K; ++k) {
        ref += (double)A[m * K + k] * B[n * K + k];
        uint32_t ua, ub;
        memcpy(&ua, &A[m * K + k], 4);
        memcpy(&ub, &B[n * K + k], 4);
        ua &= 0xFFFFE000u;
        ub &= 0xFFFFE000u;
        float fa, fb;
        memcpy(&fa, &ua, 4);
        memcpy(&fb, &ub, 4);
        ref_t += (double)fa * fb;
      }
      max_err = fmax(max_err, fabs(D[m * N + n] - ref));
      max_err_tf32 = fmax(max_err_tf32, fabs(D[m * N + n] - ref_t));
      max_ref = fmax(max_ref, fabs(ref));
    }
  printf("M=%d N=%d K=%d split3=%d: max|D-exact|=%.3e  max|D-tf32ref|=%.3e  max|ref|=%.3f  D[0]=%f D[last]=%f\n", M_ROWS,
         N, K, split3, max_err, max_err_tf32, max_ref, D[0], D[M_ROWS * N - 1]);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return 0;
}

int main() {
  int rc = 0;
  rc |= run<128, 32, 32>(0);
  rc |= run<128, 32, 32>(1);
  rc |= run<128, 16, 64>(1);
  rc |= run<128, 96, 16>(1);
  rc |= run<128, 64, 128>(1);
  return rc;
}

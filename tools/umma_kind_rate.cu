// Microbenchmark: cycles per tcgen05.mma for kind::tf32 (K = 8 per instruction) and kind::f16 (K = 16 per instruction),
// M = 128, SWIZZLE_NONE K-major operands with 32-byte rows (the conv kernel's layout), as a function of N.
// Question: does an fp16 hi / lo operand split halve the instruction count of the issue-bound conv layers
// at the same cost per instruction?   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_kind_rate umma_kind_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
template <int F16>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (F16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <int F16>
__global__ void __launch_bounds__(128) rate_kernel(int N, int n_mma, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = F16 ? 0x3C003C00u : 0x3F800000u;  // ones
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_s;
  if (warp == 0) {
    // c_format F32 (1 << 4); a / b format: TF32 = 2, F16 = 0
    const uint32_t fmt = F16 ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // A: 2 planes of 1024 slots x 16 B (plane stride 16 KB); B: 2 planes of N rows x 16 B at 32 KB
    const uint64_t da = make_desc(smem_u32(smem), 16384, 128);
    const uint64_t da1 = da + 1;  // one slot further: the conv kernel's tap shifts
    const uint64_t db = make_desc(smem_u32(smem) + 32768, (uint32_t)N * 16, 128);
    const uint32_t d0 = tm, d1 = tm + N;
    long long t0 = clock64();
    if (tid == 0) {
      for (int i = 0; i < n_mma / 8; ++i) {
        umma<F16>(d0, da, db, idesc, 1u);
        umma<F16>(d1, da1, db, idesc, 1u);
        umma<F16>(d0, da, db, idesc, 1u);
        umma<F16>(d1, da1, db, idesc, 1u);
        umma<F16>(d0, da, db, idesc, 1u);
        umma<F16>(d1, da1, db, idesc, 1u);
        umma<F16>(d0, da, db, idesc, 1u);
        umma<F16>(d1, da1, db, idesc, 1u);
      }
    }
    __syncwarp();
    if (tid == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (tid == 0) out[0] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}
int main() {
  long long *d, h;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 4000;
  printf("kind N : cycles per instruction (M=128; K=8 tf32 / K=16 f16), cycles per 16 K-elements\n");
  const int Ns[] = {16, 32, 48, 64, 96, 128, 256};
  for (int k = 0; k < 2; ++k)
    for (int ni = 0; ni < 7; ++ni) {
      const int N = Ns[ni];
      if (k == 0) rate_kernel<0><<<1, 128, 64 * 1024>>>(N, n_mma, d); else rate_kernel<1><<<1, 128, 64 * 1024>>>(N, n_mma, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double c = (double)h / n_mma;
      printf("%s %3d : %7.1f   %7.1f\n", k ? "f16 " : "tf32", N, c, k ? c : 2 * c);
    }
  return 0;
}

// Microbenchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8, SWIZZLE_NONE K-major operands)
// as a function of N, the number of independent accumulators and the A row-offset pattern.
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__global__ void __launch_bounds__(128) rate_kernel(int N, int n_acc, int n_mma, int shift_rows, int swz, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 1.0f;
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // A: 2 planes of 1024 slots x 16 B (plane stride 16 KB); B: 2 planes of N rows at 32 KB
    uint64_t da, db;
    if (swz == 0) {
      da = make_desc(smem_u32(smem), 16384, 128);
      db = make_desc(smem_u32(smem) + 32768, (uint32_t)N * 16, 128);
    } else {  // SWIZZLE_32B K-major: rows of 32 B, 8-row groups 256 B apart (layout_type 6)
      da = make_desc(smem_u32(smem), 16, 256) | ((uint64_t)6 << 61);
      db = make_desc(smem_u32(smem) + 32768, 16, 256) | ((uint64_t)6 << 61);
    }
    long long t0 = clock64();
    // tight issue loop: loop-invariant operands, two alternating accumulators, one lane issues
    const uint32_t d0 = tm, d1 = tm + (n_acc > 1 ? N : 0);
    const uint64_t da1 = da + (uint64_t)shift_rows;
    if (tid == 0) {
      for (int i = 0; i < n_mma / 8; ++i) {
        umma(d0, da, db, idesc, 1u);
        umma(d1, da1, db, idesc, 1u);
        umma(d0, da, db, idesc, 1u);
        umma(d1, da1, db, idesc, 1u);
        umma(d0, da, db, idesc, 1u);
        umma(d1, da1, db, idesc, 1u);
        umma(d0, da, db, idesc, 1u);
        umma(d1, da1, db, idesc, 1u);
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (tid == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}
int main_rate() {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n_mma = 2000;
  printf("N n_acc shift swz : issue_cyc/mma  total_cyc/mma\n");
  int Ns[] = {16, 32, 64, 128, 256};
  for (int swz = 0; swz < 2; ++swz)
    for (int ni = 0; ni < 5; ++ni)
      for (int n_acc = 1; n_acc <= 8; n_acc *= 2)
        for (int shift = 0; shift <= 1; ++shift) {
          int N = Ns[ni];
          if (n_acc * N > 512) continue;
          rate_kernel<<<1, 128, 64 * 1024>>>(N, n_acc, n_mma, shift, swz, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%3d %d %d %d : %7.1f %7.1f\n", N, n_acc, shift, swz, (double)h[0] / n_mma, (double)h[1] / n_mma);
        }
  return 0;
}

// Pattern of the conv kernel: n_acc independent accumulators (m-tiles), per step either one instruction shape
// (mode 0: N) or the merged pair (mode 1: for acc: MMA(2N, A_hi); for acc: MMA(N, A_lo); mode 2: per acc both
// back to back).  The whole warp runs the loop (uniform operands), as in conv_umma.cu.
__global__ void __launch_bounds__(128) pattern_kernel(int N, int n_acc, int n_steps, int mode, int shift_rows, long long *out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 1.0f;
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
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_s, 0);
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc_n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_2n = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // A: 2 planes of 2048 slots x 16 B (plane stride 32 KB) = 64 KB; B: 2 planes of 2N rows at 128 KB
    const uint64_t da = make_desc(smem_u32(smem), 32768, 128);
    const uint64_t db = make_desc(smem_u32(smem) + 131072, (uint32_t)(2 * N) * 16, 128);
    const int cols = mode == 0 ? N : 2 * N;
    long long t0 = clock64();
    for (int i = 0; i < n_steps; ++i) {
      const uint64_t a_i = da + (uint64_t)((i % 9) * shift_rows);  // tap-like start offsets
      if (mode == 0) {
        for (int a = 0; a < n_acc; ++a)
          if (leader) umma(tm + a * cols, a_i + (uint64_t)(a * 128), db, idesc_n, 1u);
      } else if (mode == 1) {
        for (int a = 0; a < n_acc; ++a)
          if (leader) umma(tm + a * cols, a_i + (uint64_t)(a * 128), db, idesc_2n, 1u);
        for (int a = 0; a < n_acc; ++a)
          if (leader) umma(tm + a * cols, a_i + (uint64_t)(a * 128 + 1024), db, idesc_n, 1u);
      } else {
        for (int a = 0; a < n_acc; ++a) {
          if (leader) umma(tm + a * cols, a_i + (uint64_t)(a * 128), db, idesc_2n, 1u);
          if (leader) umma(tm + a * cols, a_i + (uint64_t)(a * 128 + 1024), db, idesc_n, 1u);
        }
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (tid == 0) {
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}
int main(int argc, char **argv) {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("conv pattern: N n_acc mode shift grid : issue_cyc/mma  total_cyc/mma\n");
  int Ns[] = {16, 32, 64, 128};
  for (int grid = 1; grid <= 148; grid += 147)
    for (int ni = 0; ni < 4; ++ni)
      for (int mode = 0; mode < 3; ++mode)
        for (int n_acc = 1; n_acc <= 8; ++n_acc)
          for (int shift = 0; shift <= 1; ++shift) {
            const int N = Ns[ni];
            const int cols = mode == 0 ? N : 2 * N;
            if (n_acc * cols > 512 || (mode > 0 && 2 * N > 256)) continue;
            if (grid > 1 && (shift == 0 || (n_acc != 1 && n_acc != 3 && n_acc != 7))) continue;
            const int n_steps = 300;
            const int n_mma = n_steps * n_acc * (mode == 0 ? 1 : 2);
            pattern_kernel<<<grid, 128, 200 * 1024>>>(N, n_acc, n_steps, mode, shift, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%3d %d %d %d %3d : %7.1f %7.1f\n", N, n_acc, mode, shift, grid, (double)h[0] / n_mma, (double)h[1] / n_mma);
          }
  if (argc > 1) return 0;
  return main_rate();
}

// Pairwise soft + hard IoU on the 5th-generation tensor cores — modellib.f_iou(pairwise=True) (modellib.py:104-155)
// for the matching (full_model.py:983) and f_iou / f_dice of the thresholded masks (:1063-1081), ONE pass over y_out:
//
//   I_soft[b,n,m] = sum_p a[b,n,p] g[b,m,p]          I_hard[b,n,m] = sum_p [a[b,n,p] > thr] g[b,m,p]
//   Sa = sum_p a,  Sa_hard = sum_p [a > thr],  Sg = sum_p g       (the per-pixel eps adds H*W*1e-5 to every union)
//
// is a K-major GEMM with K = H*W pixels: A rows = the (example, output) masks, B rows = the (example, ground truth)
// masks, both exactly as they lie in memory ([B,T,H*W], pixel contiguous).  A CTA owns a GROUP of G consecutive examples
// (G*T + 1 <= 128 rows) and a slice of the pixels; per 32-pixel step TMA lands a [G*T x 32] box of a and of g in
// shared memory (128-byte rows, SWIZZLE_128B), converter warps split a in place into hi (nearest tf32) / lo = a - hi
// and write the thresholded copy, and one thread issues per 8 pixels
//     D_soft += A_hi B^T;   D_soft += A_lo B^T;   D_hard += A_thr B^T          (tcgen05.mma kind::tf32, M = 128)
// (g and the thresholded masks are {0,1}: exact in tf32; a needs the split on ONE operand only).  Row G*T of both tiles
// is a row of ones, so column G*T of D holds the row sums of a and row G*T of D the sums of g.  Products between
// different examples of a group are computed and dropped (the tensor core has time to spare: the kernel is bound by
// the single read of a and g from HBM).  Accumulators live in TMEM (2 x NT columns); each CTA writes its partial block
// diagonal, a finalize kernel adds the pixel slices in a fixed order and forms IoU / DICE.
#include <cuda.h>
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kStepPix = 32;  // pixels per pipeline step = one 128-byte swizzle row
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * 128;  // 16 KB
constexpr int kMaxNT = 144;                  // B tile rows (G*T + 1 rounded up to 16)
constexpr int kBBytes = kMaxNT * 128;        // 18 KB
constexpr int kStageBytes = 3 * kTileBytes + kBBytes;  // A_hi (landing zone), A_lo, A_thr, B
constexpr int kStages = 3;
constexpr int kConvThreads = 256;                 // warps 2..9
constexpr int kThreads = 64 + kConvThreads;       // warp 0: TMA, warp 1: MMA
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024;  // + alignment slack

struct IouUmmaParams {
  int B, T, G, NT, HW;
  int n_groups, splits, steps_total;
  float thr;
  float *partial;  // [n_groups][splits][2][128][T+1]  then  bsum [n_groups][splits][128]
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(mbar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
// K-major, SWIZZLE_128B: 8-row groups of 1024 bytes (SBO), LBO unused (1), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
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
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ float tf32_hi(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(kThreads, 1)
    iou_umma_kernel(const __grid_constant__ IouUmmaParams p, const __grid_constant__ CUtensorMap tm_a,
                    const __grid_constant__ CUtensorMap tm_b) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_raw[kStages], bar_full[kStages], bar_empty[kStages], bar_done;
  unsigned char *smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);  // SWIZZLE_128B: 1024-byte tiles
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int gidx = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int rows = p.G * p.T;  // data rows of both tiles; row `rows` is the row of ones
  // pixel steps of this CTA
  const int per = (p.steps_total + p.splits - 1) / p.splits;
  const int s0 = split * per, s1 = min(p.steps_total, s0 + per);
  const int n_steps = max(0, s1 - s0);
  if (n_steps == 0) {  // more pixel slices than steps: this slice is empty - its partial block is all zeros
    const int T1 = p.T + 1;
    float *out = p.partial + (((size_t)gidx * p.splits + split) * 2) * kTileRows * T1;
    float *bsum = p.partial + (size_t)p.n_groups * p.splits * 2 * kTileRows * T1 +
                  ((size_t)gidx * p.splits + split) * kTileRows;
    for (int i = tid; i < 2 * kTileRows * T1; i += kThreads) out[i] = 0.f;
    for (int i = tid; i < kTileRows; i += kThreads) bsum[i] = 0.f;
    return;
  }

  if (warp == 0 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bar_raw[s]), 1);
      mbar_init(smem_u32(&bar_full[s]), kConvThreads);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows the TMA never writes: zero them once in every stage, then the rows of ones (row `rows` of A_hi / A_thr / B;
  // the converters keep A_lo's copy at zero)
  for (int s = 0; s < kStages; ++s) {
    unsigned char *st = smem + (size_t)s * kStageBytes;
    float *a_hi = reinterpret_cast<float *>(st), *b_t = reinterpret_cast<float *>(st + 3 * kTileBytes);
    for (int i = tid; i < (kTileRows - rows) * 32; i += kThreads) a_hi[rows * 32 + i] = (i < 32) ? 1.0f : 0.0f;
    for (int i = tid; i < (p.NT - rows) * 32; i += kThreads) b_t[rows * 32 + i] = (i < 32) ? 1.0f : 0.0f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    const bool leader = elect_one();
    const uint32_t tx = (uint32_t)rows * 128u * 2u;
    for (int k = 0; k < n_steps; ++k) {
      const int s = k % kStages;
      mbar_wait(smem_u32(&bar_empty[s]), (uint32_t)(((k / kStages) & 1) ^ 1));
      if (leader) {
        const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
        const uint32_t bar = smem_u32(&bar_raw[s]);
        mbar_arrive_expect_tx(bar, tx);
        tma_load_2d(st, &tm_a, (s0 + k) * kStepPix, gidx * rows, bar);
        tma_load_2d(st + 3u * kTileBytes, &tm_b, (s0 + k) * kStepPix, gidx * rows, bar);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < n_steps; ++k) {
      const int s = k % kStages;
      mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((k / kStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint32_t st = smem_u32(smem + (size_t)s * kStageBytes);
        const uint64_t d_hi = make_desc_sw128(st), d_lo = make_desc_sw128(st + kTileBytes),
                       d_thr = make_desc_sw128(st + 2u * kTileBytes), d_b = make_desc_sw128(st + 3u * kTileBytes);
#pragma unroll
        for (int k8 = 0; k8 < kStepPix / 8; ++k8) {
          const uint64_t off = (uint64_t)(k8 * 2);  // 32 bytes per k8 step, in 16-byte units
          const uint32_t acc = (k > 0 || k8 > 0) ? 1u : 0u;
          umma_tf32(tmem_base, d_hi + off, d_b + off, idesc, acc);
          umma_tf32(tmem_base, d_lo + off, d_b + off, idesc, 1u);
          umma_tf32(tmem_base + 256u, d_thr + off, d_b + off, idesc, acc);
        }
        umma_commit(smem_u32(&bar_empty[s]));
        if (k == n_steps - 1) umma_commit(smem_u32(&bar_done));
      }
      __syncwarp();
    }
  } else {
    // =============================== converters ===============================
    const int ctid = tid - 64;
    for (int k = 0; k < n_steps; ++k) {
      const int s = k % kStages;
      mbar_wait(smem_u32(&bar_raw[s]), (uint32_t)((k / kStages) & 1));
      unsigned char *st = smem + (size_t)s * kStageBytes;
      float4 *hi4 = reinterpret_cast<float4 *>(st);
      float4 *lo4 = reinterpret_cast<float4 *>(st + kTileBytes);
      float4 *th4 = reinterpret_cast<float4 *>(st + 2 * kTileBytes);
      // every 16-byte unit of the tile (the swizzle only permutes units inside a row: the same offset in all copies)
      for (int u = ctid; u < (rows + 1) * 8; u += kConvThreads) {
        const float4 v = hi4[u];
        const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        hi4[u] = h;
        lo4[u] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        th4[u] = make_float4(v.x > p.thr ? 1.f : 0.f, v.y > p.thr ? 1.f : 0.f, v.z > p.thr ? 1.f : 0.f,
                             v.w > p.thr ? 1.f : 0.f);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(smem_u32(&bar_full[s]));
    }
    // =============================== epilogue (warps 2..5: one TMEM lane quadrant each) ===============================
    if (warp < 6) {
      mbar_wait(smem_u32(&bar_done), 0u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3;           // lane quadrant this warp may read
      const int r = q * 32 + lane;      // tile row = TMEM lane
      const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
      const int T1 = p.T + 1;
      float *out = p.partial + (((size_t)gidx * p.splits + split) * 2) * kTileRows * T1;
      float *bsum = p.partial + (size_t)p.n_groups * p.splits * 2 * kTileRows * T1 +
                    ((size_t)gidx * p.splits + split) * kTileRows;
      for (int accu = 0; accu < 2; ++accu) {
        const uint32_t base = tmem_base + lane_sel + (uint32_t)(accu * 256);
        for (int c0 = 0; c0 < p.NT; c0 += 8) {
          float v[8];
          tmem_ld8(base + (uint32_t)c0, v);  // (warp-collective)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            if (r < rows) {
              const int e = r / p.T;
              if (c >= e * p.T && c < (e + 1) * p.T) out[((size_t)accu * kTileRows + r) * T1 + (c - e * p.T)] = v[j];
              if (c == rows) out[((size_t)accu * kTileRows + r) * T1 + p.T] = v[j];  // row sum of a / of [a > thr]
            } else if (r == rows && accu == 0 && c < rows) {
              bsum[c] = v[j];  // sum of g over this CTA's pixels
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// iou / dice from the partials: fixed summation order over the pixel slices
__global__ void __launch_bounds__(256) iou_umma_finalize_kernel(const float *__restrict__ partial, int B, int T, int G,
                                                                int n_groups, int splits, float hw_eps,
                                                                float *__restrict__ iou_soft,
                                                                float *__restrict__ iou_hard,
                                                                float *__restrict__ dice_hard) {
  const int b = blockIdx.x;
  const int gidx = b / G, e = b - gidx * G;
  const int T1 = T + 1;
  const float *bs = partial + (size_t)n_groups * splits * 2 * kTileRows * T1;
  for (int idx = threadIdx.x; idx < T * T; idx += blockDim.x) {
    const int n = idx / T, m = idx - n * T;
    const int r = e * T + n;
    float i_s = 0.f, i_h = 0.f, sa = 0.f, sah = 0.f, sg = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float *o = partial + (((size_t)gidx * splits + s) * 2) * kTileRows * T1;
      i_s += o[(size_t)r * T1 + m];
      sa += o[(size_t)r * T1 + T];
      i_h += o[((size_t)kTileRows + r) * T1 + m];
      sah += o[((size_t)kTileRows + r) * T1 + T];
      sg += bs[((size_t)gidx * splits + s) * kTileRows + e * T + m];
    }
    const size_t dst = ((size_t)b * T + n) * T + m;
    if (iou_soft) iou_soft[dst] = i_s / (sa + sg - i_s + hw_eps);
    if (iou_hard) iou_hard[dst] = i_h / (sah + sg - i_h + hw_eps);
    if (dice_hard) dice_hard[dst] = 2.f * i_h / ((sah + hw_eps) + (sg + hw_eps));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

// [rows_total, HW] fp32, box = [32 pixels x box_rows], 128-byte swizzle
bool mask_map(const float *x, size_t rows_total, int HW, int box_rows, CUtensorMap *out) {
  EncodeTiledFn enc = encoder();
  if (enc == nullptr) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)rows_total};
  const cuuint64_t strides[1] = {(cuuint64_t)HW * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kStepPix, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct IouPlan {
  int G, NT, n_groups, splits, steps_total;
};

bool iou_plan(int B, int T, int HW, IouPlan *pl) {
  if (T < 1 || T > 127 || (HW % kStepPix) != 0 || B < 1) return false;
  pl->G = 127 / T;
  if (pl->G > B) pl->G = B;
  pl->NT = (pl->G * T + 1 + 15) / 16 * 16;
  if (pl->NT > kMaxNT) return false;
  pl->n_groups = (B + pl->G - 1) / pl->G;
  pl->steps_total = HW / kStepPix;
  int splits = ra::kNumSMs / pl->n_groups;
  if (splits < 1) splits = 1;
  if (splits > pl->steps_total) splits = pl->steps_total;
  pl->splits = splits;
  return true;
}

}  // namespace

extern "C" size_t ra_pairwise_iou_umma_workspace(int B, int T, int H, int W) {
  IouPlan pl;
  if (!iou_plan(B, T, H * W, &pl)) return 0;
  return ((size_t)pl.n_groups * pl.splits * 2 * kTileRows * (T + 1) + (size_t)pl.n_groups * pl.splits * kTileRows) *
         sizeof(float);
}

extern "C" int ra_pairwise_iou_umma_f32(const float *a, const float *g, int B, int T, int H, int W, float hard_thr,
                                        void *ws, float *iou_soft, float *iou_hard, float *dice_hard, void *stream) {
  if (!a || !g || !ws || B < 0 || T < 1 || H < 1 || W < 1 || !(hard_thr > 0.f && hard_thr < 1.f)) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  IouPlan pl;
  if (!iou_plan(B, T, H * W, &pl)) return RA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(g)) & 15) return RA_ERR_UNSUPPORTED;
  CUtensorMap tm_a, tm_b;
  if (!mask_map(a, (size_t)B * T, H * W, pl.G * T, &tm_a) || !mask_map(g, (size_t)B * T, H * W, pl.G * T, &tm_b))
    return RA_ERR_UNSUPPORTED;
  IouUmmaParams p;
  p.B = B; p.T = T; p.G = pl.G; p.NT = pl.NT; p.HW = H * W;
  p.n_groups = pl.n_groups; p.splits = pl.splits; p.steps_total = pl.steps_total;
  p.thr = hard_thr;
  p.partial = reinterpret_cast<float *>(ws);
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e =
        cudaFuncSetAttribute(iou_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(iou_umma_kernel)", e);
      return RA_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = ra::as_stream(stream);
  iou_umma_kernel<<<pl.n_groups * pl.splits, kThreads, kSmemBytes, s>>>(p, tm_a, tm_b);
  int rc = ra::finish_launch("iou_umma_kernel");
  if (rc != RA_OK) return rc;
  iou_umma_finalize_kernel<<<B, 256, 0, s>>>(p.partial, B, T, pl.G, pl.n_groups, pl.splits, (float)(H * W) * 1e-5f,
                                             iou_soft, iou_hard, dice_hard);
  return ra::finish_launch("iou_umma_finalize_kernel");
}

// Batched max-weight bipartite matching (Kuhn-Munkres with min vertex cover), one warp per
// example.  Replaces the reference's CPU-only TF custom op (hungarian.cc:26-30,540) and its
// caller f_segm_match (modellib.py:382-415).
//
// The algorithm is the reference's (hungarian.cc:335-488): covers c_x = row max, c_y = 0;
// every round rebuild the equality graph |c_x+c_y-w| <= 1e-6 (fp32 sums, :309-325); when
// `next_match`, rebuild a maximum matching FROM SCRATCH by repeated breadth-first
// augmentation (:179-217); otherwise grow the alternating sets S/T or lower the covers by
// the minimum slack (:406-483).  Tie-breaking is what makes results bit-identical, and it
// lives in the BFS: the reference marks vertices on pop, so its FIFO holds duplicates and a
// vertex's parent is its LAST pusher.  In the layered source->X->Y->sink network that is
// equivalent to "parent = neighbour of highest rank in the previous level; rank inside a
// level = (parent rank, vertex index)", which needs no duplicates (proof sketch and the CPU
// model that is tested against the literal restatement: oracle/hungarian_bitset.c).
//
// Mapping: lanes own columns (y = lane and lane+32) for the O(n^2) float work (row max,
// equality graph via __ballot_sync, minimum slack via warp min); the set logic runs on
// 64-bit masks replicated in every lane; the augmenting search itself is sequential and
// runs on lane 0 over shared-memory arrays.  Everything is latency-bound by design: the
// whole problem (<= 16 KB) lives in shared memory and B warps run on B SMs.
#include "common.cuh"

namespace {

constexpr int kMaxN = RA_HUNG_MAX_N;
constexpr int kMaxRounds = 1000;  // hungarian.cc:20,362
typedef unsigned long long mask_t;

struct WarpScratch {
  float cx[kMaxN];
  float cy[kMaxN];
  mask_t eq[kMaxN];
  int match_y[kMaxN];
  int match_x[kMaxN];
  int level_x[kMaxN];
  int level_y[kMaxN];
  int parent_y[kMaxN];
};

// Mask arithmetic for the two widths of the search: 32-bit masks when nx, ny <= 32 (every shuffle and bit operation is
// one instruction instead of two - the kernel is a single latency-bound warp, so instruction count is time), 64-bit
// otherwise.
template <typename M>
struct MaskOps;
template <>
struct MaskOps<uint32_t> {
  static __device__ __forceinline__ int popc(uint32_t m) { return __popc(m); }
  static __device__ __forceinline__ int lowest(uint32_t m) { return __ffs((int)m) - 1; }
  static __device__ __forceinline__ int highest(uint32_t m) { return 31 - __clz((int)m); }
  static __device__ __forceinline__ uint32_t shfl_down(uint32_t v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
  static __device__ __forceinline__ uint32_t shfl(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
};
template <>
struct MaskOps<mask_t> {
  static __device__ __forceinline__ int popc(mask_t m) { return __popcll(m); }
  static __device__ __forceinline__ int lowest(mask_t m) { return __ffsll((long long)m) - 1; }
  static __device__ __forceinline__ int highest(mask_t m) { return 63 - __clzll((long long)m); }
  static __device__ __forceinline__ mask_t shfl_down(mask_t v, int o) {
    return (mask_t)__shfl_down_sync(0xffffffffu, (unsigned long long)v, o);
  }
  static __device__ __forceinline__ mask_t shfl(mask_t v, int src) {
    return (mask_t)__shfl_sync(0xffffffffu, (unsigned long long)v, src);
  }
};

// Inclusive suffix OR over the lanes (lane l gets OR of lanes l..31).
template <typename M>
__device__ __forceinline__ M suffix_or(M v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const M t = MaskOps<M>::shfl_down(v, o);
    if (lane + o < 32) v |= t;
  }
  return v;
}
// Exclusive prefix sum over the lanes; *total = sum over the warp.
__device__ __forceinline__ int prefix_sum_excl(int v, int lane, int *total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// One ranked breadth-first search + augmentation, run by the whole warp: rank r of a level lives in lane r
// (and lane r - 32 for r >= 32).  "The highest-ranked pusher wins" is an exclusive suffix OR over the ranks,
// "next level in (parent rank, column index) order" an exclusive prefix sum of the claim counts.  free_x /
// free_y (unmatched rows / columns) are replicated in every lane.  Returns true if the matching grew.
// M = uint32_t requires nx, ny <= 32 (then a level never has more than 32 ranks).
template <typename M>
__device__ bool augment_ranked(WarpScratch &s, int lane, M &free_x, M &free_y) {
  typedef MaskOps<M> O;
  constexpr bool kWide = sizeof(M) == 8;
  const M one = 1;
  if (!free_x) return false;
  // level 1: the unmatched rows in index order
  {
    const int x0 = lane;
    if ((free_x >> x0) & one) s.level_x[O::popc(free_x & ((one << x0) - one))] = x0;
    if (kWide) {
      const int x1 = lane + 32;
      if ((free_x >> x1) & one) s.level_x[O::popc(free_x & ((one << x1) - one))] = x1;
    }
  }
  int n_lx = O::popc(free_x);
  M seen_y = 0;
  __syncwarp();

  for (;;) {
    const bool two = kWide && n_lx > 32;  // warp-uniform
    // ---- claims: c[r] = A[r] & ~(OR of A[r'] for r' > r), A[r] = residual neighbours of rank r not seen yet
    M A0 = 0, A1 = 0;
    int px0 = -1, px1 = -1;
    if (lane < n_lx) {
      px0 = s.level_x[lane];
      M adj = (M)s.eq[px0];
      const int my = s.match_y[px0];
      if (my >= 0) adj &= ~(one << my);  // a saturated edge has no residual capacity
      A0 = adj & ~seen_y;
    }
    M excl1 = 0, total1 = 0;
    if (two) {
      if (lane + 32 < n_lx) {
        px1 = s.level_x[lane + 32];
        M adj = (M)s.eq[px1];
        const int my = s.match_y[px1];
        if (my >= 0) adj &= ~(one << my);
        A1 = adj & ~seen_y;
      }
      const M S1 = suffix_or<M>(A1, lane);
      excl1 = O::shfl_down(S1, 1);
      if (lane == 31) excl1 = 0;
      total1 = O::shfl(S1, 0);
    }
    const M S0 = suffix_or<M>(A0, lane);
    M excl0 = O::shfl_down(S0, 1);
    if (lane == 31) excl0 = 0;
    excl0 |= total1;
    const M level_claim = O::shfl(S0, 0) | total1;
    if (!level_claim) return false;
    seen_y |= level_claim;
    const M c0 = A0 & ~excl0, c1 = A1 & ~excl1;

    // ---- next level in (parent rank, column index) order
    int tot0 = 0, tot1 = 0;
    int pos0 = prefix_sum_excl(O::popc(c0), lane, &tot0);
    for (M c = c0; c; c &= c - one) {
      const int y = O::lowest(c);
      s.parent_y[y] = px0;
      s.level_y[pos0++] = y;
    }
    if (two) {
      int pos1 = tot0 + prefix_sum_excl(O::popc(c1), lane, &tot1);
      for (M c = c1; c; c &= c - one) {
        const int y = O::lowest(c);
        s.parent_y[y] = px1;
        s.level_y[pos1++] = y;
      }
    }
    const int n_ly = tot0 + tot1;
    __syncwarp();

    // ---- the sink is reached from the highest-ranked free column: highest rank, then highest index
    int last_free_y = -1;
    {
      const M f0 = c0 & free_y, f1 = c1 & free_y;
      const unsigned b1 = two ? __ballot_sync(0xffffffffu, f1 != 0) : 0u;
      if (b1) {
        const int src = 31 - __clz((int)b1);
        last_free_y = O::highest(O::shfl(f1, src));
      } else {
        const unsigned b0 = __ballot_sync(0xffffffffu, f0 != 0);
        if (b0) {
          const int src = 31 - __clz((int)b0);
          last_free_y = O::highest(O::shfl(f0, src));
        }
      }
    }
    if (last_free_y >= 0) {
      int x_start = 0;
      if (lane == 0) {
        int y = last_free_y;
        for (;;) {
          const int x = s.parent_y[y];
          const int prev = s.match_y[x];
          s.match_y[x] = y;
          s.match_x[y] = x;
          if (prev < 0) {
            x_start = x;
            break;
          }
          y = prev;
        }
      }
      x_start = __shfl_sync(0xffffffffu, x_start, 0);
      free_x &= ~(one << x_start);
      free_y &= ~(one << last_free_y);
      __syncwarp();
      return true;
    }
    // ---- all columns of the level are matched: follow the matched edges back into X
    if (lane < n_ly) s.level_x[lane] = s.match_x[s.level_y[lane]];
    if (kWide && lane + 32 < n_ly) s.level_x[lane + 32] = s.match_x[s.level_y[lane + 32]];
    n_lx = n_ly;
    __syncwarp();
  }
}

// iou/s_gt == nullptr: plain op (W given).  Otherwise f_segm_match: W is built from iou and
// s_gt when it is loaded, and the matching is masked on the way out.
template <typename M>
__global__ void __launch_bounds__(32) hungarian_kernel(const float *__restrict__ W_in, const float *__restrict__ s_gt,
                                                       int nx, int ny, float *__restrict__ M_out,
                                                       float *__restrict__ cx_out, float *__restrict__ cy_out,
                                                       float *__restrict__ w_out, int *__restrict__ status_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpScratch &s = *reinterpret_cast<WarpScratch *>(smem_raw);
  float *w = reinterpret_cast<float *>(smem_raw + sizeof(WarpScratch));

  const int b = blockIdx.x;
  const int lane = threadIdx.x;
  const int y0 = lane, y1 = lane + 32;
  const bool has0 = y0 < ny, has1 = y1 < ny;
  const size_t base = (size_t)b * nx * ny;

  // ---- load (and optionally build) the weight matrix
  for (int k = lane; k < nx * ny; k += 32) {
    float v = W_in[base + k];
    if (s_gt != nullptr) {
      // modellib.py:395-406: (iou*mask_x)*mask_y, floor(v*1e6+0.5)/1e6 (TF-0.12 tf.round), +1e-5.
      // Explicit _rn intrinsics: no FMA contraction, same roundings as the fp32 reference ops.
      const int x = k / ny, y = k - x * ny;
      v = __fmul_rn(__fmul_rn(v, s_gt[(size_t)b * ny + y]), s_gt[(size_t)b * nx + x]);
      v = floorf(__fadd_rn(__fmul_rn(v, 1e6f), 0.5f));
      v = __fadd_rn(__fdiv_rn(v, 1e6f), 1e-5f);
      if (w_out != nullptr) w_out[base + k] = v;
    }
    w[k] = v;
  }
  __syncwarp();

  // ---- covers: c_x = row max, c_y = 0 (hungarian.cc:339-348)
  for (int x = 0; x < nx; ++x) {
    float v = -INFINITY;
    if (has0) v = w[x * ny + y0];
    if (has1) v = fmaxf(v, w[x * ny + y1]);
    v = ra::warp_max(v);
    if (lane == 0) s.cx[x] = v;
  }
  s.cy[y0] = 0.f;
  s.cy[y1] = 0.f;
  __syncwarp();

  mask_t S = 0, T = 0;
  bool next_match = true;
  bool eq_stale = true;  // the equality graph is a function of (covers, w): rebuild it only after a cover update
  int status = 0;
  const mask_t all_x = nx == 64 ? ~0ull : ((1ull << nx) - 1ull);
  const mask_t all_y = ny == 64 ? ~0ull : ((1ull << ny) - 1ull);

  for (int round = 0;; ++round) {
    if (round == kMaxRounds) {  // hungarian.cc:362-377: give back the unfinished matching
      status |= RA_HUNG_ST_OUTER_CAP;
      break;
    }
    // ---- equality graph (hungarian.cc:309-325).  `<= 1e-6` on the double literal is the same
    // set of floats as `<= 1e-6f` (the float below 1e-6; tests/test_hungarian_oracle.py).
    if (eq_stale) {
      const float cy0 = s.cy[y0], cy1 = s.cy[y1];
      for (int x = 0; x < nx; ++x) {
        const float c = s.cx[x];
        bool p0 = false, p1 = false;
        if (has0) {
          const float d = __fsub_rn(__fadd_rn(c, cy0), w[x * ny + y0]);
          p0 = (fabsf(d) <= 1e-6f) && (c > 0.f || cy0 > 0.f);
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, p0);
        unsigned m1 = 0u;
        if (ny > 32) {  // warp-uniform
          if (has1) {
            const float d = __fsub_rn(__fadd_rn(c, cy1), w[x * ny + y1]);
            p1 = (fabsf(d) <= 1e-6f) && (c > 0.f || cy1 > 0.f);
          }
          m1 = __ballot_sync(0xffffffffu, p1);
        }
        if (lane == 0) s.eq[x] = (mask_t)m0 | ((mask_t)m1 << 32);
      }
      eq_stale = false;
      __syncwarp();
    }

    if (next_match) {
      s.match_y[y0] = -1;
      s.match_y[y1] = -1;
      s.match_x[y0] = -1;
      s.match_x[y1] = -1;
      __syncwarp();
      M free_x = (M)all_x, free_y = (M)all_y;
      while (augment_ranked<M>(s, lane, free_x, free_y)) {
      }
      // hungarian.cc:219-248: the smaller side must be fully matched
      if (nx - MaskOps<M>::popc(free_x) == (nx >= ny ? ny : nx)) break;
      S = 1ull << MaskOps<M>::lowest(free_x);  // first unmatched row, hungarian.cc:394-403
      T = 0;
    }

    mask_t NS = 0;  // hungarian.cc:250-263
    for (mask_t r = S; r; r &= r - 1) NS |= s.eq[__ffsll((long long)r) - 1];

    if (NS == T) {
      // minimum slack over S x (Y \ T), hungarian.cc:418-426
      float a = 3.402823466e+38f;
      const float cy0 = s.cy[y0], cy1 = s.cy[y1];
      const bool t0 = has0 && !((T >> y0) & 1ull), t1 = has1 && !((T >> y1) & 1ull);
      for (mask_t r = S; r; r &= r - 1) {
        const int x = __ffsll((long long)r) - 1;
        const float c = s.cx[x];
        if (t0) a = fminf(a, __fsub_rn(__fadd_rn(c, cy0), w[x * ny + y0]));
        if (t1) a = fminf(a, __fsub_rn(__fadd_rn(c, cy1), w[x * ny + y1]));
      }
      a = ra::warp_min(a);
      if (a <= 1e-6f) {  // `a < 1e-6` on the double literal, hungarian.cc:428
        next_match = true;
        continue;
      }
      __syncwarp();
      if (y0 < nx && ((S >> y0) & 1ull)) s.cx[y0] = __fsub_rn(s.cx[y0], a);
      if (y1 < nx && ((S >> y1) & 1ull)) s.cx[y1] = __fsub_rn(s.cx[y1], a);
      if (has0 && ((T >> y0) & 1ull)) s.cy[y0] = __fadd_rn(s.cy[y0], a);
      if (has1 && ((T >> y1) & 1ull)) s.cy[y1] = __fadd_rn(s.cy[y1], a);
      eq_stale = true;
      __syncwarp();
    } else {
      // hungarian.cc:446-482: pull matched columns of N(S) into T, their rows into S
      while (__popcll(NS) > __popcll(T)) {
        const int y = __ffsll((long long)(NS & ~T)) - 1;
        const int z = s.match_x[y];
        if (z < 0) {
          next_match = true;
          break;
        }
        next_match = false;
        S |= 1ull << z;
        NS |= s.eq[z];
        T |= 1ull << y;
      }
    }
  }
  __syncwarp();

  // ---- outputs
  for (int x = 0; x < nx; ++x) {
    const int my = s.match_y[x];
    float sx = 1.f;
    if (s_gt != nullptr) sx = s_gt[(size_t)b * nx + x];
    if (has0) {
      float m = (my == y0) ? 1.f : 0.f;
      if (s_gt != nullptr) m = __fmul_rn(__fmul_rn(m, s_gt[(size_t)b * ny + y0]), sx);
      M_out[base + x * ny + y0] = m;
    }
    if (has1) {
      float m = (my == y1) ? 1.f : 0.f;
      if (s_gt != nullptr) m = __fmul_rn(__fmul_rn(m, s_gt[(size_t)b * ny + y1]), sx);
      M_out[base + x * ny + y1] = m;
    }
  }
  if (cx_out != nullptr) {
    if (y0 < nx) cx_out[(size_t)b * nx + y0] = s.cx[y0];
    if (y1 < nx) cx_out[(size_t)b * nx + y1] = s.cx[y1];
  }
  if (cy_out != nullptr) {
    if (has0) cy_out[(size_t)b * ny + y0] = s.cy[y0];
    if (has1) cy_out[(size_t)b * ny + y1] = s.cy[y1];
  }
  if (status_out != nullptr && lane == 0) status_out[b] = status;
}

int launch(const float *W, const float *s_gt, int B, int nx, int ny, float *M, float *cx, float *cy, float *w_out,
           int32_t *status, cudaStream_t stream) {
  if (B < 0 || nx < 1 || ny < 1) return RA_ERR_INVALID_ARG;
  if (nx > kMaxN || ny > kMaxN) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;  // empty batch: nothing to read or write
  if (W == nullptr || M == nullptr) return RA_ERR_INVALID_ARG;
  const size_t smem = sizeof(WarpScratch) + (size_t)nx * ny * sizeof(float);
  if (nx <= 32 && ny <= 32)
    hungarian_kernel<uint32_t><<<B, 32, smem, stream>>>(W, s_gt, nx, ny, M, cx, cy, w_out, status);
  else
    hungarian_kernel<mask_t><<<B, 32, smem, stream>>>(W, s_gt, nx, ny, M, cx, cy, w_out, status);
  return ra::finish_launch("hungarian_kernel");
}

}  // namespace

extern "C" int ra_hungarian_f32(const float *W, int B, int nx, int ny, float *M, float *cover_x, float *cover_y,
                                int32_t *status, void *stream) {
  return launch(W, nullptr, B, nx, ny, M, cover_x, cover_y, nullptr, status, ra::as_stream(stream));
}

extern "C" int ra_segm_match_f32(const float *iou, const float *s_gt, int B, int T, float *match, float *weights_out,
                                 int32_t *status, void *stream) {
  if (s_gt == nullptr) return RA_ERR_INVALID_ARG;
  return launch(iou, s_gt, B, T, T, match, nullptr, nullptr, weights_out, status, ra::as_stream(stream));
}

extern "C" int ra_hungarian_f32_host(const float *W, int B, int nx, int ny, float *M, float *cover_x, float *cover_y,
                                     int32_t *status) {
  if (B < 0 || nx < 1 || ny < 1) return RA_ERR_INVALID_ARG;
  if (nx > kMaxN || ny > kMaxN) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  if (W == nullptr || M == nullptr) return RA_ERR_INVALID_ARG;
  const size_t nW = (size_t)B * nx * ny, nX = (size_t)B * nx, nY = (size_t)B * ny;
  float *d = nullptr;
  int32_t *dst = nullptr;
  cudaError_t e = cudaMalloc(&d, (2 * nW + nX + nY) * sizeof(float));
  if (e != cudaSuccess) {
    ra::set_last_error("cudaMalloc", e);
    return RA_ERR_CUDA;
  }
  e = cudaMalloc(&dst, (size_t)B * sizeof(int32_t));
  if (e != cudaSuccess) {
    cudaFree(d);
    ra::set_last_error("cudaMalloc", e);
    return RA_ERR_CUDA;
  }
  float *dW = d, *dM = d + nW, *dX = dM + nW, *dY = dX + nX;
  int rc = RA_OK;
  e = cudaMemcpy(dW, W, nW * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = launch(dW, nullptr, B, nx, ny, dM, dX, dY, nullptr, dst, nullptr);
    if (rc == RA_OK) e = cudaMemcpy(M, dM, nW * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == RA_OK && e == cudaSuccess && cover_x) e = cudaMemcpy(cover_x, dX, nX * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == RA_OK && e == cudaSuccess && cover_y) e = cudaMemcpy(cover_y, dY, nY * sizeof(float), cudaMemcpyDeviceToHost);
    if (rc == RA_OK && e == cudaSuccess && status) e = cudaMemcpy(status, dst, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  cudaFree(dst);
  if (e != cudaSuccess) {
    ra::set_last_error("ra_hungarian_f32_host", e);
    return RA_ERR_CUDA;
  }
  return rc;
}

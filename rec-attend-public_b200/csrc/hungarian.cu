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
  mask_t claimed[kMaxN];
  int match_y[kMaxN];
  int match_x[kMaxN];
  int level_x[kMaxN];
  int level_y[kMaxN];
  int parent_y[kMaxN];
  int n_matched;
  int first_free;
};

// One ranked breadth-first search + augmentation (lane 0 only). Returns true if augmented.
__device__ bool augment_ranked(WarpScratch &s, int nx) {
  int n_lx = 0;
  mask_t seen_y = 0;
  for (int x = 0; x < nx; ++x)
    if (s.match_y[x] < 0) s.level_x[n_lx++] = x;

  while (n_lx > 0) {
    // the highest-ranked pusher wins: claim columns from the back of the level
    mask_t level_claim = 0;
    for (int r = n_lx - 1; r >= 0; --r) {
      const int x = s.level_x[r];
      mask_t adj = s.eq[x];
      const int my = s.match_y[x];
      if (my >= 0) adj &= ~(1ull << my);  // a saturated edge has no residual capacity
      const mask_t c = adj & ~seen_y & ~level_claim;
      s.claimed[r] = c;
      level_claim |= c;
    }
    if (!level_claim) return false;
    seen_y |= level_claim;

    // next level in (parent rank, column index) order; remember the last free column
    int n_ly = 0;
    int last_free_y = -1;
    for (int r = 0; r < n_lx; ++r) {
      mask_t c = s.claimed[r];
      const int px = s.level_x[r];
      while (c) {
        const int y = __ffsll((long long)c) - 1;
        c &= c - 1;
        s.parent_y[y] = px;
        s.level_y[n_ly++] = y;
        if (s.match_x[y] < 0) last_free_y = y;
      }
    }

    if (last_free_y >= 0) {  // the sink is reached from the highest-ranked free column
      int y = last_free_y;
      for (;;) {
        const int x = s.parent_y[y];
        const int prev = s.match_y[x];
        s.match_y[x] = y;
        s.match_x[y] = x;
        if (prev < 0) break;
        y = prev;
      }
      return true;
    }
    // all columns of the level are matched: follow the matched edges back into X
    n_lx = 0;
    for (int r = 0; r < n_ly; ++r) s.level_x[n_lx++] = s.match_x[s.level_y[r]];
  }
  return false;
}

// iou/s_gt == nullptr: plain op (W given).  Otherwise f_segm_match: W is built from iou and
// s_gt when it is loaded, and the matching is masked on the way out.
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
  int status = 0;

  for (int round = 0;; ++round) {
    if (round == kMaxRounds) {  // hungarian.cc:362-377: give back the unfinished matching
      status |= RA_HUNG_ST_OUTER_CAP;
      break;
    }
    // ---- equality graph (hungarian.cc:309-325).  `<= 1e-6` on the double literal is the same
    // set of floats as `<= 1e-6f` (the float below 1e-6; tests/test_hungarian_oracle.py).
    {
      const float cy0 = s.cy[y0], cy1 = s.cy[y1];
      for (int x = 0; x < nx; ++x) {
        const float c = s.cx[x];
        bool p0 = false, p1 = false;
        if (has0) {
          const float d = __fsub_rn(__fadd_rn(c, cy0), w[x * ny + y0]);
          p0 = (fabsf(d) <= 1e-6f) && (c > 0.f || cy0 > 0.f);
        }
        if (has1) {
          const float d = __fsub_rn(__fadd_rn(c, cy1), w[x * ny + y1]);
          p1 = (fabsf(d) <= 1e-6f) && (c > 0.f || cy1 > 0.f);
        }
        const unsigned m0 = __ballot_sync(0xffffffffu, p0);
        const unsigned m1 = __ballot_sync(0xffffffffu, p1);
        if (lane == 0) s.eq[x] = (mask_t)m0 | ((mask_t)m1 << 32);
      }
    }
    __syncwarp();

    if (next_match) {
      if (lane == 0) {
        for (int x = 0; x < nx; ++x) s.match_y[x] = -1;
        for (int y = 0; y < ny; ++y) s.match_x[y] = -1;
        while (augment_ranked(s, nx)) {
        }
        int n_matched = 0, first_free = -1;
        for (int x = 0; x < nx; ++x) {
          if (s.match_y[x] >= 0)
            ++n_matched;
          else if (first_free < 0)
            first_free = x;
        }
        s.n_matched = n_matched;
        s.first_free = first_free;
      }
      __syncwarp();
      // hungarian.cc:219-248: the smaller side must be fully matched
      if (s.n_matched == (nx >= ny ? ny : nx)) break;
      S = 1ull << s.first_free;
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
  hungarian_kernel<<<B, 32, smem, stream>>>(W, s_gt, nx, ny, M, cx, cy, w_out, status);
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

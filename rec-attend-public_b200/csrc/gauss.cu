// DRAW-style separable Gaussian attention: filter construction, glimpse extraction and
// paste-back — modellib.get_gaussian_filter (modellib.py:581-612), modellib.extract_patch
// (:615-641) as used at full_model.py:728-741 (attention box), :788-789 (glimpse) and
// :810-818,845 (mask paste-back + canvas update).
//
// The reference materialises dense [B,L,F] filters and runs 2 batched GEMMs per channel.
// Here the filters are built once per step as tap-major profiles fy [B,F,H], fx [B,F,W]
// together with each tap's support band: entries with exp argument below -kCut are exactly
// zero (their sum over a whole row is below fp32 resolution of the kept terms), so every
// consumer only walks the band — the work is proportional to sigma, the HBM traffic to one
// pass over the image.
//   glimpse:   tmp[b,i,(x,d)] = sum_{y in band_i} fy[i,y] X[b,y,(x,d)]         (row pass)
//              patch[b,i,j,d] = gamma * sum_{x in band_j} tmp[b,i,(x,d)] fx[j,x]  (column pass)
//   paste-back: out[b,y,x] = sum_j (sum_i fy[i,y] P[i,j]) fx[j,x], then the fused epilogue.
#include "common.cuh"

namespace {

constexpr float kCut = 30.0f;  // drop filter entries below exp(-30) ~ 9e-14 of the tap's peak
constexpr int kMaxF = 64;
constexpr int kTapsPerCta = 8;

// ------------------------------------------------------------------ filter construction
__global__ void build_filters_kernel(const float *__restrict__ box, int H, int W, int F, float *__restrict__ fy,
                                     float *__restrict__ fx, int *__restrict__ band) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  const int b = blockIdx.x;
  const float *bo = box + (size_t)b * RA_BOX_STRIDE;
  {
    const int axis = blockIdx.y;  // one CTA per (example, axis, group of kTapsPerCta taps)
    const int t0 = blockIdx.z * kTapsPerCta, t1 = min(F, t0 + kTapsPerCta);
    const int L = axis == 0 ? H : W;
    float *f = axis == 0 ? fy + (size_t)b * F * H : fx + (size_t)b * F * W;
    const float ctr = bo[RA_BOX_CTR_Y + axis];
    const float size = bo[RA_BOX_SIZE_Y + axis];
    const float var = expf(bo[RA_BOX_LGVAR_Y + axis]);
    // modellib.py:610: 1 / sqrt(exp(lg_var)) / sqrt(2*pi)
    const float norm = 1.0f / sqrtf(var) / sqrtf(2.0f * 3.14159265358979323846f);
    const float step = (size + 1.0f) / (float)F;  // modellib.py:599
    for (int idx = t0 * L + threadIdx.x; idx < t1 * L; idx += blockDim.x) {
      const int t = idx / L, l = idx - t * L;
      const float mu = ctr + step * ((float)t - (float)(F - 1) / 2.0f);
      const float d = (float)l - mu;
      const float e = ((-0.5f * d) * d) / var;  // modellib.py:611
      f[idx] = (e >= -kCut) ? norm * expf(e) : 0.f;
    }
    // support band of every tap (a superset of the non-zero entries; empty if hi < lo)
    for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
      const float mu = ctr + step * ((float)t - (float)(F - 1) / 2.0f);
      const float R = sqrtf(2.0f * var * kCut) + 1.0f;
      float lo = ceilf(mu - R), hi = floorf(mu + R);
      int ilo = 0, ihi = L - 1;
      if (lo == lo && hi == hi) {  // not NaN
        lo = fminf(fmaxf(lo, 0.f), (float)L);
        hi = fmaxf(fminf(hi, (float)(L - 1)), -1.f);
        ilo = (int)lo;
        ihi = (int)hi;
      }
      int *bd = band + (((size_t)b * 2 + axis) * F + t) * 2;
      bd[0] = ilo;
      bd[1] = ihi;
    }
  }
}

// ------------------------------------------------------------------ glimpse: row pass
// Superset of the union of the F taps' support bands along one axis, recomputed from the box (no memory traffic,
// no synchronisation): taps are monotone in t, so the union is [lo(tap 0), hi(tap F-1)].  Every thread of every
// consumer CTA runs this, so it uses the fast exp / sqrt intrinsics; two extra pixels on each side absorb their
// error and any last-bit difference to build_filters_kernel's arithmetic.  Filter entries outside a tap's band
// are exact zeros, so consumers may skip everything outside [lo, hi] (empty: lo > hi).
__device__ __forceinline__ void band_union(const float *__restrict__ bo, int axis, int F, int L, int *lo_out,
                                           int *hi_out) {
  const float ctr = bo[RA_BOX_CTR_Y + axis];
  const float size = bo[RA_BOX_SIZE_Y + axis];
  const float var = __expf(bo[RA_BOX_LGVAR_Y + axis]);
  const float step = (size + 1.0f) / (float)F;
  const float half = (float)(F - 1) / 2.0f;
  const float R = __fsqrt_rn(2.0f * var * kCut) * 1.0001f + 1.0f;
  const float m0 = ctr - step * half, m1 = ctr + step * half;
  float lo = floorf(fminf(m0, m1) - R) - 2.0f, hi = ceilf(fmaxf(m0, m1) + R) + 2.0f;
  int ilo = 0, ihi = L - 1;
  if (lo == lo && hi == hi) {  // not NaN (NaN boxes keep the full range, like the per-tap bands)
    lo = fminf(fmaxf(lo, 0.f), (float)L);
    hi = fmaxf(fminf(hi, (float)(L - 1)), -1.f);
    ilo = (int)lo;
    ihi = (int)hi;
  }
  *lo_out = ilo;
  *hi_out = ihi;
}

// One launch for both sources: CTAs [0, nx_s) of grid.x serve the static stack xs viewed as [B][H][W*Cs],
// the rest the canvas [B][H][W]; tmp_s [B][F][W*Cs], tmp_c [B][F][W].  Only the columns inside the union of
// the x-bands are produced (the column pass reads nothing else) and only the rows of the taps' y-bands are read.
// grid (chunks of 4*blockDim floats, F/IB tap groups, B).
template <int IB>
__global__ void __launch_bounds__(128) extract_rows_kernel(const float *__restrict__ xs, int Cs,
                                                           const float *__restrict__ canvas, int H, int W,
                                                           const float *__restrict__ fy, const int *__restrict__ band,
                                                           const float *__restrict__ box, int F,
                                                           float *__restrict__ tmp_s, float *__restrict__ tmp_c,
                                                           int nx_s) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  extern __shared__ float wsm[];  // [IB][H]
  int xr[2];
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * IB;
  const bool is_c = (int)blockIdx.x >= nx_s;
  const float *src = is_c ? canvas : xs;
  float *tmp = is_c ? tmp_c : tmp_s;
  const int C = is_c ? 1 : Cs;
  const int rowlen = W * C;
  const int bx = is_c ? (int)blockIdx.x - nx_s : (int)blockIdx.x;
  const int e0 = (bx * (int)blockDim.x + (int)threadIdx.x) * 4;
  band_union(box + (size_t)b * RA_BOX_STRIDE, 1, F, W, &xr[0], &xr[1]);
  const int elo = xr[0] * C, ehi = (xr[1] + 1) * C;  // floats [elo, ehi) of a row are needed
  {
    const int c0 = bx * (int)blockDim.x * 4, c1 = c0 + (int)blockDim.x * 4;
    if (c1 <= elo || c0 >= ehi) return;  // whole CTA outside the box columns (uniform)
  }
  const int *bd = band + ((size_t)b * 2 + 0) * F * 2;
  int ylo = H, yhi = -1;
#pragma unroll
  for (int k = 0; k < IB; ++k) {
    if (i0 + k < F) {
      const int lo = bd[(i0 + k) * 2], hi = bd[(i0 + k) * 2 + 1];
      if (hi >= lo) {
        ylo = min(ylo, lo);
        yhi = max(yhi, hi);
      }
    }
  }
  const int ny = yhi - ylo + 1;
  for (int idx = threadIdx.x; idx < IB * max(ny, 0); idx += blockDim.x) {
    const int k = idx / ny, yy = idx - k * ny;
    wsm[k * H + yy] = (i0 + k < F) ? fy[((size_t)b * F + i0 + k) * H + ylo + yy] : 0.f;
  }
  __syncthreads();
  if (e0 >= rowlen || e0 + 4 <= elo || e0 >= ehi) return;

  float4 acc[IB];
#pragma unroll
  for (int k = 0; k < IB; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float *sp = src + ((size_t)b * H + ylo) * rowlen + e0;
  int yy = 0;
  for (; yy + 4 <= ny; yy += 4) {  // four independent row loads in flight
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4 *>(sp + (size_t)(yy + u) * rowlen));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int k = 0; k < IB; ++k) {
        const float w0 = wsm[k * H + yy + u];
        acc[k].x = fmaf(w0, v[u].x, acc[k].x);
        acc[k].y = fmaf(w0, v[u].y, acc[k].y);
        acc[k].z = fmaf(w0, v[u].z, acc[k].z);
        acc[k].w = fmaf(w0, v[u].w, acc[k].w);
      }
    }
  }
  for (; yy < ny; ++yy) {
    const float4 v0 = __ldg(reinterpret_cast<const float4 *>(sp + (size_t)yy * rowlen));
#pragma unroll
    for (int k = 0; k < IB; ++k) {
      const float w0 = wsm[k * H + yy];
      acc[k].x = fmaf(w0, v0.x, acc[k].x);
      acc[k].y = fmaf(w0, v0.y, acc[k].y);
      acc[k].z = fmaf(w0, v0.z, acc[k].z);
      acc[k].w = fmaf(w0, v0.w, acc[k].w);
    }
  }
#pragma unroll
  for (int k = 0; k < IB; ++k)
    if (i0 + k < F) *reinterpret_cast<float4 *>(tmp + ((size_t)b * F + i0 + k) * rowlen + e0) = acc[k];
}

// ------------------------------------------------------------------ glimpse: column pass
// grid (F taps i, B); slab [W][D] of tmp rows in the reference's channel order.
__global__ void __launch_bounds__(256) extract_cols_kernel(const float *__restrict__ tmp_s, int Cs,
                                                           const float *__restrict__ tmp_c,
                                                           const int *__restrict__ chan_map,
                                                           const float *__restrict__ fx, const int *__restrict__ band,
                                                           const float *__restrict__ box, int W, int F, int Dp,
                                                           float *__restrict__ patch) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  extern __shared__ float slab[];  // [W][D]
  const int i = blockIdx.x, b = blockIdx.y;
  const int D = Cs + (tmp_c != nullptr ? 1 : 0);
  // only columns inside the union of the x-bands were produced by the row pass and are read below
  int xlo, xhi;
  band_union(box + (size_t)b * RA_BOX_STRIDE, 1, F, W, &xlo, &xhi);
  if (Cs > 0) {
    const float *ts = tmp_s + ((size_t)b * F + i) * (size_t)W * Cs;
    for (int idx = xlo * Cs + threadIdx.x; idx < (xhi + 1) * Cs; idx += blockDim.x) {
      const int x = idx / Cs, c = idx - x * Cs;
      slab[x * D + chan_map[c]] = ts[idx];
    }
  }
  if (tmp_c != nullptr) {
    const float *tc = tmp_c + ((size_t)b * F + i) * W;
    const int cc = chan_map[Cs];
    for (int x = xlo + threadIdx.x; x <= xhi; x += blockDim.x) slab[x * D + cc] = tc[x];
  }
  __syncthreads();
  const float gamma = box[(size_t)b * RA_BOX_STRIDE + RA_BOX_GAMMA_ATTN];
  const int *bd = band + ((size_t)b * 2 + 1) * F * 2;
  const float *fxb = fx + (size_t)b * F * W;
  // Dp >= D is the channel stride of the patch; channels D..Dp-1 are written as zeros (padding to a multiple of
  // 4 channels lets the tensor-core convolutions read the patch with TMA)
  for (int idx = threadIdx.x; idx < F * Dp; idx += blockDim.x) {
    const int j = idx / Dp, d = idx - j * Dp;
    if (d >= D) {
      patch[(((size_t)b * F + i) * F + j) * Dp + d] = 0.f;
      continue;
    }
    const int lo = bd[j * 2], hi = bd[j * 2 + 1];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float *fj = fxb + (size_t)j * W;
    int x = lo;
    for (; x + 3 <= hi; x += 4) {
      const float f0 = __ldg(fj + x), f1 = __ldg(fj + x + 1), f2 = __ldg(fj + x + 2), f3 = __ldg(fj + x + 3);
      a0 = fmaf(slab[x * D + d], f0, a0);
      a1 = fmaf(slab[(x + 1) * D + d], f1, a1);
      a2 = fmaf(slab[(x + 2) * D + d], f2, a2);
      a3 = fmaf(slab[(x + 3) * D + d], f3, a3);
    }
    for (; x <= hi; ++x) a0 = fmaf(slab[x * D + d], __ldg(fj + x), a0);
    patch[(((size_t)b * F + i) * F + j) * Dp + d] = gamma * ((a0 + a1) + (a2 + a3));  // full_model.py:788
  }
}

// ------------------------------------------------------------------ paste-back
constexpr int kPbTY = 8;      // rows per tile.  MEASURED (tools/pb_prof.py): 16 rows per tile (half the CTAs pay the patch load + Fy^T P
                              // prologue) is no faster - 23.1 vs 22.1 us per launch
constexpr int kPbTX = 128;    // columns per tile
constexpr int kPbTilesY = 1;  // consecutive row tiles served by one CTA.  MEASURED: 4 tiles per CTA (fewer, fatter CTAs,
                              // patch loaded once) is slower - 54 us vs 37 us per launch at 256x512, B=32 - because
                              // the tiles inside the box serialise within a CTA; 1 keeps them spread over the SMs.

#ifdef RA_PB_PROF
__device__ unsigned long long g_pb_prof[4096 * 8];
#define PB_PROF(i)                                                                                         \
  do {                                                                                                     \
    if (threadIdx.x == 0) {                                                                                \
      unsigned long long t_;                                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                               \
      g_pb_prof[((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) % 4096 * 8 + (i)] = t_;   \
    }                                                                                                      \
  } while (0)
#else
#define PB_PROF(i) do {} while (0)
#endif
__global__ void __launch_bounds__(256) paste_back_kernel(const float *__restrict__ patch,
                                                         const float *__restrict__ fy, const float *__restrict__ fx,
                                                         const int *__restrict__ band,
                                                         const float *__restrict__ box, int H, int W, int F,
                                                         int disable_overwrite, float *__restrict__ attn_box,
                                                         float *__restrict__ y_out, size_t out_bstride,
                                                         float *__restrict__ canvas) {
  PB_PROF(0);
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  PB_PROF(1);
  __shared__ float P_s[kMaxF * kMaxF];       // patch [F][F]
  __shared__ float wy_s[kPbTY][kMaxF];       // fy[i][y] for the tile rows
  __shared__ float t2_s[kPbTY][kMaxF + 1];   // sum_i fy[i][y] P[i][j]
  __shared__ float sy_s[kPbTY];              // sum_i fy[i][y]
  int yr[2] = {0, H - 1}, xr[2] = {0, W - 1};
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kPbTX;
  const int tid = threadIdx.x;
  const float *bo = box + (size_t)b * RA_BOX_STRIDE;
  const bool has_patch = patch != nullptr;

  // Tiles that no tap's support band reaches (most of the image for a small box): every filter entry is an exact
  // zero there, so attn_box = y_out = sigmoid(-5) - write the constants (float4) and update the canvas.
  // (band != NULL says the filters come from build_filters_kernel, i.e. are exact zeros outside the bands)
  if (band != nullptr) {
    band_union(bo, 0, F, H, &yr[0], &yr[1]);
    band_union(bo, 1, F, W, &xr[0], &xr[1]);
  }
  const bool x_outside = band != nullptr && (x0 > xr[1] || x0 + kPbTX - 1 < xr[0]);

  constexpr int kRows = kPbTY / 2;
  const int col = tid % kPbTX, rh = tid / kPbTX;  // 2 row-halves of kRows rows
  const int x = x0 + col;
  bool p_loaded = false;

  for (int tyi = 0; tyi < kPbTilesY; ++tyi) {
    const int y0 = (blockIdx.y * kPbTilesY + tyi) * kPbTY;
    if (y0 >= H) break;
    if (x_outside || (band != nullptr && (y0 > yr[1] || y0 + kPbTY - 1 < yr[0]))) {
      // ---------------- constant tile
      const float c5 = ra::sigmoidf_acc(-5.0f);
      if ((W & 3) == 0) {
        const float4 c4 = make_float4(c5, c5, c5, c5);
        for (int idx = tid; idx < kPbTY * (kPbTX / 4); idx += blockDim.x) {
          const int ty = idx / (kPbTX / 4), xx = x0 + (idx - ty * (kPbTX / 4)) * 4, y = y0 + ty;
          if (y >= H || xx >= W) continue;
          const size_t pix = (size_t)y * W + xx;
          if (attn_box != nullptr) *reinterpret_cast<float4 *>(attn_box + (size_t)b * out_bstride + pix) = c4;
          if (has_patch) {
            float4 *cp = reinterpret_cast<float4 *>(canvas + (size_t)b * H * W + pix);
            const float4 cv = *cp;
            float4 v = c4;
            if (disable_overwrite)
              v = make_float4(c5 * (1.0f - cv.x), c5 * (1.0f - cv.y), c5 * (1.0f - cv.z), c5 * (1.0f - cv.w));
            *reinterpret_cast<float4 *>(y_out + (size_t)b * out_bstride + pix) = v;
            *cp = make_float4(fmaxf(cv.x, v.x), fmaxf(cv.y, v.y), fmaxf(cv.z, v.z), fmaxf(cv.w, v.w));
          }
        }
      } else {
        for (int idx = tid; idx < kPbTY * kPbTX; idx += blockDim.x) {
          const int ty = idx / kPbTX, xx = x0 + (idx - ty * kPbTX), y = y0 + ty;
          if (y >= H || xx >= W) continue;
          const size_t pix = (size_t)y * W + xx;
          if (attn_box != nullptr) attn_box[(size_t)b * out_bstride + pix] = c5;
          if (has_patch) {
            const size_t cpix = (size_t)b * H * W + pix;
            const float cv = canvas[cpix];
            float v = c5;
            if (disable_overwrite) v *= (1.0f - cv);
            y_out[(size_t)b * out_bstride + pix] = v;
            canvas[cpix] = fmaxf(cv, v);
          }
        }
      }
      PB_PROF(7);
      continue;
    }

    // ---------------- tile inside the box: Fy^T P for its rows, then the band of Fx per column
    // taps whose band can reach column x (analytic superset; entries outside a band are 0)
    int jlo = 0, jhi = F - 1;
    {
      const float ctr = bo[RA_BOX_CTR_X], size = bo[RA_BOX_SIZE_X];
      const float var = expf(bo[RA_BOX_LGVAR_X]);
      const float step = (size + 1.0f) / (float)F;
      const float R = sqrtf(2.0f * var * kCut) + 1.0f;
      const float half = (float)(F - 1) / 2.0f;
      const float a = ((float)x - R - ctr) / step + half, c = ((float)x + R - ctr) / step + half;
      if (a == a && c == c && fabsf(a) < 1e9f && fabsf(c) < 1e9f) {
        jlo = max(0, (int)floorf(a));
        jhi = min(F - 1, (int)ceilf(c));
      }
    }
    // warp-uniform tap range so that t2_s reads are broadcasts
    for (int o = 16; o > 0; o >>= 1) {
      jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, o));
      jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, o));
    }
    const float g_box = bo[RA_BOX_GAMMA_BOX], g_y = bo[RA_BOX_GAMMA_Y];
    PB_PROF(2);
    __syncthreads();  // the previous tile's readers of wy_s / t2_s / sy_s are done
    if (has_patch && !p_loaded) {
      for (int idx = tid; idx < F * F; idx += blockDim.x) P_s[idx] = patch[(size_t)b * F * F + idx];
      p_loaded = true;
    }
    for (int idx = tid; idx < kPbTY * F; idx += blockDim.x) {
      const int ty = idx / F, i = idx - ty * F;
      const int y = y0 + ty;
      wy_s[ty][i] = (y < H) ? fy[((size_t)b * F + i) * H + y] : 0.f;
    }
    __syncthreads();
    PB_PROF(3);
    if (has_patch) {
      for (int idx = tid; idx < kPbTY * F; idx += blockDim.x) {
        const int ty = idx / F, j = idx - ty * F;
        float a = 0.f;
        for (int i = 0; i < F; ++i) a = fmaf(wy_s[ty][i], P_s[i * F + j], a);
        t2_s[ty][j] = a;
      }
    }
    if (tid < kPbTY) {
      float a = 0.f;
      for (int i = 0; i < F; ++i) a += wy_s[tid][i];
      sy_s[tid] = a;
    }
    __syncthreads();
    PB_PROF(4);

    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.f;
    float sx = 0.f;
    if (x < W) {
      const float *fxp = fx + (size_t)b * F * W + x;
      // the canvas values of the epilogue are requested before the tap loop: their L2 / HBM round trip overlaps it
      float cvp[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int y = y0 + rh * kRows + r;
        cvp[r] = (has_patch && y < H) ? canvas[(size_t)b * H * W + (size_t)y * W + x] : 0.f;
      }
      // four filter loads in flight per round (same summation order as the plain loop).  MEASURED: staging these Fx
      // rows in shared memory (24 KB per CTA, coalesced loads) is slower - 45 us vs 34 us per launch - the lost
      // occupancy costs more than the dependent L2 round trips.
      for (int j = jlo; j <= jhi; j += 4) {
        float wx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wx[u] = (j + u <= jhi) ? __ldg(fxp + (size_t)(j + u) * W) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + u > jhi) break;
          sx += wx[u];
          if (has_patch) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r] = fmaf(t2_s[rh * kRows + r][j + u], wx[u], acc[r]);
          }
        }
      }
      PB_PROF(5);
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const int y = y0 + rh * kRows + r;
        if (y >= H) continue;
        const size_t pix = (size_t)y * W + x;
        if (attn_box != nullptr)  // full_model.py:738-741
          attn_box[(size_t)b * out_bstride + pix] = ra::sigmoidf_acc(g_box * (sy_s[rh * kRows + r] * sx) - 5.0f);
        if (has_patch) {  // full_model.py:810-818, 845
          const size_t cpix = (size_t)b * H * W + pix;
          const float cv = cvp[r];
          float v = ra::sigmoidf_acc(g_y * acc[r] - 5.0f);
          if (disable_overwrite) v *= (1.0f - cv);
          y_out[(size_t)b * out_bstride + pix] = v;
          canvas[cpix] = fmaxf(cv, v);
        }
      }
    }
    PB_PROF(6);
  }
}

}  // namespace

#ifdef RA_PB_PROF
extern "C" int ra_debug_pb_prof(unsigned long long *host_out) {
  return cudaMemcpyFromSymbol(host_out, g_pb_prof, sizeof(unsigned long long) * 4096 * 8) == cudaSuccess ? 0 : 1;
}
extern "C" int ra_debug_pb_prof_clear() {
  static unsigned long long z[4096 * 8];
  return cudaMemcpyToSymbol(g_pb_prof, z, sizeof(z)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int ra_gaussian_filters_f32(const float *box, int B, int H, int W, int F, float *fy, float *fx,
                                       int32_t *band, void *stream) {
  if (!box || !fy || !fx || !band || B < 0 || H < 1 || W < 1 || F < 1) return RA_ERR_INVALID_ARG;
  if (F > kMaxF) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  const cudaError_t le = ra::launch_pdl(build_filters_kernel, dim3(B, 2, (F + kTapsPerCta - 1) / kTapsPerCta), dim3(256),
                                        (size_t)0, ra::as_stream(stream), box, H, W, F, fy, fx, band);
  if (le != cudaSuccess) {
    ra::set_last_error("cudaLaunchKernelEx(build_filters_kernel)", le);
    return RA_ERR_CUDA;
  }
  return ra::finish_launch("build_filters_kernel");
}

extern "C" int ra_gaussian_extract_f32(const float *xs, int Cs, const float *canvas, const int32_t *chan_map,
                                       const float *box, const float *fy, const float *fx, const int32_t *band, int B,
                                       int H, int W, int F, float *tmp, float *x_patch, int patch_cstride,
                                       void *stream) {
  if (!chan_map || !box || !fy || !fx || !band || !tmp || !x_patch || B < 0 || Cs < 0 || (Cs > 0 && !xs) ||
      (Cs == 0 && !canvas) || patch_cstride < Cs + (canvas != nullptr ? 1 : 0))
    return RA_ERR_INVALID_ARG;
  if (F > kMaxF || (W % 4) != 0) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  cudaStream_t s = ra::as_stream(stream);
  // taps per CTA of the row pass.  MEASURED: 8 taps per CTA (half the CTAs re-reading the same image rows, RA_EXTRACT_IB=8)
  // is no faster - 1.05 vs 1.01 ms per forward at KITTI B=32: the wider union of the taps' row bands costs what the
  // re-reads save.
  static const int IB = (getenv("RA_EXTRACT_IB") && atoi(getenv("RA_EXTRACT_IB")) == 8) ? 8 : 4;
  const int groups = (F + IB - 1) / IB;
  float *tmp_s = tmp;
  float *tmp_c = tmp + (size_t)B * F * W * Cs;
  const size_t smem_rows = (size_t)IB * H * sizeof(float);
  {
    // one launch for both sources (the static stack and the canvas)
    const int nx_s = Cs > 0 ? (W * Cs / 4 + 127) / 128 : 0;
    const int nx_c = canvas != nullptr ? (W / 4 + 127) / 128 : 0;
    dim3 grid(nx_s + nx_c, groups, B);
    const cudaError_t le =
        IB == 4 ? ra::launch_pdl(extract_rows_kernel<4>, grid, dim3(128), smem_rows, s, xs, Cs, canvas, H, W, fy, band, box,
                                 F, tmp_s, tmp_c, nx_s)
                : ra::launch_pdl(extract_rows_kernel<8>, grid, dim3(128), smem_rows, s, xs, Cs, canvas, H, W, fy, band, box,
                                 F, tmp_s, tmp_c, nx_s);
    if (le != cudaSuccess) {
      ra::set_last_error("cudaLaunchKernelEx(extract_rows_kernel)", le);
      return RA_ERR_CUDA;
    }
    const int rc = ra::finish_launch("extract_rows_kernel");
    if (rc != RA_OK) return rc;
  }
  const int D = Cs + (canvas != nullptr ? 1 : 0);
  const size_t smem_cols = (size_t)W * D * sizeof(float);
  if (smem_cols > 200 * 1024) return RA_ERR_UNSUPPORTED;
  static size_t attr_bytes = 48 * 1024;
  if (smem_cols > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(extract_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(extract_cols_kernel)", e);
      return RA_ERR_CUDA;
    }
    attr_bytes = 200 * 1024;
  }
  {
    const float *ts_arg = Cs > 0 ? tmp_s : nullptr, *tc_arg = canvas != nullptr ? tmp_c : nullptr;
    const cudaError_t le = ra::launch_pdl(extract_cols_kernel, dim3(F, B), dim3(256), smem_cols, s, ts_arg, Cs, tc_arg,
                                          chan_map, fx, band, box, W, F, patch_cstride, x_patch);
    if (le != cudaSuccess) {
      ra::set_last_error("cudaLaunchKernelEx(extract_cols_kernel)", le);
      return RA_ERR_CUDA;
    }
  }
  return ra::finish_launch("extract_cols_kernel");
}

extern "C" int ra_paste_back_f32(const float *patch, const float *box, const float *fy, const float *fx,
                                 const int32_t *band, int B, int H, int W, int F, int disable_overwrite,
                                 float *attn_box, float *y_out, size_t out_bstride, float *canvas, void *stream) {
  if (!box || !fy || !fx || B < 0 || H < 1 || W < 1) return RA_ERR_INVALID_ARG;
  if (patch != nullptr && (!y_out || !canvas)) return RA_ERR_INVALID_ARG;
  if (patch == nullptr && attn_box == nullptr) return RA_ERR_INVALID_ARG;
  if (F > kMaxF) return RA_ERR_UNSUPPORTED;
  if (B == 0) return RA_OK;
  dim3 grid((W + kPbTX - 1) / kPbTX, ((H + kPbTY - 1) / kPbTY + kPbTilesY - 1) / kPbTilesY, B);
  const cudaError_t le = ra::launch_pdl(paste_back_kernel, grid, dim3(256), (size_t)0, ra::as_stream(stream), patch, fy, fx,
                                        band, box, H, W, F, disable_overwrite, attn_box, y_out, out_bstride, canvas);
  if (le != cudaSuccess) {
    ra::set_last_error("cudaLaunchKernelEx(paste_back_kernel)", le);
    return RA_ERR_CUDA;
  }
  return ra::finish_launch("paste_back_kernel");
}

// Loss-side kernels of the hot path: ground-truth attention boxes, pairwise soft IoU / DICE,
// the scalar loss block, the score head and the box model's per-step ground-truth interaction.
// modellib.py:71-155 (f_dice/f_inter/f_union/f_iou), :268-339 (coverage, conf loss),
// :366-379 (greedy match), :482-511 (count metrics), :663-749 (GT boxes);
// full_model.py:821-822 (score), :916-1081 (loss block); box_model.py:484-504.
#include <mutex>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------ GT boxes
// One CTA per (b,t) mask: a single pass for the four index extrema and the area, then the
// padded rectangle and (optionally) the filled box.  The reference materialises five
// [B,T,H,W,2] temporaries for this (modellib.py:679-686,745-747).
__global__ void __launch_bounds__(256) gt_box_kernel(const float *__restrict__ y_gt, int H, int W, float padding_ratio,
                                                     float min_padding, float *__restrict__ top_left,
                                                     float *__restrict__ bot_right, float *__restrict__ rect,
                                                     float *__restrict__ box, float *__restrict__ area) {
  __shared__ float red[32];
  __shared__ float res[8];
  const size_t m = blockIdx.x;
  const float *g = y_gt + m * (size_t)H * W;
  const float hw = (float)(H * W);
  float mny = INFINITY, mnx = INFINITY, mxy = -INFINITY, mxx = -INFINITY, sum = 0.f;
  auto visit = [&](float v, float fy, float fx) {
    const float off = (1.0f - v) * hw;  // modellib.py:682-683
    mny = fminf(mny, fy + off);
    mnx = fminf(mnx, fx + off);
    mxy = fmaxf(mxy, fy * v);
    mxx = fmaxf(mxx, fx * v);
    sum += v;
  };
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    // streamed float4 loads, four in flight per thread, one row/column split per 16 bytes
    const int W4 = W >> 2, n4 = H * W4;
    for (int base = threadIdx.x; base < n4; base += blockDim.x * 4) {
      float4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i4 = base + u * blockDim.x;
        q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i4 < n4) q[u] = ra::ldg_stream4(g + (size_t)i4 * 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i4 = base + u * blockDim.x;
        if (i4 >= n4) break;
        const int yy = i4 / W4, xx = (i4 - yy * W4) * 4;
        const float fy = (float)yy, fx = (float)xx;
        visit(q[u].x, fy, fx);
        visit(q[u].y, fy, fx + 1.0f);
        visit(q[u].z, fy, fx + 2.0f);
        visit(q[u].w, fy, fx + 3.0f);
      }
    }
  } else {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) visit(g[i], (float)(i / W), (float)(i % W));
  }
  mny = ra::block_min(mny, red);
  mnx = ra::block_min(mnx, red);
  mxy = ra::block_max(mxy, red);
  mxx = ra::block_max(mxx, red);
  sum = ra::block_sum(sum, red);
  if (threadIdx.x == 0) {
    float tl[2] = {mny, mnx}, br[2] = {mxy, mxx};
    const float nz = sum > 0.f ? 1.f : 0.f;
    for (int d = 0; d < 2; ++d) {
      const float size = br[d] - tl[d];
      const float pad = fmaxf(padding_ratio * size, min_padding);  // modellib.py:689-693 (shift ratio 0)
      tl[d] -= pad;
      br[d] += pad;
      res[d] = tl[d];
      res[2 + d] = br[d];
      top_left[m * 2 + d] = tl[d] * nz;  // modellib.py:697-699
      bot_right[m * 2 + d] = nz * br[d] + (1.f - nz) * (2.f * min_padding);
    }
    if (rect != nullptr) {
      rect[m * 4 + 0] = tl[0];
      rect[m * 4 + 1] = tl[1];
      rect[m * 4 + 2] = br[0];
      rect[m * 4 + 3] = br[1];
    }
    if (area != nullptr) area[m] = sum;
  }
  if (box == nullptr) return;
  __syncthreads();
  const float ty = res[0], tx = res[1], by = res[2], bx = res[3];
  float *o = box + m * (size_t)H * W;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const float fy = (float)(i / W), fx = (float)(i % W);
    o[i] = (fy >= ty && fx >= tx && fy <= by && fx <= bx) ? 1.f : 0.f;  // modellib.py:745-747
  }
}

// ------------------------------------------------------------------ pairwise IoU
// C[n][m] = sum_k A[n][k] B[m][k] with a ones row appended to both operands, so row N / column
// M of C carry sum(b_m) / sum(a_n).  Split-K over the pixels: a CTA stages KC pixels of all
// rows in shared memory; thread (nb, mb, ks) owns a 6x6 block of C over its k-slice.
constexpr int kR = 6;     // register block
constexpr int kKC = 256;  // pixels per staged chunk
constexpr int kPad = 4;

struct IouParams {
  const float *a;
  const float *b;
  const float *b_rect;
  float *partial;
  int N, M, H, W;
  int NB, MB;    // register blocks along n and m
  int KS;        // k-slices per CTA
  int chunks_per_cta;
  int ctas_per_ex;
  float hard_thr;
};

__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}

// Stage chunk k0 of both operands into (A_s, B_s): rows of real masks by cp.async (16 B, L2 only: the data is read
// once), the ones row / zero padding rows / rectangle indicators / ragged tail by plain stores.
__device__ __forceinline__ void iou_stage(const IouParams &p, int b, int k0, int NP, int MP, float *A_s, float *B_s) {
  const int HW = p.H * p.W, ld = kKC + kPad, tid = threadIdx.x;
  for (int idx = tid; idx < NP * (kKC / 4); idx += blockDim.x) {
    const int r = idx / (kKC / 4), q = (idx - r * (kKC / 4)) * 4;
    const int k = k0 + q;
    float *dst = A_s + r * ld + q;
    if (r < p.N && k < HW) {
      cp_async16(dst, p.a + ((size_t)b * p.N + r) * HW + k);
    } else {
      const float o = (r == p.N && k < HW) ? 1.f : 0.f;
      *reinterpret_cast<float4 *>(dst) = make_float4(o, o, o, o);
    }
  }
  for (int idx = tid; idx < MP * (kKC / 4); idx += blockDim.x) {
    const int r = idx / (kKC / 4), q = (idx - r * (kKC / 4)) * 4;
    const int k = k0 + q;
    float *dst = B_s + r * ld + q;
    if (r < p.M && k < HW && p.b_rect == nullptr) {
      cp_async16(dst, p.b + ((size_t)b * p.M + r) * HW + k);
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < HW) {
        if (r < p.M) {
          const float *rc = p.b_rect + ((size_t)b * p.M + r) * 4;
          const float ty = rc[0], tx = rc[1], by = rc[2], bx = rc[3];
          float o[4];
          if ((p.W & 3) == 0) {  // the four pixels share a row: one division
            const int yy = k / p.W, xx = k - yy * p.W;
            const float fy = (float)yy;
            const bool row_in = fy >= ty && fy <= by;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float fx = (float)(xx + e);
              o[e] = (row_in && fx >= tx && fx <= bx) ? 1.f : 0.f;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float fy = (float)((k + e) / p.W), fx = (float)((k + e) % p.W);
              o[e] = (fy >= ty && fx >= tx && fy <= by && fx <= bx) ? 1.f : 0.f;
            }
          }
          v = make_float4(o[0], o[1], o[2], o[3]);
        } else if (r == p.M) {
          v = make_float4(1.f, 1.f, 1.f, 1.f);
        }
      }
      *reinterpret_cast<float4 *>(dst) = v;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(256) pairwise_iou_kernel(IouParams p) {
  extern __shared__ __align__(16) float smem[];
  const int NP = p.NB * kR, MP = p.MB * kR;
  const int ld = kKC + kPad;
  const int buf_floats = (NP + MP) * ld;  // one stage: A rows [NP][ld] then B rows [MP][ld]; two stages
  const int b = blockIdx.y;
  const int HW = p.H * p.W;
  const int tid = threadIdx.x;
  const int ks = tid % p.KS;
  const int blk = tid / p.KS;
  const int nb = blk / p.MB, mb = blk % p.MB;
  const bool active = blk < p.NB * p.MB;
  const bool hard = p.hard_thr > 0.f;

  float acc[kR][kR];
#pragma unroll
  for (int i = 0; i < kR; ++i)
#pragma unroll
    for (int j = 0; j < kR; ++j) acc[i][j] = 0.f;

  const int chunk0 = blockIdx.x * p.chunks_per_cta;
  int n_ch = 0;
  for (int ch = 0; ch < p.chunks_per_cta; ++ch)
    if ((chunk0 + ch) * kKC < HW) n_ch = ch + 1;
  if (n_ch > 0) iou_stage(p, b, chunk0 * kKC, NP, MP, smem, smem + NP * ld);
  for (int ch = 0; ch < n_ch; ++ch) {
    float *A_s = smem + (ch & 1) * buf_floats;
    float *B_s = A_s + NP * ld;
    if (ch + 1 < n_ch) {  // prefetch the next chunk into the other stage while this one is consumed
      float *A_n = smem + ((ch + 1) & 1) * buf_floats;
      iou_stage(p, b, (chunk0 + ch + 1) * kKC, NP, MP, A_n, A_n + NP * ld);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      for (int q = ks * 4; q < kKC; q += p.KS * 4) {
        float4 av[kR], bv[kR];
#pragma unroll
        for (int i = 0; i < kR; ++i) av[i] = *reinterpret_cast<const float4 *>(A_s + (nb * kR + i) * ld + q);
#pragma unroll
        for (int j = 0; j < kR; ++j) bv[j] = *reinterpret_cast<const float4 *>(B_s + (mb * kR + j) * ld + q);
        if (hard) {  // thresholded masks (the ones row stays 1, padding stays 0: thr is in (0, 1))
#pragma unroll
          for (int i = 0; i < kR; ++i) {
            av[i].x = av[i].x > p.hard_thr ? 1.f : 0.f;
            av[i].y = av[i].y > p.hard_thr ? 1.f : 0.f;
            av[i].z = av[i].z > p.hard_thr ? 1.f : 0.f;
            av[i].w = av[i].w > p.hard_thr ? 1.f : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < kR; ++i)
#pragma unroll
          for (int j = 0; j < kR; ++j) {
            acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
            acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
            acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
            acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
          }
      }
    }
    __syncthreads();  // the stage is refilled by the next iteration's prefetch
  }
  // reduce the KS k-slices of every (nb, mb) block through shared memory, fixed order
  __syncthreads();
  float *red = smem;  // [blocks][KS][36] fits: blocks*KS <= 256 -> 256*36 floats
  if (active) {
#pragma unroll
    for (int i = 0; i < kR; ++i)
#pragma unroll
      for (int j = 0; j < kR; ++j) red[((size_t)blk * p.KS + ks) * (kR * kR) + i * kR + j] = acc[i][j];
  }
  __syncthreads();
  float *out = p.partial + ((size_t)b * p.ctas_per_ex + blockIdx.x) * NP * MP;
  for (int idx = tid; idx < NP * MP; idx += blockDim.x) {
    const int n = idx / MP, m = idx - n * MP;
    const int bk = (n / kR) * p.MB + (m / kR);
    const int e = (n % kR) * kR + (m % kR);
    float s = 0.f;
    for (int k = 0; k < p.KS; ++k) s += red[((size_t)bk * p.KS + k) * (kR * kR) + e];
    out[idx] = s;
  }
}

__global__ void pairwise_iou_finalize_kernel(const float *__restrict__ partial, int N, int M, int NP, int MP,
                                             int ctas_per_ex, float hw_eps, float *__restrict__ iou,
                                             float *__restrict__ dice) {
  const int b = blockIdx.x;
  __shared__ float C[40 * 40];
  for (int idx = threadIdx.x; idx < NP * MP; idx += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < ctas_per_ex; ++c) s += partial[((size_t)b * ctas_per_ex + c) * NP * MP + idx];
    C[idx] = s;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * M; idx += blockDim.x) {
    const int n = idx / M, m = idx - n * M;
    const float inter = C[n * MP + m];
    const float sa = C[n * MP + M];  // a_n . ones
    const float sb = C[N * MP + m];  // ones . b_m
    // modellib.py:110-114: sum(a + b - a*b + eps) over the pixels
    iou[((size_t)b * N + n) * M + m] = inter / (sa + sb - inter + hw_eps);
    // modellib.py:90-95: 2*inter / (sum(a + 1e-5) + sum(b + 1e-5))
    if (dice != nullptr) dice[((size_t)b * N + n) * M + m] = 2.f * inter / ((sa + hw_eps) + (sb + hw_eps));
  }
}

// ------------------------------------------------------------------ loss block
__global__ void __launch_bounds__(256) loss_block_kernel(const float *__restrict__ iou_box,
                                                         const float *__restrict__ match_box,
                                                         const float *__restrict__ iou_soft,
                                                         const float *__restrict__ match,
                                                         const float *__restrict__ iou_hard,
                                                         const float *__restrict__ dice_hard,
                                                         const float *__restrict__ s_out,
                                                         const float *__restrict__ s_gt,
                                                         const float *__restrict__ gt_area, int B, int T,
                                                         float mix, float wd_term, float segm_coeff,
                                                         float *__restrict__ out) {
  __shared__ float red[32];
  float v[RA_LOSS_COUNT];
#pragma unroll
  for (int k = 0; k < RA_LOSS_COUNT; ++k) v[k] = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const size_t o2 = (size_t)b * T * T, o1 = (size_t)b * T;
    float s_ib = 0.f, c_b = 0.f, s_is = 0.f, c_s = 0.f, s_ih = 0.f, s_d = 0.f;
    for (int k = 0; k < T * T; ++k) {
      const float mb = match_box[o2 + k], ms = match[o2 + k];
      s_ib += iou_box[o2 + k] * mb;
      c_b += mb;
      s_is += iou_soft[o2 + k] * ms;
      c_s += ms;
      if (iou_hard != nullptr) s_ih += iou_hard[o2 + k] * ms;
      if (dice_hard != nullptr) s_d += dice_hard[o2 + k] * ms;
    }
    c_b = fmaxf(c_b, 1.f);  // full_model.py:944,994
    c_s = fmaxf(c_s, 1.f);
    v[RA_LOSS_BOX] -= s_ib / c_b;
    v[RA_LOSS_IOU_SOFT] += s_is / c_s;
    v[RA_LOSS_IOU_HARD] += s_ih / c_s;
    v[RA_LOSS_DICE] += s_d / c_s;
    // coverage (modellib.py:268-313)
    float tot = 0.f;
    for (int m = 0; m < T; ++m) tot += gt_area[o1 + m];
    float wcs = 0.f, ucs = 0.f, wch = 0.f, uch = 0.f;
    for (int m = 0; m < T; ++m) {
      float cs = -INFINITY, chd = -INFINITY;
      for (int n = 0; n < T; ++n) {
        cs = fmaxf(cs, iou_soft[o2 + n * T + m]);
        if (iou_hard != nullptr) chd = fmaxf(chd, iou_hard[o2 + n * T + m]);
      }
      const float ar = gt_area[o1 + m];
      const float wt = ar / (tot + (ar == 0.f ? 1.f : 0.f));
      wcs += cs * wt;
      ucs += cs;
      if (iou_hard != nullptr) {
        wch += chd * wt;
        uch += chd;
      }
    }
    v[RA_LOSS_WT_COV_SOFT] += wcs;
    v[RA_LOSS_UNWT_COV_SOFT] += ucs / c_s;
    v[RA_LOSS_WT_COV_HARD] += wch;
    v[RA_LOSS_UNWT_COV_HARD] += uch / c_s;
    // confidence loss with cumulative min / reversed cumulative max (modellib.py:316-339,430-437)
    float conf = 0.f;
    {
      float run_min = INFINITY;
      float cnt_out = 0.f, cnt_gt = 0.f;
      for (int t = 0; t < T; ++t) {
        const float s = s_out[o1 + t];
        run_min = fminf(run_min, s);
        float gsum = 0.f;
        for (int m = 0; m < T; ++m) gsum += match[o2 + t * T + m];
        float run_max = -INFINITY;
        for (int u = t; u < T; ++u) run_max = fmaxf(run_max, s_out[o1 + u]);
        conf += -gsum * logf(run_min + 1e-5f) - (1.f - gsum) * logf(1.f - run_max + 1e-5f);
        cnt_out += s > 0.5f ? 1.f : 0.f;
        cnt_gt += s_gt[o1 + t];
      }
      v[RA_LOSS_COUNT_ACC] += (cnt_out == cnt_gt) ? 1.f : 0.f;
      v[RA_LOSS_DIC] += cnt_out - cnt_gt;
      v[RA_LOSS_DIC_ABS] += fabsf(cnt_out - cnt_gt);
    }
    v[RA_LOSS_CONF] += conf;
  }
  const float fB = (float)B;
#pragma unroll
  for (int k = 0; k < RA_LOSS_COUNT; ++k) v[k] = ra::block_sum(v[k], red);
  if (threadIdx.x == 0) {
    out[RA_LOSS_BOX] = v[RA_LOSS_BOX] / fB;
    out[RA_LOSS_IOU_SOFT] = v[RA_LOSS_IOU_SOFT] / fB;
    out[RA_LOSS_SEGM] = -out[RA_LOSS_IOU_SOFT];
    out[RA_LOSS_CONF] = v[RA_LOSS_CONF] / fB / (float)T;
    out[RA_LOSS_IOU_HARD] = v[RA_LOSS_IOU_HARD] / fB;
    out[RA_LOSS_WT_COV_SOFT] = v[RA_LOSS_WT_COV_SOFT] / fB;
    out[RA_LOSS_UNWT_COV_SOFT] = v[RA_LOSS_UNWT_COV_SOFT] / fB;
    out[RA_LOSS_WT_COV_HARD] = v[RA_LOSS_WT_COV_HARD] / fB;
    out[RA_LOSS_UNWT_COV_HARD] = v[RA_LOSS_UNWT_COV_HARD] / fB;
    out[RA_LOSS_DICE] = v[RA_LOSS_DICE] / fB;
    out[RA_LOSS_COUNT_ACC] = v[RA_LOSS_COUNT_ACC] / fB;
    out[RA_LOSS_DIC] = v[RA_LOSS_DIC] / fB;
    out[RA_LOSS_DIC_ABS] = v[RA_LOSS_DIC_ABS] / fB;
    out[RA_LOSS_TOTAL] = out[RA_LOSS_BOX] + segm_coeff * out[RA_LOSS_SEGM] + mix * out[RA_LOSS_CONF] + wd_term;
    out[14] = 0.f;
    out[15] = 0.f;
  }
}

// ------------------------------------------------------------------ score head
__global__ void score_kernel(const float *__restrict__ h, int Hd, const float *__restrict__ core, int Cd,
                             const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ s_out,
                             int s_stride) {
  ra::pdl_wait();     // PDL: the previous kernel of the stream has completed, its results are visible
  ra::pdl_trigger();  // the next kernel may be scheduled (it waits the same way)
  __shared__ float red[32];
  const int b = blockIdx.x;
  float a = 0.f;
  for (int k = threadIdx.x; k < Hd; k += blockDim.x) a = fmaf(h[(size_t)b * Hd + k], w[k], a);
  for (int k = threadIdx.x; k < Cd; k += blockDim.x) a = fmaf(core[(size_t)b * Cd + k], w[Hd + k], a);
  a = ra::block_sum(a, red);
  if (threadIdx.x == 0) s_out[(size_t)b * s_stride] = ra::sigmoidf_acc(a + bias[0]);
}

// ------------------------------------------------------------------ box model GT interaction
// Phase 1: per example, soft IoU of this step's attention box against every GT rectangle, then the greedy match.
// Two launches: (row chunks x examples) CTAs accumulate, per GT rectangle, the sum of the box inside it (and the total),
// one small CTA per example finishes - the rectangle areas are analytic.  (The first version used ONE CTA per example
// for all H*W pixels: 1.27 ms per launch at 256x512, B = 32.)
constexpr int kBgChunks = 16;  // row chunks per example

// One pass over the chunk's pixels: a thread walks COLUMNS (x = tid, tid + 256, ..), so which rectangles contain its
// column is a bit mask computed once per column; which contain the row is a per-row mask in shared memory; the T sums
// live in registers.  (The first single-kernel version made ceil(T / 8) passes with four compares per pixel and
// rectangle: 72 us per launch at 256x512, B = 32.)
constexpr int kBgMaxT = 32;
__global__ void __launch_bounds__(256) box_gt_partial_kernel(const float *__restrict__ attn_box, size_t box_bstride,
                                                             const float *__restrict__ gt_rect, int T, int H, int W,
                                                             float *__restrict__ partial /* [B][kBgChunks][T+1] */) {
  extern __shared__ float sm[];  // rc_s [T][4] | ymask_s [rows] (as uint) | red_s [8][kBgMaxT + 1]
  const int b = blockIdx.y, ch = blockIdx.x;
  const int rows = (H + kBgChunks - 1) / kBgChunks;
  float *rc_s = sm;
  unsigned *ymask_s = reinterpret_cast<unsigned *>(sm + 4 * T);
  float *red_s = sm + 4 * T + rows;
  for (int i = threadIdx.x; i < T * 4; i += blockDim.x) rc_s[i] = gt_rect[(size_t)b * T * 4 + i];
  __syncthreads();
  const int y0 = ch * rows, y1 = min(H, y0 + rows);
  for (int r = threadIdx.x; r < y1 - y0; r += blockDim.x) {
    const float fy = (float)(y0 + r);
    unsigned m = 0;
    for (int k = 0; k < T; ++k) m |= (fy >= rc_s[k * 4 + 0] && fy <= rc_s[k * 4 + 2]) ? (1u << k) : 0u;
    ymask_s[r] = m;
  }
  __syncthreads();
  const float *bx = attn_box + (size_t)b * box_bstride + (size_t)y0 * W;
  float *out = partial + ((size_t)b * kBgChunks + ch) * (T + 1);
  float acc[kBgMaxT];
#pragma unroll
  for (int k = 0; k < kBgMaxT; ++k) acc[k] = 0.f;
  float tot = 0.f;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const float fx = (float)x;
    unsigned xm = 0;
    for (int k = 0; k < T; ++k) xm |= (fx >= rc_s[k * 4 + 1] && fx <= rc_s[k * 4 + 3]) ? (1u << k) : 0u;
    int r = 0;
    for (; r + 4 <= y1 - y0; r += 4) {  // four independent loads in flight
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = bx[(size_t)(r + u) * W + x];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        tot += v[u];
        const unsigned m = xm & ymask_s[r + u];
#pragma unroll
        for (int k = 0; k < kBgMaxT; ++k) acc[k] += ((m >> k) & 1u) ? v[u] : 0.f;
      }
    }
    for (; r < y1 - y0; ++r) {
      const float v = bx[(size_t)r * W + x];
      tot += v;
      const unsigned m = xm & ymask_s[r];
#pragma unroll
      for (int k = 0; k < kBgMaxT; ++k) acc[k] += ((m >> k) & 1u) ? v : 0.f;
    }
  }
  // warp sums, then the 8 warps through shared memory (fixed order)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kBgMaxT; ++k) {
    const float s1 = ra::warp_sum(acc[k]);
    if (lane == 0) red_s[warp * (kBgMaxT + 1) + k] = s1;
  }
  tot = ra::warp_sum(tot);
  if (lane == 0) red_s[warp * (kBgMaxT + 1) + kBgMaxT] = tot;
  __syncthreads();
  for (int k = threadIdx.x; k <= T; k += blockDim.x) {
    const int src = k < T ? k : kBgMaxT;
    float s1 = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s1 += red_s[w * (kBgMaxT + 1) + src];
    out[k] = s1;
  }
}

// T > kBgMaxT (up to 64 rectangles): ceil(T / 8) passes over the chunk, eight sums per pass
__global__ void __launch_bounds__(256) box_gt_partial_wide_kernel(const float *__restrict__ attn_box, size_t box_bstride,
                                                             const float *__restrict__ gt_rect, int T, int H, int W,
                                                             float *__restrict__ partial /* [B][kBgChunks][T+1] */) {
  extern __shared__ float sm[];  // rc_s [T][4]
  __shared__ float red[32];
  const int b = blockIdx.y, ch = blockIdx.x;
  float *rc_s = sm;
  for (int i = threadIdx.x; i < T * 4; i += blockDim.x) rc_s[i] = gt_rect[(size_t)b * T * 4 + i];
  __syncthreads();
  const int rows = (H + kBgChunks - 1) / kBgChunks;
  const int y0 = ch * rows, y1 = min(H, y0 + rows);
  const int n = max(0, y1 - y0) * W;
  const float *bx = attn_box + (size_t)b * box_bstride + (size_t)y0 * W;
  float *out = partial + ((size_t)b * kBgChunks + ch) * (T + 1);
  float tot = 0.f;
  for (int m0 = 0; m0 < T; m0 += 8) {
    float in8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) in8[e] = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = bx[i];
      const int yy = i / W;
      const float fy = (float)(y0 + yy), fx = (float)(i - yy * W);
      if (m0 == 0) tot += v;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (m0 + e < T) {
          const float *rc = rc_s + (m0 + e) * 4;
          const bool in = fy >= rc[0] && fx >= rc[1] && fy <= rc[2] && fx <= rc[3];
          in8[e] += in ? v : 0.f;
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float s1 = ra::block_sum(in8[e], red);
      if (threadIdx.x == 0 && m0 + e < T) out[m0 + e] = s1;
    }
  }
  tot = ra::block_sum(tot, red);
  if (threadIdx.x == 0) out[T] = tot;
}

__global__ void box_gt_finalize_kernel(const float *__restrict__ partial, const float *__restrict__ gt_rect, int T, int H,
                                       int W, float *__restrict__ iou_t, int iou_bstride, float *__restrict__ grd) {
  __shared__ float iou_s[64];
  const int b = blockIdx.x, m = threadIdx.x;
  const float *pp = partial + (size_t)b * kBgChunks * (T + 1);
  float tot = 0.f;
  for (int c = 0; c < kBgChunks; ++c) tot += pp[c * (T + 1) + T];
  if (m < T) {
    float inter = 0.f;
    for (int c = 0; c < kBgChunks; ++c) inter += pp[c * (T + 1) + m];
    // number of pixels (y, x) with ty <= y <= by, tx <= x <= bx inside the image
    const float *rc = gt_rect + ((size_t)b * T + m) * 4;
    const float ylo = fmaxf(0.f, ceilf(rc[0])), yhi = fminf((float)(H - 1), floorf(rc[2]));
    const float xlo = fmaxf(0.f, ceilf(rc[1])), xhi = fminf((float)(W - 1), floorf(rc[3]));
    const float area = fmaxf(0.f, yhi - ylo + 1.f) * fmaxf(0.f, xhi - xlo + 1.f);
    const float hw_eps = (float)(H * W) * 1e-5f;
    const float v = inter / (tot + area - inter + hw_eps);  // box_model.py:494-495, full_model.py:756-758
    iou_t[(size_t)b * iou_bstride + m] = v;
    iou_s[m] = v;
  }
  __syncthreads();
  if (m < T) {  // modellib.py:366-379 with matched == 0: one-hot of the row max, ties share 1/k
    float mx = -INFINITY, cnt = 0.f;
    for (int k = 0; k < T; ++k) mx = fmaxf(mx, iou_s[k]);
    for (int k = 0; k < T; ++k) cnt += (iou_s[k] == mx) ? 1.f : 0.f;
    grd[(size_t)b * T + m] = (iou_s[m] == mx) ? 1.f / cnt : 0.f;
  }
}

// grd_ws doubles as the workspace of the partial sums: it must hold B * max(T, kBgChunks * (T + 1)) floats
int box_gt_iou_launch(const float *attn_box, size_t box_bstride, const float *gt_rect, int B, int T, int H, int W,
                      float *iou_t, int iou_bstride, float *grd, float *partial, cudaStream_t s) {
  if (T > kBgMaxT) {
    box_gt_partial_wide_kernel<<<dim3(kBgChunks, B), 256, (size_t)(4 * T) * sizeof(float), s>>>(attn_box, box_bstride,
                                                                                              gt_rect, T, H, W, partial);
  } else {
    const size_t bg_smem = ((size_t)4 * T + (H + kBgChunks - 1) / kBgChunks + 8 * (kBgMaxT + 1)) * sizeof(float);
    box_gt_partial_kernel<<<dim3(kBgChunks, B), 256, bg_smem, s>>>(attn_box, box_bstride, gt_rect, T, H, W, partial);
  }
  int rc = ra::finish_launch("box_gt_partial_kernel");
  if (rc != RA_OK) return rc;
  box_gt_finalize_kernel<<<B, 64, 0, s>>>(partial, gt_rect, T, H, W, iou_t, iou_bstride, grd);
  return ra::finish_launch("box_gt_finalize_kernel");
}

// use_iou_box (full_model.py:750-754, box_model.py:487-491): the knob / box-model greedy match on the coordinate IoU
// of modellib.f_iou_box (modellib.py:206-238; no eps, strict overlap test) between this step's box and every (clean)
// GT box, instead of the soft IoU of the pasted attention box.  One CTA per example, thread m = GT box m.
__global__ void greedy_iou_box_kernel(const float *__restrict__ box, const float *__restrict__ tl_gt,
                                      const float *__restrict__ br_gt, int T, float *__restrict__ iou_t,
                                      int iou_bstride, float *__restrict__ grd) {
  __shared__ float iou_s[64];
  const int b = blockIdx.x, m = threadIdx.x;
  const float *bo = box + (size_t)b * RA_BOX_STRIDE;
  if (m < T) {
    const float y1a = bo[RA_BOX_TL_Y], x1a = bo[RA_BOX_TL_X], y2a = bo[RA_BOX_BR_Y], x2a = bo[RA_BOX_BR_X];
    // the GT corners are get_gt_box's RETURNED top_left / bot_right, i.e. after the empty-mask fix (an empty mask
    // gives the box (0,0)-(2*min_padding), modellib.py:697-699) - not the rectangle of the filled box
    const float *tg = tl_gt + ((size_t)b * T + m) * 2, *bg = br_gt + ((size_t)b * T + m) * 2;
    const float y1b = tg[0], x1b = tg[1], y2b = bg[0], x2b = bg[1];
    const float x1 = fmaxf(x1a, x1b), y1 = fmaxf(y1a, y1b), x2 = fminf(x2a, x2b), y2 = fminf(y2a, y2b);
    const float flag = ((x1 < x2) ? 1.f : 0.f) * ((y1 < y2) ? 1.f : 0.f);
    const float inter = __fmul_rn(__fmul_rn(flag, x2 - x1), y2 - y1);
    const float area_a = __fmul_rn(x2a - x1a, y2a - y1a), area_b = __fmul_rn(x2b - x1b, y2b - y1b);
    const float v = __fdiv_rn(inter, __fadd_rn(area_a, area_b) - inter);
    iou_t[(size_t)b * iou_bstride + m] = v;
    iou_s[m] = v;
  }
  __syncthreads();
  if (m < T) {  // modellib.py:366-379 with matched == 0; a NaN score (0/0) poisons the row like a NaN-propagating max
    float mx = -INFINITY, cnt = 0.f;
    bool nan = false;
    for (int k = 0; k < T; ++k) {
      nan |= (iou_s[k] != iou_s[k]);
      mx = fmaxf(mx, iou_s[k]);
    }
    for (int k = 0; k < T; ++k) cnt += (iou_s[k] == mx) ? 1.f : 0.f;
    grd[(size_t)b * T + m] = nan ? __int_as_float(0x7fc00000) : ((iou_s[m] == mx) ? 1.f / cnt : 0.f);
  }
}

// Phase 2: canvas = max(canvas, sum_m grd[m] * y_gt[m] * (1 - noise))   (box_model.py:497-503)
__global__ void box_gt_canvas_kernel(const float *__restrict__ grd, const float *__restrict__ y_gt,
                                     const float *__restrict__ noise, size_t noise_bstride, int T, int HW,
                                     float *__restrict__ canvas) {
  const int b = blockIdx.y;
  __shared__ float g_s[64];
  for (int m = threadIdx.x; m < T; m += blockDim.x) g_s[m] = grd[(size_t)b * T + m];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int m = 0; m < T; ++m) {
      const float g = g_s[m];
      if (g != 0.f) v += g * y_gt[((size_t)b * T + m) * HW + i];
    }
    if (noise != nullptr) v = v - v * noise[(size_t)b * noise_bstride + i];
    const size_t ci = (size_t)b * HW + i;
    canvas[ci] = fmaxf(canvas[ci], v);
  }
}

// ------------------------------------------------------------------ scheduled-sampling knob (training mode)
// Noisy GT attention boxes, full_model.py:568-580 = modellib.get_gt_attn / get_gt_box (modellib.py:644-701) with a
// per-(example, object) padding ratio and centre shift: from the raw extrema of the mask (rect_raw = get_gt_box with
// zero padding) to centre and size.  One thread per (b, t).
__global__ void gt_attn_noise_kernel(const float *__restrict__ rect_raw, const float *__restrict__ area,
                                     const float *__restrict__ pad, const float *__restrict__ shift, float min_padding,
                                     int n, float *__restrict__ ctr, float *__restrict__ size) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float nz = area[i] > 0.f ? 1.f : 0.f;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    float tl = rect_raw[i * 4 + d], br = rect_raw[i * 4 + 2 + d];
    const float sz = br - tl;                                 // modellib.py:689
    tl += shift[i * 2 + d] * sz;                              // :690
    tl -= fmaxf(pad[i] * sz, min_padding);                    // :691
    br += shift[i * 2 + d] * sz;                              // :692
    br += fmaxf(pad[i] * sz, min_padding);                    // :693
    tl *= nz;                                                 // :697-699 (empty mask -> top-left corner box)
    br = nz * br + (1.f - nz) * (2.f * min_padding);
    ctr[i * 2 + d] = (tl + br) / 2.0f;                        // get_box_ctr_size
    size[i * 2 + d] = br - tl;
  }
}

// full_model.py:760-773: where the Bernoulli switch of this step is on, centre and size of the attention box become
// those of the greedily matched noisy GT box (grd = one-hot / tie-shared weights); top-left / bottom-right follow.
__global__ void knob_mix_box_kernel(float *__restrict__ box, const float *__restrict__ grd,
                                    const float *__restrict__ ctr_gt, const float *__restrict__ size_gt,
                                    const float *__restrict__ knob, int knob_stride, int B, int T) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float *bo = box + (size_t)b * RA_BOX_STRIDE;
  const float kb = knob[(size_t)b * knob_stride];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    float c = 0.f, s = 0.f;
    for (int m = 0; m < T; ++m) {
      const float g = grd[(size_t)b * T + m];
      c += g * ctr_gt[((size_t)b * T + m) * 2 + d];
      s += g * size_gt[((size_t)b * T + m) * 2 + d];
    }
    const float cn = kb * c + (1.f - kb) * bo[RA_BOX_CTR_Y + d];
    const float sn = kb * s + (1.f - kb) * bo[RA_BOX_SIZE_Y + d];
    bo[RA_BOX_CTR_Y + d] = cn;
    bo[RA_BOX_SIZE_Y + d] = sn;
    bo[RA_BOX_TL_Y + d] = cn - sn / 2.0f;  // modellib.get_box_coord
    bo[RA_BOX_BR_Y + d] = cn + sn / 2.0f;
  }
}

// full_model.py:826-845: canvas = max(canvas, knob ? (sum_m grd*y_gt) * (1 - noise) : y_out)
__global__ void knob_canvas_kernel(const float *__restrict__ grd, const float *__restrict__ y_gt,
                                   const float *__restrict__ noise, size_t noise_bstride,
                                   const float *__restrict__ knob, int knob_stride, const float *__restrict__ y_out,
                                   size_t out_bstride, int T, int HW, float *__restrict__ canvas) {
  const int b = blockIdx.y;
  __shared__ float g_s[64];
  for (int m = threadIdx.x; m < T; m += blockDim.x) g_s[m] = grd[(size_t)b * T + m];
  __syncthreads();
  const float ks = knob[(size_t)b * knob_stride];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (ks != 0.f) {
      for (int m = 0; m < T; ++m) {
        const float g = g_s[m];
        if (g != 0.f) v += g * y_gt[((size_t)b * T + m) * HW + i];
      }
      v = v - v * noise[(size_t)b * noise_bstride + i];
    }
    const float yo = y_out[(size_t)b * out_bstride + i];
    const float w = ks * v + (1.f - ks) * yo;
    const size_t ci = (size_t)b * HW + i;
    canvas[ci] = fmaxf(canvas[ci], w);
  }
}

int iou_plan(int N, int M, int HW, IouParams *p) {
  const int rows = (N > M ? N : M) + 1;
  const int nb = (rows + kR - 1) / kR;
  if (nb * kR > 40) return RA_ERR_UNSUPPORTED;
  p->NB = p->MB = nb;
  p->KS = 256 / (nb * nb);
  if (p->KS > 64) p->KS = 64;
  if (p->KS < 1) return RA_ERR_UNSUPPORTED;
  const int chunks = (HW + kKC - 1) / kKC;
  p->ctas_per_ex = chunks < 32 ? chunks : 32;
  p->chunks_per_cta = (chunks + p->ctas_per_ex - 1) / p->ctas_per_ex;
  p->ctas_per_ex = (chunks + p->chunks_per_cta - 1) / p->chunks_per_cta;
  return RA_OK;
}

}  // namespace

// Device scratch of the box/GT IoU partial sums, grown on demand and kept (a few KB; one per process and device).
// Allocation happens outside stream capture in practice: the first (eager warm-up) call sizes it.
static float *box_gt_workspace(size_t floats) {
  static std::mutex mu;
  static float *buf = nullptr;
  static size_t cap = 0;
  std::lock_guard<std::mutex> lock(mu);
  if (floats > cap) {
    float *nb = nullptr;
    if (cudaMalloc(&nb, floats * sizeof(float)) != cudaSuccess) {
      ra::set_last_error("cudaMalloc(box_gt workspace)", cudaGetLastError());
      return nullptr;
    }
    buf = nb;  // the old, smaller buffer is leaked on purpose: an in-flight kernel or a captured graph may still use it
    cap = floats;
  }
  return buf;
}

extern "C" int ra_gt_box_f32(const float *y_gt, int B, int T, int H, int W, float padding_ratio, float min_padding,
                             float *top_left, float *bot_right, float *rect, float *box, float *area, void *stream) {
  if (!y_gt || !top_left || !bot_right || B < 0 || T < 1 || H < 1 || W < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  gt_box_kernel<<<B * T, 256, 0, ra::as_stream(stream)>>>(y_gt, H, W, padding_ratio, min_padding, top_left, bot_right,
                                                          rect, box, area);
  return ra::finish_launch("gt_box_kernel");
}

extern "C" size_t ra_pairwise_iou_workspace(int B, int N, int M, int HW) {
  IouParams p;
  if (iou_plan(N, M, HW, &p) != RA_OK) return 0;
  return (size_t)B * p.ctas_per_ex * (p.NB * kR) * (p.MB * kR);
}

extern "C" int ra_pairwise_iou_f32(const float *a, const float *b, const float *b_rect, int B, int N, int M, int H,
                                   int W, float hard_threshold, float *partial, float *iou, float *dice, void *stream) {
  if (!a || (!b && !b_rect) || !partial || !iou || B < 0 || N < 1 || M < 1 || H < 1 || W < 1)
    return RA_ERR_INVALID_ARG;
  if (((size_t)H * W) % 4 != 0 || hard_threshold >= 1.0f) return RA_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) & 15) != 0 || (b != nullptr && (reinterpret_cast<uintptr_t>(b) & 15) != 0))
    return RA_ERR_INVALID_ARG;  // 16-byte cp.async
  IouParams p;
  int rc = iou_plan(N, M, H * W, &p);
  if (rc != RA_OK) return rc;
  if (B == 0) return RA_OK;
  p.a = a;
  p.b = b;
  p.b_rect = b_rect;
  p.partial = partial;
  p.N = N;
  p.M = M;
  p.H = H;
  p.W = W;
  p.hard_thr = hard_threshold;
  const int NP = p.NB * kR, MP = p.MB * kR;
  size_t smem = (size_t)2 * (NP + MP) * (kKC + kPad) * sizeof(float);  // two stages
  const size_t red_bytes = (size_t)256 * kR * kR * sizeof(float);
  if (smem < red_bytes) smem = red_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pairwise_iou_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      ra::set_last_error("cudaFuncSetAttribute(pairwise_iou_kernel)", e);
      return RA_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = ra::as_stream(stream);
  const int threads = p.NB * p.MB * p.KS;
  pairwise_iou_kernel<<<dim3(p.ctas_per_ex, B), threads, smem, s>>>(p);
  rc = ra::finish_launch("pairwise_iou_kernel");
  if (rc != RA_OK) return rc;
  pairwise_iou_finalize_kernel<<<B, 256, 0, s>>>(partial, N, M, NP, MP, p.ctas_per_ex, (float)(H * W) * 1e-5f, iou,
                                                 dice);
  return ra::finish_launch("pairwise_iou_finalize_kernel");
}

extern "C" int ra_loss_block_f32(const float *iou_box, const float *match_box, const float *iou_soft,
                                 const float *match, const float *iou_hard, const float *dice_hard,
                                 const float *s_out, const float *s_gt, const float *gt_area, int B, int T,
                                 float loss_mix_ratio, float weight_decay_term, float segm_coeff, float *out,
                                 void *stream) {
  if (!iou_box || !match_box || !iou_soft || !match || !s_out || !s_gt || !gt_area || !out || B < 1 || T < 1)
    return RA_ERR_INVALID_ARG;
  loss_block_kernel<<<1, 256, 0, ra::as_stream(stream)>>>(iou_box, match_box, iou_soft, match, iou_hard, dice_hard,
                                                          s_out, s_gt, gt_area, B, T, loss_mix_ratio,
                                                          weight_decay_term, segm_coeff, out);
  return ra::finish_launch("loss_block_kernel");
}

extern "C" int ra_score_f32(const float *h, int Hd, const float *core, int Cd, const float *w, const float *bias,
                            int B, float *s_out, int s_stride, void *stream) {
  if (!h || !w || !bias || !s_out || Hd < 1 || Cd < 0 || (Cd > 0 && !core) || B < 0) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  const cudaError_t le = ra::launch_pdl(score_kernel, dim3(B), dim3(256), (size_t)0, ra::as_stream(stream), h, Hd, core, Cd,
                                        w, bias, s_out, s_stride);
  if (le != cudaSuccess) {
    ra::set_last_error("cudaLaunchKernelEx(score_kernel)", le);
    return RA_ERR_CUDA;
  }
  return ra::finish_launch("score_kernel");
}

extern "C" int ra_box_gt_step_f32(const float *attn_box, size_t box_bstride, const float *gt_rect, const float *y_gt,
                                  const float *noise, size_t noise_bstride, int B, int T, int H, int W, float *iou_t,
                                  int iou_bstride, float *grd_ws, float *canvas, void *stream) {
  if (!attn_box || !gt_rect || !y_gt || !iou_t || !grd_ws || !canvas || B < 0 || T < 1 || T > 64)
    return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  cudaStream_t s = ra::as_stream(stream);
  float *partial = box_gt_workspace((size_t)B * kBgChunks * (T + 1));
  if (partial == nullptr) return RA_ERR_CUDA;
  int rc = box_gt_iou_launch(attn_box, box_bstride, gt_rect, B, T, H, W, iou_t, iou_bstride, grd_ws, partial, s);
  if (rc != RA_OK) return rc;
  const int HW = H * W;
  int bx = (HW + 255) / 256;
  if (bx > 64) bx = 64;
  box_gt_canvas_kernel<<<dim3(bx, B), 256, 0, s>>>(grd_ws, y_gt, noise, noise_bstride, T, HW, canvas);
  return ra::finish_launch("box_gt_canvas_kernel");
}

extern "C" int ra_gt_attn_noise_f32(const float *rect_raw, const float *area, const float *pad, const float *shift,
                                    float min_padding, int B, int T, float *ctr, float *size, void *stream) {
  if (B < 0 || T < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!rect_raw || !area || !pad || !shift || !ctr || !size) return RA_ERR_INVALID_ARG;
  const int n = B * T;
  gt_attn_noise_kernel<<<(n + 127) / 128, 128, 0, ra::as_stream(stream)>>>(rect_raw, area, pad, shift, min_padding, n, ctr,
                                                                           size);
  return ra::finish_launch("gt_attn_noise_kernel");
}

extern "C" int ra_knob_greedy_box_f32(const float *attn_box, size_t box_bstride, const float *gt_rect, int B, int T, int H,
                                      int W, float *iou_t, int iou_bstride, float *grd, void *stream) {
  if (B < 0 || T < 1 || T > 64 || H < 1 || W < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!attn_box || !gt_rect || !iou_t || !grd) return RA_ERR_INVALID_ARG;
  float *partial = box_gt_workspace((size_t)B * kBgChunks * (T + 1));
  if (partial == nullptr) return RA_ERR_CUDA;
  return box_gt_iou_launch(attn_box, box_bstride, gt_rect, B, T, H, W, iou_t, iou_bstride, grd, partial,
                           ra::as_stream(stream));
}

extern "C" int ra_greedy_iou_box_f32(const float *box, const float *tl_gt, const float *br_gt, int B, int T, float *iou_t,
                                     int iou_bstride, float *grd, void *stream) {
  if (B < 0 || T < 1 || T > 64 || iou_bstride < T) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!box || !tl_gt || !br_gt || !iou_t || !grd) return RA_ERR_INVALID_ARG;
  greedy_iou_box_kernel<<<B, 64, 0, ra::as_stream(stream)>>>(box, tl_gt, br_gt, T, iou_t, iou_bstride, grd);
  return ra::finish_launch("greedy_iou_box_kernel");
}

extern "C" int ra_box_gt_canvas_f32(const float *grd, const float *y_gt, const float *noise, size_t noise_bstride, int B,
                                    int T, int H, int W, float *canvas, void *stream) {
  if (B < 0 || T < 1 || T > 64 || H < 1 || W < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!grd || !y_gt || !canvas) return RA_ERR_INVALID_ARG;
  const int HW = H * W;
  int bx = (HW + 255) / 256;
  if (bx > 64) bx = 64;
  box_gt_canvas_kernel<<<dim3(bx, B), 256, 0, ra::as_stream(stream)>>>(grd, y_gt, noise, noise_bstride, T, HW, canvas);
  return ra::finish_launch("box_gt_canvas_kernel");
}

extern "C" int ra_knob_mix_box_f32(float *box, const float *grd, const float *ctr_gt, const float *size_gt,
                                   const float *knob, int knob_stride, int B, int T, void *stream) {
  if (B < 0 || T < 1 || knob_stride < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!box || !grd || !ctr_gt || !size_gt || !knob) return RA_ERR_INVALID_ARG;
  knob_mix_box_kernel<<<(B + 63) / 64, 64, 0, ra::as_stream(stream)>>>(box, grd, ctr_gt, size_gt, knob, knob_stride, B, T);
  return ra::finish_launch("knob_mix_box_kernel");
}

extern "C" int ra_knob_canvas_f32(const float *grd, const float *y_gt, const float *noise, size_t noise_bstride,
                                  const float *knob, int knob_stride, const float *y_out, size_t out_bstride, int B, int T,
                                  int H, int W, float *canvas, void *stream) {
  if (B < 0 || T < 1 || T > 64 || H < 1 || W < 1 || knob_stride < 1) return RA_ERR_INVALID_ARG;
  if (B == 0) return RA_OK;
  if (!grd || !y_gt || !noise || !knob || !y_out || !canvas) return RA_ERR_INVALID_ARG;
  const int HW = H * W;
  int bx = (HW + 255) / 256;
  if (bx > 64) bx = 64;
  knob_canvas_kernel<<<dim3(bx, B), 256, 0, ra::as_stream(stream)>>>(grd, y_gt, noise, noise_bstride, knob, knob_stride, y_out,
                                                                     out_bstride, T, HW, canvas);
  return ra::finish_launch("knob_canvas_kernel");
}

/*
 * rec_attend_b200.h — C ABI of librecattend_b200.so (hand-written sm_100a CUDA).
 *
 * The drop-in boundary for the recurrent-attention decoding hot path of
 * renmengye/rec-attend-public.  Plain pointers and sizes only (no torch / TF types).
 * Conventions:
 *   - every `ra_*` entry point returns 0 on success or a negative RA_ERR_* code; nothing
 *     here aborts the process (the reference's LOG(FATAL) paths become status codes);
 *   - unless the name ends in `_host`, all data pointers are DEVICE pointers to contiguous
 *     row-major fp32 (int32 where noted) owned by the caller; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = the legacy default stream) and is
 *     asynchronous; the entry points are stateless and re-entrant;
 *   - `_host` variants take HOST pointers, do their own H2D/D2H copies and synchronise
 *     before returning (this is what a TF-style CPU custom-op shim binds);
 *   - layouts follow the reference: images/features NHWC, mask stacks [B,T,H,W], index 0
 *     of any 2-vector is y (rows), index 1 is x (cols).
 * Reference citations are relative to the reference repository root.
 */
#ifndef REC_ATTEND_B200_H_
#define REC_ATTEND_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RA_OK 0
#define RA_ERR_INVALID_ARG (-1)
#define RA_ERR_CUDA (-2)
#define RA_ERR_UNSUPPORTED (-3)
#define RA_ERR_NO_DEVICE (-4)

/* per-example status bits written by ra_hungarian_f32 */
#define RA_HUNG_ST_OUTER_CAP 1 /* 1000 outer rounds reached: unfinished matching returned,
                                  like the reference's LOG(ERROR) path (hungarian.cc:362-377) */

#define RA_HUNG_MAX_N 64 /* nx, ny <= 64 (vertex sets are 64-bit masks) */

/* Library / device probe. ra_version() = 10000*major + 100*minor + patch. */
int ra_version(void);
int ra_device_count(void);
/* Last CUDA error string seen by this thread inside the library ("" if none). */
const char *ra_last_error(void);
/* Number of kernels this library has launched in this process (monotonic). */
unsigned long long ra_launch_count(void);

/* --------------------------------------------------------------------------------------
 * Hungarian matching — replaces the TF custom op
 *   REGISTER_OP("Hungarian").Input("weights: float").Output("matching: float")
 *       .Output("cover_x: float").Output("cover_y: float")            (hungarian.cc:26-30)
 * registered for DEVICE_CPU only (hungarian.cc:540) and called as
 *   hungarian_module.hungarian(W)[0]                                   (modellib.py:406).
 * W [B,nx,ny] (B = 1 for the rank-2 form, hungarian.cc:58-60) -> M [B,nx,ny] in {0,1},
 * cover_x [B,nx] (op shape [B,nx,1]), cover_y [B,ny] (op shape [B,1,ny]).
 * status [B] int32 (may be NULL) receives RA_HUNG_ST_* bits.
 * One warp per example; results are bit-identical to the reference algorithm wherever the
 * reference terminates (its 1000-pop BFS abort, hungarian.cc:124-127, does not exist here).
 * -------------------------------------------------------------------------------------- */
int ra_hungarian_f32(const float *W, int B, int nx, int ny, float *M, float *cover_x, float *cover_y,
                     int32_t *status, void *stream);
int ra_hungarian_f32_host(const float *W, int B, int nx, int ny, float *M, float *cover_x, float *cover_y,
                          int32_t *status);

/* f_segm_match (modellib.py:382-415): W = floor(iou*mx*my*1e6+0.5)/1e6 + 1e-5, Hungarian,
 * mask again.  iou [B,T,T] (rows = outputs, cols = ground truth), s_gt [B,T] -> match [B,T,T].
 * weights_out (may be NULL) receives the fp32 matrix handed to the matcher. */
int ra_segm_match_f32(const float *iou, const float *s_gt, int B, int T, float *match, float *weights_out,
                      int32_t *status, void *stream);

/* --------------------------------------------------------------------------------------
 * Convolution block — nnlib.py:214-255 (run_cnn) and :339-402 (run_dcnn) in eval mode:
 *   y = act( (conv(x) ) * scale + shift ), optional 2x2 max-pool; BN(eval)+bias are folded
 *   into per-channel scale/shift by the caller (nnlib.py:119, eps 1e-3).
 * x1 [B,Hin,Win,C1] (+ optional x2 [B,Hin,Win,C2], channel-concatenated after x1, the
 * skip connection of nnlib.py:362-367), w [3,3,C1+C2,Cout] in HWIO order (for transposed
 * layers the caller passes the flipped/transposed filter), 3x3, SAME.
 * upsample = 1: plain conv (stride 1).  upsample = 2: conv2d_transpose stride 2
 * (nnlib.py:372-376): the input is zero-inserted to 2Hin x 2Win and the window starts 2
 * before the output pixel.  pool in {1,2}; relu in {0,1}.
 * y [B,Hout/pool,Wout/pool,Cout].  add_to (may be NULL) [B,Hout,Wout,Cout] is added to the
 * raw convolution before scale/shift (used to split the first controller layer into a
 * static part and a per-step canvas part, full_model.py:640-663).
 * -------------------------------------------------------------------------------------- */
int ra_conv3x3_f32(const float *x1, int C1, const float *x2, int C2, const float *w, const float *scale,
                   const float *shift, const float *add_to, int B, int Hin, int Win, int Cout, int upsample,
                   int pool, int relu, float *y, void *stream);

/* First controller layer, per-step half (full_model.py:640-663): the layer is linear in its
 * input channels and only the canvas channel changes between decode steps, so
 *   y = pool(relu((pre + conv3x3(canvas; w)) * scale + shift))
 * with pre [B,H,W,C0] = the raw convolution of the step-invariant channels (computed once per
 * forward), canvas [B,H,W], w [3,3,1,C0].  C0 % 4 == 0, pool in {1,2}.  y [B,H/pool,W/pool,C0]. */
int ra_canvas_conv_f32(const float *pre, const float *canvas, const float *w, const float *scale,
                       const float *shift, int B, int H, int W, int C0, int pool, int relu, float *y, void *stream);

/* --------------------------------------------------------------------------------------
 * The same convolution block on the tcgen05 tensor cores (kind::tf32, accumulators in TMEM)
 * with the 3xTF32 split (hi*hi + hi*lo + lo*hi, ~2^-22 relative) that the 1e-3 parity bar
 * needs; persistent warp-specialised pipeline (csrc/conv_umma.cu).  Arguments as
 * ra_conv3x3_f32 except: no add_to, and the filter is pre-packed by the caller into the
 * kernel's shared-memory image
 *   wpack [n_split][n_chunks][9 taps][KC/4][2*NPc][4]
 * element [s][ch][tap][c4][r][j] = part(w[tap][ch*KC + 4*c4 + j][s*NPc + (r mod NPc)]), part = hi
 * (w rounded to the nearest tf32) for r < NPc and lo = w - hi for r >= NPc, zero padded,
 * with KC, NPc, n_split, n_chunks given by ra_conv3x3_umma_plan for the layer's (Cin, Cout,
 * un-pooled output size, pool, batch).  When the plan reports rowstack = 1 (narrow layers: the three kx taps of a
 * filter row share one read of the A operand) the image is
 *   wpack [n_split][n_chunks][3 ky][KC/4][6*NPc][4]
 * with rows [hi kx0 | hi kx1 | hi kx2 | lo kx0 | lo kx1 | lo kx2] (NPc each).
 * Supported: Cout <= 256, even output width.
 * -------------------------------------------------------------------------------------- */
int ra_conv3x3_umma_plan(int Cin, int Cout, int Hout, int Wout, int pool, int B, int *KC, int *NPc, int *n_split,
                         int *n_chunks, int *rowstack);
/* The same for a layer whose input is the channel concatenation [x1 (C1) | x2 (C2)] (C2 = 0: ra_conv3x3_umma_plan with
 * Cin = C1): the fp16 operand split (ra_conv3x3_umma_set_f16) feeds the kernel in 16-channel TMA boxes, which must not
 * straddle the two inputs - such layers keep the 3xTF32 plan.  ra_conv3x3_umma_f32 plans with the (C1, C2) it is given:
 * pack the filter image for THIS plan when C2 > 0. */
int ra_conv3x3_umma_plan_split(int C1, int C2, int Cout, int Hout, int Wout, int pool, int B, int *KC, int *NPc,
                               int *n_split, int *n_chunks, int *rowstack);
/* Diagnostics: info[19] = KC, NPc, n_split, n_chunks, TH, TW, n_mt, stages, merged, w_resident, grid, smem_bytes,
 * acc_cols, stage_bytes, w_res_bytes, slots_alloc, ksplit, nbuf, rowstack of the tile plan. */
int ra_conv3x3_umma_plan_info(int Cin, int Cout, int Hout, int Wout, int pool, int B, int *info);
/* Filter image of a plan whose layout flags (`rowstack` of ra_conv3x3_umma_plan) have bit 1 set (fp16 hi / lo operand
 * split, RA_UMMA_F16): v [rows][KC/4][NPc][4] fp32 = the filter in the operand order of one half of the tf32 image
 * (rows = n_split * n_chunks * 9), out [rows][KC/8][2*NPc][8 halves] (the same number of bytes as v): rows 0..NPc-1 of a
 * plane = fp16(w) of 8 consecutive input channels, rows NPc..2NPc-1 = fp16((w - hi) * 2^11).  Device pointers. */
/* Operand format of the tile plans made from now on: 1 (default) = fp16 hi / lo split on the layers whose plan allows it
 * (<= 64 output channels per CTA, 16-channel chunks, TMA feed in 16-channel boxes), 0 = 3xTF32 everywhere, 2 = like 1,
 * forcing 16-channel chunks.  Both formats meet the same error bound (11 + 11 significand bits per operand).
 * + 4 (modes 5, 6): the tensor core also folds the 2^11-scaled correction half of every accumulator into its main half at
 * the end of a tile (tcgen05.mma with the correction columns as TMEM A operand): the epilogue reads half the columns.
 * Initial value: environment variable RA_UMMA_F16.  Returns the previous mode (mode < 0: query only).  Filter images are
 * packed per plan: set this before a model packs its filters. */
int ra_conv3x3_umma_set_f16(int mode);
int ra_umma_pack_f16(const float *v, long long rows, int KC, int NPc, float *out, void *stream);
/* A CHAIN of conv layers in ONE launch (the 6 + 7 layers of the patch network of a decode step, full_model.py:792-807;
 * controller layers 1-7, :663): a persistent grid of one CTA per SM runs the layers back to back with a grid-wide
 * barrier between them instead of paying CTA start-up, TMEM allocation and pipeline fill / drain per launch.
 *  ra_conv3x3_umma_chain_prepare: layers[n_layers <= 16] with the arguments of ra_conv3x3_umma_f32 (HOST array; all
 *    data pointers are device pointers) -> builds every layer's tile plan and tensor maps into desc_host (HOST memory
 *    of ra_conv3x3_umma_chain_desc_bytes(n_layers) bytes: the blob travels as the kernel's argument);
 *    *grid_out / *smem_out = launch geometry for the run call.  Layer l+1 may read what layer l wrote.
 *  ra_conv3x3_umma_chain_run: one launch on `stream`; counter_dev = 16 device uint32 (one barrier counter per layer
 *    boundary, zeroed here).  Results are identical to n_layers calls of ra_conv3x3_umma_f32. */
typedef struct {
  const float *x1;
  int C1;
  const float *x2;
  int C2;
  const float *wpack, *scale, *shift;
  int B, Hin, Win, Cout, upsample, pool, relu;
  float *y;
} ra_conv_layer_t;
size_t ra_conv3x3_umma_chain_desc_bytes(int n_layers);
int ra_conv3x3_umma_chain_prepare(const ra_conv_layer_t *layers, int n_layers, void *desc_host, int *grid_out,
                                  size_t *smem_out);
int ra_conv3x3_umma_chain_run(const void *desc_host, int n_layers, int grid, size_t smem_bytes,
                              unsigned int *counter_dev, void *stream);
/* Diagnostics: device buffer of 8 int64 per CTA (148 CTAs max) receiving clock64() stamps of the pipeline
 * phases of the next ra_conv3x3_umma_f32 launches; NULL switches it off. */
int ra_debug_conv_timeline(long long *device_buf);
int ra_conv3x3_umma_f32(const float *x1, int C1, const float *x2, int C2, const float *wpack, const float *scale,
                        const float *shift, int B, int Hin, int Win, int Cout, int upsample, int pool, int relu,
                        float *y, void *stream);

/* --------------------------------------------------------------------------------------
 * Controller step — full_model.py:668-725 / box_model.py:417-470: 5 soft-attention glimpse
 * read-outs of the controller feature map, LSTM (nnlib.py:637-649, state reset to 0),
 * glimpse MLP + softmax (nnlib.py:476-493), controller head, box-parameter maths
 * (modellib.py:752-856).  One launch per decode step for the whole batch.
 * feat [B,P,Cf] (P = h'w', p = y*w'+x).  Weights in the reference's layouts:
 * lstm_wx [4][Cf][Hd], lstm_wh [4][Hd][Hd], lstm_b [4][Hd] in gate order i,f,o,u;
 * gmlp_w0 [Hd,Hd], gmlp_b0 [Hd], gmlp_w1 [Hd,P], gmlp_b1 [P]; cmlp_w [Hd,9], cmlp_b [9].
 * Outputs: h_out [B,Hd]; ctrl_out [B,9]; glimpse_map [B,n_iter,P];
 * box [B,RA_BOX_STRIDE] = see RA_BOX_* offsets.
 * flags: RA_CTRL_* bits.
 * -------------------------------------------------------------------------------------- */
#define RA_CTRL_SQUASH 1      /* squash_ctrl_params (full_model.py:695-697) */
#define RA_CTRL_FIXED_VAR 2   /* fixed_var (:702-703) */
#define RA_CTRL_DYNAMIC_VAR 4 /* dynamic_var (:708-709) */
#define RA_CTRL_FIXED_GAMMA 8 /* fixed_gamma (:711-713) */

#define RA_BOX_CTR_Y 0
#define RA_BOX_CTR_X 1
#define RA_BOX_SIZE_Y 2
#define RA_BOX_SIZE_X 3
#define RA_BOX_LGVAR_Y 4
#define RA_BOX_LGVAR_X 5
#define RA_BOX_GAMMA_ATTN 6 /* exp(lg_gamma_attn) */
#define RA_BOX_GAMMA_BOX 7  /* exp(lg_gamma_box) */
#define RA_BOX_GAMMA_Y 8    /* exp(lg_gamma_y) */
#define RA_BOX_TL_Y 9       /* ctr - size/2 (modellib.py:850-852) */
#define RA_BOX_TL_X 10
#define RA_BOX_BR_Y 11      /* ctr + size/2 */
#define RA_BOX_BR_X 12
#define RA_BOX_STRIDE 16

int ra_controller_step_f32(const float *feat, int B, int P, int Cf, int Hd, int n_iter, const float *lstm_wx,
                           const float *lstm_wh, const float *lstm_b, const float *gmlp_w0, const float *gmlp_b0,
                           const float *gmlp_w1, const float *gmlp_b1, const float *cmlp_w, const float *cmlp_b,
                           int inp_height, int inp_width, int filter_height, int filter_width, int flags,
                           float *h_out, float *ctrl_out, float *glimpse_map, float *box, void *stream);

/* --------------------------------------------------------------------------------------
 * Gaussian attention filters — modellib.get_gaussian_filter (modellib.py:581-612), built
 * once per decode step from `box` (RA_BOX_* layout) as tap-major profiles
 *   fy [B,F,H], fx [B,F,W]:  f[b,t,l] = exp(-(l-mu_t)^2 / (2 e^{lg_var})) / sqrt(2 pi e^{lg_var}),
 *   mu_t = ctr + (size+1)/F * (t - (F-1)/2)
 * (the reference's [B,L,F] filter transposed) plus the support band of every tap:
 * band [B,2,F,2] int32 = (lo, hi) inclusive pixel range per axis (0 = y, 1 = x) and tap;
 * entries whose exponent is below -30 are stored as exact zeros.
 * -------------------------------------------------------------------------------------- */
int ra_gaussian_filters_f32(const float *box, int B, int H, int W, int F, float *fy, float *fx, int32_t *band,
                            void *stream);

/* --------------------------------------------------------------------------------------
 * Gaussian glimpse — modellib.extract_patch (modellib.py:615-641) as called at
 * full_model.py:788-789:
 *   x_patch[b,i,j,d] = gamma_attn[b] * sum_{y,x} fy[b,i,y] X[b,y,x,d] fx[b,j,x]
 * X is given as the two channel groups that the reference concatenates
 * (full_model.py:645-660): xs [B,H,W,Cs] step-invariant channels and canvas [B,H,W]
 * (either may be absent: Cs = 0 / canvas = NULL).  chan_map [Cs+1] int32 gives, for every
 * source channel (xs 0..Cs-1, then the canvas), its channel position in the reference's
 * concat order.  x_patch [B,F,F,patch_cstride] with patch_cstride >= Cs+1: channels Cs+1..
 * patch_cstride-1 are written as zeros (padding the patch to a multiple of 4 channels lets
 * ra_conv3x3_umma_f32 read it with TMA).  tmp: caller workspace of B*F*W*(Cs+1) floats.
 * W must be a multiple of 4.
 * -------------------------------------------------------------------------------------- */
int ra_gaussian_extract_f32(const float *xs, int Cs, const float *canvas, const int32_t *chan_map,
                            const float *box, const float *fy, const float *fx, const int32_t *band, int B, int H,
                            int W, int F, float *tmp, float *x_patch, int patch_cstride, void *stream);

/* --------------------------------------------------------------------------------------
 * Paste-back — extract_patch(., Fy^T, Fx^T, 1) as used for the attention box
 * (full_model.py:738-741) and the mask (:810-818), fused with the canvas update (:845):
 *   attn_box[b,y,x] = sigmoid(gamma_box * sum_ij fy[i,y] fx[j,x] - 5)
 *   y_out[b,y,x]    = sigmoid(gamma_y * sum_ij fy[i,y] P[b,i,j] fx[j,x] - 5) [* (1-canvas)]
 *   canvas          = max(canvas, y_out)
 * patch [B,F,F] (may be NULL: attention box only, box_model.py:484-487); attn_box may be
 * NULL (mask only).  attn_box / y_out are written at [b*out_bstride + y*W + x]
 * (out_bstride = T*H*W lets the caller write step t of a [B,T,H,W] stack in place);
 * canvas [B,H,W] is updated in place.  band [B,2,F,2] (from ra_gaussian_filters_f32; may be
 * NULL) lets tiles outside every tap's support skip the filter arithmetic: the result is the
 * same (entries outside a band are exact zeros), only faster.
 * -------------------------------------------------------------------------------------- */
int ra_paste_back_f32(const float *patch, const float *box, const float *fy, const float *fx, const int32_t *band,
                      int B, int H, int W, int F, int disable_overwrite, float *attn_box, float *y_out,
                      size_t out_bstride, float *canvas, void *stream);

/* Score head, full_model.py:821-822: s = sigmoid(w . concat(h, core) + b); h [B,Hd],
 * core [B,Cd] (may be NULL with Cd = 0: box_model.py:508-511), w [Hd+Cd], s_out at
 * s_out[b*s_stride]. */
int ra_score_f32(const float *h, int Hd, const float *core, int Cd, const float *w, const float *bias, int B,
                 float *s_out, int s_stride, void *stream);

/* --------------------------------------------------------------------------------------
 * Ground-truth boxes — modellib.get_gt_box (modellib.py:663-701) with get_idx_map /
 * get_filled_box_idx (:704-749).  y_gt [B,T,H,W] -> top_left [B,T,2], bot_right [B,T,2]
 * (after the empty-mask fix-up, :697-699), rect [B,T,4] = (tl_y, tl_x, br_y, br_x) BEFORE
 * the fix-up (what the filled box is drawn from, :694; may be NULL), box [B,T,H,W] (filled
 * rectangle, may be NULL), area [B,T] = sum of y_gt (may be NULL).
 * -------------------------------------------------------------------------------------- */
int ra_gt_box_f32(const float *y_gt, int B, int T, int H, int W, float padding_ratio, float min_padding,
                  float *top_left, float *bot_right, float *rect, float *box, float *area, void *stream);

/* --------------------------------------------------------------------------------------
 * Pairwise soft IoU / DICE — modellib.f_iou(pairwise=True) (modellib.py:138-153) with
 * f_inter / f_union (:104-114, eps 1e-5 added per pixel) and f_dice (:81-97).
 * a [B,N,HW], b [B,M,HW] -> iou [B,N,M]; dice (may be NULL) [B,N,M].
 * hard_threshold > 0 binarises a on the fly (a > thr), full_model.py:1063.
 * b_rect (may be NULL) [B,M,4] = (tl_y,tl_x,br_y,br_x) as written by ra_gt_box_f32: when
 * given, b is not read and is taken as the rectangle indicator tl <= idx <= br (the filled
 * GT boxes of full_model.py:931).  H*W must be a multiple of 4; N, M <= 39.
 * partial: caller workspace of ra_pairwise_iou_workspace(B,N,M,HW) floats.
 * -------------------------------------------------------------------------------------- */
size_t ra_pairwise_iou_workspace(int B, int N, int M, int HW);
int ra_pairwise_iou_f32(const float *a, const float *b, const float *b_rect, int B, int N, int M, int H, int W,
                        float hard_threshold, float *partial, float *iou, float *dice, void *stream);

/* --------------------------------------------------------------------------------------
 * Loss block — full_model.py:942-1081 (fixed_order = False, loss fns = 'iou'):
 * scalars out[RA_LOSS_*] from iou matrices, matchings, s_out, s_gt, GT areas.
 * -------------------------------------------------------------------------------------- */
#define RA_LOSS_BOX 0
#define RA_LOSS_SEGM 1
#define RA_LOSS_CONF 2
#define RA_LOSS_IOU_SOFT 3
#define RA_LOSS_IOU_HARD 4
#define RA_LOSS_WT_COV_SOFT 5
#define RA_LOSS_UNWT_COV_SOFT 6
#define RA_LOSS_WT_COV_HARD 7
#define RA_LOSS_UNWT_COV_HARD 8
#define RA_LOSS_DICE 9
#define RA_LOSS_COUNT_ACC 10
#define RA_LOSS_DIC 11
#define RA_LOSS_DIC_ABS 12
#define RA_LOSS_TOTAL 13 /* box + segm_coeff*segm + mix*conf + weight_decay_term (segm_coeff = 0: box_model.py:558-629) */
#define RA_LOSS_COUNT 16
int ra_loss_block_f32(const float *iou_box, const float *match_box, const float *iou_soft, const float *match,
                      const float *iou_hard, const float *dice_hard, const float *s_out, const float *s_gt,
                      const float *gt_area, int B, int T, float loss_mix_ratio, float weight_decay_term,
                      float segm_coeff, float *out, void *stream);

/* Greedy GT interaction of box_model.py:484-504 for one decode step: iou_t[b,m] =
 * f_inter(attn_box_t, box_gt_m)/f_union (rectangle form of box_gt), grd = one-hot of the row
 * max with ties sharing 1/k (modellib.py:366-379), canvas = max(canvas,
 * sum_m grd*y_gt_m*(1-noise)).  attn_box_t at attn_box[b*box_bstride + p]; noise may be NULL;
 * grd_ws: caller workspace of B*T floats (receives the greedy match). */
int ra_box_gt_step_f32(const float *attn_box, size_t box_bstride, const float *gt_rect, const float *y_gt,
                       const float *noise, size_t noise_bstride, int B, int T, int H, int W, float *iou_t,
                       int iou_bstride, float *grd_ws, float *canvas, void *stream);

/* The canvas half of ra_box_gt_step_f32 alone (box_model.py:497-503), for callers that computed grd themselves
 * (use_iou_box): canvas = max(canvas, sum_m grd[b,m]*y_gt[b,m]*(1-noise)); noise may be NULL. */
int ra_box_gt_canvas_f32(const float *grd, const float *y_gt, const float *noise, size_t noise_bstride, int B, int T,
                         int H, int W, float *canvas, void *stream);

/* --------------------------------------------------------------------------------------
 * Scheduled sampling ("knob", training mode only) — full_model.py:561-625,744-785,826-845, with
 * the random draws as inputs:
 *  ra_gt_attn_noise_f32: noisy GT attention boxes (modellib.get_gt_attn with per-object padding
 *    ratio pad [B,T] and centre shift [B,T,2]) from the raw mask extrema rect_raw [B,T,4]
 *    (ra_gt_box_f32 with zero padding) and area [B,T] -> ctr, size [B,T,2].
 *  ra_knob_greedy_box_f32: iou_t[b,m] = f_inter(attn_box_t, box_gt_m) / f_union (rectangle form)
 *    and grd = f_greedy_match(iou_t, 0) (modellib.py:366-379), :756-759.
 *  ra_greedy_iou_box_f32: the opt['use_iou_box'] form of the same match (full_model.py:750-754,
 *    box_model.py:487-491): iou_t[b,m] = modellib.f_iou_box (modellib.py:206-238, coordinate IoU, strict overlap
 *    test, no eps) of the box record's top-left / bottom-right against tl_gt, br_gt [B,T,2] = the top_left /
 *    bot_right OUTPUTS of ra_gt_box_f32 (after the empty-mask fix of modellib.py:697-699; not `rect`, which is the
 *    rectangle of the filled box), then the greedy match.  A NaN score (0/0 for two zero-area boxes) makes the
 *    whole grd row NaN.
 *  ra_knob_mix_box_f32: box record (RA_BOX_*) of step t <- knob ? matched noisy GT box : itself, :760-776.
 *  ra_knob_canvas_f32: canvas = max(canvas, knob ? (sum_m grd*y_gt)*(1-noise) : y_out_t), :826-845.
 * knob points at the [B] switches of this step (element b at knob[b*knob_stride]).
 * -------------------------------------------------------------------------------------- */
int ra_gt_attn_noise_f32(const float *rect_raw, const float *area, const float *pad, const float *shift,
                         float min_padding, int B, int T, float *ctr, float *size, void *stream);
int ra_knob_greedy_box_f32(const float *attn_box, size_t box_bstride, const float *gt_rect, int B, int T, int H, int W,
                           float *iou_t, int iou_bstride, float *grd, void *stream);
int ra_greedy_iou_box_f32(const float *box, const float *tl_gt, const float *br_gt, int B, int T, float *iou_t,
                          int iou_bstride, float *grd, void *stream);
int ra_knob_mix_box_f32(float *box, const float *grd, const float *ctr_gt, const float *size_gt, const float *knob,
                        int knob_stride, int B, int T, void *stream);
int ra_knob_canvas_f32(const float *grd, const float *y_gt, const float *noise, size_t noise_bstride, const float *knob,
                       int knob_stride, const float *y_out, size_t out_bstride, int B, int T, int H, int W,
                       float *canvas, void *stream);

/* out[b,h,w,:] = concat(a[..,:Ca], b[..,:Cb], c[..,:Cc]) over npix = B*H*W pixels — the
 * step-invariant part of the input stack of full_model.py:640-661 (Cb, Cc may be 0). */
int ra_concat_channels_f32(const float *a, int Ca, const float *b, int Cb, const float *c, int Cc, size_t npix,
                           float *out, void *stream);

/* --------------------------------------------------------------------------------------
 * Backward of one training-mode conv block  y = pool(relu(bn_batch(conv3x3(x [,x2]) + b)))  — the
 * gradients TensorFlow's autodiff derives for nnlib.py:229-253 / :372-400 with phase_train = True
 * (optimizer.compute_gradients, full_model.py:1049).  First building block of the backward pass; fp32
 * CUDA-core kernels (csrc/conv_bwd.cu).
 *  ra_bn_train_block_bwd_f32: raw [B,H,W,C] (conv output incl. bias, kept from the forward), the batch
 *    statistics mean / var [C] (ra_bn_train_block_f32's batch_mean / batch_var), dy [B,H/pool,W/pool,C]
 *    -> d_raw [B,H,W,C], dgamma [C], dbeta [C]; max-pool routes to the first maximum of the window, relu'(0) = 0,
 *    the batch moments are differentiated (tf.nn.moments).  ws: ra_bn_train_block_bwd_workspace() bytes.
 *  ra_conv3x3_bwd_weight_f32: dw [3,3,C1+C2,Cout] = gradient of the CONV-FORM filter ra_conv3x3_f32 consumes
 *    (for upsample = 2 the input is zero-inserted, Z[2y,2x] = x[y,x], and a tap reads Z[Y+ky-2, X+kx-2] — the
 *    forward kernel's convention, which makes it TensorFlow's SAME conv2d_transpose), db [Cout] (may be NULL) = column sums of
 *    d_out [B,Hin*up,Win*up,Cout].  ws: ra_conv3x3_bwd_weight_workspace() bytes.
 *  ra_filter_flip_transpose_f32: out[ky,kx,co,ci] = w[2-ky,2-kx,ci,co].  The data gradient of the block is
 *    ra_conv3x3_f32(d_out, flip_transpose(w)) (then ra_subsample2_f32 with off = 1 for upsample = 2); the same map converts
 *    between the conv-form filter of a transposed-conv layer and TensorFlow's [kh,kw,Cout,Cin] layout.
 *  ra_subsample2_f32: dst [B,H,W,C] = src [B,2H,2W,C] at positions (2y+off, 2x+off), off in {0,1}.
 * -------------------------------------------------------------------------------------- */
size_t ra_bn_train_block_bwd_workspace(int B, int H, int W, int C, int pool);
int ra_bn_train_block_bwd_f32(const float *raw, const float *dy, const float *gamma, const float *beta,
                              const float *mean, const float *var, int B, int H, int W, int C, int pool, int relu,
                              float eps, void *ws, float *d_raw, float *dgamma, float *dbeta, void *stream);
size_t ra_conv3x3_bwd_weight_workspace(int B, int Hin, int Win, int Cin, int Cout, int upsample);
int ra_conv3x3_bwd_weight_f32(const float *x1, int C1, const float *x2, int C2, const float *d_out, int B, int Hin,
                              int Win, int Cout, int upsample, void *ws, float *dw, float *db, void *stream);
/* Grouped / shared-input forms used by the training step, where the T decode steps are processed as ONE batch of
 * G = T groups of B examples (no gradient flows between steps: the canvas is behind tf.stop_gradient,
 * full_model.py:846-848):
 *  ra_bn_train_block_bwd_grouped_f32: raw [G,B,H,W,C], dy [G,B,H/p,W/p,C], gamma / beta / mean / var [G,C] (one BN copy
 *    per group = per decode step, nnlib.py:212-254) -> d_raw [G,B,H,W,C], dgamma / dbeta [G,C].
 *  ra_conv3x3_bwd_weight_ex_f32: x1_bmod > 0: x1 holds x1_bmod examples shared by every group (example n reads
 *    x1[n % x1_bmod]) — the step-invariant input channels of the first controller layer. */
size_t ra_bn_train_block_bwd_grouped_workspace(int G, int B, int H, int W, int C, int pool);
int ra_bn_train_block_bwd_grouped_f32(const float *raw, const float *dy, const float *gamma, const float *beta,
                                      const float *mean, const float *var, int G, int B, int H, int W, int C, int pool,
                                      int relu, float eps, void *ws, float *d_raw, float *dgamma, float *dbeta,
                                      void *stream);
int ra_conv3x3_bwd_weight_ex_f32(const float *x1, int C1, int x1_bmod, const float *x2, int C2, const float *d_out,
                                 int B, int Hin, int Win, int Cout, int upsample, void *ws, float *dw, float *db,
                                 void *stream);
int ra_filter_flip_transpose_f32(const float *w, int Ci, int Co, float *out, void *stream);
int ra_subsample2_f32(const float *src, int B, int H, int W, int C, int off, float *dst, void *stream);

/* --------------------------------------------------------------------------------------
 * Backward of the matching loss block (full_model.py:942-1034) to the model outputs — what TensorFlow's
 * autodiff hands to y_out / attn_box / s_out; the matchings are constants (modellib.py:11).
 *  ra_iou_loss_bwd_f32: da [B,N,H,W] (batch stride a_bstride, like a) = gradient of
 *    -(scale/B) sum_b (1/max(1, sum match[b])) sum_nm match[b,n,m] * f_iou(a_n, g_m)  (modellib.py:104-155, eps per
 *    pixel) with g = b_masks [B,M,H,W] (segmentation loss, :983-1012) or the filled rectangles b_rect [B,M,4]
 *    (box loss, :931-973; ra_gt_box_f32's `rect`) — exactly one of the two non-NULL.  match [B,N,M].
 *    ws: ra_iou_loss_bwd_workspace() bytes.
 *  ra_conf_loss_bwd_f32: ds [B,T] = gradient of scale * f_conf_loss(s_out, match) (modellib.py:316-339 with the
 *    cumulative min / max of :39-68): flows to the arg-min of every prefix and the arg-max of every suffix.
 * -------------------------------------------------------------------------------------- */
size_t ra_iou_loss_bwd_workspace(int B, int N, int M);
int ra_iou_loss_bwd_f32(const float *a, size_t a_bstride, const float *b_masks, const float *b_rect,
                        const float *match, int B, int N, int M, int H, int W, float scale, void *ws, float *da,
                        void *stream);
int ra_conf_loss_bwd_f32(const float *s_out, const float *match, int B, int T, int M, float scale, float *ds,
                         void *stream);

/* --------------------------------------------------------------------------------------
 * Backward of the paste-back and of the Gaussian filters — TensorFlow's autodiff of
 * out = sigmoid(gamma * (Fy P Fx^T) - 5) (full_model.py:738-741 with P = ones, :810-814 with the mask-head patch;
 * modellib.py:615-641) and of modellib.get_gaussian_filter (modellib.py:581-612).
 *  ra_paste_back_bwd_f32: d_out, out [B,H,W] (batch stride out_bstride), patch [B,F,F] or NULL (= ones),
 *    fy [B,F,H], fx [B,F,W] tap-major, gamma at gamma[b*gamma_stride] (the exp'd gain of the box record)
 *    -> d_patch [B,F,F] (NULL iff patch is NULL), d_fy [B,F,H], d_fx [B,F,W] (added to the existing contents when
 *    accumulate != 0: the filters also feed the glimpse and the attention box), d_gamma [B] = dL/dgamma
 *    (dL/d lg_gamma = d_gamma * gamma).  ws: ra_paste_back_bwd_workspace() bytes.  disable_overwrite is not handled.
 *  ra_gaussian_filters_bwd_f32: d_fy, d_fx -> d_box [B,6] = (d_ctr_y, d_ctr_x, d_size_y, d_size_x, d_lg_var_y,
 *    d_lg_var_x) for the box record `box` the filters were built from.
 * -------------------------------------------------------------------------------------- */
 /* ra_gaussian_extract_bwd_f32: backward of the glimpse x_patch = gamma_attn * Fy^T X Fx (ra_gaussian_extract_f32,
 *    full_model.py:788-789): d_patch, x_patch [B,F,F,patch_cstride] (patch channel order), X as xs / canvas /
 *    chan_map like the forward -> d_fy, d_fx (accumulated when accumulate != 0), d_gamma [B] = dL/dgamma_attn.
 *    No gradient goes to X (data; the canvas is behind tf.stop_gradient, full_model.py:846-848).
 *    ws: ra_gaussian_extract_bwd_workspace(B, W, F, Cs + 1 or Cs) bytes. */
size_t ra_gaussian_extract_bwd_workspace(int B, int W, int F, int D);
int ra_gaussian_extract_bwd_f32(const float *xs, int Cs, const float *canvas, const int32_t *chan_map, const float *fy,
                                const float *fx, const float *gamma, int gamma_stride, const float *d_patch,
                                const float *x_patch, int patch_cstride, int B, int H, int W, int F, int accumulate,
                                void *ws, float *d_fy, float *d_fx, float *d_gamma, void *stream);
size_t ra_paste_back_bwd_workspace(int B, int H, int W, int F);
int ra_paste_back_bwd_f32(const float *d_out, const float *out, size_t out_bstride, const float *patch, const float *fy,
                          const float *fx, const float *gamma, int gamma_stride, int B, int H, int W, int F,
                          int accumulate, void *ws, float *d_patch, float *d_fy, float *d_fx, float *d_gamma,
                          void *stream);
int ra_gaussian_filters_bwd_f32(const float *box, const float *fy, const float *fx, const float *d_fy, const float *d_fx,
                                int B, int H, int W, int F, float *d_box, void *stream);
/* Step-batched forms (example n = t * n_inner + b of a [T, n_inner, ...] stack):
 *  ra_paste_back_bwd_ex_f32: d_out / out of example n at (n % n_inner) * out_bstride + (n / n_inner) * out_ostride
 *    (a [B,T,H,W] stack read in [T,B] order: out_bstride = T*H*W, out_ostride = H*W).
 *  ra_gaussian_extract_bwd_ex_f32: xs_bmod > 0: xs holds xs_bmod examples shared by every step (n % xs_bmod).
 *  box (both; may be NULL) = the [B,RA_BOX_STRIDE] records the filters were built from by ra_gaussian_filters_f32:
 *    the kernels then only walk the support bands of the taps (entries outside are exact zeros) - the work is
 *    proportional to the attended window instead of the image. */
int ra_paste_back_bwd_ex_f32(const float *d_out, const float *out, size_t out_bstride, int n_inner, size_t out_ostride,
                             const float *patch, const float *fy, const float *fx, const float *gamma, int gamma_stride,
                             const float *box, int B, int H, int W, int F, int accumulate, void *ws, float *d_patch,
                             float *d_fy, float *d_fx, float *d_gamma, void *stream);
int ra_gaussian_extract_bwd_ex_f32(const float *xs, int Cs, int xs_bmod, const float *canvas, const int32_t *chan_map,
                                   const float *fy, const float *fx, const float *gamma, int gamma_stride,
                                   const float *box, const float *d_patch, const float *x_patch, int patch_cstride,
                                   int B, int H, int W, int F, int accumulate, void *ws, float *d_fy, float *d_fx,
                                   float *d_gamma, void *stream);

/* --------------------------------------------------------------------------------------
 * Backward of the controller (full_model.py:668-725) — TensorFlow's autodiff of the soft read-out, the LSTM, the
 * glimpse-MLP softmax (n_iter iterations), the linear head and the box-parameter maths.  Weight layouts as in
 * ra_controller_step_f32 (gate order i, f, o, u).
 *  ra_controller_tape_f32: re-runs the forward of one decode step and records per (example, iteration) the record
 *    [map_k | glimpse_k | h_{k-1} | c_{k-1} | gates i,f,o,u | c_k | h_k | a1_k | map_{k+1}] (field offsets and record
 *    size from ra_controller_tape_layout; ra_controller_tape_floats() floats per example); also h_out, ctrl_out.
 *  ra_controller_head_bwd_f32: d_box [B,6] (ra_gaussian_filters_bwd_f32's layout, summed over the filters' consumers)
 *    and d_gamma3 [B,3] = dL/dgamma (attn, box, y) -> d_ctrl_out [B,9], honouring the RA_CTRL_* flags.
 *  ra_controller_bwd_f32: d_h [B,Hd] (may be NULL; the score head's share) and d_ctrl_out -> d_feat [B,P,Cf] and the
 *    pre-activation deltas dG [B,n_iter,4,Hd], dA1 [B,n_iter,Hd], dLog [B,n_iter,P].
 *  ra_outer_sum_f32: dW [n_in,n_out] = sum_r A[r*a_stride + :]^T D[r*d_stride + :], db [n_out] (may be NULL) = column
 *    sums of D — every dense-layer weight gradient of the controller (rows = (example, iteration)).
 * -------------------------------------------------------------------------------------- */
size_t ra_controller_tape_floats(int P, int Cf, int Hd, int n_iter);
int ra_controller_tape_layout(int P, int Cf, int Hd, int *offsets);
int ra_controller_tape_f32(const float *feat, int B, int P, int Cf, int Hd, int n_iter, const float *lstm_wx,
                           const float *lstm_wh, const float *lstm_b, const float *gmlp_w0, const float *gmlp_b0,
                           const float *gmlp_w1, const float *gmlp_b1, const float *cmlp_w, const float *cmlp_b,
                           float *tape, float *h_out, float *ctrl_out, void *stream);
int ra_controller_head_bwd_f32(const float *ctrl_out, const float *box, const float *d_box, const float *d_gamma3, int B,
                               int inp_height, int inp_width, int flags, float *d_ctrl_out, void *stream);
int ra_controller_bwd_f32(const float *feat, int B, int P, int Cf, int Hd, int n_iter, const float *lstm_wx,
                          const float *lstm_wh, const float *gmlp_w0, const float *gmlp_w1, const float *cmlp_w,
                          const float *tape, const float *d_h, const float *d_ctrl_out, float *d_feat, float *dG,
                          float *dA1, float *dLog, void *stream);
int ra_outer_sum_f32(const float *A, size_t a_stride, int n_in, const float *D, size_t d_stride, int n_out, int R,
                     float *dW, float *db, void *stream);
/* The same sum with the rows split over up to 32 chunks of CTAs (a [64,256] gradient over 3200 rows is 4 CTAs
 * otherwise); ws = ra_outer_sum_workspace(n_in, n_out, R) bytes of scratch for the per-chunk partials (0: no split
 * needed, ws may be NULL), which are added in chunk order - deterministic. */
size_t ra_outer_sum_workspace(int n_in, int n_out, int R);
int ra_outer_sum_ex_f32(const float *A, size_t a_stride, int n_in, const float *D, size_t d_stride, int n_out, int R,
                        void *ws, float *dW, float *db, void *stream);

/* --------------------------------------------------------------------------------------
 * Foreground / orientation FCN head + loss block — fg_model.py:174-236 (the
 * FCN's conv stack runs on ra_conv3x3_umma_f32 / ra_conv3x3_f32).  logits [npix, nsc+nori] = last
 * DCNN layer (no BN, no activation, fg_model.py:121,148):
 *   y_out [npix,nsc] = sigmoid (nsc == 1) or softmax over the nsc semantic channels (:183-188);
 *   d_out [npix,nori] = softmax over the orientation channels (:176-181; nori may be 0);
 *   y_hard [npix,nsc] (may be NULL) = y_out > 0.5, or the one-hot of the per-pixel maximum (:201-204).
 * With y_gt [npix,nsc] (and d_gt [npix,nori] when nori > 0) also out[RA_FG_*]: f_iou_all soft / hard
 * (modellib.py:171-181; over channels 1.. when nsc > 1), the per-pixel mean of f_bce / f_ce
 * (modellib.py:418-427), foreground_loss = that mean if loss_is_bce else -iou_soft (:223-226), the
 * masked orientation cross-entropy and accuracy (:229-239) and loss = foreground_loss + orientation_ce.
 * ws: device scratch of ra_fg_head_workspace() bytes (needed with y_gt).  y_gt == NULL: inference only.
 * -------------------------------------------------------------------------------------- */
#define RA_FG_IOU_SOFT 0
#define RA_FG_IOU_HARD 1
#define RA_FG_SEGLOSS 2
#define RA_FG_FOREGROUND_LOSS 3
#define RA_FG_ORIENTATION_CE 4
#define RA_FG_ORIENTATION_ACC 5
#define RA_FG_LOSS 6
size_t ra_fg_head_workspace(void);
int ra_fg_head_f32(const float *logits, size_t npix, int nsc, int nori, const float *y_gt, const float *d_gt,
                   int loss_is_bce, float *y_out, float *d_out, float *y_hard, float *out, void *ws, void *stream);
/* Gradient of loss = foreground_loss [+ orientation_ce] (fg_model.py:223-250) at the logits: d_logits [npix, nsc+nori].
 * y_out / d_out = the head's outputs, ws = the workspace ra_fg_head_f32 filled on the same inputs (it holds the global
 * sums the IoU loss and the masked cross-entropy divide by).  Training mode of the FCN. */
int ra_fg_head_bwd_f32(const float *y_out, const float *d_out, size_t npix, int nsc, int nori, const float *y_gt,
                       const float *d_gt, int loss_is_bce, const void *ws, float *d_logits, void *stream);

/* --------------------------------------------------------------------------------------
 * Instance-label post-processing — utils/postprocess.py as chained by
 * full_model_eval.py:112-125: apply_confidence (postprocess.py:17-31),
 * apply_one_label (:34-55), apply_threshold (:6-14), mask_foreground (:146-155),
 * remove_tiny (:106-143).  y_out [B,T,H,W], s_out [B,T], fg [B,H,W] (may be NULL):
 *   v[t]     = y_out[b,t,p] * s_out[b,t]                       (fp32)
 *   k        = first arg max_t v[t]                             (np.argmax)
 *   on       = (double)v[k] > thresh                            (Python-float threshold)
 *   label[p] = k + 1 if on (and fg[p] != 0) else 0              int32 [B,H,W]
 *   area[b,t] = sum_p on * fg[p] over pixels with k == t        (remove_tiny's y.sum)
 *   remove_tiny > 0: instances with area <= remove_tiny are erased from the label map
 *   conf[b,t] = (s_out > 0.5) * (area > remove_tiny or remove_tiny == 0)
 *   y_hard (may be NULL) [B,T,H,W] = the reference's dense float masks (label == t+1) * fg.
 * workspace: ra_postprocess_workspace(B, T) bytes of device memory.  cv2-based steps
 * (upsample + bilateral filter, morph) are outside this entry point.
 * -------------------------------------------------------------------------------------- */
size_t ra_postprocess_workspace(int B, int T);
int ra_postprocess_f32(const float *y_out, const float *s_out, const float *fg, int B, int T, int H, int W,
                       double thresh, float remove_tiny, void *workspace, int32_t *label, float *y_hard, float *conf,
                       float *area, void *stream);

/* --------------------------------------------------------------------------------------
 * In-graph augmentation — image_ops.random_transformation (image_ops.py:9-113) with the
 * random draws as arguments: dst = transpose?(flip?(crop(pad(src, padding), off_y, off_x))).
 * src / dst [N,H,W,C] (NHWC images: N = B; mask stacks [B,T,H,W]: N = B*T, C = 1); off_* in
 * [0, 2*padding] (tf.random_uniform(maxval = 2*padding)); transpose needs H == W.
 * phase_train = False is off_y = off_x = padding with no flips (the identity).
 * -------------------------------------------------------------------------------------- */
int ra_random_transformation_f32(const float *src, size_t N, int H, int W, int C, int padding, int off_y, int off_x,
                                 int vflip, int hflip, int transpose, float *dst, void *stream);

/* dst[i] = (float)src[i] over n bytes: ground-truth masks handed over as uint8 {0,1} (the datasets store PNG masks,
 * data_api/ins_seg_dataset.py:169-172; the reference converts on the host before feeding, runner.py:91-96) are
 * expanded to the fp32 [B,T,H,W] stack of full_model.py:165-194 on the device. */
int ra_u8_to_f32(const uint8_t *src, size_t n, float *dst, void *stream);

/* --------------------------------------------------------------------------------------
 * Training-mode batch normalisation of one conv block — nnlib.batch_norm with
 * phase_train = True (nnlib.py:65-128) + ReLU + max-pool (nnlib.py:229-253):
 *   mean, var = moments of x over (B,H,W) (biased);  ema -= (1-decay)*(ema - batch)  (decay 0.9)
 *   y = pool(relu(x*inv + (beta - mean*inv))),  inv = gamma * rsqrt(var + eps)       (eps 1e-3)
 * x [B,H,W,C] = raw convolution output incl. bias (e.g. ra_conv3x3_umma_f32 with scale 1,
 * shift = bias, relu 0, pool 1); C <= 256 (float4 path when C % 4 == 0); ema_* updated in place
 * (may be NULL), batch_* [C] out (may be NULL); workspace: ra_bn_train_workspace() floats.
 * -------------------------------------------------------------------------------------- */
size_t ra_bn_train_workspace(int B, int H, int W, int C);
int ra_bn_train_block_f32(const float *x, int B, int H, int W, int C, const float *gamma, const float *beta, float eps,
                          float decay, int pool, int relu, float *workspace, float *ema_mean, float *ema_var,
                          float *batch_mean, float *batch_var, float *y, void *stream);

/* --------------------------------------------------------------------------------------
 * Optimiser block — full_model.py:1039-1057 / box_model.py:635-652 on one flat fp32 bucket of
 * n trainable elements (the buffer the gradient all-reduce of SURVEY §8e runs on):
 *   g  = grad * grad_scale (1/world after a SUM all-reduce) + wd[i] * param   (wd may be NULL;
 *        wd[i] = weight_decay on conv/mlp/lstm weight matrices, 0 elsewhere, nnlib.py:59-61)
 *   g  = clip(g, -clip, clip)                  (tf.clip_by_value per element; clip <= 0: off)
 *   m += (g - m)(1-beta1);  v += (g*g - v)(1-beta2)
 *   param -= lr_t * m / (sqrt(v) + eps),  lr_t = lr * sqrt(1-beta2^t) / (1-beta1^t), t = step_t >= 1
 * (TensorFlow-0.12 AdamOptimizer; the reference uses eps = 1e-7).  lr is the already decayed
 * learning rate base * decay^floor(step / steps_per_decay) (tf.train.exponential_decay, staircase).
 * -------------------------------------------------------------------------------------- */
int ra_adam_step_f32(float *param, const float *grad, float *m, float *v, const float *wd, size_t n,
                     float grad_scale, float lr, float beta1, float beta2, float eps, float clip, int step_t,
                     void *stream);

/* --------------------------------------------------------------------------------------
 * Pairwise soft + hard IoU on the tcgen05 tensor cores, one pass over the masks (csrc/iou_umma.cu) —
 * modellib.f_iou(a, g, pairwise=True) (modellib.py:104-155) for the matching (full_model.py:983) together with
 * f_iou / f_dice of the thresholded masks a > hard_thr (full_model.py:1063-1081, modellib.py:71-101).
 * a, g [B,T,H,W] contiguous (g = ground-truth masks), N = M = T <= 127, H*W % 32 == 0, 16-byte aligned;
 * iou_soft / iou_hard / dice_hard [B,T,T] (each may be NULL).  K-major GEMM over the pixels: TMA (SWIZZLE_128B) ->
 * hi / lo tf32 split of a + thresholded copy in shared memory -> tcgen05.mma kind::tf32, accumulators in TMEM, row /
 * column sums through a row of ones; partial block diagonals per pixel slice summed in a fixed order.
 * ws: ra_pairwise_iou_umma_workspace() bytes (0: shape not supported - use ra_pairwise_iou_f32).
 * -------------------------------------------------------------------------------------- */
size_t ra_pairwise_iou_umma_workspace(int B, int T, int H, int W);
int ra_pairwise_iou_umma_f32(const float *a, const float *g, int B, int T, int H, int W, float hard_thr, void *ws,
                             float *iou_soft, float *iou_hard, float *dice_hard, void *stream);

/* --------------------------------------------------------------------------------------
 * Training-step glue (csrc/train.cu) — what sits between the per-block gradients and the optimiser in
 * `sess.run([loss, train_step])` (runner.py:98-105; full_model.py:1039-1057, box_model.py:635-652).
 *  ra_param_gather_f32: flat trainable bucket -> every device-side weight image in one launch.  `codes[i]` describes
 *    destination element i of the concatenation of nseg destination tensors (seg_start [nseg] = first element of each
 *    segment, seg_dst [nseg] = device pointers; both arrays in DEVICE memory): 0 = constant zero (padding), else
 *    ((flat index + 1) << 2) | kind with kind 0 = the value, 1 = hi (value rounded to the nearest tf32), 2 = lo =
 *    value - hi (the two operand images of the 3xTF32 tcgen05 convolution).
 *  ra_param_scatter_f32: the inverse for gradients (kind ignored, code 0 skipped): flat[index] = src element.
 *  ra_bn_fold_f32: eval-mode BN + bias folded per (step, channel): scale = gamma / sqrt(ema_var + eps),
 *    shift = beta - ema_mean * scale + bias * scale (nnlib.py:113-119); gamma .. ema_var, scale, shift [T,C], bias [C].
 *  ra_weight_decay_f32: out[0] = sum_i wd[i] * param[i]^2 / 2 (nnlib.py:59-61); ws: ra_weight_decay_workspace() bytes.
 *  ra_sum_groups_f32: dst [n] = sum_g src [G,n].   ra_add_f32: dst += src.
 *  ra_split_channels_f32: src [npix, C1+C2] -> dst1 [npix,C1] (may be NULL), dst2 [npix,C2] (may be NULL; added to
 *    when accumulate2 != 0): the input gradient of a layer fed by concat(prev, skip) (nnlib.py:362-367).
 *  ra_score_bwd_f32: s = sigmoid([h, core] w + b) (full_model.py:821-822): s_out, d_s [B,T]; rows n = t*B + b:
 *    dpre [T*B], d_h [T*B,Hd] = dpre * w[:Hd], d_core [T*B,Cd] (+)= dpre * w[Hd:] (Cd may be 0: box_model.py:508-511).
 *  ra_knob_box_bwd_f32: scheduled sampling (full_model.py:760-776): d_box [T*B,6] = d_mixed with its centre / size
 *    entries scaled by 1 - knob_box[b,t], plus d_pre (may be NULL; the pre-mix filters' share).
 *  ra_iou_box_coord_bwd_f32: d_box [T*B,6] += gradient of scale * box_loss through modellib.f_iou_box
 *    (modellib.py:206-238) of the controller's own box record `box` [T*B,RA_BOX_STRIDE] against tl_gt / br_gt [B,T,2]
 *    with the constant matching match_box [B,T,T] (opt['use_iou_box'], full_model.py:750-754,926-929).
 * -------------------------------------------------------------------------------------- */
 /*  ra_wt_cov_f32: modellib.f_weighted_coverage (modellib.py:268-302) of iou [B,N,M] with GT sizes `area` [B,M] (or, rect
 *    != NULL, the pixel counts of the filled GT rectangles [B,M,4] on an H x W grid) -> cov[0]; coeff [B,N,M] (may be
 *    NULL) = the weight of every GT at its first arg-max output, the gradient coefficients ra_iou_loss_bwd_f32 takes in
 *    place of the matching when segm_loss_fn / box_loss_fn == 'wt_cov' (full_model.py:967,1013-1014).
 *  ra_loss_select_f32: scal[RA_LOSS_BOX] = -box_cov[0], scal[RA_LOSS_SEGM] = -segm_cov[0] (each optional), total rebuilt
 *    without the weight-decay term. */
int ra_wt_cov_f32(const float *iou, const float *area, const float *rect, int H, int W, int B, int N, int M, float *cov,
                  float *coeff, void *stream);
int ra_loss_select_f32(float *scal, const float *box_cov, const float *segm_cov, float mix, float segm_coeff,
                       void *stream);
int ra_param_gather_f32(const float *flat, const int32_t *codes, const long long *seg_start, float *const *seg_dst,
                        int nseg, long long total, void *stream);
int ra_param_scatter_f32(float *flat, const int32_t *codes, const long long *seg_start, const float *const *seg_src,
                         int nseg, long long total, void *stream);
int ra_bn_fold_f32(const float *gamma, const float *beta, const float *ema_mean, const float *ema_var, const float *bias,
                   int T, int C, float eps, float *scale, float *shift, void *stream);
size_t ra_weight_decay_workspace(void);
int ra_weight_decay_f32(const float *param, const float *wd, size_t n, void *ws, float *out, void *stream);
int ra_sum_groups_f32(const float *src, int G, size_t n, float *dst, void *stream);
int ra_add_f32(float *dst, const float *src, size_t n, void *stream);
int ra_split_channels_f32(const float *src, size_t npix, int C1, int C2, float *dst1, float *dst2, int accumulate2,
                          void *stream);
int ra_score_bwd_f32(const float *s_out, const float *d_s, int B, int T, const float *w, int Hd, int Cd, float *dpre,
                     float *d_h, float *d_core, int accumulate_core, void *stream);
int ra_knob_box_bwd_f32(const float *d_mixed, const float *d_pre, const float *knob_box, int B, int T, float *d_box,
                        void *stream);
int ra_iou_box_coord_bwd_f32(const float *box, const float *tl_gt, const float *br_gt, const float *match_box, int B,
                             int T, float scale, float *d_box, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* REC_ATTEND_B200_H_ */

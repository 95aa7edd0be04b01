"""Operator-level host mirror of the reference's modellib.py / nnlib.py for the hot path.

Same names and argument meaning as the reference functions they replace; every call goes
through the C ABI of librecattend_b200.so on the current CUDA stream (torch is used only for
device memory and streams).  Inputs must be fp32 CUDA tensors; there is no CPU fallback.
"""
import ctypes

import torch

from . import _lib

_c = ctypes


def _stream():
  return _c.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
  if t is None:
    return _c.c_void_p(0)
  return _c.c_void_p(t.data_ptr())


def _chk(*tensors):
  for t in tensors:
    if t is None:
      continue
    if not t.is_cuda:
      raise _lib.RecAttendError('rec_attend_b200 ops need CUDA tensors (no CPU fallback)')
    if not t.is_contiguous():
      raise _lib.RecAttendError('rec_attend_b200 ops need contiguous tensors')
    if t.dtype not in (torch.float32, torch.int32):
      raise _lib.RecAttendError('rec_attend_b200 ops are fp32 (int32 for index tensors)')


# ----------------------------------------------------------------------------- matching
def hungarian(weights):
  """``hungarian_module.hungarian(W)`` (hungarian.cc:26-30, modellib.py:406).
  W [B,nx,ny] or [nx,ny] -> (matching, cover_x [..,nx,1], cover_y [..,1,ny]); the extra
  attribute-free 4th return is the int32 status word per example."""
  _chk(weights)
  if weights.dim() == 2:
    w3 = weights.unsqueeze(0)
  elif weights.dim() == 3:
    w3 = weights
  else:
    raise ValueError('Must have dimension 3 or 2.')  # hungarian.cc:61-64
  B, nx, ny = w3.shape
  M = torch.empty_like(w3)
  cx = torch.empty((B, nx, 1), device=w3.device, dtype=torch.float32)
  cy = torch.empty((B, 1, ny), device=w3.device, dtype=torch.float32)
  st = torch.zeros((B,), device=w3.device, dtype=torch.int32)
  _lib.call('ra_hungarian_f32', _p(w3), B, nx, ny, _p(M), _p(cx), _p(cy), _p(st), _stream())
  if weights.dim() == 2:
    return M[0], cx[0], cy[0], st
  return M, cx, cy, st


def f_segm_match(iou, s_gt, return_weights=False):
  """modellib.f_segm_match (modellib.py:382-415): iou [B,T,T], s_gt [B,T] -> match [B,T,T]."""
  _chk(iou, s_gt)
  B, T, T2 = iou.shape
  assert T == T2 and tuple(s_gt.shape) == (B, T)
  match = torch.empty_like(iou)
  w = torch.empty_like(iou) if return_weights else None
  st = torch.zeros((B,), device=iou.device, dtype=torch.int32)
  _lib.call('ra_segm_match_f32', _p(iou), _p(s_gt), B, T, _p(match), _p(w), _p(st), _stream())
  if return_weights:
    return match, w, st
  return match


# ----------------------------------------------------------------------------- conv blocks
def conv3x3_block(x, w, scale, shift, pool=1, relu=True, x2=None, upsample=1, add_to=None, out=None):
  """One fused layer of nnlib.run_cnn / run_dcnn in eval mode (nnlib.py:229-253, :372-400).
  x [B,H,W,C1] (+x2 [B,H,W,C2]), w [3,3,C1+C2,Cout] HWIO conv-form filter."""
  _chk(x, w, scale, shift, x2, add_to, out)
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  Cout = w.shape[3]
  assert tuple(w.shape) == (3, 3, C1 + C2, Cout), (tuple(w.shape), C1, C2)
  Ho, Wo = H * upsample // pool, W * upsample // pool
  if out is None:
    out = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
  assert tuple(out.shape) == (B, Ho, Wo, Cout)
  _lib.call('ra_conv3x3_f32', _p(x), C1, _p(x2), C2, _p(w), _p(scale), _p(shift), _p(add_to), B, H, W, Cout,
            upsample, pool, 1 if relu else 0, _p(out), _stream())
  return out


def canvas_conv(pre, canvas, w, scale, shift, pool=1, relu=True, out=None):
  """Per-step half of the first controller layer: pool(relu((pre + conv(canvas)) * scale + shift))."""
  _chk(pre, canvas, w, scale, shift, out)
  B, H, W, C0 = pre.shape
  if out is None:
    out = torch.empty((B, H // pool, W // pool, C0), device=pre.device, dtype=torch.float32)
  _lib.call('ra_canvas_conv_f32', _p(pre), _p(canvas), _p(w), _p(scale), _p(shift), B, H, W, C0, pool,
            1 if relu else 0, _p(out), _stream())
  return out


def umma_plan(Cin, Cout, Hout, Wout, pool, B, C2=0):
  """(KC, NPc, n_split, n_chunks, layout flags) of the tcgen05 conv kernel for one layer shape and batch size; the
  filter image must be packed for exactly this plan (umma_filter_image / pack_umma_weights / params.pack_umma).
  C2: channels of the second input when the layer reads a concatenation [x1 | x2] (Cin = C1 + C2).  Flags: bit 0
  row-stacked taps, bit 1 fp16 hi / lo split."""
  kc, npc, nsp, nch, rs = _c.c_int(0), _c.c_int(0), _c.c_int(0), _c.c_int(0), _c.c_int(0)
  _lib.call('ra_conv3x3_umma_plan_split', Cin - C2, C2, Cout, Hout, Wout, pool, B, _c.byref(kc), _c.byref(npc),
            _c.byref(nsp), _c.byref(nch), _c.byref(rs))
  return kc.value, npc.value, nsp.value, nch.value, rs.value


def umma_plan_info(Cin, Cout, Hout, Wout, pool, B):
  """The full tile plan of the tcgen05 conv kernel as a dict (diagnostics)."""
  info = (_c.c_int * 19)()
  _lib.call('ra_conv3x3_umma_plan_info', Cin, Cout, Hout, Wout, pool, B, info)
  keys = ['KC', 'NPc', 'n_split', 'n_chunks', 'TH', 'TW', 'n_mt', 'stages', 'merged', 'w_resident', 'grid',
          'smem_bytes', 'acc_cols', 'stage_bytes', 'w_res_bytes', 'slots_alloc', 'ksplit', 'nbuf', 'rowstack']
  return dict(zip(keys, list(info)))


def pack_umma_weights(w_hwio, KC, NPc, n_split, rowstack=0):
  """HWIO conv filter [3,3,Cin,Cout] (numpy) -> the kernel's shared-memory image
  [n_split][n_chunks][9][KC/4][2*NPc][4]: rows 0..NPc-1 = hi (w rounded to the nearest tf32),
  rows NPc..2NPc-1 = lo = w - hi (exact); rowstack: [n_split][n_chunks][3][KC/4][6*NPc][4] with rows
  [hi kx0 | hi kx1 | hi kx2 | lo kx0 | lo kx1 | lo kx2] (values only; params.pack_umma also carries the indices)."""
  import numpy as np
  from . import params as PM
  if rowstack & 2:
    raise _lib.RecAttendError('fp16 hi / lo plan (RA_UMMA_F16): the filter image is made on the device, use '
                              'umma_filter_image')
  w = np.asarray(w_hwio, np.float32)
  return PM.pack_umma(PM.WI(w, np.zeros(w.shape, np.int64)), KC, NPc, n_split, rowstack).val


def umma_set_f16(mode):
  """Operand format of the tile plans made from now on (0 = 3xTF32, 1 / 2 = fp16 hi / lo split, + 4 = correction half
  folded on the tensor core); returns the previous mode; mode < 0 only queries."""
  return _lib.lib().ra_conv3x3_umma_set_f16(int(mode))


def umma_pack_f16(src, KC, NPc, out=None):
  """Device tensor src [..., KC/4, NPc, 4] fp32 (params.umma_f16_source) -> the fp16 hi / lo filter image of an
  RA_UMMA_F16 plan (same byte size, carried as a float32 tensor of raw bits)."""
  _chk(src, out)
  if out is None:
    out = torch.empty_like(src)
  rows = src.numel() // (KC * NPc)
  _lib.call('ra_umma_pack_f16', _p(src), rows, KC, NPc, _p(out), _stream())
  return out


def umma_filter_image(w_hwio, KC, NPc, n_split, flags, device):
  """HWIO filter (numpy) -> the device-side filter image for a plan with layout `flags` (bit 0 row-stacked taps, bit 1
  fp16 hi / lo split) as given by umma_plan."""
  import numpy as np
  from . import params as PM
  w = np.asarray(w_hwio, np.float32)
  if flags & 2:
    src = PM.umma_f16_source(PM.WI(w, np.zeros(w.shape, np.int64)), KC, NPc, n_split).val
    return umma_pack_f16(torch.from_numpy(src).to(device), KC, NPc)
  return torch.from_numpy(pack_umma_weights(w, KC, NPc, n_split, flags & 1)).to(device)


def conv3x3_block_umma(x, wpack, Cout, scale, shift, pool=1, relu=True, x2=None, upsample=1, out=None):
  """conv3x3_block on the tensor cores (3xTF32); wpack from pack_umma_weights for this layer's plan."""
  _chk(x, wpack, scale, shift, x2, out)
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  Ho, Wo = H * upsample // pool, W * upsample // pool
  if out is None:
    out = torch.empty((B, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
  assert tuple(out.shape) == (B, Ho, Wo, Cout)
  _lib.call('ra_conv3x3_umma_f32', _p(x), C1, _p(x2), C2, _p(wpack), _p(scale), _p(shift), B, H, W, Cout, upsample,
            pool, 1 if relu else 0, _p(out), _stream())
  return out


class _ConvLayerSpec(_c.Structure):
  """ra_conv_layer_t of include/rec_attend_b200.h."""
  _fields_ = [('x1', _c.c_void_p), ('C1', _c.c_int), ('x2', _c.c_void_p), ('C2', _c.c_int), ('wpack', _c.c_void_p),
              ('scale', _c.c_void_p), ('shift', _c.c_void_p), ('B', _c.c_int), ('Hin', _c.c_int), ('Win', _c.c_int),
              ('Cout', _c.c_int), ('upsample', _c.c_int), ('pool', _c.c_int), ('relu', _c.c_int), ('y', _c.c_void_p)]


class ConvChain(object):
  """Several conv3x3_block_umma layers run by ONE launch of a persistent grid (ra_conv3x3_umma_chain_*): the patch
  network of a decode step, controller layers 1-7.  `layers`: list of dicts with the arguments of conv3x3_block_umma
  (x, wpack, Cout, scale, shift, pool, relu, x2, upsample, out - `out` is required).  The tile plans and tensor maps are
  built here (host side; construct it once, outside the timed path); `run()` is one launch on the current
  stream.  The tensors are kept alive by the object."""

  def __init__(self, layers):
    n = len(layers)
    arr = (_ConvLayerSpec * n)()
    self._keep = []
    dev = layers[0]['x'].device
    for i, L in enumerate(layers):
      x, x2, out = L['x'], L.get('x2'), L['out']
      _chk(x, L['wpack'], L['scale'], L['shift'], x2, out)
      B, H, W, C1 = x.shape
      up, pool = L.get('upsample', 1), L.get('pool', 1)
      assert tuple(out.shape) == (B, H * up // pool, W * up // pool, L['Cout']), (i, tuple(out.shape))
      arr[i] = _ConvLayerSpec(x.data_ptr(), C1, 0 if x2 is None else x2.data_ptr(), 0 if x2 is None else x2.shape[3],
                              L['wpack'].data_ptr(), L['scale'].data_ptr(), L['shift'].data_ptr(), B, H, W, L['Cout'],
                              up, pool, 1 if L.get('relu', True) else 0, out.data_ptr())
      self._keep.extend([x, x2, out, L['wpack'], L['scale'], L['shift']])
    self.n = n
    nbytes = int(_lib.lib().ra_conv3x3_umma_chain_desc_bytes(n))
    if nbytes == 0:
      raise _lib.RecAttendError('conv chain: unsupported number of layers {}'.format(n))
    self.desc = _c.create_string_buffer(nbytes)  # host blob: plans + tensor maps, passed as the kernel argument
    self.counter = torch.zeros(16, device=dev, dtype=torch.int32)  # one barrier counter per layer boundary
    grid, smem = _c.c_int(0), _c.c_size_t(0)
    _lib.call('ra_conv3x3_umma_chain_prepare', _c.byref(arr), n, self.desc, _c.byref(grid), _c.byref(smem))
    self.grid, self.smem = grid.value, smem.value

  def run(self):
    _lib.call('ra_conv3x3_umma_chain_run', self.desc, self.n, self.grid, self.smem, _p(self.counter), _stream())


def concat_channels(a, b=None, c=None, out=None):
  _chk(a, b, c, out)
  B, H, W, Ca = a.shape
  Cb = 0 if b is None else b.shape[3]
  Cc = 0 if c is None else c.shape[3]
  if out is None:
    out = torch.empty((B, H, W, Ca + Cb + Cc), device=a.device, dtype=torch.float32)
  _lib.call('ra_concat_channels_f32', _p(a), Ca, _p(b), Cb, _p(c), Cc, B * H * W, _p(out), _stream())
  return out


# ----------------------------------------------------------------------------- controller
def controller_step(feat, lstm_wx, lstm_wh, lstm_b, gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1, cmlp_w, cmlp_b, inp_height,
                    inp_width, filter_height, filter_width, flags, n_iter=5, h_out=None, ctrl_out=None,
                    glimpse_map=None, box=None):
  """full_model.py:668-725: feat [B,P,Cf] -> (h [B,256], ctrl_out [B,9], glimpse_map [B,n_iter,P], box [B,12])."""
  _chk(feat, lstm_wx, lstm_wh, lstm_b, gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1, cmlp_w, cmlp_b)
  B, P, Cf = feat.shape
  Hd = lstm_wh.shape[-1]
  dev = feat.device
  if h_out is None:
    h_out = torch.empty((B, Hd), device=dev, dtype=torch.float32)
  if ctrl_out is None:
    ctrl_out = torch.empty((B, 9), device=dev, dtype=torch.float32)
  if glimpse_map is None:
    glimpse_map = torch.empty((B, n_iter, P), device=dev, dtype=torch.float32)
  if box is None:
    box = torch.empty((B, _lib.BOX_STRIDE), device=dev, dtype=torch.float32)
  _chk(h_out, ctrl_out, glimpse_map, box)
  _lib.call('ra_controller_step_f32', _p(feat), B, P, Cf, Hd, n_iter, _p(lstm_wx), _p(lstm_wh), _p(lstm_b),
            _p(gmlp_w0), _p(gmlp_b0), _p(gmlp_w1), _p(gmlp_b1), _p(cmlp_w), _p(cmlp_b), inp_height, inp_width,
            filter_height, filter_width, flags, _p(h_out), _p(ctrl_out), _p(glimpse_map), _p(box), _stream())
  return h_out, ctrl_out, glimpse_map, box


# ----------------------------------------------------------------------------- attention
def get_gaussian_filter(box, H, W, F, fy=None, fx=None, band=None):
  """modellib.get_gaussian_filter (modellib.py:581-612) for both axes at once.
  Returns tap-major fy [B,F,H], fx [B,F,W] and the tap support bands [B,2,F,2] int32."""
  _chk(box, fy, fx, band)
  B = box.shape[0]
  dev = box.device
  if fy is None:
    fy = torch.empty((B, F, H), device=dev, dtype=torch.float32)
  if fx is None:
    fx = torch.empty((B, F, W), device=dev, dtype=torch.float32)
  if band is None:
    band = torch.empty((B, 2, F, 2), device=dev, dtype=torch.int32)
  _lib.call('ra_gaussian_filters_f32', _p(box), B, H, W, F, _p(fy), _p(fx), _p(band), _stream())
  return fy, fx, band


def extract_patch(xs, canvas, chan_map, box, fy, fx, band, tmp=None, out=None):
  """modellib.extract_patch (modellib.py:615-641) with the attention gain of
  full_model.py:788: xs [B,H,W,Cs] + canvas [B,H,W] -> x_patch [B,F,F,Cs+1].  `out` may have a wider last
  dimension (channel stride); the extra channels are zero-filled."""
  _chk(xs, canvas, chan_map, box, fy, fx, band, tmp, out)
  B, F, H = fy.shape
  W = fx.shape[2]
  Cs = 0 if xs is None else xs.shape[3]
  D = Cs + (1 if canvas is not None else 0)
  dev = fy.device
  if tmp is None:
    tmp = torch.empty((B * F * W * D,), device=dev, dtype=torch.float32)
  if out is None:
    out = torch.empty((B, F, F, D), device=dev, dtype=torch.float32)
  assert tmp.numel() >= B * F * W * D and tuple(out.shape[:3]) == (B, F, F) and out.shape[3] >= D
  assert out.is_contiguous()
  _lib.call('ra_gaussian_extract_f32', _p(xs), Cs, _p(canvas), _p(chan_map), _p(box), _p(fy), _p(fx), _p(band), B,
            H, W, F, _p(tmp), _p(out), out.shape[3], _stream())
  return out


def paste_back(patch, box, fy, fx, canvas, attn_box=None, y_out=None, out_bstride=None, disable_overwrite=False,
               band=None):
  """full_model.py:738-741 (attention box), :810-818 (mask) and :845 (canvas = max) fused.
  attn_box / y_out may be views of step t of a [B,T,H,W] stack (pass out_bstride = T*H*W).
  `band` (from get_gaussian_filter) enables the constant fast path for tiles outside the box."""
  B, F, H = fy.shape
  W = fx.shape[2]
  if out_bstride is None:
    out_bstride = H * W
  _chk(patch, box, fy, fx, canvas, band)
  _lib.call('ra_paste_back_f32', _p(patch), _p(box), _p(fy), _p(fx), _p(band), B, H, W, F, 1 if disable_overwrite else 0,
            _p(attn_box), _p(y_out), out_bstride, _p(canvas), _stream())


def score(h, core, w, bias, s_out, s_stride):
  """full_model.py:821-822 / box_model.py:508-511."""
  _chk(h, core, w, bias)
  B, Hd = h.shape
  Cd = 0 if core is None else core.numel() // B
  _lib.call('ra_score_f32', _p(h), Hd, _p(core), Cd, _p(w), _p(bias), B, _p(s_out), s_stride, _stream())


# ----------------------------------------------------------------------------- loss side
def get_gt_box(y_gt, padding_ratio=0.0, min_padding=10.0, want_box=True):
  """modellib.get_gt_box (modellib.py:663-701), center_shift_ratio = 0.
  Returns top_left [B,T,2], bot_right [B,T,2], box [B,T,H,W] or None, rect [B,T,4], area [B,T]."""
  _chk(y_gt)
  B, T, H, W = y_gt.shape
  dev = y_gt.device
  tl = torch.empty((B, T, 2), device=dev, dtype=torch.float32)
  br = torch.empty((B, T, 2), device=dev, dtype=torch.float32)
  rect = torch.empty((B, T, 4), device=dev, dtype=torch.float32)
  area = torch.empty((B, T), device=dev, dtype=torch.float32)
  box = torch.empty((B, T, H, W), device=dev, dtype=torch.float32) if want_box else None
  _lib.call('ra_gt_box_f32', _p(y_gt), B, T, H, W, float(padding_ratio), float(min_padding), _p(tl), _p(br),
            _p(rect), _p(box), _p(area), _stream())
  return tl, br, box, rect, area


def f_iou(a, b=None, pairwise=True, b_rect=None, hard_threshold=0.0, want_dice=False, H=None, W=None, out=None):
  """modellib.f_iou(a, b, timespan, pairwise=True) (modellib.py:138-153): a [B,N,H,W],
  b [B,M,H,W] (or b_rect [B,M,4] rectangles) -> iou [B,N,M] (and f_dice, :81-97)."""
  assert pairwise, 'only the pairwise form is on the hot path'
  _chk(a, b, b_rect)
  B, N, H, W = a.shape
  M = b.shape[1] if b is not None else b_rect.shape[1]
  dev = a.device
  n_ws = _lib.lib().ra_pairwise_iou_workspace(B, N, M, H * W)
  if n_ws == 0:
    raise _lib.RecAttendError('ra_pairwise_iou: unsupported N/M')
  ws = torch.empty((n_ws,), device=dev, dtype=torch.float32)
  iou = torch.empty((B, N, M), device=dev, dtype=torch.float32) if out is None else out
  assert tuple(iou.shape) == (B, N, M) and iou.is_contiguous()
  dice = torch.empty((B, N, M), device=dev, dtype=torch.float32) if want_dice else None
  _lib.call('ra_pairwise_iou_f32', _p(a), _p(b), _p(b_rect), B, N, M, H, W, float(hard_threshold), _p(ws), _p(iou),
            _p(dice), _stream())
  if want_dice:
    return iou, dice
  return iou


def f_iou_soft_hard(a, b, hard_threshold=0.5, out_soft=None):
  """modellib.f_iou(a, b, pairwise=True) (modellib.py:138-153) of the soft masks AND f_iou / f_dice of the thresholded
  masks a > hard_threshold (full_model.py:1063-1081) in ONE pass over a and b, on the tensor cores
  (ra_pairwise_iou_umma_f32).  a, b [B,T,H,W].  Returns (iou_soft, iou_hard, dice_hard) [B,T,T], or None when the
  shape is outside the kernel's range (T > 127, H*W % 32 != 0): the caller falls back to f_iou."""
  _chk(a, b)
  B, T, H, W = a.shape
  if tuple(b.shape) != (B, T, H, W):
    return None
  n_ws = _lib.lib().ra_pairwise_iou_umma_workspace(B, T, H, W)
  if n_ws == 0:
    return None
  dev = a.device
  ws = torch.empty((n_ws // 4,), device=dev, dtype=torch.float32)
  out = [torch.empty((B, T, T), device=dev, dtype=torch.float32) for _ in range(3)]
  if out_soft is not None:
    assert tuple(out_soft.shape) == (B, T, T) and out_soft.is_contiguous()
    out[0] = out_soft
  _lib.call('ra_pairwise_iou_umma_f32', _p(a), _p(b), B, T, H, W, float(hard_threshold), _p(ws), _p(out[0]), _p(out[1]),
            _p(out[2]), _stream())
  return tuple(out)


def loss_block(iou_box, match_box, iou_soft, match, iou_hard, dice_hard, s_out, s_gt, gt_area, loss_mix_ratio,
               weight_decay_term, segm_coeff=1.0):
  """full_model.py:942-1081 scalars -> dict keyed like the reference's model dict."""
  _chk(iou_box, match_box, iou_soft, match, iou_hard, dice_hard, s_out, s_gt, gt_area)
  B, T = s_out.shape
  out = torch.empty((_lib.LOSS_COUNT,), device=s_out.device, dtype=torch.float32)
  _lib.call('ra_loss_block_f32', _p(iou_box), _p(match_box), _p(iou_soft), _p(match), _p(iou_hard), _p(dice_hard),
            _p(s_out), _p(s_gt), _p(gt_area), B, T, float(loss_mix_ratio), float(weight_decay_term),
            float(segm_coeff), _p(out), _stream())
  return out


def box_gt_step(attn_box_t, box_bstride, gt_rect, y_gt, noise_t, noise_bstride, iou_t, iou_bstride, grd_ws, canvas):
  """box_model.py:484-504 for one decode step (greedy GT match drives the canvas)."""
  B, T, H, W = y_gt.shape
  _lib.call('ra_box_gt_step_f32', _p(attn_box_t), box_bstride, _p(gt_rect), _p(y_gt), _p(noise_t), noise_bstride, B,
            T, H, W, _p(iou_t), iou_bstride, _p(grd_ws), _p(canvas), _stream())


# ----------------------------------------------------------------------------- training-mode conv block
def batch_norm_train_block(x_raw, gamma, beta, ema_mean=None, ema_var=None, pool=1, relu=True, eps=1e-3, decay=0.9,
                           out=None, batch_mean=None, batch_var=None):
  """nnlib.batch_norm with phase_train=True (nnlib.py:65-128) + activation + max-pool (nnlib.py:229-253) on the raw
  convolution output x_raw [B,H,W,C] (bias included): batch moments, EMA shadows updated IN PLACE
  (shadow -= (1-decay)(shadow - batch)), y = pool(relu(bn(x))).  Returns (y, batch_mean, batch_var)."""
  _chk(x_raw, gamma, beta, ema_mean, ema_var, out)
  B, H, W, C = x_raw.shape
  dev = x_raw.device
  n_ws = _lib.lib().ra_bn_train_workspace(B, H, W, C) if B > 0 else 1
  if n_ws == 0:
    raise _lib.RecAttendError('ra_bn_train_block: unsupported channel count {}'.format(C))
  ws = torch.empty((n_ws,), device=dev, dtype=torch.float32)
  if out is None:
    out = torch.empty((B, H // pool, W // pool, C), device=dev, dtype=torch.float32)
  bm = torch.empty((C,), device=dev, dtype=torch.float32) if batch_mean is None else batch_mean
  bv = torch.empty((C,), device=dev, dtype=torch.float32) if batch_var is None else batch_var
  _chk(bm, bv)
  _lib.call('ra_bn_train_block_f32', _p(x_raw), B, H, W, C, _p(gamma), _p(beta), float(eps), float(decay), pool,
            1 if relu else 0, _p(ws), _p(ema_mean), _p(ema_var), _p(bm), _p(bv), _p(out), _stream())
  return out, bm, bv


def conv3x3_block_train(x, wpack, bias, gamma, beta, ema_mean=None, ema_var=None, pool=1, relu=True, x2=None,
                        upsample=1):
  """One nn.cnn / nn.dcnn layer in TRAINING mode (nnlib.py:229-253, :372-400): tensor-core convolution + bias, then
  batch-statistics BN + ReLU + max-pool.  wpack from pack_umma_weights for this layer's plan."""
  Cout = bias.shape[0]
  ones = torch.ones_like(bias)
  raw = conv3x3_block_umma(x, wpack, Cout, ones, bias, pool=1, relu=False, x2=x2, upsample=upsample)
  return batch_norm_train_block(raw, gamma, beta, ema_mean, ema_var, pool=pool, relu=relu)


# ----------------------------------------------------------------------------- augmentation
def random_transformation(x, padding, offset, vflip=False, hflip=False, transpose=False, y=None, d=None, c=None):
  """image_ops.random_transformation (image_ops.py:9-113) in training mode with the random draws supplied:
  offset = (off_y, off_x) in [0, 2*padding], flips / transpose as booleans (the reference draws them only when no
  orientation input `d` is given, :46-49).  x [B,H,W,3], y [B,T,H,W], d [B,H,W,8], c [B,H,W,C'].  Returns a dict with
  the keys of the reference ('x', 'y', 'd', 'c').  phase_train=False == offset (padding, padding), no flips."""
  if d is not None and (vflip or hflip or transpose):
    raise _lib.RecAttendError('orientation mode is on: no random flips / transpose (image_ops.py:46-49)')
  out = {}
  for key, t, is_stack in (('x', x, False), ('y', y, True), ('d', d, False), ('c', c, False)):
    if t is None:
      continue
    _chk(t)
    if is_stack:
      B, T, H, W = t.shape
      N, C = B * T, 1
    else:
      N, H, W, C = t.shape
    dst = torch.empty_like(t)
    _lib.call('ra_random_transformation_f32', _p(t), N, H, W, C, int(padding), int(offset[0]), int(offset[1]),
              1 if vflip else 0, 1 if hflip else 0, 1 if transpose else 0, _p(dst), _stream())
    out[key] = dst
  return out


# ----------------------------------------------------------------------------- backward of a training-mode conv block
def _ws(nbytes, device):
  return torch.empty(max(int(nbytes), 8), device=device, dtype=torch.uint8)


def batch_norm_train_block_bwd(raw, dy, gamma, beta, mean, var, pool=1, relu=True, eps=1e-3):
  """Gradient of y = pool(relu(bn_batch(raw))) (ops.batch_norm_train_block): raw [B,H,W,C] = the conv output incl.
  bias, mean / var [C] = the batch statistics the forward returned, dy [B,H/pool,W/pool,C].
  Returns (d_raw [B,H,W,C], dgamma [C], dbeta [C])."""
  _chk(raw, dy, gamma, beta, mean, var)
  B, H, W, C = raw.shape
  assert tuple(dy.shape) == (B, H // pool, W // pool, C), (tuple(dy.shape), tuple(raw.shape), pool)
  dev = raw.device
  d_raw = torch.empty_like(raw)
  dgamma = torch.empty(C, device=dev, dtype=torch.float32)
  dbeta = torch.empty(C, device=dev, dtype=torch.float32)
  ws = _ws(_lib.lib().ra_bn_train_block_bwd_workspace(B, H, W, C, pool), dev)
  _lib.call('ra_bn_train_block_bwd_f32', _p(raw), _p(dy), _p(gamma), _p(beta), _p(mean), _p(var), B, H, W, C, pool,
            1 if relu else 0, float(eps), _p(ws), _p(d_raw), _p(dgamma), _p(dbeta), _stream())
  return d_raw, dgamma, dbeta


def conv3x3_bwd_weight(x, d_out, x2=None, upsample=1, want_db=True):
  """Gradient of the conv-form filter w [3,3,C1+C2,Cout] of conv3x3_block(x [,x2], w, upsample=...) and of the bias,
  given d_out [B,H*up,W*up,Cout] = gradient of the raw convolution output.  Returns (dw, db or None)."""
  _chk(x, d_out, x2)
  B, H, W, C1 = x.shape
  C2 = 0 if x2 is None else x2.shape[3]
  Cout = d_out.shape[3]
  assert tuple(d_out.shape) == (B, H * upsample, W * upsample, Cout)
  dev = x.device
  dw = torch.empty((3, 3, C1 + C2, Cout), device=dev, dtype=torch.float32)
  db = torch.empty(Cout, device=dev, dtype=torch.float32) if want_db else None
  ws = _ws(_lib.lib().ra_conv3x3_bwd_weight_workspace(B, H, W, C1 + C2, Cout, upsample), dev)
  _lib.call('ra_conv3x3_bwd_weight_f32', _p(x), C1, _p(x2), C2, _p(d_out), B, H, W, Cout, upsample, _p(ws), _p(dw),
            _p(db), _stream())
  return dw, db


def filter_flip_transpose(w):
  """out[ky,kx,co,ci] = w[2-ky,2-kx,ci,co] on the device (w [3,3,Ci,Co])."""
  _chk(w)
  _, _, Ci, Co = w.shape
  out = torch.empty((3, 3, Co, Ci), device=w.device, dtype=torch.float32)
  _lib.call('ra_filter_flip_transpose_f32', _p(w), Ci, Co, _p(out), _stream())
  return out


def conv3x3_bwd_data(d_out, w, upsample=1):
  """Gradient of the input of conv3x3_block(x, w, upsample=...) (x = concat of x1, x2 along channels) given d_out:
  a SAME convolution of d_out with the flipped, transposed filter; for the transposed-conv layer (upsample = 2) the
  input gradient is that result at the odd positions.  w [3,3,Cin,Cout] conv-form -> dx [B,H,W,Cin]."""
  _chk(d_out, w)
  Cin = w.shape[2]
  wb = filter_flip_transpose(w)
  one = torch.ones(Cin, device=w.device, dtype=torch.float32)
  zero = torch.zeros(Cin, device=w.device, dtype=torch.float32)
  full = conv3x3_block(d_out, wb, one, zero, pool=1, relu=False)
  if upsample == 1:
    return full
  B, H2, W2, _ = full.shape
  dx = torch.empty((B, H2 // 2, W2 // 2, Cin), device=w.device, dtype=torch.float32)
  _lib.call('ra_subsample2_f32', _p(full), B, H2 // 2, W2 // 2, Cin, 1, _p(dx), _stream())
  return dx


def conv3x3_block_train_bwd(x, w, raw, dy, gamma, beta, mean, var, pool=1, relu=True, x2=None, upsample=1, eps=1e-3,
                            want_dx=True):
  """Backward of one training-mode layer of nn.cnn / nn.dcnn (forward: raw = conv3x3_block(x [,x2], w, 1, bias,
  relu=False, upsample=...); y = batch_norm_train_block(raw, ...)).  Returns a dict: dx (channels of x then x2),
  dw (conv-form), db, dgamma, dbeta.  db is exactly the column sum of d_raw - analytically zero behind a
  batch-statistics BN (the mean removes the bias), kept because TensorFlow computes it too."""
  d_raw, dgamma, dbeta = batch_norm_train_block_bwd(raw, dy, gamma, beta, mean, var, pool=pool, relu=relu, eps=eps)
  dw, db = conv3x3_bwd_weight(x, d_raw, x2=x2, upsample=upsample)
  out = {'dw': dw, 'db': db, 'dgamma': dgamma, 'dbeta': dbeta, 'd_raw': d_raw}
  if want_dx:
    out['dx'] = conv3x3_bwd_data(d_raw, w, upsample=upsample)
  return out


# ----------------------------------------------------------------------------- backward of the loss block
def iou_loss_bwd(a, match, b_masks=None, b_rect=None, scale=1.0, out=None):
  """Gradient wrt a [B,N,H,W] of -(scale/B) sum_b 1/cnt_b sum_nm match * f_iou(a_n, g_m) (full_model.py:942-1012):
  g = b_masks [B,M,H,W] (segmentation loss) or b_rect [B,M,4] (box loss, get_gt_box's rect)."""
  _chk(a, match, b_masks, b_rect, out)
  B, N, H, W = a.shape
  M = match.shape[2]
  if out is None:
    out = torch.empty_like(a)
  ws = _ws(_lib.lib().ra_iou_loss_bwd_workspace(B, N, M), a.device)
  _lib.call('ra_iou_loss_bwd_f32', _p(a), N * H * W, _p(b_masks), _p(b_rect), _p(match), B, N, M, H, W, float(scale),
            _p(ws), _p(out), _stream())
  return out


def conf_loss_bwd(s_out, match, scale=1.0):
  """Gradient wrt s_out [B,T] of scale * f_conf_loss(s_out, match) (modellib.py:316-339)."""
  _chk(s_out, match)
  B, T = s_out.shape
  ds = torch.empty_like(s_out)
  _lib.call('ra_conf_loss_bwd_f32', _p(s_out), _p(match), B, T, match.shape[2], float(scale), _p(ds), _stream())
  return ds


# ----------------------------------------------------------------------------- backward of the attention write
def paste_back_bwd(d_out, out, box, fy, fx, gamma_index, patch=None, d_fy=None, d_fx=None, band=True):
  """Gradient of out = sigmoid(gamma * (Fy P Fx^T) - 5) (full_model.py:738-741, :810-814): d_out, out [B,H,W]
  (contiguous), box [B,RA_BOX_STRIDE], fy [B,F,H], fx [B,F,W], gamma_index = _lib.BOX_GAMMA_Y or BOX_GAMMA_BOX,
  patch [B,F,F] or None (= ones).  Passing d_fy / d_fx accumulates into them (the filters have several consumers).
  Returns (d_patch or None, d_fy, d_fx, d_gamma [B])."""
  _chk(d_out, out, box, fy, fx, patch, d_fy, d_fx)
  B, H, W = d_out.shape
  F = fy.shape[1]
  dev = d_out.device
  acc = 1 if d_fy is not None else 0
  if d_fy is None:
    d_fy = torch.empty_like(fy)
    d_fx = torch.empty_like(fx)
  d_patch = torch.empty((B, F, F), device=dev, dtype=torch.float32) if patch is not None else None
  d_gamma = torch.empty(B, device=dev, dtype=torch.float32)
  ws = _ws(_lib.lib().ra_paste_back_bwd_workspace(B, H, W, F), dev)
  gamma = box.view(-1)[gamma_index:]
  # band=True: fy / fx were built from `box` by get_gaussian_filter (exact zeros outside the taps' support bands), so
  # the kernels only walk the bands; band=False walks the full axes (filters of any origin)
  _lib.call('ra_paste_back_bwd_ex_f32', _p(d_out), _p(out), H * W, max(B, 1), 0, _p(patch), _p(fy), _p(fx), _p(gamma),
            box.shape[1], _p(box if band else None), B, H, W, F, acc, _p(ws), _p(d_patch), _p(d_fy), _p(d_fx),
            _p(d_gamma), _stream())
  return d_patch, d_fy, d_fx, d_gamma


def extract_patch_bwd(d_patch, x_patch, xs, canvas, chan_map, box, fy, fx, d_fy=None, d_fx=None, band=True):
  """Gradient of the glimpse x_patch = gamma_attn * Fy^T X Fx (ops.extract_patch; full_model.py:788-789) w.r.t. the
  filters and the gain: d_patch, x_patch [B,F,F,cstride]; xs / canvas / chan_map as in the forward call.
  Passing d_fy / d_fx accumulates into them.  Returns (d_fy, d_fx, d_gamma [B])."""
  _chk(d_patch, x_patch, xs, canvas, chan_map, box, fy, fx, d_fy, d_fx)
  B, F, H = fy.shape
  W = fx.shape[2]
  Cs = 0 if xs is None else xs.shape[3]
  D = Cs + (1 if canvas is not None else 0)
  dev = fy.device
  acc = 1 if d_fy is not None else 0
  if d_fy is None:
    d_fy = torch.empty_like(fy)
    d_fx = torch.empty_like(fx)
  d_gamma = torch.empty(B, device=dev, dtype=torch.float32)
  ws = _ws(_lib.lib().ra_gaussian_extract_bwd_workspace(B, W, F, D), dev)
  gamma = box.view(-1)[_lib.BOX_GAMMA_ATTN:]
  _lib.call('ra_gaussian_extract_bwd_ex_f32', _p(xs), Cs, 0, _p(canvas), _p(chan_map), _p(fy), _p(fx), _p(gamma),
            box.shape[1], _p(box if band else None), _p(d_patch), _p(x_patch), d_patch.shape[3], B, H, W, F, acc,
            _p(ws), _p(d_fy), _p(d_fx), _p(d_gamma), _stream())
  return d_fy, d_fx, d_gamma


def gaussian_filters_bwd(box, fy, fx, d_fy, d_fx):
  """Gradient of modellib.get_gaussian_filter (modellib.py:581-612) for both axes: -> d_box [B,6] =
  (d_ctr_y, d_ctr_x, d_size_y, d_size_x, d_lg_var_y, d_lg_var_x)."""
  _chk(box, fy, fx, d_fy, d_fx)
  B, F, H = fy.shape
  W = fx.shape[2]
  d_box = torch.empty((B, 6), device=box.device, dtype=torch.float32)
  _lib.call('ra_gaussian_filters_bwd_f32', _p(box), _p(fy), _p(fx), _p(d_fy), _p(d_fx), B, H, W, F, _p(d_box),
            _stream())
  return d_box


# ----------------------------------------------------------------------------- backward of the controller
TAPE_FIELDS = ('map', 'glimpse', 'h_prev', 'c_prev', 'gates', 'c', 'h', 'a1', 'map_next')


def controller_tape(feat, lstm_wx, lstm_wh, lstm_b, gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1, cmlp_w, cmlp_b, n_iter=5):
  """Forward of the controller for one decode step with everything the backward needs recorded
  (ra_controller_tape_f32).  Returns (tape [B,n_iter,rec], offsets dict, h_out [B,Hd], ctrl_out [B,9])."""
  _chk(feat, lstm_wx, lstm_wh, lstm_b, gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1, cmlp_w, cmlp_b)
  B, P, Cf = feat.shape
  Hd = lstm_wh.shape[-1]
  off = (_c.c_int * 10)()
  _lib.call('ra_controller_tape_layout', P, Cf, Hd, off)
  offsets = dict(zip(TAPE_FIELDS + ('rec',), list(off)))
  dev = feat.device
  tape = torch.empty((B, n_iter, offsets['rec']), device=dev, dtype=torch.float32)
  h_out = torch.empty((B, Hd), device=dev, dtype=torch.float32)
  ctrl_out = torch.empty((B, 9), device=dev, dtype=torch.float32)
  _lib.call('ra_controller_tape_f32', _p(feat), B, P, Cf, Hd, n_iter, _p(lstm_wx), _p(lstm_wh), _p(lstm_b),
            _p(gmlp_w0), _p(gmlp_b0), _p(gmlp_w1), _p(gmlp_b1), _p(cmlp_w), _p(cmlp_b), _p(tape), _p(h_out),
            _p(ctrl_out), _stream())
  return tape, offsets, h_out, ctrl_out


def outer_sum(A, a_stride, n_in, D, d_stride, n_out, R, want_bias=True):
  """dW [n_in,n_out] = sum_r A_r^T D_r over R rows (A_r at A.data + r*a_stride floats), db = column sums of D."""
  dev = D.device
  dW = torch.empty((n_in, n_out), device=dev, dtype=torch.float32)
  db = torch.empty(n_out, device=dev, dtype=torch.float32) if want_bias else None
  nb = _lib.lib().ra_outer_sum_workspace(n_in, n_out, R)  # > 0: the rows are split over chunks of CTAs
  ws = _ws(nb, dev) if nb else None
  _lib.call('ra_outer_sum_ex_f32', _p(A), a_stride, n_in, _p(D), d_stride, n_out, R, _p(ws), _p(dW), _p(db),
            _stream())
  return dW, db


def controller_bwd(feat, box, lstm_wx, lstm_wh, lstm_b, gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1, cmlp_w, cmlp_b, inp_height,
                   inp_width, flags, d_box, d_gamma3, d_h=None, n_iter=5):
  """Backward of ops.controller_step for one decode step: d_box [B,6] = (d_ctr, d_size, d_lg_var) summed over the
  consumers of the filters, d_gamma3 [B,3] = dL/dgamma (attn, box, y), d_h [B,Hd] = the score head's gradient.
  Returns a dict: d_feat [B,P,Cf] and the weight gradients in the device layouts of the forward call
  (lstm_wx [4,Cf,Hd], lstm_wh [4,Hd,Hd], lstm_b [4,Hd], gmlp_w0/b0, gmlp_w1/b1, cmlp_w/b)."""
  _chk(feat, box, d_box, d_gamma3, d_h)
  B, P, Cf = feat.shape
  Hd = lstm_wh.shape[-1]
  dev = feat.device
  tape, off, h_out, ctrl_out = controller_tape(feat, lstm_wx, lstm_wh, lstm_b, gmlp_w0, gmlp_b0, gmlp_w1, gmlp_b1,
                                               cmlp_w, cmlp_b, n_iter=n_iter)
  d_ctrl = torch.empty((B, 9), device=dev, dtype=torch.float32)
  _lib.call('ra_controller_head_bwd_f32', _p(ctrl_out), _p(box), _p(d_box), _p(d_gamma3), B, inp_height, inp_width,
            flags, _p(d_ctrl), _stream())
  d_feat = torch.empty_like(feat)
  dG = torch.empty((B, n_iter, 4, Hd), device=dev, dtype=torch.float32)
  dA1 = torch.empty((B, n_iter, Hd), device=dev, dtype=torch.float32)
  dLog = torch.empty((B, n_iter, P), device=dev, dtype=torch.float32)
  _lib.call('ra_controller_bwd_f32', _p(feat), B, P, Cf, Hd, n_iter, _p(lstm_wx), _p(lstm_wh), _p(gmlp_w0),
            _p(gmlp_w1), _p(cmlp_w), _p(tape), _p(d_h), _p(d_ctrl), _p(d_feat), _p(dG), _p(dA1), _p(dLog), _stream())
  R, rec = B * n_iter, off['rec']
  flat = tape.view(-1)
  out = {'d_feat': d_feat, 'd_ctrl_out': d_ctrl, 'h': h_out, 'ctrl_out': ctrl_out}
  wx, wh, bg = [], [], []
  for g in range(4):
    Dg = dG.view(-1)[g * Hd:]
    w_, b_ = outer_sum(flat[off['glimpse']:], rec, Cf, Dg, 4 * Hd, Hd, R)
    wx.append(w_)
    bg.append(b_)
    wh.append(outer_sum(flat[off['h_prev']:], rec, Hd, Dg, 4 * Hd, Hd, R, want_bias=False)[0])
  out['lstm_wx'], out['lstm_wh'], out['lstm_b'] = torch.stack(wx), torch.stack(wh), torch.stack(bg)
  out['gmlp_w0'], out['gmlp_b0'] = outer_sum(flat[off['h']:], rec, Hd, dA1, Hd, Hd, R)
  out['gmlp_w1'], out['gmlp_b1'] = outer_sum(flat[off['a1']:], rec, Hd, dLog, P, P, R)
  out['cmlp_w'], out['cmlp_b'] = outer_sum(h_out, Hd, Hd, d_ctrl, 9, 9, B)
  return out


# ----------------------------------------------------------------------------- scheduled sampling (training mode)
def gt_attn_noise(rect_raw, area, pad, shift, min_padding):
  """Noisy GT attention boxes (full_model.py:568-580): rect_raw [B,T,4] raw mask extrema, area [B,T], pad [B,T(,1)],
  shift [B,T,2] -> (ctr [B,T,2], size [B,T,2])."""
  _chk(rect_raw, area, pad, shift)
  B, T = area.shape
  ctr = torch.empty((B, T, 2), device=area.device, dtype=torch.float32)
  size = torch.empty((B, T, 2), device=area.device, dtype=torch.float32)
  _lib.call('ra_gt_attn_noise_f32', _p(rect_raw), _p(area), _p(pad), _p(shift), float(min_padding), B, T, _p(ctr),
            _p(size), _stream())
  return ctr, size


def knob_greedy_box(attn_box_t, box_bstride, gt_rect, H, W, iou_t, iou_bstride, grd):
  """full_model.py:756-759: IoU of this step's attention box with every GT box + greedy match."""
  B, T = grd.shape
  _lib.call('ra_knob_greedy_box_f32', _p(attn_box_t), box_bstride, _p(gt_rect), B, T, H, W, _p(iou_t), iou_bstride,
            _p(grd), _stream())


def greedy_iou_box(box_t, tl_gt, br_gt, iou_t, iou_bstride, grd):
  """opt['use_iou_box'] (full_model.py:750-754, box_model.py:487-491): modellib.f_iou_box of this step's box record
  against every GT box (tl_gt, br_gt [B,T,2] = get_gt_box's returned corners) + f_greedy_match."""
  _chk(box_t, tl_gt, br_gt, grd)
  B, T = grd.shape
  _lib.call('ra_greedy_iou_box_f32', _p(box_t), _p(tl_gt), _p(br_gt), B, T, _p(iou_t), iou_bstride, _p(grd),
            _stream())


def box_gt_canvas(grd, y_gt, noise_t, noise_bstride, canvas):
  """box_model.py:497-503: canvas = max(canvas, sum_m grd*y_gt_m*(1-noise)) (noise_t may be None)."""
  B, T, H, W = y_gt.shape
  _lib.call('ra_box_gt_canvas_f32', _p(grd), _p(y_gt), _p(noise_t), noise_bstride, B, T, H, W, _p(canvas), _stream())


def knob_mix_box(box_t, grd, ctr_gt, size_gt, knob_t, knob_stride):
  """full_model.py:760-776: mix the matched noisy GT box into the box record of this step (in place)."""
  B, T = grd.shape
  _lib.call('ra_knob_mix_box_f32', _p(box_t), _p(grd), _p(ctr_gt), _p(size_gt), _p(knob_t), knob_stride, B, T,
            _stream())


def knob_canvas(grd, y_gt, noise_t, noise_bstride, knob_t, knob_stride, y_out_t, out_bstride, canvas):
  """full_model.py:826-845: canvas = max(canvas, knob ? matched GT mask * (1 - noise) : y_out_t)."""
  B, T, H, W = y_gt.shape
  _lib.call('ra_knob_canvas_f32', _p(grd), _p(y_gt), _p(noise_t), noise_bstride, _p(knob_t), knob_stride, _p(y_out_t),
            out_bstride, B, T, H, W, _p(canvas), _stream())

"""Host side of the foreground / orientation FCN — the drop-in for ``fg_model.get_model`` (fg_model.py:11) in eval
mode (the producer of the hot path's ``d_in`` / ``y_in`` inputs for KITTI and Cityscapes).

Same contract as the reference: an ``opt`` dict (fg_model_train.py:470-499), inputs ``x [B,H,W,3]``, ``y_gt
[B,H,W,nsc]``, ``d_gt [B,H,W,8]``; outputs ``y_out``, ``d_out``, ``iou_soft``, ``iou_hard``, ``foreground_loss``,
``orientation_ce``, ``orientation_acc``, ``loss`` (fg_model.py:181-243); weights under the keys the graph registers
(``cnn_w_i``, ``cnn_{i}_0_gamma`` ..., ``dcnn_w_i``; fg_model.get_save_var, fg_model.py:266-285).

Every FLOP runs in librecattend_b200.so: the CNN / DCNN layers are the same fused conv blocks as the hot path's
(tcgen05 3xTF32 kernel with folded EMA batch norm, ReLU, 2x2 max-pool, transposed conv as zero-inserted conv, skip
concat as a second input pointer); layers whose shape has no tile plan (e.g. 512 output channels) run on the CUDA-core
fp32 conv kernel; the head + loss block is one pass over the logits (csrc/fg.cu).  Eval mode only
(``phase_train=False``: image_ops.random_transformation is the identity, BN uses the EMA shadows).
"""
import ctypes as _c
import os

import numpy as np
import torch

from . import _lib, ops
from .config import fg_skip_wiring
from .full_model import _deconv_to_conv, _fold_bn

FG_SCALARS = {'iou_soft': 0, 'iou_hard': 1, 'segloss': 2, 'foreground_loss': 3, 'orientation_ce': 4,
              'orientation_acc': 5, 'loss': 6}  # RA_FG_* slots of include/rec_attend_b200.h


class FgModel(object):
  """``fg_model.get_model(opt)`` replacement (eval mode)."""

  def __init__(self, opt, device=None):
    self.opt = dict(opt)
    if not torch.cuda.is_available():
      raise _lib.RecAttendError('rec_attend_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    _lib.lib()
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    o = self.opt
    if not o.get('use_bn', True):
      raise _lib.RecAttendError('use_bn=False is not used by any shipped fg_model config')
    self.H, self.W, self.Cin = o['inp_height'], o['inp_width'], o['inp_depth']
    self.nsc = o.get('num_semantic_classes', 1)
    self.nori = o['num_orientation_classes'] if o.get('add_orientation', False) else 0
    self.cnn_pool, self.dcnn_pool = list(o['cnn_pool']), list(o['dcnn_pool'])
    self.cnn_ch = [self.Cin] + list(o['cnn_depth'])
    self.dcnn_ch = [self.cnn_ch[-1]] + list(o['dcnn_depth'])
    if self.dcnn_ch[-1] != self.nsc + self.nori:
      raise _lib.RecAttendError('Expecting last channel to be {}'.format(self.nsc + self.nori))  # fg_model.py:162-172
    if o.get('segm_loss_fn', 'iou') not in ('iou', 'bce'):
      raise _lib.RecAttendError("segm_loss_fn must be 'iou' or 'bce' (fg_model.py:223-226)")
    self.skip_src, self.skip_ch = fg_skip_wiring(o)
    # spatial sizes: input of CNN layer i / output of DCNN layer i; the skip tensors must match their consumers
    sub = int(np.prod(self.cnn_pool))
    if self.H % sub or self.W % sub:
      raise _lib.RecAttendError('input size must be divisible by {} (SAME pooling on even sizes only)'.format(sub))
    self.cnn_in = []
    h, w = self.H, self.W
    for p in self.cnn_pool:
      self.cnn_in.append((h, w))
      h, w = h // p, w // p
    self.dcnn_out = []
    for p in self.dcnn_pool:
      h, w = h * p, w * p
      self.dcnn_out.append((h, w))
    if self.dcnn_out[-1] != (self.H, self.W):
      raise _lib.RecAttendError('the DCNN must return to the input resolution')
    for i, j in enumerate(self.skip_src):
      if j is not None:
        prev = self.dcnn_out[i - 1]
        if self.cnn_in[j] != prev:
          raise _lib.RecAttendError('skip {} -> DCNN layer {}: {} vs {}'.format(j, i, self.cnn_in[j], prev))
    self.w = None
    self._bufs = {}

  # ------------------------------------------------------------------ weights
  def _dev(self, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

  def load_weights(self, weights):
    self._raw_weights = {k: np.asarray(v, np.float32) for k, v in weights.items()}
    layers = []
    for i, pool in enumerate(self.cnn_pool):
      wi = np.asarray(weights['cnn_w_%d' % i], np.float32)
      if wi.shape != (3, 3, self.cnn_ch[i], self.cnn_ch[i + 1]):
        raise _lib.RecAttendError('cnn_w_{} has shape {}'.format(i, wi.shape))
      sc, sh = _fold_bn(weights, 'cnn', i, 1, np.asarray(weights['cnn_b_%d' % i], np.float32))
      layers.append(self._layer(wi, sc[0], sh[0], self.cnn_in[i], pool, 1, True))
    n_d = len(self.dcnn_pool)
    for i, up in enumerate(self.dcnn_pool):
      wt = np.asarray(weights['dcnn_w_%d' % i], np.float32)
      cin = self.dcnn_ch[i] + self.skip_ch[i]
      if wt.shape != (3, 3, self.dcnn_ch[i + 1], cin):
        raise _lib.RecAttendError('dcnn_w_{} has shape {}, expected [3,3,{},{}] (nnlib.py:320-325)'.format(
            i, wt.shape, self.dcnn_ch[i + 1], cin))
      bias = np.asarray(weights['dcnn_b_%d' % i], np.float32)
      if i < n_d - 1:
        sc, sh = _fold_bn(weights, 'dcnn', i, 1, bias)
        sc, sh, relu = sc[0], sh[0], True
      else:  # last layer: no BN, no activation (fg_model.py:121,148)
        sc, sh, relu = np.ones_like(bias), bias, False
      layers.append(self._layer(_deconv_to_conv(wt), sc, sh, self.dcnn_out[i], 1, up, relu))
    self.w = layers
    return self

  def _layer(self, w_hwio, scale, shift, out_hw, pool, upsample, relu):
    return {'w': np.ascontiguousarray(w_hwio, np.float32), 'scale': self._dev(scale), 'shift': self._dev(shift),
            'Hout': out_hw[0], 'Wout': out_hw[1], 'pool': pool, 'up': upsample, 'relu': relu, 'packed': {}}

  def _conv(self, L, x, x2, out):
    """One fused layer.  tcgen05 kernel when the layer shape has a tile plan, else the CUDA-core fp32 kernel
    (also with RA_CONV_FP32=1, the precision reference)."""
    B = x.shape[0]
    if B not in L['packed']:
      packed = None
      if not os.environ.get('RA_CONV_FP32'):
        w = L['w']
        try:
          KC, NPc, nsp, _, rs = ops.umma_plan(w.shape[2], w.shape[3], L['Hout'], L['Wout'], L['pool'], B)
          packed = self._dev(ops.pack_umma_weights(w, KC, NPc, nsp, rs))
        except _lib.RecAttendError:
          packed = None  # no tile plan for this shape (RA_ERR_UNSUPPORTED): fp32 kernel below
      if packed is None and 'w_dev' not in L:
        L['w_dev'] = self._dev(L['w'])
      L['packed'][B] = packed
    packed = L['packed'][B]
    if packed is not None:
      return ops.conv3x3_block_umma(x, packed, L['w'].shape[3], L['scale'], L['shift'], pool=L['pool'], relu=L['relu'],
                                    x2=x2, upsample=L['up'], out=out)
    return ops.conv3x3_block(x, L['w_dev'], L['scale'], L['shift'], pool=L['pool'], relu=L['relu'], x2=x2,
                             upsample=L['up'], out=out)

  def conv_kernels(self, B):
    """Which kernel each layer runs on at batch size B ('umma' / 'fp32'), after a forward at that size."""
    return ['umma' if L['packed'].get(B) is not None else 'fp32' for L in self.w]

  # ------------------------------------------------------------------ forward
  def _alloc(self, B):
    dev, f32 = self.device, torch.float32
    bufs = {'cnn': [], 'dcnn': []}
    for i, pool in enumerate(self.cnn_pool):
      h, w = self.cnn_in[i]
      bufs['cnn'].append(torch.empty((B, h // pool, w // pool, self.cnn_ch[i + 1]), device=dev, dtype=f32))
    for i in range(len(self.dcnn_pool)):
      h, w = self.dcnn_out[i]
      bufs['dcnn'].append(torch.empty((B, h, w, self.dcnn_ch[i + 1]), device=dev, dtype=f32))
    npix = B * self.H * self.W
    bufs['y_out'] = torch.empty((B, self.H, self.W, self.nsc), device=dev, dtype=f32)
    bufs['y_hard'] = torch.empty((B, self.H, self.W, self.nsc), device=dev, dtype=f32)
    bufs['d_out'] = torch.empty((B, self.H, self.W, self.nori), device=dev, dtype=f32) if self.nori else None
    bufs['scal'] = torch.zeros(8, device=dev, dtype=f32)
    bufs['ws'] = torch.empty(int(_lib.lib().ra_fg_head_workspace()), device=dev, dtype=torch.uint8)
    bufs['npix'] = npix
    return bufs

  def forward(self, batch, outputs=None, phase_train=False):
    """``sess.run`` replacement.  batch: x [B,H,W,3] and, for the loss block, y_gt [B,H,W,nsc] (+ d_gt [B,H,W,8] with
    the orientation head); without y_gt only y_out / d_out / y_out_hard are produced (fg_model_pack.py's use)."""
    if phase_train:
      raise _lib.RecAttendError('fg_model: only the eval-mode forward is built')
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')

    def dev(v):
      if isinstance(v, np.ndarray):
        v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
      return v.to(self.device, dtype=torch.float32).contiguous()

    x = dev(batch['x'])
    B = x.shape[0]
    if tuple(x.shape) != (B, self.H, self.W, self.Cin):
      raise _lib.RecAttendError('x must be [B,{},{},{}], got {}'.format(self.H, self.W, self.Cin, tuple(x.shape)))
    y_gt = dev(batch['y_gt']) if batch.get('y_gt') is not None else None
    d_gt = dev(batch['d_gt']) if (self.nori and batch.get('d_gt') is not None) else None
    if y_gt is not None and tuple(y_gt.shape) != (B, self.H, self.W, self.nsc):
      raise _lib.RecAttendError('y_gt must be [B,H,W,{}]'.format(self.nsc))
    if y_gt is not None and self.nori and (d_gt is None or tuple(d_gt.shape) != (B, self.H, self.W, self.nori)):
      raise _lib.RecAttendError('the orientation head needs d_gt [B,H,W,{}] next to y_gt'.format(self.nori))
    if B not in self._bufs:
      self._bufs[B] = self._alloc(B)
    bufs = self._bufs[B]
    n_c = len(self.cnn_pool)
    _lib.TAG = 'fg_cnn'
    acts = [x]  # [x] + h_cnn
    for i in range(n_c):
      acts.append(self._conv(self.w[i], acts[-1], None, bufs['cnn'][i]))
    _lib.TAG = 'fg_dcnn'
    prev = acts[-1]
    for i in range(len(self.dcnn_pool)):
      j = self.skip_src[i]
      prev = self._conv(self.w[n_c + i], prev, None if j is None else acts[j], bufs['dcnn'][i])
    _lib.TAG = 'fg_head'
    bce = 1 if self.opt.get('segm_loss_fn', 'iou') == 'bce' else 0
    _lib.call('ra_fg_head_f32', ops._p(prev), bufs['npix'], self.nsc, self.nori, ops._p(y_gt), ops._p(d_gt), bce,
              ops._p(bufs['y_out']), ops._p(bufs['d_out']), ops._p(bufs['y_hard']), ops._p(bufs['scal']),
              _c.c_void_p(bufs['ws'].data_ptr()), ops._stream())
    out = {'y_out': bufs['y_out'], 'y_out_hard': bufs['y_hard'], 'logits': prev}
    if self.nori:
      out['d_out'] = bufs['d_out']
    if y_gt is not None:
      for k, i in FG_SCALARS.items():
        if k in ('orientation_ce', 'orientation_acc') and not self.nori:
          continue
        out[k] = bufs['scal'][i]
    if outputs is not None:
      out = {k: out[k] for k in outputs}
    return out

  def pack_outputs(self, out):
    """What fg_model_pack.py writes back for the instance model: y_in = y_out, d_in = d_out (the hot path's extra
    input channels, full_model.py:165-194)."""
    return {'y_in': out['y_out'], 'd_in': out.get('d_out')}


def get_model(opt, device=None):
  """Same call as the reference's ``fg_model.get_model(opt)``."""
  return FgModel(opt, device=device)

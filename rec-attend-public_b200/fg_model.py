"""Host side of the foreground / orientation FCN — the drop-in for ``fg_model.get_model`` (fg_model.py:11) in eval
mode (the producer of the hot path's ``d_in`` / ``y_in`` inputs for KITTI and Cityscapes).

Same contract as the reference: an ``opt`` dict (fg_model_train.py:470-499), inputs ``x [B,H,W,3]``, ``y_gt
[B,H,W,nsc]``, ``d_gt [B,H,W,8]``; outputs ``y_out``, ``d_out``, ``iou_soft``, ``iou_hard``, ``foreground_loss``,
``orientation_ce``, ``orientation_acc``, ``loss`` (fg_model.py:181-243); weights under the keys the graph registers
(``cnn_w_i``, ``cnn_{i}_0_gamma`` ..., ``dcnn_w_i``; fg_model.get_save_var, fg_model.py:266-285).

Every FLOP runs in librecattend_b200.so: the CNN / DCNN layers are the same fused conv blocks as the hot path's
(tcgen05 3xTF32 kernel with folded EMA batch norm, ReLU, 2x2 max-pool, transposed conv as zero-inserted conv, skip
concat as a second input pointer); layers whose shape has no tile plan (e.g. 512 output channels) run on the CUDA-core
fp32 conv kernel; the head + loss block is one pass over the logits (csrc/fg.cu).

Training mode (``forward(phase_train=True)``, ``train_step``; fg_model.py:96-172,249-266 with the identity crop of
image_ops.random_transformation): batch-statistics BN with the EMA shadows moved in place, the backward pass through
head, DCNN (skip gradients handed back to the CNN activations) and CNN from the per-block backward kernels, Adam
without gradient clipping (``optim.minimize``) on one flat bucket with the data-parallel all-reduce of `optim.py`.
The trainable tensors live in that bucket; the training-mode convolutions run on the fp32 CUDA-core conv kernel with the
filters in their reference layouts, so a step needs no host round trip and no re-packing (one fused device copy).
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
  """``fg_model.get_model(opt)`` replacement."""

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
    self._trainer = None
    self._eval_dirty = False

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
    self._trainer = None  # the flat bucket is rebuilt from the new weights at the next training call
    self._eval_dirty = False
    return self

  def _layer(self, w_hwio, scale, shift, out_hw, pool, upsample, relu):
    return {'w': np.ascontiguousarray(w_hwio, np.float32), 'scale': self._dev(scale), 'shift': self._dev(shift),
            'Hout': out_hw[0], 'Wout': out_hw[1], 'pool': pool, 'up': upsample, 'relu': relu, 'packed': {}}

  def _conv(self, L, x, x2, out):
    """One fused layer.  tcgen05 kernel when the layer shape has a tile plan, else the CUDA-core fp32 kernel
    (also with RA_CONV_FP32=1, the precision reference)."""
    B = x.shape[0]
    C2 = 0 if x2 is None else int(x2.shape[3])
    if B not in L['packed']:
      packed = None
      if not os.environ.get('RA_CONV_FP32'):
        w = L['w']
        try:
          KC, NPc, nsp, _, rs = ops.umma_plan(w.shape[2], w.shape[3], L['Hout'], L['Wout'], L['pool'], B, C2=C2)
          packed = ops.umma_filter_image(w, KC, NPc, nsp, rs, self.device)
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
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')
    if phase_train:
      return self._forward_train(batch, outputs)[0]
    if self._eval_dirty:  # weights / EMA shadows moved by training: refold the BN and re-pack the eval-mode filters
      self.load_weights(self.export_weights())

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

  # ------------------------------------------------------------------ training mode
  def _train_state(self):
    """The flat parameter bucket (optim.AdamOptimizer; no gradient clipping: fg_model.py:258-265 minimises total_loss
    directly) with per-key views for the kernels, and the EMA shadows as device tensors."""
    if self._trainer is None:
      from . import optim
      o = dict(self.opt, clip_gradient=0.0)
      adam = optim.AdamOptimizer(o, self._raw_weights, device=self.device)
      views, tens = [], {}
      for k, (off, shape) in adam.flat.layout.items():
        n = int(np.prod(shape)) if shape else 1
        v = adam.params[off:off + n].view(shape)
        views.append(v)
        tens[k] = v.clone()  # own allocation: 256-byte aligned whatever the offset inside the bucket
      ema = {k: self._dev(v) for k, v in self._raw_weights.items() if k.endswith(('_ema_mean', '_ema_var'))}
      self._trainer = {'adam': adam, 'p': tens, 'views': views, 'ema': ema, 'grad': torch.zeros_like(adam.params)}
    return self._trainer

  def _sync_params(self):
    """Bucket -> the per-tensor copies the kernels read, after an optimiser step (device-side, one fused copy)."""
    ts = self._trainer
    torch._foreach_copy_([ts['p'][k] for k in ts['adam'].flat.layout], ts['views'])

  def _stage(self, batch):
    def dev(v):
      if isinstance(v, np.ndarray):
        v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
      return v.to(self.device, dtype=torch.float32).contiguous()

    x = dev(batch['x'])
    B = x.shape[0]
    if tuple(x.shape) != (B, self.H, self.W, self.Cin):
      raise _lib.RecAttendError('x must be [B,{},{},{}], got {}'.format(self.H, self.W, self.Cin, tuple(x.shape)))
    if batch.get('y_gt') is None:
      raise _lib.RecAttendError('training mode needs y_gt')
    y_gt = dev(batch['y_gt'])
    if tuple(y_gt.shape) != (B, self.H, self.W, self.nsc):
      raise _lib.RecAttendError('y_gt must be [B,H,W,{}]'.format(self.nsc))
    d_gt = None
    if self.nori:
      if batch.get('d_gt') is None:
        raise _lib.RecAttendError('the orientation head needs d_gt [B,H,W,{}] next to y_gt'.format(self.nori))
      d_gt = dev(batch['d_gt'])
    return x, y_gt, d_gt, B

  def _forward_train(self, batch, outputs=None):
    """Training-mode forward (nnlib.py:229-253,372-400 with phase_train=True): raw convolution + bias, batch-statistics
    BN (EMA shadows moved in place) + ReLU + max-pool; the last DCNN layer has no BN and no activation.  Returns the
    output dict and the tape the backward pass needs."""
    ts = self._train_state()
    P, ema = ts['p'], ts['ema']
    x, y_gt, d_gt, B = self._stage(batch)
    n_c, n_d = len(self.cnn_pool), len(self.dcnn_pool)
    tape = {'cnn': [], 'dcnn': []}
    acts = [x]
    _lib.TAG = 'fg_cnn'
    for i in range(n_c):
      w, b = P['cnn_w_%d' % i], P['cnn_b_%d' % i]
      raw = ops.conv3x3_block(acts[-1], w, torch.ones_like(b), b, pool=1, relu=False)
      y, mean, var = ops.batch_norm_train_block(raw, P['cnn_%d_0_gamma' % i], P['cnn_%d_0_beta' % i],
                                                ema['cnn_%d_0_ema_mean' % i], ema['cnn_%d_0_ema_var' % i],
                                                pool=self.cnn_pool[i], relu=True)
      tape['cnn'].append({'raw': raw, 'mean': mean, 'var': var})
      acts.append(y)
    _lib.TAG = 'fg_dcnn'
    prev = acts[-1]
    for i in range(n_d):
      j = self.skip_src[i]
      skip = None if j is None else acts[j]
      wc = ops.filter_flip_transpose(P['dcnn_w_%d' % i])  # TF [3,3,Cout,Cin] -> conv-form [3,3,Cin,Cout] (nnlib.py:320-325)
      b = P['dcnn_b_%d' % i]
      raw = ops.conv3x3_block(prev, wc, torch.ones_like(b), b, pool=1, relu=False, x2=skip, upsample=self.dcnn_pool[i])
      rec = {'x': prev, 'x2': skip, 'wc': wc, 'raw': raw}
      if i < n_d - 1:
        prev, rec['mean'], rec['var'] = ops.batch_norm_train_block(
            raw, P['dcnn_%d_0_gamma' % i], P['dcnn_%d_0_beta' % i], ema['dcnn_%d_0_ema_mean' % i],
            ema['dcnn_%d_0_ema_var' % i], pool=1, relu=True)
      else:
        prev = raw
      tape['dcnn'].append(rec)
    _lib.TAG = 'fg_head'
    if B not in self._bufs:
      self._bufs[B] = self._alloc(B)
    bufs = self._bufs[B]
    bce = 1 if self.opt.get('segm_loss_fn', 'iou') == 'bce' else 0
    _lib.call('ra_fg_head_f32', ops._p(prev), bufs['npix'], self.nsc, self.nori, ops._p(y_gt), ops._p(d_gt), bce,
              ops._p(bufs['y_out']), ops._p(bufs['d_out']), ops._p(bufs['y_hard']), ops._p(bufs['scal']),
              _c.c_void_p(bufs['ws'].data_ptr()), ops._stream())
    out = {'y_out': bufs['y_out'], 'y_out_hard': bufs['y_hard'], 'logits': prev}
    if self.nori:
      out['d_out'] = bufs['d_out']
    for k, i in FG_SCALARS.items():
      if k in ('orientation_ce', 'orientation_acc') and not self.nori:
        continue
      out[k] = bufs['scal'][i]
    self._eval_dirty = True  # the EMA shadows moved
    tape.update({'acts': acts, 'y_gt': y_gt, 'd_gt': d_gt, 'B': B, 'bce': bce})
    if outputs is not None:
      out = {k: out[k] for k in outputs}
    return out, tape

  def _backward(self, tape):
    """Gradients of loss = foreground_loss [+ orientation_ce] into the flat bucket (data term only; the optimiser adds
    wd * w): head, then the DCNN last layer first - the skip half of every input gradient goes back to the CNN
    activation it came from -, then the CNN."""
    ts = self._train_state()
    P, adam, grad = ts['p'], ts['adam'], ts['grad']
    layout = adam.flat.layout
    grad.zero_()

    def put(key, g):
      off, shape = layout[key]
      grad[off:off + g.numel()].copy_(g.reshape(-1))

    B, acts = tape['B'], tape['acts']
    bufs = self._bufs[B]
    n_c, n_d = len(self.cnn_pool), len(self.dcnn_pool)
    _lib.TAG = 'fg_bwd_head'
    d_cur = torch.empty((B, self.H, self.W, self.nsc + self.nori), device=self.device, dtype=torch.float32)
    _lib.call('ra_fg_head_bwd_f32', ops._p(bufs['y_out']), ops._p(bufs['d_out']), bufs['npix'], self.nsc, self.nori,
              ops._p(tape['y_gt']), ops._p(tape['d_gt']), tape['bce'], _c.c_void_p(bufs['ws'].data_ptr()), ops._p(d_cur),
              ops._stream())
    from .train import _add, _split
    d_acts = [None] * len(acts)  # gradient at acts[j] from the skip connections

    _lib.TAG = 'fg_bwd_dcnn'
    for i in range(n_d - 1, -1, -1):
      rec = tape['dcnn'][i]
      up = self.dcnn_pool[i]
      if i == n_d - 1:  # no BN, no activation: d_raw = d_logits
        dw, db = ops.conv3x3_bwd_weight(rec['x'], d_cur, x2=rec['x2'], upsample=up)
        dx = ops.conv3x3_bwd_data(d_cur, rec['wc'], upsample=up)
      else:
        r = ops.conv3x3_block_train_bwd(rec['x'], rec['wc'], rec['raw'], d_cur, P['dcnn_%d_0_gamma' % i],
                                        P['dcnn_%d_0_beta' % i], rec['mean'], rec['var'], pool=1, relu=True,
                                        x2=rec['x2'], upsample=up)
        dw, db, dx = r['dw'], r['db'], r['dx']
        put('dcnn_%d_0_gamma' % i, r['dgamma'])
        put('dcnn_%d_0_beta' % i, r['dbeta'])
      put('dcnn_w_%d' % i, ops.filter_flip_transpose(dw))  # the layout map is a permutation: it carries the gradient
      put('dcnn_b_%d' % i, db)
      if rec['x2'] is not None:  # channels of x, then of the skip tensor (ra_split_channels_f32)
        j = self.skip_src[i]
        d_cur, d_acts[j] = _split(dx, rec['x'].shape[3], rec['x2'].shape[3], dst2=d_acts[j],
                                  accumulate2=d_acts[j] is not None)
      else:
        d_cur = dx
    _lib.TAG = 'fg_bwd_cnn'
    for i in range(n_c - 1, -1, -1):
      if d_acts[i + 1] is not None:
        _add(d_cur, d_acts[i + 1])
      rec = tape['cnn'][i]
      r = ops.conv3x3_block_train_bwd(acts[i], P['cnn_w_%d' % i], rec['raw'], d_cur, P['cnn_%d_0_gamma' % i],
                                      P['cnn_%d_0_beta' % i], rec['mean'], rec['var'], pool=self.cnn_pool[i], relu=True,
                                      want_dx=i > 0)
      put('cnn_w_%d' % i, r['dw'])
      put('cnn_b_%d' % i, r['db'])
      put('cnn_%d_0_gamma' % i, r['dgamma'])
      put('cnn_%d_0_beta' % i, r['dbeta'])
      if i > 0:
        d_cur = r['dx']
    return grad

  def train_step(self, batch):
    """``sess.run([loss, train_step], feed_dict)`` for the graph of fg_model.py:249-266: training-mode forward,
    backward, gradient all-reduce over the data-parallel ranks, Adam(eps 1e-7) with the staircase learning-rate decay
    and weight decay on the conv filters.  Returns the loss scalars of the forward (BEFORE the update)."""
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')
    out, tape = self._forward_train(batch)
    res = {k: out[k].clone() for k in FG_SCALARS if k in out}
    grad = self._backward(tape)
    adam = self._train_state()['adam']
    res['learn_rate'] = adam.step(grad)
    res['global_step'] = adam.global_step
    self._sync_params()
    return res

  @property
  def optimizer(self):
    return None if self._trainer is None else self._trainer['adam']

  def export_weights(self):
    """The current weights under the reference's keys (trained tensors from the bucket, moved EMA shadows)."""
    if self._trainer is None:
      return dict(self._raw_weights)
    out = self._trainer['adam'].export_weights(self._raw_weights)
    for k, v in self._trainer['ema'].items():
      out[k] = v.cpu().numpy()
    return out

  def pack_outputs(self, out):
    """What fg_model_pack.py writes back for the instance model: y_in = y_out, d_in = d_out (the hot path's extra
    input channels, full_model.py:165-194)."""
    return {'y_in': out['y_out'], 'd_in': out.get('d_out')}


def get_model(opt, device=None):
  """Same call as the reference's ``fg_model.get_model(opt)``."""
  return FgModel(opt, device=device)

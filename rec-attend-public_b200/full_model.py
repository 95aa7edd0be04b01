"""Host side of the recurrent-attention model — the drop-in for ``full_model.get_model``
(full_model.py:13) and ``box_model.get_model`` (box_model.py:11) in eval mode.

Same contract as the reference: an ``opt`` dict in (full_model_train.py:581-658), named
inputs (``x [B,H,W,3]``, ``y_gt [B,T,H,W]``, ``s_gt [B,T]``, optional ``d_in [B,H,W,8]``,
``y_in [B,H,W,C]``) and named outputs (``y_out``, ``s_out``, ``attn_box``, ``match``, ``loss`` …,
full_model.py:853-910,941-1081), weights imported by the flat ``weights.h5`` key names
(full_model_read.py:32-70).  Python only sequences launches: every FLOP runs in
librecattend_b200.so through the C ABI (ops.py); torch provides device memory, streams and
(optionally) CUDA-graph capture of the whole T-step decode.
"""
import os

import numpy as np
import torch

from . import _lib, ops, params as PM
from .config import input_depths
from .synthetic import dcnn_skip_channels

BN_EPS = 1e-3  # nnlib.py:119

LOSS_KEYS = _lib.LOSS_NAMES


def _fold_bn(weights, scope, layer, T, bias):
  """Eval-mode batch norm (nnlib.py:113-119) + conv bias folded into y = conv*scale + shift,
  one row per decode step (the reference keeps a separate BN copy per step, nnlib.py:212-254)."""
  scale = np.empty((T, bias.shape[0]), np.float32)
  shift = np.empty((T, bias.shape[0]), np.float32)
  for t in range(T):
    k = '{}_{}_{}_'.format(scope, layer, t)
    inv = (1.0 / np.sqrt(weights[k + 'ema_var'].astype(np.float32) + np.float32(BN_EPS))).astype(np.float32)
    inv = inv * weights[k + 'gamma'].astype(np.float32)
    scale[t] = inv
    shift[t] = weights[k + 'beta'] - weights[k + 'ema_mean'] * inv + bias * inv
  return scale, shift


def _bn_params(wi, scope, layer, T):
  """Per-(layer, step) BN parameters and EMA shadows as [T, C] arrays + the conv bias (training-mode forward);
  `wi(key)` returns the weight with its source indices (params.WI)."""
  out = {'bias': wi('{}_b_{}'.format(scope, layer))}
  for n in ('gamma', 'beta', 'ema_mean', 'ema_var'):
    out[n] = PM.WI.stack([wi('{}_{}_{}_{}'.format(scope, layer, t, n)) for t in range(T)])
  return out


def _deconv_to_conv(w):
  """conv2d_transpose filter [kh,kw,Cout,Cin] (nnlib.py:320-325,372-376) -> the HWIO filter of
  the equivalent stride-1 convolution over the (zero-inserted) input: spatial flip + swap."""
  return PM.deconv_to_conv(np.asarray(w, np.float32))


def _pad_cin(w_hwio, cin):
  """Zero rows for padded input channels (appended after the real ones)."""
  return PM.pad_cin(w_hwio, cin)


def capture_graph(fn):
  """Capture fn() in a CUDA graph.  The cyclic garbage collector is held off for the duration: collecting an
  unreachable model of an earlier run destroys ITS graphs and frees their private pools - a cudaFree that
  invalidates the capture in progress ("operation not permitted when stream is capturing")."""
  import gc
  g = torch.cuda.CUDAGraph()
  gc.collect()
  was = gc.isenabled()
  gc.disable()
  try:
    with torch.cuda.graph(g):
      out = fn()
  finally:
    if was:
      gc.enable()
  return g, out


class _ModelBase(object):

  def __init__(self, opt, device=None):
    self.opt = dict(opt)
    if not torch.cuda.is_available():
      raise _lib.RecAttendError('rec_attend_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    _lib.lib()  # fail loudly now if the extension is missing
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    o = self.opt
    self.T, self.H, self.W = o['timespan'], o['inp_height'], o['inp_width']
    self.F = o['filter_height']
    if o['filter_width'] != self.F:
      raise _lib.RecAttendError('square attention filters only')
    self.Hd = o['ctrl_rnn_hid_dim']
    self.n_iter = o['num_ctrl_rnn_iter']
    if o['num_ctrl_mlp_layers'] != 1 or o['num_glimpse_mlp_layers'] != 2:
      raise _lib.RecAttendError('unsupported controller/glimpse MLP depth (all shipped configs use 1 / 2)')
    if not o.get('use_bn', True):
      raise _lib.RecAttendError('use_bn=False is not used by any shipped config')
    if o.get('fixed_order', False):
      raise _lib.RecAttendError('fixed_order is not on the hot path (no shipped config uses it)')
    self.add_d = bool(o.get('add_d_out', False))
    self.nsc = o.get('num_semantic_classes', 1)
    self.D = input_depths(o)[0]
    self.Cs = self.D - 1  # step-invariant channels: x (3) [+ d_in (8) + y_in (nsc)]
    # reference concat order is x, canvas, d_in, y_in (full_model.py:643-658)
    self.static_idx = [0, 1, 2] + list(range(4, self.D))
    self.chan_map = torch.tensor(self.static_idx + [3], dtype=torch.int32, device=self.device)
    flags = 0
    if o.get('squash_ctrl_params', False):
      flags |= _lib.CTRL_SQUASH
    if self._fixed_var():
      flags |= _lib.CTRL_FIXED_VAR
    if o.get('dynamic_var', False):
      flags |= _lib.CTRL_DYNAMIC_VAR
    if self._fixed_gamma():
      flags |= _lib.CTRL_FIXED_GAMMA
    self.ctrl_flags = flags
    self.ctrl_pool = list(o['ctrl_cnn_pool'])
    sub = int(np.prod(self.ctrl_pool))
    self.gh, self.gw = self.H // sub, self.W // sub
    self.P = self.gh * self.gw
    self.w = None
    self.wd_term = 0.0
    self._bufs = {}
    self._synced = []      # (device tensor, source index array, kind array) of every weight image (see _dev)
    self._f16_images = []  # (fp32 source, fp16 hi / lo filter image, KC, NPc) of every RA_UMMA_F16 layer (see _packed)
    self._sync_table = None
    self._trainer = None
    self._tape_on = None   # the tape dict while a train_step forward is being enqueued (see train.py)
    self._bn_layers = []   # (device-key prefix, weight-dict scope, layer) of every BN layer
    self._bn_dirty = False  # a training-mode forward has moved the EMA shadows: refold before the next eval forward
    self.n_chains = int(os.environ.get('RA_CHAINS', '1'))  # sub-batch chains of the decode loop (see _chains)

  def _fixed_var(self):
    return bool(self.opt.get('fixed_var', False))

  def _fixed_gamma(self):
    return bool(self.opt.get('fixed_gamma', False))

  # ------------------------------------------------------------------ weights
  def _dev(self, a):
    """Host array -> device tensor.  A params.WI (values + source indices) is also REGISTERED: after an optimiser
    step `sync_weights` rewrites every registered tensor from the flat parameter bucket in one launch."""
    if isinstance(a, PM.WI):
      t = torch.from_numpy(np.ascontiguousarray(a.val, dtype=np.float32)).to(self.device)
      self._synced.append((t, a.idx.reshape(-1), None if a.kind is None else a.kind.reshape(-1)))
      return t
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)

  def _wi(self, key):
    return PM.wi_of(self._raw_weights, self._all_layout, key)

  def _begin_load(self, weights):
    """Common head of load_weights: host copy of the weight dict, the all-weights index layout, and a clean slate
    for everything derived from the previous weights (captured CUDA graphs hold raw pointers to the old tensors)."""
    self._raw_weights = {k: np.asarray(v, np.float32) for k, v in weights.items()}
    self._all_layout = PM.AllLayout(self._raw_weights)
    self._synced = []
    self._f16_images = []
    self._sync_table = None
    self._bn_layers = []
    self._bn_dirty = False
    self._trainer = None
    for b in self._bufs.values():
      b.pop('graphs', None)
      b.pop('chains', None)  # conv chains hold the old weight pointers in their device-side descriptors

  def _load_controller(self, weights):
    T = self.T
    w = {}
    n = len(self.opt['ctrl_cnn_filter_size'])
    w0 = self._wi('ctrl_cnn_w_0')
    if w0.val.shape[2] != self.D:
      raise _lib.RecAttendError('ctrl_cnn_w_0 has {} input channels, expected {}'.format(w0.val.shape[2], self.D))
    w0c = w0.map(lambda a: a[:, :, 3:4, :])
    w['ccnn_w0_canvas'] = self._dev(w0c)
    w['ccnn_w0_canvas_idx'] = w0c.idx
    c0 = w0.val.shape[3]
    w['one_c0'] = torch.ones(c0, device=self.device)
    w['zero_c0'] = torch.zeros(c0, device=self.device)
    h, wd_ = self.H, self.W
    w['ccnn_w0_static_umma'] = self._pack(w0.map(lambda a: a[:, :, self.static_idx, :]), h, wd_, 1)
    for i in range(n):
      if i > 0:
        w['ccnn_w%d' % i] = self._pack(self._wi('ctrl_cnn_w_%d' % i), h, wd_, self.ctrl_pool[i])
      h, wd_ = h // self.ctrl_pool[i], wd_ // self.ctrl_pool[i]
      sc, sh = _fold_bn(weights, 'ctrl_cnn', i, T, np.asarray(weights['ctrl_cnn_b_%d' % i], np.float32))
      w['ccnn_scale%d' % i] = self._dev(sc)
      w['ccnn_shift%d' % i] = self._dev(sh)
      self._load_bn(w, 'ccnn', 'ctrl_cnn', i)
    gates = 'ifou'  # kernel gate order i, f, o, u
    for dst, src in (('lstm_wx', 'ctrl_lstm_w_x'), ('lstm_wh', 'ctrl_lstm_w_h'), ('lstm_b', 'ctrl_lstm_b_')):
      st = PM.WI.stack([self._wi(src + g) for g in gates])
      w[dst] = self._dev(st)
      w[dst + '_idx'] = st.idx
    for k in ('glimpse_mlp_w_0', 'glimpse_mlp_b_0', 'glimpse_mlp_w_1', 'glimpse_mlp_b_1', 'ctrl_mlp_w_0',
              'ctrl_mlp_b_0', 'score_mlp_b_0'):
      wk = self._wi(k)
      w[k] = self._dev(wk)
      w[k + '_idx'] = wk.idx
    ws = self._wi('score_mlp_w_0').map(lambda a: a.reshape(-1))
    w['score_mlp_w_0'] = self._dev(ws)
    w['score_mlp_w_0_idx'] = ws.idx
    if w['glimpse_mlp_w_1'].shape[1] != self.P:
      raise _lib.RecAttendError('glimpse_mlp_w_1 maps to {} positions, the feature map has {}'.format(
          w['glimpse_mlp_w_1'].shape[1], self.P))
    return w

  def _load_bn(self, w, prefix, scope, i):
    """Unfolded BN parameters of layer i for the training-mode forward (batch statistics, EMA update in place)."""
    bp = _bn_params(self._wi, scope, i, self.T)
    for n, a in bp.items():
      w['{}_{}{}'.format(prefix, n, i)] = self._dev(a)
      w['{}_{}{}_idx'.format(prefix, n, i)] = a.idx
    w['{}_one{}'.format(prefix, i)] = torch.ones(bp['bias'].val.shape[0], device=self.device)
    self._bn_layers.append((prefix, scope, i))

  def _pack(self, wi, Hout, Wout, pool):
    """Register a conv-form HWIO filter (params.WI) for the tcgen05 kernel; the shared-memory image depends on the
    tile plan (hence on the batch size) and is packed on first use per batch size."""
    return {'w': np.ascontiguousarray(wi.val, dtype=np.float32), 'idx': wi.idx, 'Hout': Hout, 'Wout': Wout,
            'pool': pool, 'packed': {}}

  def _w_dev(self, wp):
    """The plain conv-form filter on the device (CUDA-core convolution, weight-gradient layout)."""
    if 'w_dev' not in wp:
      wp['w_dev'] = self._dev(PM.WI(wp['w'], wp['idx']))
    return wp['w_dev']

  def _wb_dev(self, wp):
    """The filter of the layer's data-gradient convolution: flipped, channel axes swapped."""
    if 'wb_dev' not in wp:
      wp['wb_dev'] = self._dev(PM.WI(wp['w'], wp['idx']).map(PM.flip_transpose))
    return wp['wb_dev']

  def _conv(self, x, wp, scale, shift, pool, relu=True, x2=None, upsample=1, out=None):
    """One conv block on the tensor cores (csrc/conv_umma.cu).  RA_CONV_FP32=1 (diagnostics) runs the same block
    on the CUDA cores in plain fp32 FMA arithmetic (csrc/conv.cu) - the precision reference for the 3xTF32 path."""
    if os.environ.get('RA_CONV_FP32'):
      return ops.conv3x3_block(x, self._w_dev(wp), scale, shift, pool=pool, relu=relu, x2=x2, upsample=upsample,
                               out=out)
    return ops.conv3x3_block_umma(x, self._packed(wp, x.shape[0], pool, x2), wp['w'].shape[3], scale, shift, pool=pool,
                                  relu=relu, x2=x2, upsample=upsample, out=out)

  def _packed(self, wp, B, pool, x2=None):
    """The tcgen05 filter image of a registered filter for batch size B (packed on first use: the tile plan, hence
    the image, depends on the batch size, on the pooling and on how the input channels split over x1 / x2)."""
    C2 = 0 if x2 is None else int(x2.shape[3])
    key = (B, pool, C2)
    if key not in wp['packed']:
      w = wp['w']
      KC, NPc, nsp, _, rs = ops.umma_plan(w.shape[2], w.shape[3], wp['Hout'], wp['Wout'], pool, B, C2=C2)
      if rs & 2:  # fp16 hi / lo split (RA_UMMA_F16): the registered tensor is the fp32 source, the image is made from it
        src = self._dev(PM.umma_f16_source(PM.WI(w, wp['idx']), KC, NPc, nsp))
        img = ops.umma_pack_f16(src, KC, NPc)
        self._f16_images.append((src, img, KC, NPc))
        wp['packed'][key] = img
      else:
        wp['packed'][key] = self._dev(PM.pack_umma(PM.WI(w, wp['idx']), KC, NPc, nsp, rs))
    return wp['packed'][key]

  def _use_chain(self):
    """RA_CHAIN=1 (opt-in): eval mode runs a whole conv stack of a decode step (controller layers 1-7; the 6 + 7
    patch-network layers) as ONE launch of a persistent grid (ops.ConvChain; bit-identical results).  MEASURED on B200
    (KITTI B=32): the summed conv time falls from 11.3 ms (one eager launch per layer) to 9.8 ms, but inside the CUDA
    graph the per-layer launches already overlap through programmatic dependent launch (8.4 ms effective) and the
    chain cannot overlap with its neighbours: 17.4 ms per step against 15.7 ms - so one launch per layer stays the
    default (profiles/r02f_chain.txt)."""
    # (the chain kernel runs the 3xTF32 variant only: it also needs RA_UMMA_F16=0)
    return (bool(os.environ.get('RA_CHAIN')) and not os.environ.get('RA_CONV_FP32') and int(self.n_chains) <= 1
            and (ops.umma_set_f16(-1) & 3) == 0)

  def _chain_spec(self, x, wp, scale, shift, pool, relu=True, x2=None, upsample=1, out=None):
    return {'x': x, 'x2': x2, 'wpack': self._packed(wp, x.shape[0], pool, x2), 'Cout': wp['w'].shape[3], 'scale': scale,
            'shift': shift, 'pool': pool, 'relu': relu, 'upsample': upsample, 'out': out}

  def _block(self, train, x, wp, prefix, i, t, pool, relu=True, x2=None, upsample=1, out=None):
    """One nn.cnn / nn.dcnn layer.  Eval: conv with the folded EMA batch norm in its epilogue.  Training
    (nnlib.py:96-119): raw conv + bias, then batch-statistics BN + ReLU + pool, EMA shadows of copy t moved in place."""
    w = self.w
    if not train:
      return self._conv(x, wp, w['%s_scale%d' % (prefix, i)][t], w['%s_shift%d' % (prefix, i)][t], pool, relu=relu,
                        x2=x2, upsample=upsample, out=out)
    tp = self._tape_on
    raw_out = bm = bv = None
    if tp is not None:  # train_step: the backward needs the raw conv output and the batch statistics of every layer
      raw_out, bm, bv = tp['%s_raw%d' % (prefix, i)][t], tp['%s_mean%d' % (prefix, i)][t], tp['%s_var%d' % (prefix, i)][t]
    raw = self._conv(x, wp, w['%s_one%d' % (prefix, i)], w['%s_bias%d' % (prefix, i)], 1, relu=False, x2=x2,
                     upsample=upsample, out=raw_out)
    y, _, _ = ops.batch_norm_train_block(raw, w['%s_gamma%d' % (prefix, i)][t], w['%s_beta%d' % (prefix, i)][t],
                                         w['%s_ema_mean%d' % (prefix, i)][t], w['%s_ema_var%d' % (prefix, i)][t],
                                         pool=pool, relu=relu, eps=BN_EPS, out=out, batch_mean=bm, batch_var=bv)
    return y

  def _act(self, bufs, net, i, t):
    """Activation buffer of layer i of `net` at step t: the per-step tape slot during a train_step forward (the
    backward reads every layer's input), else the buffer shared by all steps."""
    tp = self._tape_on
    if tp is not None:
      return tp['%s_out%d' % (net, i)][t]
    return bufs[net][i]

  def _filt(self, bufs, t):
    tp = self._tape_on
    if tp is not None:
      return tp['fy'][t], tp['fx'][t]
    return bufs['fy'], bufs['fx']

  def _refold_bn(self):
    """After training-mode forwards / optimiser steps: fold gamma, beta, the (moved) EMA shadows and the conv bias
    back into the eval-mode scale / shift rows (nnlib.py:113-119), on the device."""
    w = self.w
    for prefix, _, i in self._bn_layers:
      g = w['%s_gamma%d' % (prefix, i)]
      _lib.call('ra_bn_fold_f32', ops._p(g), ops._p(w['%s_beta%d' % (prefix, i)]),
                ops._p(w['%s_ema_mean%d' % (prefix, i)]), ops._p(w['%s_ema_var%d' % (prefix, i)]),
                ops._p(w['%s_bias%d' % (prefix, i)]), g.shape[0], g.shape[1], BN_EPS,
                ops._p(w['%s_scale%d' % (prefix, i)]), ops._p(w['%s_shift%d' % (prefix, i)]), ops._stream())
    self._bn_dirty = False

  def average_ema_across_ranks(self):
    """Data-parallel training normalises with per-rank batch statistics (all-reduce on gradients only),
    so the EMA shadows drift apart between ranks while the trained parameters stay identical.  Call this on EVERY
    rank before a checkpoint / evaluation: the shadows become the mean over ranks (each rank saw an equal share of
    the global batch), and the eval-mode scale / shift rows are refolded."""
    from . import dist_util
    tensors = [self.w['%s_%s%d' % (prefix, n, i)] for prefix, _, i in self._bn_layers for n in ('ema_mean', 'ema_var')]
    dist_util.average_(tensors)
    self._bn_dirty = True

  def _set_wd_term(self, weights):
    """The weight-decay part of the loss value (nnlib.py:59-61) lives in a device scalar: an optimiser step
    recomputes it on the device (ra_weight_decay_f32), captured graphs read the same address."""
    self.wd_term = self._weight_decay_term(weights)
    self.w['wd_dev'] = torch.full((1,), float(self.wd_term), device=self.device, dtype=torch.float32)

  def _add_wd(self, scal):
    """loss += the device-side weight-decay term (the loss block itself is called with 0)."""
    _lib.call('ra_add_f32', ops._p(scal[_lib.LOSS_NAMES.index('loss'):]), ops._p(self.w['wd_dev']), 1, ops._stream())

  # ------------------------------------------------------------------ optimiser <-> device weight images
  def sync_weights(self, flat_params, tmap):
    """Rewrite every registered device weight image from the flat trainable bucket (`optim.AdamOptimizer.params`)
    in ONE launch of ra_param_gather_f32 - the device-side counterpart of load_weights(export_weights())."""
    n = len(self._synced)
    if self._sync_table is None or self._sync_table['n'] != n:
      codes, starts, ptrs, keep = [], [], [], []
      total = 0
      for t, idx, kind in self._synced:
        code = PM.encode(idx, kind, tmap)
        if code is None:
          continue
        assert code.size == t.numel()
        codes.append(code)
        starts.append(total)
        ptrs.append(t.data_ptr())
        keep.append(t)
        total += code.size
      dev = self.device
      self._sync_table = {
          'n': n, 'total': total, 'nseg': len(starts), 'keep': keep,
          'codes': torch.from_numpy(np.concatenate(codes) if codes else np.zeros(0, np.int32)).to(dev),
          'starts': torch.tensor(starts, dtype=torch.int64, device=dev),
          'ptrs': torch.tensor(ptrs, dtype=torch.int64, device=dev),
      }
    tb = self._sync_table
    _lib.call('ra_param_gather_f32', ops._p(flat_params), ops._p(tb['codes']), ops._p(tb['starts']), ops._p(tb['ptrs']),
              tb['nseg'], tb['total'], ops._stream())
    for src, img, KC, NPc in self._f16_images:
      ops.umma_pack_f16(src, KC, NPc, out=img)
    self._bn_dirty = True  # gamma / beta / bias moved: refold before the next eval forward

  def _weight_decay_term(self, weights):
    """nnlib.py:59-61: sum over conv/mlp/lstm weight matrices of wd * ||w||^2 / 2 (a constant
    of the weights; computed once on the host in float64 then rounded)."""
    wd = float(self.opt['weight_decay'])
    tot = 0.0
    for k, v in weights.items():
      if '_w_' in k and not k.endswith(('_beta', '_gamma', '_ema_mean', '_ema_var')):
        v = np.asarray(v, np.float32)
        tot += wd * float(np.sum(v.astype(np.float64)**2)) / 2.0
    return tot

  def export_weights(self):
    """The weights in the reference key schema (TF layouts); the BN EMA shadows reflect every training-mode forward
    run since load_weights, the trainable tensors every train_step."""
    out = {k: np.array(v) for k, v in self._raw_weights.items()}
    if self.w is not None:
      for prefix, scope, i in self._bn_layers:
        for n in ('ema_mean', 'ema_var'):
          a = self.w['%s_%s%d' % (prefix, n, i)].cpu().numpy()
          for t in range(self.T):
            out['{}_{}_{}_{}'.format(scope, i, t, n)] = a[t].copy()
    if self._trainer is not None:  # the parameters train_step has moved
      opt_ = self._trainer.optim
      out.update({k: np.array(v) for k, v in opt_.flat.unflatten(opt_.params.cpu().numpy()).items()})
    return out

  # ------------------------------------------------------------------ shared pieces
  def _inputs(self, batch):
    def dev(v, allow_u8=False):
      # host arrays stay on the host here: forward() copies them straight into its static device buffers
      if isinstance(v, np.ndarray):
        if allow_u8 and v.dtype == np.uint8:
          return torch.from_numpy(np.ascontiguousarray(v))
        v = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
      if allow_u8 and v.dtype == torch.uint8:
        return v.contiguous()  # {0,1} masks as bytes: expanded to fp32 on the device (ra_u8_to_f32)
      if v.dtype != torch.float32:
        v = v.float()
      return v.contiguous()

    x = dev(batch['x'])
    B = x.shape[0]
    if tuple(x.shape) != (B, self.H, self.W, 3):
      raise _lib.RecAttendError('x must be [B,{},{},3], got {}'.format(self.H, self.W, tuple(x.shape)))
    d_in = y_in = None
    if self.add_d:
      d_in, y_in = dev(batch['d_in']), dev(batch['y_in'])
      if tuple(d_in.shape) != (B, self.H, self.W, 8) or tuple(y_in.shape) != (B, self.H, self.W, self.nsc):
        raise _lib.RecAttendError('d_in / y_in have the wrong shape')
    y_gt = dev(batch['y_gt'], allow_u8=True) if 'y_gt' in batch else None
    s_gt = dev(batch['s_gt']) if 's_gt' in batch else None
    return x, d_in, y_in, y_gt, s_gt

  def _stage_inputs(self, bufs, slot, tensors, stream):
    """Copy the step's inputs into static input set `slot` on `stream` (H2D for host tensors)."""
    sets = bufs.setdefault('static_in_sets', [{}, {}])
    st = sets[slot]
    with torch.cuda.stream(stream):
      for k, v in tensors.items():
        if v is None:
          continue
        if k not in st:
          st[k] = torch.empty(v.shape, device=self.device, dtype=torch.float32)
        if v.dtype == torch.uint8:
          k8 = k + '_u8'
          if k8 not in st:
            st[k8] = torch.empty(v.shape, device=self.device, dtype=torch.uint8)
          st[k8].copy_(v, non_blocking=True)
          _lib.call('ra_u8_to_f32', ops._p(st[k8]), st[k8].numel(), ops._p(st[k]), ops._stream())
        else:
          st[k].copy_(v, non_blocking=True)
    return st

  def _buffers(self, B):
    if B not in self._bufs:
      self._bufs[B] = self._alloc(B)
    return self._bufs[B]

  def _alloc_controller(self, B, bufs):
    dev, T, H, W = self.device, self.T, self.H, self.W
    f32 = torch.float32
    bufs['xs'] = torch.empty((B, H, W, self.Cs), device=dev, dtype=f32)
    c0 = self.opt['ctrl_cnn_depth'][0]
    bufs['static_pre'] = torch.empty((B, H, W, c0), device=dev, dtype=f32)
    bufs['canvas'] = torch.empty((B, H, W), device=dev, dtype=f32)
    h, w = H, W
    bufs['ccnn'] = []
    for i, (ch, pl) in enumerate(zip(self.opt['ctrl_cnn_depth'], self.ctrl_pool)):
      h, w = h // pl, w // pl
      bufs['ccnn'].append(torch.empty((B, h, w, ch), device=dev, dtype=f32))
    bufs['h_all'] = torch.empty((T, B, self.Hd), device=dev, dtype=f32)
    bufs['ctrl_out_all'] = torch.empty((T, B, 9), device=dev, dtype=f32)
    bufs['gmap_all'] = torch.empty((T, B, self.n_iter, self.P), device=dev, dtype=f32)
    bufs['box_all'] = torch.empty((T, B, _lib.BOX_STRIDE), device=dev, dtype=f32)
    bufs['fy'] = torch.empty((B, self.F, H), device=dev, dtype=f32)
    bufs['fx'] = torch.empty((B, self.F, W), device=dev, dtype=f32)
    bufs['band'] = torch.empty((B, 2, self.F, 2), device=dev, dtype=torch.int32)
    bufs['attn_box'] = torch.empty((B, T, H, W), device=dev, dtype=f32)
    bufs['s_out'] = torch.empty((B, T), device=dev, dtype=f32)

  def _prepare(self, bufs, x, d_in, y_in):
    """Step-invariant work, once per forward: the static input stack and the static half of
    the first controller layer (conv is linear in its input channels; only the canvas
    channel changes between decode steps, full_model.py:640-663)."""
    w = self.w
    _lib.TAG = 'prepare'
    ops.concat_channels(x, d_in, y_in, out=bufs['xs'])
    self._conv(bufs['xs'], w['ccnn_w0_static_umma'], w['one_c0'], w['zero_c0'], 1, relu=False,
               out=bufs['static_pre'])
    bufs['canvas'].zero_()  # full_model.py:239

  def _controller(self, bufs, t, train=False):
    """full_model.py:663-725: controller CNN (BN copy t) + glimpse LSTM + head -> box params."""
    w = self.w
    B, H, W = bufs['canvas'].shape
    tp = self._tape_on
    _lib.TAG = 'ctrl_cnn'
    if tp is not None:
      tp['canvas'][t].copy_(bufs['canvas'])  # the canvas this step starts from (glimpse / first-layer gradients)
    if not train:
      ops.canvas_conv(bufs['static_pre'], bufs['canvas'], w['ccnn_w0_canvas'], w['ccnn_scale0'][t],
                      w['ccnn_shift0'][t], pool=self.ctrl_pool[0], relu=True, out=bufs['ccnn'][0])
    else:
      # raw layer-0 output = static part + canvas part + bias at full resolution, then batch-statistics BN
      raw_out = bm = bv = None
      if tp is not None:
        raw_out, bm, bv = tp['ccnn_raw0'][t], tp['ccnn_mean0'][t], tp['ccnn_var0'][t]
      raw0 = ops.canvas_conv(bufs['static_pre'], bufs['canvas'], w['ccnn_w0_canvas'], w['ccnn_one0'], w['ccnn_bias0'],
                             pool=1, relu=False, out=raw_out)
      ops.batch_norm_train_block(raw0, w['ccnn_gamma0'][t], w['ccnn_beta0'][t], w['ccnn_ema_mean0'][t],
                                 w['ccnn_ema_var0'][t], pool=self.ctrl_pool[0], relu=True, eps=BN_EPS,
                                 out=self._act(bufs, 'ccnn', 0, t), batch_mean=bm, batch_var=bv)
    n_c = len(self.ctrl_pool)
    if not train and self._use_chain() and 'ccnn' in bufs and isinstance(bufs.get('chains', {}), dict):
      chains = bufs.setdefault('chains', {})
      if ('ctrl', t) not in chains:  # (built in the warm-up pass, outside graph capture)
        chains[('ctrl', t)] = ops.ConvChain([
            self._chain_spec(bufs['ccnn'][i - 1], w['ccnn_w%d' % i], w['ccnn_scale%d' % i][t], w['ccnn_shift%d' % i][t],
                             self.ctrl_pool[i], out=bufs['ccnn'][i]) for i in range(1, n_c)])
      chains[('ctrl', t)].run()
    else:
      for i in range(1, n_c):
        self._block(train, self._act(bufs, 'ccnn', i - 1, t), w['ccnn_w%d' % i], 'ccnn', i, t, self.ctrl_pool[i],
                    out=self._act(bufs, 'ccnn', i, t))
    feat = self._act(bufs, 'ccnn', n_c - 1, t).view(B, self.P, -1)
    _lib.TAG = 'controller'
    ops.controller_step(feat, w['lstm_wx'], w['lstm_wh'], w['lstm_b'], w['glimpse_mlp_w_0'], w['glimpse_mlp_b_0'],
                        w['glimpse_mlp_w_1'], w['glimpse_mlp_b_1'], w['ctrl_mlp_w_0'], w['ctrl_mlp_b_0'], self.H,
                        self.W, self.F, self.F, self.ctrl_flags, n_iter=self.n_iter, h_out=bufs['h_all'][t],
                        ctrl_out=bufs['ctrl_out_all'][t], glimpse_map=bufs['gmap_all'][t], box=bufs['box_all'][t])
    fy, fx = self._filt(bufs, t)
    ops.get_gaussian_filter(bufs['box_all'][t], self.H, self.W, self.F, fy=fy, fx=fx, band=bufs['band'])

  def _controller_outputs(self, bufs, out):
    box = bufs['box_all'].permute(1, 0, 2)  # [B,T,12]
    ctr = box[:, :, _lib.BOX_CTR_Y:_lib.BOX_CTR_X + 1]
    size = box[:, :, _lib.BOX_SIZE_Y:_lib.BOX_SIZE_X + 1]
    out['attn_ctr'] = ctr.contiguous()
    out['attn_size'] = size.contiguous()
    out['attn_top_left'] = box[:, :, _lib.BOX_TL_Y:_lib.BOX_TL_X + 1].contiguous()
    out['attn_bot_right'] = box[:, :, _lib.BOX_BR_Y:_lib.BOX_BR_X + 1].contiguous()
    out['attn_lg_var'] = box[:, :, _lib.BOX_LGVAR_Y:_lib.BOX_LGVAR_X + 1].contiguous()
    out['ctrl_out'] = bufs['ctrl_out_all'].permute(1, 0, 2).contiguous()
    out['h_ctrl'] = bufs['h_all'].permute(1, 0, 2).contiguous()
    B = box.shape[0]
    out['ctrl_rnn_glimpse_map'] = bufs['gmap_all'].permute(1, 0, 2, 3).reshape(B, self.T, self.n_iter, self.gh,
                                                                              self.gw)
    out['attn_box'] = bufs['attn_box']
    out['s_out'] = bufs['s_out']
    out['canvas'] = bufs['canvas'].view(B, self.H, self.W, 1)


class FullModel(_ModelBase):
  """``full_model.get_model(opt)`` replacement (eval mode: phase_train=False)."""

  def __init__(self, opt, device=None):
    super(FullModel, self).__init__(opt, device)
    o = self.opt
    for k in ('ctrl_add_inp', 'ctrl_add_canvas', 'attn_add_inp', 'attn_add_canvas'):
      if not o.get(k, True):
        raise _lib.RecAttendError('{}=False is not used by any shipped config'.format(k))
    for k in ('ctrl_add_d_out', 'ctrl_add_y_out', 'attn_add_d_out', 'attn_add_y_out', 'add_y_out'):
      if bool(o.get(k, self.add_d)) != self.add_d:
        raise _lib.RecAttendError('mixed d_out/y_out input flags are not used by any shipped config')
    for k in ('segm_loss_fn', 'box_loss_fn'):
      if o[k] not in ('iou', 'wt_cov'):  # the other switches are broken in the reference itself
        raise _lib.RecAttendError("{}: only 'iou' and 'wt_cov' are supported (SURVEY §9.13)".format(k))
    self.attn_pool = list(o['attn_cnn_pool'])
    self.dcnn_pool = list(o['attn_dcnn_pool'])
    self.skip_ch = dcnn_skip_channels(o)
    self.use_skip = bool(o.get('add_skip_conn', True))
    self.disable_overwrite = bool(o.get('disable_overwrite', True))  # full_model.py:117-120
    self.min_padding = float(o['padding'] + 4)  # full_model.py:567
    # the glimpse is stored with its channel count padded to a multiple of 4 (zero channels, zero filter rows):
    # 16-byte pixel strides are what a TMA tensor map over it needs (csrc/conv_umma.cu)
    self.Dp = (self.D + 3) // 4 * 4

  def load_weights(self, weights):
    T = self.T
    self._begin_load(weights)
    w = self._load_controller(weights)
    sz = self.F
    for i in range(len(self.attn_pool)):
      wi = self._wi('attn_cnn_w_%d' % i)
      if i == 0:
        wi = wi.map(lambda a: _pad_cin(a, self.Dp))
      w['acnn_w%d' % i] = self._pack(wi, sz, sz, self.attn_pool[i])
      sz //= self.attn_pool[i]
      sc, sh = _fold_bn(weights, 'attn_cnn', i, T, np.asarray(weights['attn_cnn_b_%d' % i], np.float32))
      w['acnn_scale%d' % i], w['acnn_shift%d' % i] = self._dev(sc), self._dev(sh)
      self._load_bn(w, 'acnn', 'attn_cnn', i)
    for i in range(len(self.dcnn_pool)):
      sz *= self.dcnn_pool[i]
      wi = self._wi('attn_dcnn_w_%d' % i).map(PM.deconv_to_conv)
      if i == len(self.attn_pool) and self.use_skip and self.skip_ch[i] > 0:
        cin = wi.val.shape[2] - self.D + self.Dp  # its skip input is the padded glimpse
        wi = wi.map(lambda a: _pad_cin(a, cin))
      w['adcnn_w%d' % i] = self._pack(wi, sz, sz, 1)
      sc, sh = _fold_bn(weights, 'attn_dcnn', i, T, np.asarray(weights['attn_dcnn_b_%d' % i], np.float32))
      w['adcnn_scale%d' % i], w['adcnn_shift%d' % i] = self._dev(sc), self._dev(sh)
      self._load_bn(w, 'adcnn', 'attn_dcnn', i)
    self.w = w
    self._set_wd_term(weights)
    return self

  def _alloc(self, B):
    dev, T, H, W, F = self.device, self.T, self.H, self.W, self.F
    f32 = torch.float32
    bufs = {}
    self._alloc_controller(B, bufs)
    bufs['extract_tmp'] = torch.empty((B * F * W * self.D,), device=dev, dtype=f32)
    bufs['x_patch_all'] = torch.empty((T, B, F, F, self.Dp), device=dev, dtype=f32)
    bufs['y_patch_all'] = torch.empty((T, B, F, F, 1), device=dev, dtype=f32)
    s = F
    bufs['acnn'] = []
    for ch, pl in zip(self.opt['attn_cnn_depth'], self.attn_pool):
      s = s // pl
      bufs['acnn'].append(torch.empty((B, s, s, ch), device=dev, dtype=f32))
    bufs['adcnn'] = []
    for ch, pl in zip(self.opt['attn_dcnn_depth'][:-1], self.dcnn_pool[:-1]):
      s = s * pl
      bufs['adcnn'].append(torch.empty((B, s, s, ch), device=dev, dtype=f32))
    bufs['y_out'] = torch.empty((B, T, H, W), device=dev, dtype=f32)
    return bufs

  def _stage_draws(self, bufs, draws):
    """Copy the scheduled-sampling draws into static device buffers (one set per batch size)."""
    st = bufs.setdefault('static_draws', {})
    for k in ('gt_knob_box', 'gt_knob_segm', 'gt_box_pad', 'gt_box_ctr_shift', 'gt_segm_noise'):
      v = draws[k]
      v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
      if k not in st:
        st[k] = torch.empty(v.shape, device=self.device, dtype=torch.float32)
      st[k].copy_(v, non_blocking=True)
    return st

  def _knob_setup(self, bufs, y_gt, draws):
    """Scheduled sampling, once per forward (full_model.py:561-625): clean GT rectangles for the greedy match, noisy
    GT boxes to mix in, the draws on the device."""
    o = self.opt
    dev = lambda a: a if isinstance(a, torch.Tensor) and a.is_cuda else self._dev(np.asarray(a, np.float32))
    B, T = y_gt.shape[0], self.T
    tl_clean, br_clean, _, rect_clean, area = ops.get_gt_box(y_gt, padding_ratio=o['attn_box_padding_ratio'],
                                                             min_padding=self.min_padding, want_box=False)
    _, _, _, rect_raw, _ = ops.get_gt_box(y_gt, padding_ratio=0.0, min_padding=0.0, want_box=False)
    ctr_n, size_n = ops.gt_attn_noise(rect_raw, area, dev(draws['gt_box_pad']).reshape(B, T).contiguous(),
                                      dev(draws['gt_box_ctr_shift']).contiguous(), self.min_padding)
    return {
        'rect': rect_clean, 'tl': tl_clean, 'br': br_clean, 'ctr': ctr_n, 'size': size_n, 'y_gt': y_gt,
        'knob_box': dev(draws['gt_knob_box']).contiguous(), 'knob_segm': dev(draws['gt_knob_segm']).contiguous(),
        'noise': dev(draws['gt_segm_noise']).contiguous(),
        'iou_steps': torch.empty((B, T, T), device=self.device, dtype=torch.float32),
        'grd': torch.empty((B, T), device=self.device, dtype=torch.float32),
        'canvas_tmp': torch.empty_like(bufs['canvas']),
    }

  def _decode(self, bufs, B, train=False, knob=None, eval_iou=None):
    """The T-step decode loop, full_model.py:638-848 (eval mode; train=True: batch-statistics BN; knob: the
    scheduled-sampling state of _knob_setup, training only; eval_iou = (tl_gt, br_gt, iou_steps, grd): the per-step
    coordinate IoUs a use_knob + use_iou_box graph feeds to the box loss in evaluation too, :750-754,926-929)."""
    w = self.w
    T, H, W, F = self.T, self.H, self.W, self.F
    thw = T * H * W
    fork_score = int(self.n_chains) <= 1  # with sub-batch chains the side streams belong to the chains
    tp = self._tape_on
    n_a = len(self.attn_pool)
    for t in range(T):
      self._controller(bufs, t, train)
      box_t = bufs['box_all'][t]
      fy, fx = self._filt(bufs, t)
      if knob is not None:
        # full_model.py:738-785: the attention-box OUTPUT comes from the controller's own box; then the matched noisy
        # GT box may replace centre / size, and the filters are rebuilt from the mixed box
        ops.paste_back(None, box_t, fy, fx, None, attn_box=bufs['attn_box'][:, t], y_out=None,
                       out_bstride=thw, band=bufs['band'])
        if tp is not None:  # the backward of the attention box runs through the PRE-mix box and filters
          tp['box_pre'][t].copy_(box_t)
          tp['fy0'][t].copy_(fy)
          tp['fx0'][t].copy_(fx)
        if self.opt.get('use_iou_box', False):  # coordinate IoU of the (pre-mix) box, full_model.py:750-754
          ops.greedy_iou_box(box_t, knob['tl'], knob['br'], knob['iou_steps'][:, t], T * T, knob['grd'])
        else:
          ops.knob_greedy_box(bufs['attn_box'][:, t], thw, knob['rect'], H, W, knob['iou_steps'][:, t], T * T,
                              knob['grd'])
        ops.knob_mix_box(box_t, knob['grd'], knob['ctr'], knob['size'], knob['knob_box'][:, t], T)
        ops.get_gaussian_filter(box_t, H, W, F, fy=fy, fx=fx, band=bufs['band'])
      if eval_iou is not None:
        ops.greedy_iou_box(box_t, eval_iou[0], eval_iou[1], eval_iou[2][:, t], T * T, eval_iou[3])
      x_patch = bufs['x_patch_all'][t]
      _lib.TAG = 'extract'
      ops.extract_patch(bufs['xs'], bufs['canvas'], self.chan_map, box_t, fy, fx, bufs['band'],
                        tmp=bufs['extract_tmp'], out=x_patch)
      _lib.TAG = 'attn_cnn'
      n_d = len(self.dcnn_pool)
      chain_mode = (not train) and self._use_chain()
      acts = [self._act(bufs, 'acnn', i, t) for i in range(n_a)]
      core = acts[-1]
      # full_model.py:797-807: skip list [None, h_acnn[4..0], x_patch]
      skips = [None] + (acts[::-1][1:] + [x_patch])
      dsts = [bufs['y_patch_all'][t] if i == n_d - 1 else self._act(bufs, 'adcnn', i, t) for i in range(n_d)]
      if chain_mode:
        # the whole patch network of this step - attention CNN (full_model.py:792) and deconv mask head (:797-807) -
        # as ONE persistent launch
        _lib.TAG = 'patch_net'
        chains = bufs.setdefault('chains', {})
        if ('patch', t) not in chains:
          layers = []
          prev = x_patch
          for i, pl in enumerate(self.attn_pool):
            layers.append(self._chain_spec(prev, w['acnn_w%d' % i], w['acnn_scale%d' % i][t], w['acnn_shift%d' % i][t],
                                           pl, out=acts[i]))
            prev = acts[i]
          for i, pl in enumerate(self.dcnn_pool):
            sk = skips[i] if (self.use_skip and self.skip_ch[i] > 0) else None
            layers.append(self._chain_spec(prev, w['adcnn_w%d' % i], w['adcnn_scale%d' % i][t],
                                           w['adcnn_shift%d' % i][t], 1, x2=sk, upsample=pl, out=dsts[i]))
            prev = dsts[i]
          chains[('patch', t)] = ops.ConvChain(layers)
        chains[('patch', t)].run()
      else:
        prev = x_patch
        for i, pl in enumerate(self.attn_pool):  # full_model.py:792
          self._block(train, prev, w['acnn_w%d' % i], 'acnn', i, t, pl, out=acts[i])
          prev = acts[i]
      # the score head (full_model.py:821-822) feeds only the loss: a parallel graph branch beside the mask head /
      # the paste-back
      score_side = self._side_stream(bufs, 2) if fork_score else None
      if score_side is None:
        ops.score(bufs['h_all'][t], core, w['score_mlp_w_0'], w['score_mlp_b_0'], bufs['s_out'][:, t], T)
      else:
        with torch.cuda.stream(score_side):
          ops.score(bufs['h_all'][t], core, w['score_mlp_w_0'], w['score_mlp_b_0'], bufs['s_out'][:, t], T)
      if not chain_mode:
        prev = core
        _lib.TAG = 'attn_dcnn'
        for i, pl in enumerate(self.dcnn_pool):
          sk = skips[i] if (self.use_skip and self.skip_ch[i] > 0) else None
          self._block(train, prev, w['adcnn_w%d' % i], 'adcnn', i, t, 1, x2=sk, upsample=pl, out=dsts[i])
          prev = dsts[i]
      _lib.TAG = 'paste_back'
      if knob is None:
        ops.paste_back(bufs['y_patch_all'][t].view(B, F, F), box_t, fy, fx, bufs['canvas'],
                       attn_box=bufs['attn_box'][:, t], y_out=bufs['y_out'][:, t], out_bstride=thw,
                       disable_overwrite=self.disable_overwrite, band=bufs['band'])
      else:
        # the mask goes through the MIXED filters; the canvas write is decided per example by the mask switch, so
        # the fused canvas update runs on a scratch copy and knob_canvas writes the real one (full_model.py:826-845)
        knob['canvas_tmp'].copy_(bufs['canvas'])
        ops.paste_back(bufs['y_patch_all'][t].view(B, F, F), box_t, fy, fx, knob['canvas_tmp'],
                       attn_box=None, y_out=bufs['y_out'][:, t], out_bstride=thw,
                       disable_overwrite=self.disable_overwrite, band=bufs['band'])
        ops.knob_canvas(knob['grd'], knob['y_gt'], knob['noise'][:, t], T * H * W, knob['knob_segm'][:, t], T,
                        bufs['y_out'][:, t], thw, bufs['canvas'])
      if score_side is not None:
        torch.cuda.current_stream().wait_stream(score_side)  # before the next step overwrites `core`

  def _side_stream(self, bufs, i):
    """Side streams become parallel branches of the captured CUDA graph; eager runs stay on one stream (the
    caching allocator would need record_stream bookkeeping for cross-stream temporaries)."""
    if not torch.cuda.is_current_stream_capturing():
      return None
    ss = bufs.setdefault('_side_streams', [])
    while len(ss) <= i:
      ss.append(torch.cuda.Stream())
    ss[i].wait_stream(torch.cuda.current_stream())
    return ss[i]

  def _chains(self, B):
    """The decode loop is a strictly serial chain of ~30 mostly latency-bound launches per step, but examples are
    independent (eval mode): the batch can be cut into `n_chains` contiguous sub-batches whose chains run as
    parallel branches of the CUDA graph.  MEASURED SLOWER on B200 at the bench config (18.1 ms -> 24.7 ms with 2
    chains, 28.3 ms with 4; profiles/r01d_*): the per-launch cost of the persistent conv kernels does not overlap
    across branches, so doubling the launches costs more than the concurrency gains.  Default 1; RA_CHAINS=n or
    `model.n_chains` enables it (kept because it is the natural way to use a second GPU-side queue later)."""
    n = max(1, int(self.n_chains))
    while n > 1 and (B % n != 0):
      n -= 1
    step = B // n
    return [(i * step, (i + 1) * step) for i in range(n)]

  def _chain_views(self, bufs, st, b0, b1):
    """Views of the per-batch buffers for examples [b0, b1) (batch-major buffers are sliced on dim 0, the
    time-major per-step records on dim 1; a step's slice stays contiguous)."""
    cb = {}
    for k in ('xs', 'static_pre', 'canvas', 'fy', 'fx', 'band', 'attn_box', 's_out', 'y_out'):
      cb[k] = bufs[k][b0:b1]
    for k in ('ccnn', 'acnn', 'adcnn'):
      cb[k] = [t[b0:b1] for t in bufs[k]]
    for k in ('h_all', 'ctrl_out_all', 'gmap_all', 'box_all', 'x_patch_all', 'y_patch_all'):
      cb[k] = bufs[k][:, b0:b1]
    per_ex = bufs['extract_tmp'].numel() // bufs['canvas'].shape[0]
    cb['extract_tmp'] = bufs['extract_tmp'][b0 * per_ex:b1 * per_ex]
    cb['static_in'] = {k: v[b0:b1] for k, v in st.items() if k in ('x', 'd_in', 'y_in')}
    return cb

  def _gt_boxes(self, bufs, y_gt, want_gt_box):
    """get_gt_attn (full_model.py:561-567) depends on y_gt only: it runs beside the decode loop."""
    o = self.opt
    side = self._side_stream(bufs, 0)
    _lib.TAG = 'loss'
    if side is None:
      return ops.get_gt_box(y_gt, padding_ratio=o['attn_box_padding_ratio'], min_padding=self.min_padding,
                            want_box=want_gt_box), None
    with torch.cuda.stream(side):
      res = ops.get_gt_box(y_gt, padding_ratio=o['attn_box_padding_ratio'], min_padding=self.min_padding,
                           want_box=want_gt_box)
    return res, side

  def _loss(self, bufs, y_gt, s_gt, out, gt, iou_box_steps=None):
    """full_model.py:916-1081 (matching on soft IoU, 'iou' losses, hard statistics)."""
    o = self.opt
    _lib.TAG = 'loss'
    cur = torch.cuda.current_stream()
    (tl, br, box_gt, rect, area), gt_side = gt
    B, T = s_gt.shape
    iou_both = torch.empty((2 * B, T, T), device=s_gt.device, dtype=torch.float32)
    # The box branch (pixel IoU of the attention boxes against the GT rectangles -> Hungarian) does not depend on the
    # mask branch (soft IoU -> Hungarian): inside a captured graph it runs as a parallel branch, so the two
    # latency-bound matchings (one warp per example) overlap each other and the mask IoU.
    def box_branch():
      if iou_box_steps is None:
        ib = ops.f_iou(bufs['attn_box'], None, b_rect=rect, out=iou_both[:B])
      else:  # use_knob: the per-step IoUs of the decode loop (full_model.py:926-929)
        iou_both[:B].copy_(iou_box_steps)
        ib = iou_both[:B]
      return ib, ops.f_segm_match(ib, s_gt)

    box_side = self._side_stream(bufs, 2)
    if box_side is not None:
      with torch.cuda.stream(box_side):
        if gt_side is not None:
          box_side.wait_stream(gt_side)
        iou_box, match_box = box_branch()
    # soft IoU (matching) + hard IoU / DICE (statistics, full_model.py:1063-1081) in one pass over y_out and y_gt on
    # the tensor cores; shapes outside that kernel's range take the CUDA-core kernel twice (the hard statistics do
    # not feed the matching: a parallel branch)
    side = None
    fused = None
    if not os.environ.get('RA_IOU_NO_UMMA'):
      fused = ops.f_iou_soft_hard(bufs['y_out'], y_gt, hard_threshold=0.5, out_soft=iou_both[B:])
    if fused is None:
      side = self._side_stream(bufs, 1)  # the chains have joined: their streams are free again
      if side is None:
        iou_hard, dice = ops.f_iou(bufs['y_out'], y_gt, hard_threshold=0.5, want_dice=True)
      else:
        with torch.cuda.stream(side):
          iou_hard, dice = ops.f_iou(bufs['y_out'], y_gt, hard_threshold=0.5, want_dice=True)
      iou_soft = ops.f_iou(bufs['y_out'], y_gt, out=iou_both[B:])
    else:
      iou_soft, iou_hard, dice = fused
    if gt_side is not None:
      cur.wait_stream(gt_side)
    match = ops.f_segm_match(iou_soft, s_gt)
    if box_side is None:
      iou_box, match_box = box_branch()
    else:
      cur.wait_stream(box_side)
    if side is not None:
      cur.wait_stream(side)
    scal = ops.loss_block(iou_box, match_box, iou_soft, match, iou_hard, dice, bufs['s_out'], s_gt, area,
                          o['loss_mix_ratio'], 0.0)
    if o['segm_loss_fn'] == 'wt_cov' or o['box_loss_fn'] == 'wt_cov':
      # full_model.py:967,1013-1014: the losses are the negative weighted coverages; the gradient coefficients
      # (GT weight at the arg-max output) take the place of the matching in the backward pass
      H_, W_ = self.H, self.W
      box_cov = segm_cov = None
      if o['box_loss_fn'] == 'wt_cov':
        box_cov = torch.empty(1, device=s_gt.device)
        out['_box_coeff'] = torch.empty_like(iou_box)
        _lib.call('ra_wt_cov_f32', ops._p(iou_box), ops._p(None), ops._p(rect), H_, W_, B, T, T, ops._p(box_cov),
                  ops._p(out['_box_coeff']), ops._stream())
      if o['segm_loss_fn'] == 'wt_cov':
        segm_cov = torch.empty(1, device=s_gt.device)
        out['_segm_coeff'] = torch.empty_like(iou_soft)
        _lib.call('ra_wt_cov_f32', ops._p(iou_soft), ops._p(area), ops._p(None), H_, W_, B, T, T, ops._p(segm_cov),
                  ops._p(out['_segm_coeff']), ops._stream())
      _lib.call('ra_loss_select_f32', ops._p(scal), ops._p(box_cov), ops._p(segm_cov), float(o['loss_mix_ratio']), 1.0,
                ops._stream())
    self._add_wd(scal)
    out.update({
        '_gt_rect': rect, 'attn_top_left_gt': tl, 'attn_bot_right_gt': br,
        'iou_soft_box_pairwise': iou_box, 'match_box': match_box, 'iou_soft_pairwise': iou_soft, 'match': match,
        'iou_hard_pairwise': iou_hard, 'loss_scalars': scal
    })
    if box_gt is not None:
      out['attn_box_gt'] = box_gt

  def _run(self, bufs, B, with_loss, want_all, train=False, draws=None, tape=False):
    """Enqueue one full forward on the current stream; returns the dict of (static) output tensors.
    tape=True (train_step): the forward records what the backward needs and the backward pass + the scatter of
    the gradients into the flat bucket are enqueued right behind it."""
    self._tape_on = bufs['tape'] if tape else None
    try:
      return self._run_inner(bufs, B, with_loss, want_all, train, draws, tape)
    finally:
      self._tape_on = None

  def _run_inner(self, bufs, B, with_loss, want_all, train, draws, tape):
    st = bufs['static_in']
    gt = self._gt_boxes(bufs, st['y_gt'], want_all) if with_loss else None
    chains = self._chains(B)
    knob = None
    eval_iou = None
    o = self.opt
    if (with_loss and not train and o.get('use_knob', False) and o.get('use_iou_box', False)):
      # a use_knob graph hands the per-step IoUs of the decode loop to the box loss in evaluation too
      # (full_model.py:926-929); with use_iou_box those are coordinate IoUs, not the pixel IoUs of attn_box
      (tl_e, br_e, _, _, _), gt_side = gt
      if gt_side is not None:
        torch.cuda.current_stream().wait_stream(gt_side)
      eval_iou = (tl_e, br_e, torch.empty((B, self.T, self.T), device=self.device, dtype=torch.float32),
                  torch.empty((B, self.T), device=self.device, dtype=torch.float32))
      chains = chains[:1] if len(chains) == 1 else [(0, B)]
    if len(chains) == 1 or train:  # batch statistics couple the examples: training mode never splits the batch
      self._prepare(bufs, st['x'], st.get('d_in'), st.get('y_in'))
      if train and draws is not None:
        knob = self._knob_setup(bufs, st['y_gt'], draws)
      self._decode(bufs, B, train, knob, eval_iou)
    else:
      cur = torch.cuda.current_stream()
      forked = []
      for i, (b0, b1) in enumerate(chains):
        cb = self._chain_views(bufs, st, b0, b1)
        cst = cb['static_in']
        side = self._side_stream(bufs, 1 + i) if i > 0 else None  # chain 0 stays on the capturing stream
        if side is None:
          self._prepare(cb, cst['x'], cst.get('d_in'), cst.get('y_in'))
          self._decode(cb, b1 - b0)
        else:
          with torch.cuda.stream(side):
            self._prepare(cb, cst['x'], cst.get('d_in'), cst.get('y_in'))
            self._decode(cb, b1 - b0)
          forked.append(side)
      for side in forked:
        cur.wait_stream(side)
    out = {}
    self._controller_outputs(bufs, out)
    out['y_out'] = bufs['y_out']
    if want_all:
      out['x_patch'] = bufs['x_patch_all'][..., :self.D].permute(1, 0, 2, 3, 4).contiguous()
      out['y_out_patch'] = bufs['y_patch_all'].permute(1, 0, 2, 3, 4).contiguous()
    if with_loss:
      iou_steps = knob['iou_steps'] if knob is not None else (eval_iou[2] if eval_iou is not None else None)
      self._loss(bufs, st['y_gt'], st['s_gt'], out, gt, iou_steps)
      scal = out['loss_scalars']
      for i, k in enumerate(LOSS_KEYS):
        out[k] = scal[i]
    if tape:
      from . import train as TR
      gs = bufs['grads']
      TR.full_model_backward(self, gs, bufs, B, out, knob)
      self._trainer.scatter(B, gs.entries)
    return out

  def train_step(self, batch, draws=None, frozen=(), use_graph=True, grad_scale=None):
    """``sess.run([loss, train_step], feed_dict)`` of runner.py:98-105 for the graph of full_model.py:1039-1057:
    training-mode forward (batch-statistics BN, EMA shadows moved; scheduled sampling with the supplied `draws`),
    backward pass, gradient all-reduce over the data-parallel ranks, clip_by_value(-1, 1), Adam(eps 1e-7) with the
    staircase learning-rate decay, and the device-side refresh of every weight image.  `frozen`: weight keys the
    freeze_* flags exclude (checkpoint.apply_pretrained); fixed at the first call.
    Returns the loss scalars of the forward (evaluated BEFORE the update, like the reference's fetch) plus
    'learn_rate' and 'global_step'."""
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')
    if self.disable_overwrite:
      raise _lib.RecAttendError('train_step: disable_overwrite=True is not used by any shipped config (SURVEY §9.7)')
    if 'y_gt' not in batch or 's_gt' not in batch:
      raise _lib.RecAttendError('train_step needs y_gt / s_gt')
    if self._trainer is None:
      from . import train as TR
      self._trainer = TR.Trainer(self, frozen=frozen)
    out = self.forward(batch, phase_train=True, use_graph=use_graph, draws=draws, _tape=True)
    lr = self._trainer.apply(grad_scale=grad_scale)
    res = {k: out[k] for k in LOSS_KEYS}
    res['learn_rate'] = lr
    res['global_step'] = self._trainer.optim.global_step
    return res

  @property
  def optimizer(self):
    """The AdamOptimizer behind train_step (None before the first step)."""
    return None if self._trainer is None else self._trainer.optim

  def prefetch(self, batch):
    """Start copying the NEXT step's inputs (pinned host memory -> the idle static input set) on a copy
    stream while the current step computes; the next ``forward(batch)`` with the same dict picks them up.
    This is the double-buffered input pipeline of the reference's ConcurrentBatchIterator
    (utils/concurrent_batch_iter.py) moved to the H2D link."""
    x, d_in, y_in, y_gt, s_gt = self._inputs(batch)
    bufs = self._buffers(x.shape[0])
    slot = 1 - bufs.get('cur_slot', 0)
    if '_copy_stream' not in bufs:
      bufs['_copy_stream'] = torch.cuda.Stream()
    cs = bufs['_copy_stream']
    # the idle set was last read by the forward before the current one: wait for it
    ev = bufs.get('slot_done', [None, None])[slot]
    if ev is not None:
      cs.wait_event(ev)
    self._stage_inputs(bufs, slot, {'x': x, 'd_in': d_in, 'y_in': y_in, 'y_gt': y_gt, 's_gt': s_gt}, cs)
    done = torch.cuda.Event()
    done.record(cs)
    # the batch OBJECT is kept (not its id, which a freed dict could hand to a new one): forward() picks the staged
    # inputs up only for this very object; it must not be mutated between prefetch() and forward()
    bufs['prefetched'] = (batch, slot, done, y_gt is not None)

  def forward(self, batch, outputs=None, phase_train=False, with_loss=True, use_graph=True, draws=None, _tape=False):
    """``sess.run([model[k] for k in outputs], feed_dict)`` of runner.py:98-105.
    Returns a dict of CUDA tensors (all keys when ``outputs`` is None).  The inputs are copied into
    static device buffers; with ``use_graph`` the ~750 kernel launches of the T-step decode + loss block are
    captured once per (batch size, output set, input set) in a CUDA graph and replayed (the returned tensors
    are the graph's static outputs: they are overwritten by the next forward of the same batch size)."""
    if self.w is None:
      raise _lib.RecAttendError('load_weights() first')
    if phase_train:
      # Training-mode forward: every BN layer normalises with the statistics of THIS batch and moves
      # the EMA shadows of its (layer, step) copy in place (nnlib.py:96-119).  The identity draw of
      # random_transformation is assumed (apply ops.random_transformation to the batch beforehand for other draws).
      # opt['use_knob'] (scheduled sampling, full_model.py:589-625,744-785,826-845) needs its random draws as inputs
      # (synthetic.make_knob_draws): Bernoulli switches, noisy-GT-box parameters and the canvas noise.
      if self.opt.get('use_knob', False):
        if draws is None:
          raise _lib.RecAttendError('use_knob=True: pass the random draws (synthetic.make_knob_draws) as `draws`')
        if 'y_gt' not in batch:
          raise _lib.RecAttendError('use_knob=True needs y_gt')
      else:
        draws = None
      self._bn_dirty = True
    elif self._bn_dirty:
      self._refold_bn()
    cur = torch.cuda.current_stream()
    pf = None
    for bb in self._bufs.values():
      if bb.get('prefetched') is not None and bb['prefetched'][0] is batch:
        pf, bufs = bb.pop('prefetched'), bb
        bufs['prefetched'] = None
    if pf is not None:
      _, slot, done, has_gt = pf
      cur.wait_event(done)
      st = bufs['static_in_sets'][slot]
      B = st['x'].shape[0]
    else:
      x, d_in, y_in, y_gt, s_gt = self._inputs(batch)
      B = x.shape[0]
      bufs = self._buffers(B)
      slot = bufs.get('cur_slot', 0)
      has_gt = y_gt is not None
      st = self._stage_inputs(bufs, slot, {'x': x, 'd_in': d_in, 'y_in': y_in, 'y_gt': y_gt, 's_gt': s_gt}, cur)
    bufs['cur_slot'] = slot
    bufs['static_in'] = st
    with_loss = bool(with_loss and has_gt)
    want = None if outputs is None else set(outputs)
    want_all = want is None or bool(want & {'x_patch', 'y_out_patch', 'attn_box_gt'})
    key = (with_loss, want_all, slot, bool(phase_train), draws is not None, bool(_tape))
    if phase_train and draws is not None:
      draws = self._stage_draws(bufs, draws)  # static device copies: the captured graph reads the same buffers
    if _tape and 'tape' not in bufs:
      from . import train as TR
      bufs['tape'] = TR.alloc_tape(self, B, full=True)
      bufs['grads'] = TR._alloc_grads(self, full=True)
    if not use_graph:
      out = self._run(bufs, B, with_loss, want_all, train=bool(phase_train), draws=draws if phase_train else None,
                      tape=_tape)
    else:
      graphs = bufs.setdefault('graphs', {})
      if key not in graphs:
        # warm-up on a side stream (lazy weight packing, cudaFuncSetAttribute, allocator), then capture
        side = torch.cuda.Stream()
        train = bool(phase_train)
        ema_keep = None
        if train:  # the warm-up and the capture run must not move the EMA shadows: only replays count
          ema_keep = {k: v.clone() for k, v in self.w.items()
                      if '_ema_' in k and isinstance(v, torch.Tensor)}
        side.wait_stream(cur)  # after the snapshot clones are enqueued on `cur`
        with torch.cuda.stream(side):
          self._run(bufs, B, with_loss, want_all, train=train, draws=draws if train else None, tape=_tape)
        cur.wait_stream(side)
        g, static_out = capture_graph(lambda: self._run(bufs, B, with_loss, want_all, train=train,
                                                        draws=draws if train else None, tape=_tape))
        if ema_keep is not None:
          for k, v in ema_keep.items():
            self.w[k].copy_(v)
        graphs[key] = (g, static_out)
      g, out = graphs[key]
      g.replay()
    ev = torch.cuda.Event()
    ev.record(cur)
    bufs.setdefault('slot_done', [None, None])[slot] = ev
    if want is not None:
      out = {k: out[k] for k in outputs}
    return out


def get_model(opt, is_training=True, device=None):
  """Same call as the reference's ``full_model.get_model(opt, is_training)``; returns the
  model object instead of a dict of TF tensors."""
  return FullModel(opt, device=device)

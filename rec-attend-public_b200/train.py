"""The TRAINING STEP on the GPU — ``sess.run([loss, train_step])`` of runner.py:98-105 for the graph built at
full_model.py:1039-1057 / box_model.py:635-652:

    grads = optimizer.compute_gradients(total_loss)      # TensorFlow autodiff of the whole T-step graph
    g     = clip_by_value(grads, -1, 1)
    train_step = optimizer.apply_gradients(g)            # Adam(eps 1e-7), staircase learning-rate decay

Here: a taped training-mode forward (batch-statistics BN, EMA shadows moved), the backward pass assembled from the
per-block kernels exactly as oracle/backward_manual.py specifies it, one flat gradient bucket, one NCCL all-reduce
(`optim.AdamOptimizer.step`), the fused clip + Adam launch and ONE gather launch that rewrites every device weight
image (incl. the hi / lo tcgen05 filter images) from the updated bucket.

Structure of the backward.  The canvas is behind tf.stop_gradient (full_model.py:846-848) and the matchings are
constants (modellib.py:11), so no gradient flows between decode steps: the T steps are processed as ONE batch of
N = T*B examples, layer by layer (every tape stack is [T,B,...], i.e. [N,...] in memory).  Shared conv / dense weights
then get their sum over steps from a single weight-gradient launch; the per-(layer, step) BN copies are the groups of
the grouped BN backward.
"""
import numpy as np
import torch

from . import _lib, ops, params as PM
from .optim import AdamOptimizer

BN_EPS = 1e-3


def _empty(shape, dev, dtype=torch.float32):
  return torch.empty(tuple(int(v) for v in shape), device=dev, dtype=dtype)


class Trainer(object):
  """Optimiser state + gradient plumbing of one model instance (FullModel or BoxModel)."""

  def __init__(self, model, frozen=()):
    self.m = model
    self.optim = AdamOptimizer(model.opt, model._raw_weights, device=model.device, frozen=frozen)
    self.tmap = PM.to_train_map(model._all_layout, self.optim.flat)
    self.grad_flat = torch.zeros_like(self.optim.params)
    self._wd_ws = torch.empty(int(_lib.lib().ra_weight_decay_workspace()), device=model.device, dtype=torch.uint8)
    self._scatter = {}  # batch size -> scatter table

  # ------------------------------------------------------------------ gradient bucket
  def scatter_table(self, B, entries):
    """entries: list of (gradient tensor in device layout, source-index array).  Built once per batch size (the
    gradient tensors are static buffers of the captured graph)."""
    if B not in self._scatter:
      codes, starts, ptrs, keep = [], [], [], []
      total = 0
      for t, idx in entries:
        code = PM.encode(idx, None, self.tmap)
        if code is None:
          continue  # frozen tensor: not in the bucket
        assert code.size == t.numel(), (code.size, tuple(t.shape))
        codes.append(code)
        starts.append(total)
        ptrs.append(t.data_ptr())
        keep.append(t)
        total += code.size
      dev = self.m.device
      allc = np.concatenate(codes) if codes else np.zeros(0, np.int32)
      covered = np.unique(allc[allc != 0] >> 2)
      self._scatter[B] = {
          'total': total, 'nseg': len(starts), 'keep': keep, 'covered': int(covered.size),
          'codes': torch.from_numpy(allc).to(dev),
          'starts': torch.tensor(starts, dtype=torch.int64, device=dev),
          'ptrs': torch.tensor(ptrs, dtype=torch.int64, device=dev),
      }
    return self._scatter[B]

  def scatter(self, B, entries):
    tb = self.scatter_table(B, entries)
    self.grad_flat.zero_()  # 'grad is None' entries stay zero (full_model.py:1051-1055)
    _lib.call('ra_param_scatter_f32', ops._p(self.grad_flat), ops._p(tb['codes']), ops._p(tb['starts']),
              ops._p(tb['ptrs']), tb['nseg'], tb['total'], ops._stream())

  # ------------------------------------------------------------------ optimiser tail
  def apply(self, grad_scale=None):
    """all-reduce + clip + Adam on the bucket, then every device weight image and the weight-decay scalar."""
    lr = self.optim.step(self.grad_flat, grad_scale=grad_scale)
    m = self.m
    m.sync_weights(self.optim.params, self.tmap)
    _lib.call('ra_weight_decay_f32', ops._p(self.optim.params), ops._p(self.optim.wd), self.optim.params.numel(),
              ops._p(self._wd_ws), ops._p(m.w['wd_dev']), ops._stream())
    return lr


# ----------------------------------------------------------------------------------------------- tape
def alloc_tape(m, B, full):
  """Per-step copies of everything the backward reads: per (layer, step) the layer input, the raw
  conv output and the batch statistics; per step the canvas it started from, the filters (and, with scheduled
  sampling, the pre-mix filters / box record)."""
  dev, T, H, W, F = m.device, m.T, m.H, m.W, m.F
  o = m.opt
  tp = {}

  def conv_stack(net, depths, pools, h, w, up=None):
    for i, (ch, pl) in enumerate(zip(depths, pools)):
      if up is not None:
        h, w = h * up[i], w * up[i]
      tp['%s_raw%d' % (net, i)] = _empty((T, B, h, w, ch), dev)
      h, w = h // pl, w // pl
      tp['%s_out%d' % (net, i)] = _empty((T, B, h, w, ch), dev)
      tp['%s_mean%d' % (net, i)] = _empty((T, ch), dev)
      tp['%s_var%d' % (net, i)] = _empty((T, ch), dev)
    return h, w

  conv_stack('ccnn', o['ctrl_cnn_depth'], m.ctrl_pool, H, W)
  tp['canvas'] = _empty((T, B, H, W), dev)
  tp['fy'] = _empty((T, B, F, H), dev)
  tp['fx'] = _empty((T, B, F, W), dev)
  if o.get('use_knob', False):
    tp['fy0'] = _empty((T, B, F, H), dev)
    tp['fx0'] = _empty((T, B, F, W), dev)
    tp['box_pre'] = _empty((T, B, _lib.BOX_STRIDE), dev)
  if full:
    s, _ = conv_stack('acnn', o['attn_cnn_depth'], m.attn_pool, F, F)
    n_d = len(m.dcnn_pool)
    conv_stack('adcnn', o['attn_dcnn_depth'], [1] * n_d, s, s, up=m.dcnn_pool)
    del tp['adcnn_out%d' % (n_d - 1)]  # the last layer writes y_patch_all
  return tp


# ----------------------------------------------------------------------------------------------- op wrappers
def _bn_bwd_grouped(raw, dy, gamma, beta, mean, var, pool, relu):
  """raw [T,B,H,W,C], dy [T*B,H/p,W/p,C] -> (d_raw [T*B,H,W,C], dgamma [T,C], dbeta [T,C])."""
  T, B, H, W, C = raw.shape
  dev = raw.device
  d_raw = _empty((T * B, H, W, C), dev)
  dgamma, dbeta = _empty((T, C), dev), _empty((T, C), dev)
  ws = ops._ws(_lib.lib().ra_bn_train_block_bwd_grouped_workspace(T, B, H, W, C, pool), dev)
  ops._chk(raw, dy, gamma, beta, mean, var)
  assert dy.numel() == T * B * (H // pool) * (W // pool) * C, (tuple(dy.shape), tuple(raw.shape), pool)
  _lib.call('ra_bn_train_block_bwd_grouped_f32', ops._p(raw), ops._p(dy), ops._p(gamma), ops._p(beta), ops._p(mean),
            ops._p(var), T, B, H, W, C, pool, 1 if relu else 0, BN_EPS, ops._p(ws), ops._p(d_raw), ops._p(dgamma),
            ops._p(dbeta), ops._stream())
  return d_raw, dgamma, dbeta


def _wgrad(x1, d_out, dw, db=None, x2=None, upsample=1, x1_bmod=0):
  """dw [3,3,C1+C2,Cout] (+ db) of the conv-form filter over N = d_out.shape[0] examples."""
  ops._chk(x1, d_out, x2, dw, db)
  N, Ho, Wo, Cout = d_out.shape
  Hin, Win = Ho // upsample, Wo // upsample
  C1 = x1.shape[-1]
  C2 = 0 if x2 is None else x2.shape[-1]
  assert tuple(dw.shape) == (3, 3, C1 + C2, Cout), (tuple(dw.shape), C1, C2, Cout)
  ws = ops._ws(_lib.lib().ra_conv3x3_bwd_weight_workspace(N, Hin, Win, C1 + C2, Cout, upsample), d_out.device)
  _lib.call('ra_conv3x3_bwd_weight_ex_f32', ops._p(x1), C1, x1_bmod, ops._p(x2), C2, ops._p(d_out), N, Hin, Win, Cout,
            upsample, ops._p(ws), ops._p(dw), ops._p(db), ops._stream())


def _dgrad(m, wp, d_raw, upsample):
  """Input gradient of the layer: SAME convolution with the flipped / transposed filter (odd positions for the
  transposed-conv layers) - on the tensor cores, like the forward layers (3xTF32; the packed image of the flipped
  filter is one more registered weight image).  d_raw [N,Ho,Wo,Cout] -> [N,Ho/up,Wo/up,Cin]."""
  if 'bwd' not in wp:
    _, Ho, Wo, _ = d_raw.shape
    wi = PM.WI(wp['w'], wp['idx']).map(PM.flip_transpose)
    wp['bwd'] = m._pack(wi, Ho, Wo, 1)
    wp['bwd_one'] = torch.ones(wi.val.shape[3], device=m.device)
    wp['bwd_zero'] = torch.zeros(wi.val.shape[3], device=m.device)
  Cin = wp['bwd']['w'].shape[3]
  # gradients span many orders of magnitude below 1: they keep the 3xTF32 operands (8-bit exponent); the fp16 hi / lo split
  # of the forward layers would lose the small ones (fp16 subnormals).  The plan - and with it the packed image of the
  # flipped filter - is made at this call, under mode 0.
  prev = ops.umma_set_f16(0)
  try:
    full = m._conv(d_raw, wp['bwd'], wp['bwd_one'], wp['bwd_zero'], 1, relu=False)
  finally:
    ops.umma_set_f16(prev)
  if upsample == 1:
    return full
  N, H2, W2, _ = full.shape
  dx = _empty((N, H2 // 2, W2 // 2, Cin), m.device)
  _lib.call('ra_subsample2_f32', ops._p(full), N, H2 // 2, W2 // 2, Cin, 1, ops._p(dx), ops._stream())
  return dx


def _split(src, C1, C2, dst2=None, accumulate2=False):
  """[npix, C1+C2] -> ([npix,C1], [npix,C2])."""
  shp = tuple(src.shape[:-1])
  npix = int(np.prod(shp))
  d1 = _empty(shp + (C1,), src.device)
  if dst2 is None:
    dst2 = _empty(shp + (C2,), src.device)
  _lib.call('ra_split_channels_f32', ops._p(src), npix, C1, C2, ops._p(d1), ops._p(dst2), 1 if accumulate2 else 0,
            ops._stream())
  return d1, dst2


def _add(dst, src):
  assert dst.numel() == src.numel()
  _lib.call('ra_add_f32', ops._p(dst), ops._p(src), dst.numel(), ops._stream())


def _outer_sum(A, a_stride, n_in, D, d_stride, n_out, R, dW, db=None):
  nb = _lib.lib().ra_outer_sum_workspace(n_in, n_out, R)  # > 0: the rows are split over chunks of CTAs
  ws = ops._ws(nb, dW.device) if nb else None
  _lib.call('ra_outer_sum_ex_f32', ops._p(A), a_stride, n_in, ops._p(D), d_stride, n_out, R, ops._p(ws), ops._p(dW),
            ops._p(db), ops._stream())


def _paste_back_bwd(d_out, out, B, T, H, W, patch, fy, fx, gamma, box, d_fy=None, d_fx=None):
  """Step-batched paste-back backward: d_out / out [B,T,H,W] read in [T,B] order; patch [N,F,F] or None;
  fy [N,F,H], fx [N,F,W]; gamma = view starting at the gain slot of the [N,RA_BOX_STRIDE] box records `box` the
  filters were built from (the kernels only walk the taps' support bands)."""
  N = T * B
  F = fy.shape[1]
  dev = fy.device
  acc = 1 if d_fy is not None else 0
  if d_fy is None:
    d_fy, d_fx = torch.empty_like(fy), torch.empty_like(fx)
  d_patch = _empty((N, F, F), dev) if patch is not None else None
  d_gamma = _empty((N,), dev)
  ws = ops._ws(_lib.lib().ra_paste_back_bwd_workspace(N, H, W, F), dev)
  _lib.call('ra_paste_back_bwd_ex_f32', ops._p(d_out), ops._p(out), T * H * W, B, H * W, ops._p(patch), ops._p(fy),
            ops._p(fx), ops._p(gamma), _lib.BOX_STRIDE, ops._p(box), N, H, W, F, acc, ops._p(ws), ops._p(d_patch),
            ops._p(d_fy), ops._p(d_fx), ops._p(d_gamma), ops._stream())
  return d_patch, d_fy, d_fx, d_gamma


class _GradSet(object):
  """Static gradient tensors in device layout + the index arrays that place them in the flat bucket."""

  def __init__(self, dev):
    self.dev = dev
    self.t = {}
    self.entries = []

  def new(self, name, shape, idx):
    t = _empty(shape, self.dev)
    assert int(np.prod(t.shape)) == np.asarray(idx).size, (name, tuple(t.shape), np.asarray(idx).shape)
    self.t[name] = t
    self.entries.append((t, np.asarray(idx).reshape(-1)))
    return t


def _alloc_grads(m, full):
  w, o, T = m.w, m.opt, m.T
  gs = _GradSet(m.device)

  def conv_net(prefix, n, first=0):
    for i in range(first, n):
      wp = w['%s_w%d' % (prefix, i)]
      gs.new('%s_w%d' % (prefix, i), wp['w'].shape, wp['idx'])
      C = wp['w'].shape[3]
      gs.new('%s_b%d' % (prefix, i), (C,), w['%s_bias%d_idx' % (prefix, i)])
      gs.new('%s_gamma%d' % (prefix, i), (T, C), w['%s_gamma%d_idx' % (prefix, i)])
      gs.new('%s_beta%d' % (prefix, i), (T, C), w['%s_beta%d_idx' % (prefix, i)])

  n_c = len(m.ctrl_pool)
  conv_net('ccnn', n_c, first=1)
  wp0 = w['ccnn_w0_static_umma']
  C0 = wp0['w'].shape[3]
  gs.new('ccnn_w0_static', wp0['w'].shape, wp0['idx'])
  gs.new('ccnn_w0_canvas', (3, 3, 1, C0), w['ccnn_w0_canvas_idx'])
  gs.new('ccnn_b0', (C0,), w['ccnn_bias0_idx'])
  gs.new('ccnn_gamma0', (T, C0), w['ccnn_gamma0_idx'])
  gs.new('ccnn_beta0', (T, C0), w['ccnn_beta0_idx'])
  for k in ('lstm_wx', 'lstm_wh', 'lstm_b', 'glimpse_mlp_w_0', 'glimpse_mlp_b_0', 'glimpse_mlp_w_1', 'glimpse_mlp_b_1',
            'ctrl_mlp_w_0', 'ctrl_mlp_b_0', 'score_mlp_w_0', 'score_mlp_b_0'):
    gs.new(k, w[k].shape, w[k + '_idx'])
  if full:
    conv_net('acnn', len(m.attn_pool))
    conv_net('adcnn', len(m.dcnn_pool))
  return gs


# ----------------------------------------------------------------------------------------------- backward
def _conv_layer_bwd(m, gs, prefix, i, x, raw, dy, mean, var, pool, x2=None, upsample=1, want_dx=True):
  """One training-mode conv block over all T*B examples (oracle.backward_manual.conv_block_bwd)."""
  w = m.w
  wp = w['%s_w%d' % (prefix, i)]
  d_raw, dgamma, dbeta = _bn_bwd_grouped(raw, dy, w['%s_gamma%d' % (prefix, i)], w['%s_beta%d' % (prefix, i)], mean, var,
                                         pool, True)
  gs.t['%s_gamma%d' % (prefix, i)].copy_(dgamma)
  gs.t['%s_beta%d' % (prefix, i)].copy_(dbeta)
  _wgrad(x, d_raw, gs.t['%s_w%d' % (prefix, i)], gs.t['%s_b%d' % (prefix, i)], x2=x2, upsample=upsample)
  return _dgrad(m, wp, d_raw, upsample) if want_dx else None


def _controller_side_bwd(m, gs, bufs, tp, B, d_box, d_gamma3, d_h, box_head):
  """Controller (box maths, head, BPTT, read-out) and the controller CNN, all steps at once."""
  w, T = m.w, m.T
  N = T * B
  n_c = len(m.ctrl_pool)
  feat = tp['ccnn_out%d' % (n_c - 1)].view(N, m.P, -1)
  res = ops.controller_bwd(feat, box_head, w['lstm_wx'], w['lstm_wh'], w['lstm_b'], w['glimpse_mlp_w_0'],
                           w['glimpse_mlp_b_0'], w['glimpse_mlp_w_1'], w['glimpse_mlp_b_1'], w['ctrl_mlp_w_0'],
                           w['ctrl_mlp_b_0'], m.H, m.W, m.ctrl_flags, d_box, d_gamma3, d_h=d_h, n_iter=m.n_iter)
  for dst, src in (('lstm_wx', 'lstm_wx'), ('lstm_wh', 'lstm_wh'), ('lstm_b', 'lstm_b'), ('glimpse_mlp_w_0', 'gmlp_w0'),
                   ('glimpse_mlp_b_0', 'gmlp_b0'), ('glimpse_mlp_w_1', 'gmlp_w1'), ('glimpse_mlp_b_1', 'gmlp_b1'),
                   ('ctrl_mlp_w_0', 'cmlp_w'), ('ctrl_mlp_b_0', 'cmlp_b')):
    gs.t[dst].copy_(res[src])
  Cf = feat.shape[2]
  dcur = res['d_feat'].view(N, m.gh, m.gw, Cf)
  for i in range(n_c - 1, 0, -1):
    x = tp['ccnn_out%d' % (i - 1)]
    x = x.view((N,) + tuple(x.shape[2:]))
    dcur = _conv_layer_bwd(m, gs, 'ccnn', i, x, tp['ccnn_raw%d' % i], dcur, tp['ccnn_mean%d' % i], tp['ccnn_var%d' % i],
                           m.ctrl_pool[i])
  # first layer: weight gradient only (image / stop-gradient canvas).  The static channels see the SAME input at every
  # step, so their gradient is one pass over sum_t d_raw[t]; the canvas channel needs the per-step inputs.
  d_raw0, dgamma, dbeta = _bn_bwd_grouped(tp['ccnn_raw0'], dcur, w['ccnn_gamma0'], w['ccnn_beta0'], tp['ccnn_mean0'],
                                          tp['ccnn_var0'], m.ctrl_pool[0], True)
  gs.t['ccnn_gamma0'].copy_(dgamma)
  gs.t['ccnn_beta0'].copy_(dbeta)
  C0 = d_raw0.shape[3]
  gsum = _empty((B, m.H, m.W, C0), m.device)
  _lib.call('ra_sum_groups_f32', ops._p(d_raw0), T, gsum.numel(), ops._p(gsum), ops._stream())
  _wgrad(bufs['xs'], gsum, gs.t['ccnn_w0_static'], gs.t['ccnn_b0'])
  _wgrad(tp['canvas'].view(N, m.H, m.W, 1), d_raw0, gs.t['ccnn_w0_canvas'])


def full_model_backward(m, gs, bufs, B, out, knob):
  """oracle.backward_manual.full_model_backward on the device; fills the gradient tensors of `gs`."""
  w, o, tp = m.w, m.opt, bufs['tape']
  T, H, W, F = m.T, m.H, m.W, m.F
  N = T * B
  dev = m.device
  st = bufs['static_in']
  y_gt = st['y_gt']
  match, match_box = out['match'], out['match_box']
  rect, tl_gt, br_gt = out['_gt_rect'], out['attn_top_left_gt'], out['attn_bot_right_gt']
  coord = knob is not None and bool(o.get('use_iou_box', False))
  _lib.TAG = 'bwd_loss'
  # 'wt_cov' losses (full_model.py:967,1013-1014): the gradient coefficients of the weighted coverage replace the
  # matching (the confidence loss below still uses the matching)
  d_y = ops.iou_loss_bwd(bufs['y_out'], out.get('_segm_coeff', match), b_masks=y_gt)
  if coord and '_box_coeff' in out:
    raise _lib.RecAttendError("box_loss_fn='wt_cov' with use_iou_box + use_knob is not supported")
  d_ab = None if coord else ops.iou_loss_bwd(bufs['attn_box'], out.get('_box_coeff', match_box), b_rect=rect)
  d_s = ops.conf_loss_bwd(bufs['s_out'], match, scale=float(o['loss_mix_ratio']))
  # ---- mask write and attention box
  _lib.TAG = 'bwd_attn'
  fy, fx = tp['fy'].view(N, F, H), tp['fx'].view(N, F, W)
  box = bufs['box_all'].view(N, _lib.BOX_STRIDE)
  box_head = tp['box_pre'].view(N, _lib.BOX_STRIDE) if knob is not None else box
  gam = lambda b_, slot: b_.view(-1)[slot:]
  d_P, d_fy, d_fx, dg_y = _paste_back_bwd(d_y, bufs['y_out'], B, T, H, W, bufs['y_patch_all'].view(N, F, F), fy, fx,
                                          gam(box, _lib.BOX_GAMMA_Y), box)
  d_fy0 = d_fx0 = None
  if knob is None:
    _, d_fy, d_fx, dg_box = _paste_back_bwd(d_ab, bufs['attn_box'], B, T, H, W, None, fy, fx,
                                            gam(box, _lib.BOX_GAMMA_BOX), box, d_fy=d_fy, d_fx=d_fx)
  elif not coord:
    fy0, fx0 = tp['fy0'].view(N, F, H), tp['fx0'].view(N, F, W)
    _, d_fy0, d_fx0, dg_box = _paste_back_bwd(d_ab, bufs['attn_box'], B, T, H, W, None, fy0, fx0,
                                              gam(box_head, _lib.BOX_GAMMA_BOX), box_head)
  else:
    dg_box = torch.zeros(N, device=dev)
  # ---- deconv mask head, last layer first
  _lib.TAG = 'bwd_dcnn'
  n_a, n_d = len(m.attn_pool), len(m.dcnn_pool)
  flat = lambda t_: t_.view((N,) + tuple(t_.shape[2:]))
  acnn_out = [flat(tp['acnn_out%d' % i]) for i in range(n_a)]
  x_patch = flat(bufs['x_patch_all'])
  skips = [None] + (acnn_out[::-1][1:] + [x_patch])
  skip_src = [None] + list(range(n_a - 2, -1, -1)) + ['x_patch']
  d_skip = {}
  dcur = d_P.view(N, F, F, 1)
  for i in range(n_d - 1, -1, -1):
    x = acnn_out[-1] if i == 0 else flat(tp['adcnn_out%d' % (i - 1)])
    sk = skips[i] if (m.use_skip and m.skip_ch[i] > 0) else None
    dx = _conv_layer_bwd(m, gs, 'adcnn', i, x, tp['adcnn_raw%d' % i], dcur, tp['adcnn_mean%d' % i],
                         tp['adcnn_var%d' % i], 1, x2=sk, upsample=m.dcnn_pool[i])
    if sk is None:
      dcur = dx
    else:
      dcur, d_skip[skip_src[i]] = _split(dx, x.shape[3], sk.shape[3])
  d_core = dcur
  # ---- score head
  _lib.TAG = 'bwd_score'
  Hd = m.Hd
  Cd = d_core[0].numel()
  dpre, d_h = _empty((N,), dev), _empty((N, Hd), dev)
  _lib.call('ra_score_bwd_f32', ops._p(bufs['s_out']), ops._p(d_s), B, T, ops._p(w['score_mlp_w_0']), Hd, Cd,
            ops._p(dpre), ops._p(d_h), ops._p(d_core), 1, ops._stream())
  g_sw = gs.t['score_mlp_w_0']
  _outer_sum(bufs['h_all'], Hd, Hd, dpre, 1, 1, N, g_sw, gs.t['score_mlp_b_0'])
  _outer_sum(acnn_out[-1], Cd, Cd, dpre, 1, 1, N, g_sw[Hd:])
  # ---- attention CNN
  _lib.TAG = 'bwd_acnn'
  dcur = d_core
  for i in range(n_a - 1, -1, -1):
    if i < n_a - 1 and i in d_skip:
      _add(dcur, d_skip[i])
    x = x_patch if i == 0 else acnn_out[i - 1]
    dcur = _conv_layer_bwd(m, gs, 'acnn', i, x, tp['acnn_raw%d' % i], dcur, tp['acnn_mean%d' % i], tp['acnn_var%d' % i],
                           m.attn_pool[i])
  d_xpatch = dcur
  if 'x_patch' in d_skip:
    _add(d_xpatch, d_skip['x_patch'])
  # ---- glimpse -> filters
  _lib.TAG = 'bwd_attn'
  Dp = x_patch.shape[3]
  dg_attn = _empty((N,), dev)
  ws = ops._ws(_lib.lib().ra_gaussian_extract_bwd_workspace(N, W, F, m.D), dev)
  _lib.call('ra_gaussian_extract_bwd_ex_f32', ops._p(bufs['xs']), m.Cs, B, ops._p(tp['canvas']), ops._p(m.chan_map),
            ops._p(fy), ops._p(fx), ops._p(gam(box, _lib.BOX_GAMMA_ATTN)), _lib.BOX_STRIDE, ops._p(box),
            ops._p(d_xpatch), ops._p(x_patch), Dp, N, H, W, F, 1, ops._p(ws), ops._p(d_fy), ops._p(d_fx),
            ops._p(dg_attn), ops._stream())
  d_box = ops.gaussian_filters_bwd(box, fy, fx, d_fy, d_fx)
  if knob is not None:
    d_pre = None if coord else ops.gaussian_filters_bwd(box_head, fy0, fx0, d_fy0, d_fx0)
    d_mixed = d_box
    d_box = _empty((N, 6), dev)
    _lib.call('ra_knob_box_bwd_f32', ops._p(d_mixed), ops._p(d_pre), ops._p(knob['knob_box']), B, T, ops._p(d_box),
              ops._stream())
    if coord:
      _lib.call('ra_iou_box_coord_bwd_f32', ops._p(box_head), ops._p(tl_gt), ops._p(br_gt), ops._p(match_box), B, T, 1.0,
                ops._p(d_box), ops._stream())
  d_gamma3 = torch.stack([dg_attn, dg_box, dg_y], 1)
  _lib.TAG = 'bwd_ctrl'
  _controller_side_bwd(m, gs, bufs, tp, B, d_box, d_gamma3, d_h, box_head)


def box_model_backward(m, gs, bufs, B, out):
  """Backward of box_model.get_model's loss (box_model.py:560-634): box loss on the per-step IoUs + confidence loss."""
  w, o, tp = m.w, m.opt, bufs['tape']
  T, H, W, F = m.T, m.H, m.W, m.F
  N = T * B
  dev = m.device
  match_box = out['match_box']
  coord = bool(o.get('use_iou_box', False))
  fy, fx = tp['fy'].view(N, F, H), tp['fx'].view(N, F, W)
  box = bufs['box_all'].view(N, _lib.BOX_STRIDE)
  _lib.TAG = 'bwd_loss'
  d_s = ops.conf_loss_bwd(bufs['s_out'], match_box, scale=1.0)
  if coord:
    d_box = torch.zeros((N, 6), device=dev)
    dg_box = torch.zeros(N, device=dev)
    _lib.call('ra_iou_box_coord_bwd_f32', ops._p(box), ops._p(out['attn_top_left_gt']), ops._p(out['attn_bot_right_gt']),
              ops._p(match_box), B, T, 1.0, ops._p(d_box), ops._stream())
  else:
    d_ab = ops.iou_loss_bwd(bufs['attn_box'], match_box, b_rect=out['_gt_rect'])
    _, d_fy, d_fx, dg_box = _paste_back_bwd(d_ab, bufs['attn_box'], B, T, H, W, None, fy, fx,
                                            box.view(-1)[_lib.BOX_GAMMA_BOX:], box)
    d_box = ops.gaussian_filters_bwd(box, fy, fx, d_fy, d_fx)
  Hd = m.Hd
  dpre, d_h = _empty((N,), dev), _empty((N, Hd), dev)
  _lib.call('ra_score_bwd_f32', ops._p(bufs['s_out']), ops._p(d_s), B, T, ops._p(w['score_mlp_w_0']), Hd, 0,
            ops._p(dpre), ops._p(d_h), ops._p(None), 0, ops._stream())
  _outer_sum(bufs['h_all'], Hd, Hd, dpre, 1, 1, N, gs.t['score_mlp_w_0'], gs.t['score_mlp_b_0'])
  zero = torch.zeros(N, device=dev)
  d_gamma3 = torch.stack([zero, dg_box, zero], 1)
  _lib.TAG = 'bwd_ctrl'
  _controller_side_bwd(m, gs, bufs, tp, B, d_box, d_gamma3, d_h, box)

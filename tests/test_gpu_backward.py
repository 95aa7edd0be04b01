"""GPU parity tests of the backward building blocks (csrc/conv_bwd.cu): the gradients of one training-mode conv block
y = pool(relu(bn_batch(conv3x3(x [,x2]) + b))) against torch.autograd through the ORACLE's forward functions
(oracle.model.conv2d_same / conv2d_transpose_same / batch_norm_train / max_pool_same) - i.e. what TensorFlow's autodiff
derives for nnlib.py:229-253 / :372-400.  fp32 CUDA-core kernels: tolerance 1e-4 of the gradient's scale."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import model as OM

pytestmark = pytest.mark.gpu

TOL = 2e-4


def _g(a):
  return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _block_oracle(x, w_tf, b, gamma, beta, up, pool, relu):
  """Forward of the block with autograd on; returns (y, raw, batch mean, batch var)."""
  raw = OM.conv2d_same(x, w_tf, b) if up == 1 else OM.conv2d_transpose_same(x, w_tf, b, up)
  p = {'gamma': gamma, 'beta': beta, 'ema_mean': torch.zeros_like(gamma), 'ema_var': torch.ones_like(gamma)}
  y, mean, var, _, _ = OM.batch_norm_train(raw, p)
  if relu:
    y = torch.relu(y)
  if pool == 2:
    y = OM.max_pool_same(y, 2)
  return y, raw, mean, var


BWD_CASES = [
    # B, H, W, C1, C2, Cout, up, pool, relu
    (2, 16, 24, 13, 0, 16, 1, 2, 1),    # first attention-CNN layer shape (odd Cin)
    (3, 12, 12, 64, 0, 96, 1, 2, 1),    # attn L5
    (2, 6, 6, 96, 0, 64, 2, 1, 1),      # dcnn L0: transposed conv, stride 2
    (2, 12, 12, 64, 64, 64, 1, 1, 1),   # dcnn L1: skip concat (two input pointers), 128 input channels
    (2, 24, 24, 32, 32, 16, 2, 1, 1),   # dcnn L4: transposed conv + skip
    (2, 48, 48, 16, 13, 1, 1, 1, 1),    # dcnn L6: one output channel
    (4, 32, 64, 16, 0, 16, 1, 2, 1),    # controller L1 shape (reduced size)
    (1, 10, 14, 140, 0, 100, 1, 1, 0),  # more than one channel block (Cin > 128, Cout > 96), no ReLU
]


@pytest.mark.parametrize('case', BWD_CASES)
def test_conv_block_train_backward(cuda, case):
  from rec_attend_b200 import ops
  from rec_attend_b200.full_model import _deconv_to_conv
  B, H, W, C1, C2, Cout, up, pool, relu = case
  Cin = C1 + C2
  rng = np.random.default_rng(abs(hash(case)) % 2**31)
  x_np = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
  shape_tf = (3, 3, Cin, Cout) if up == 1 else (3, 3, Cout, Cin)
  w_np = (rng.standard_normal(shape_tf) / np.sqrt(9 * Cin)).astype(np.float32)
  b_np = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
  g_np = rng.uniform(0.5, 1.5, Cout).astype(np.float32)
  be_np = (rng.standard_normal(Cout) * 0.5).astype(np.float32)
  x, w, b, gamma, beta = [torch.from_numpy(a).requires_grad_(True) for a in (x_np, w_np, b_np, g_np, be_np)]
  y, raw, mean, var = _block_oracle(x, w, b, gamma, beta, up, pool, bool(relu))
  dy_np = rng.standard_normal(tuple(y.shape)).astype(np.float32)
  gx, gw, gb, gg, gbe = torch.autograd.grad(y, [x, w, b, gamma, beta], torch.from_numpy(dy_np))
  w_conv = w_np if up == 1 else _deconv_to_conv(w_np)
  gw_conv = gw.numpy() if up == 1 else _deconv_to_conv(gw.numpy())  # the same linear map carries the gradient

  x1 = _g(x_np[..., :C1])
  x2 = _g(x_np[..., C1:]) if C2 else None
  out = ops.conv3x3_block_train_bwd(x1, _g(w_conv), _g(raw.detach().numpy()), _g(dy_np), _g(g_np), _g(be_np),
                                    _g(mean.detach().numpy()), _g(var.detach().numpy()), pool=pool, relu=bool(relu),
                                    x2=x2, upsample=up)
  torch.cuda.synchronize()
  assert rel_err(out['dgamma'].cpu().numpy(), gg.numpy()) < TOL
  assert rel_err(out['dbeta'].cpu().numpy(), gbe.numpy()) < TOL
  assert tuple(out['dw'].shape) == (3, 3, Cin, Cout)
  assert rel_err(out['dw'].cpu().numpy(), gw_conv) < TOL
  assert rel_err(out['dx'].cpu().numpy(), gx.numpy()) < TOL
  # the bias sits in front of a batch-statistics BN: its gradient is zero up to round-off, in both implementations
  scale = float(np.abs(out['d_raw'].cpu().numpy()).sum(axis=(0, 1, 2)).max())
  assert float(out['db'].abs().max()) <= 1e-4 * max(scale, 1.0) and float(gb.abs().max()) <= 1e-4 * max(scale, 1.0)


def test_bwd_weight_bias_and_filter_maps(cuda):
  """Without a BN behind it (the raw-gradient path): db = column sums; flip_transpose / subsample index maps."""
  from rec_attend_b200 import _lib, ops
  rng = np.random.default_rng(5)
  B, H, W, Cin, Cout = 2, 9, 11, 6, 5  # odd sizes: the kernels are per pixel
  x = rng.standard_normal((B, H, W, Cin)).astype(np.float32)
  g = rng.standard_normal((B, H, W, Cout)).astype(np.float32)
  xt, wt = torch.from_numpy(x).requires_grad_(True), torch.randn(3, 3, Cin, Cout, requires_grad=True)
  bt = torch.zeros(Cout, requires_grad=True)
  y = OM.conv2d_same(xt, wt, bt)
  gx, gw, gb = torch.autograd.grad(y, [xt, wt, bt], torch.from_numpy(g))
  dw, db = ops.conv3x3_bwd_weight(_g(x), _g(g))
  assert rel_err(dw.cpu().numpy(), gw.numpy()) < TOL and rel_err(db.cpu().numpy(), gb.numpy()) < TOL
  dx = ops.conv3x3_bwd_data(_g(g), _g(wt.detach().numpy()))
  assert rel_err(dx.cpu().numpy(), gx.numpy()) < TOL
  dw2, none = ops.conv3x3_bwd_weight(_g(x), _g(g), want_db=False)
  assert none is None and torch.equal(dw2, dw)  # fixed summation order: bit-identical from run to run
  w = rng.standard_normal((3, 3, Cin, Cout)).astype(np.float32)
  ft = ops.filter_flip_transpose(_g(w)).cpu().numpy()
  assert np.array_equal(ft, np.ascontiguousarray(w[::-1, ::-1].transpose(0, 1, 3, 2)))
  with pytest.raises(_lib.RecAttendError):
    ops.batch_norm_train_block_bwd(_g(g), _g(g[:, :4, :5]), *[_g(np.ones(Cout)) for _ in range(4)], pool=2)  # odd H, W
  empty = ops.conv3x3_bwd_weight(torch.zeros((0, 4, 4, 3), device='cuda'), torch.zeros((0, 4, 4, 2), device='cuda'))
  assert tuple(empty[0].shape) == (3, 3, 3, 2) and float(empty[0].abs().sum()) == 0.0 and float(empty[1].abs().sum()) == 0.0


@pytest.mark.parametrize('B,T,H,W', [(3, 5, 24, 40), (2, 20, 64, 128), (1, 8, 33, 47)])
def test_loss_block_backward(cuda, B, T, H, W):
  """Gradients of the matching loss (box + segm + mix * conf, full_model.py:942-1034) at the model outputs against
  torch.autograd through oracle.model.full_model_loss; the matchings are the oracle's (constants of the gradient)."""
  import rec_attend_b200 as ra
  from rec_attend_b200 import ops
  opt = ra.config.full_model_opt('kitti', H, W, T)
  rng = np.random.default_rng(B * 10 + T)
  yy, xx = np.mgrid[0:H, 0:W]
  y_gt = np.zeros((B, T, H, W), np.float32)
  s_gt = np.zeros((B, T), np.float32)
  for b in range(B):
    k = T if b == 0 else max(1, T // 2)  # the other examples have empty ground-truth slots
    for t in range(k):
      cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(3, H / 3)
      y_gt[b, t] = ((yy - cy)**2 + (xx - cx)**2 <= r * r)
      s_gt[b, t] = 1.0 if y_gt[b, t].sum() > 0 else 0.0
  y_np = np.clip(0.6 * y_gt[:, rng.permutation(T)] + 0.4 * rng.random((B, T, H, W)), 0, 1).astype(np.float32)
  box_np = rng.random((B, T, H, W)).astype(np.float32)
  s_np = rng.uniform(0.05, 0.95, (B, T)).astype(np.float32)
  y, box, s = [torch.from_numpy(a).requires_grad_(True) for a in (y_np, box_np, s_np)]
  r = OM.full_model_loss(opt, {}, {'y_out': y, 'attn_box': box, 's_out': s}, torch.from_numpy(y_gt),
                         torch.from_numpy(s_gt))
  gy, gbox, gs = torch.autograd.grad(r['loss'], [y, box, s])
  match, match_box = _g(r['match'].numpy()), _g(r['match_box'].numpy())
  _, _, _, rect, _ = ops.get_gt_box(_g(y_gt), padding_ratio=opt['attn_box_padding_ratio'],
                                    min_padding=opt['padding'] + 4, want_box=False)
  dy = ops.iou_loss_bwd(_g(y_np), match, b_masks=_g(y_gt))
  dbox = ops.iou_loss_bwd(_g(box_np), match_box, b_rect=rect)
  ds = ops.conf_loss_bwd(_g(s_np), match, scale=opt['loss_mix_ratio'])
  torch.cuda.synchronize()
  assert float(gy.abs().max()) > 0 and float(gbox.abs().max()) > 0
  assert rel_err(dy.cpu().numpy(), gy.numpy()) < TOL
  assert rel_err(dbox.cpu().numpy(), gbox.numpy()) < TOL
  assert rel_err(ds.cpu().numpy(), gs.numpy()) < TOL
  # rows without a match get an exactly zero gradient
  unmatched = (r['match'].sum(2) == 0).numpy()
  assert float(dy.cpu().numpy()[unmatched].__abs__().sum()) == 0.0
  from rec_attend_b200 import _lib
  with pytest.raises(_lib.RecAttendError):
    ops.iou_loss_bwd(_g(y_np), match)  # neither masks nor rectangles


@pytest.mark.parametrize('B,H,W,F,with_patch', [(3, 64, 128, 48, True), (2, 40, 56, 12, True), (2, 64, 128, 48, False),
                                               (1, 33, 47, 7, True)])
def test_paste_back_and_filter_backward(cuda, B, H, W, F, with_patch):
  """Gradients of out = sigmoid(gamma * (Fy P Fx^T) - 5) w.r.t. the patch, the gain and - through
  modellib.get_gaussian_filter - the box centre, size and log-variance, against torch.autograd through the oracle's
  get_gaussian_filter / extract_patch (full_model.py:728-741,810-814).  P = None is the attention-box form (ones)."""
  from rec_attend_b200 import _lib, ops
  rng = np.random.default_rng(B * 7 + F)
  ctr_np = np.stack([rng.uniform(0.3 * H, 0.7 * H, B), rng.uniform(0.3 * W, 0.7 * W, B)], 1).astype(np.float32)
  size_np = np.stack([rng.uniform(0.2 * H, 0.6 * H, B), rng.uniform(0.2 * W, 0.6 * W, B)], 1).astype(np.float32)
  lgv_np = rng.uniform(-0.5, 1.5, (B, 2)).astype(np.float32)
  lgg_np = rng.uniform(1.0, 3.0, B).astype(np.float32)
  P_np = rng.random((B, F, F)).astype(np.float32) if with_patch else np.ones((B, F, F), np.float32)
  ctr, size, lgv, lgg, P = [torch.from_numpy(a).requires_grad_(True) for a in (ctr_np, size_np, lgv_np, lgg_np, P_np)]
  f_y = OM.get_gaussian_filter(ctr[:, 0], size[:, 0], lgv[:, 0], H, F)
  f_x = OM.get_gaussian_filter(ctr[:, 1], size[:, 1], lgv[:, 1], W, F)
  V = OM.extract_patch(P.unsqueeze(3), f_y.transpose(1, 2), f_x.transpose(1, 2), 1)[..., 0]
  out = torch.sigmoid(torch.exp(lgg).view(-1, 1, 1) * V - 5.0)
  d_out_np = rng.standard_normal((B, H, W)).astype(np.float32)
  gP, gc, gs, gv, gg = torch.autograd.grad((out * torch.from_numpy(d_out_np)).sum(), [P, ctr, size, lgv, lgg])

  box = np.zeros((B, _lib.BOX_STRIDE), np.float32)
  box[:, 0:2], box[:, 2:4], box[:, 4:6] = ctr_np, size_np, lgv_np
  box[:, _lib.BOX_GAMMA_ATTN] = 1.0
  box[:, _lib.BOX_GAMMA_BOX] = np.exp(lgg_np)
  box[:, _lib.BOX_GAMMA_Y] = np.exp(lgg_np)
  box_d = _g(box)
  fy, fx, _ = ops.get_gaussian_filter(box_d, H, W, F)
  assert rel_err(fy.cpu().numpy(), f_y.detach().numpy().transpose(0, 2, 1)) < 1e-4  # tap-major forward filters
  d_patch, d_fy, d_fx, d_gamma = ops.paste_back_bwd(_g(d_out_np), _g(out.detach().numpy()), box_d, fy, fx,
                                                    _lib.BOX_GAMMA_Y if with_patch else _lib.BOX_GAMMA_BOX,
                                                    patch=_g(P_np) if with_patch else None)
  d_box = ops.gaussian_filters_bwd(box_d, fy, fx, d_fy, d_fx)
  torch.cuda.synchronize()
  tol = 1e-3
  if with_patch:
    assert rel_err(d_patch.cpu().numpy(), gP.numpy()) < tol
  else:
    assert d_patch is None
  assert rel_err(d_gamma.cpu().numpy() * np.exp(lgg_np), gg.numpy()) < tol  # dL/d lg_gamma = dL/dgamma * gamma
  db = d_box.cpu().numpy()
  assert rel_err(db[:, 0:2], gc.numpy()) < tol and rel_err(db[:, 2:4], gs.numpy()) < tol
  assert rel_err(db[:, 4:6], gv.numpy()) < tol
  # accumulation: a second consumer of the same filters adds its share
  d_fy2, d_fx2 = d_fy.clone(), d_fx.clone()
  ops.paste_back_bwd(_g(d_out_np), _g(out.detach().numpy()), box_d, fy, fx,
                     _lib.BOX_GAMMA_Y if with_patch else _lib.BOX_GAMMA_BOX, patch=_g(P_np) if with_patch else None,
                     d_fy=d_fy2, d_fx=d_fx2)
  torch.cuda.synchronize()
  assert rel_err(d_fy2.cpu().numpy(), 2 * d_fy.cpu().numpy()) < 1e-6 and rel_err(d_fx2.cpu().numpy(), 2 * d_fx.cpu().numpy()) < 1e-6


@pytest.mark.parametrize('B,H,W,F,D,cstride,with_canvas', [(2, 64, 128, 48, 13, 16, True), (3, 40, 56, 12, 4, 4, True),
                                                          (2, 32, 48, 8, 3, 4, False)])
def test_glimpse_extract_backward(cuda, B, H, W, F, D, cstride, with_canvas):
  """Gradients of x_patch = gamma_attn * Fy^T X Fx (full_model.py:788-789) w.r.t. the gain and - through the filters -
  the box centre / size / log-variance, against torch.autograd through the oracle's extract_patch.  X is split like
  the forward kernel's inputs (step-invariant channels + canvas, chan_map = the reference's concat order)."""
  from rec_attend_b200 import _lib, ops
  rng = np.random.default_rng(B * 11 + D)
  ctr_np = np.stack([rng.uniform(0.3 * H, 0.7 * H, B), rng.uniform(0.3 * W, 0.7 * W, B)], 1).astype(np.float32)
  size_np = np.stack([rng.uniform(0.2 * H, 0.6 * H, B), rng.uniform(0.2 * W, 0.6 * W, B)], 1).astype(np.float32)
  lgv_np = rng.uniform(-0.5, 1.5, (B, 2)).astype(np.float32)
  lgg_np = rng.uniform(-0.5, 0.5, B).astype(np.float32)
  X_np = rng.random((B, H, W, D)).astype(np.float32)
  G_np = rng.standard_normal((B, F, F, D)).astype(np.float32)
  ctr, size, lgv, lgg = [torch.from_numpy(a).requires_grad_(True) for a in (ctr_np, size_np, lgv_np, lgg_np)]
  f_y = OM.get_gaussian_filter(ctr[:, 0], size[:, 0], lgv[:, 0], H, F)
  f_x = OM.get_gaussian_filter(ctr[:, 1], size[:, 1], lgv[:, 1], W, F)
  xp = torch.exp(lgg).view(-1, 1, 1, 1) * OM.extract_patch(torch.from_numpy(X_np), f_y, f_x, D)
  gc, gs, gv, gg = torch.autograd.grad((xp * torch.from_numpy(G_np)).sum(), [ctr, size, lgv, lgg])

  # device inputs: canvas = reference channel 3 (after x), the rest static, like FullModel.chan_map
  if with_canvas:
    static_idx = [c for c in range(D) if c != 3]
    xs, canvas = _g(X_np[..., static_idx]), _g(X_np[..., 3])
    chan_map = torch.tensor(static_idx + [3], dtype=torch.int32, device='cuda')
  else:
    xs, canvas = _g(X_np), None
    chan_map = torch.arange(D + 1, dtype=torch.int32, device='cuda')  # [Cs+1] entries, the last one unused
  box = np.zeros((B, _lib.BOX_STRIDE), np.float32)
  box[:, 0:2], box[:, 2:4], box[:, 4:6] = ctr_np, size_np, lgv_np
  box[:, _lib.BOX_GAMMA_ATTN] = np.exp(lgg_np)
  box_d = _g(box)
  fy, fx, band = ops.get_gaussian_filter(box_d, H, W, F)
  pad = lambda a: np.concatenate([a, np.zeros(a.shape[:3] + (cstride - D,), np.float32)], 3)
  x_patch = ops.extract_patch(xs, canvas, chan_map, box_d, fy, fx, band,
                              out=torch.empty((B, F, F, cstride), device='cuda'))
  assert rel_err(x_patch.cpu().numpy()[..., :D], xp.detach().numpy()) < 1e-3  # forward sanity (patch channel order)
  d_fy, d_fx, d_gamma = ops.extract_patch_bwd(_g(pad(G_np)), x_patch, xs, canvas, chan_map, box_d, fy, fx)
  d_box = ops.gaussian_filters_bwd(box_d, fy, fx, d_fy, d_fx)
  torch.cuda.synchronize()
  tol = 1e-3
  assert rel_err(d_gamma.cpu().numpy() * np.exp(lgg_np), gg.numpy()) < tol
  db = d_box.cpu().numpy()
  assert rel_err(db[:, 0:2], gc.numpy()) < tol and rel_err(db[:, 2:4], gs.numpy()) < tol
  assert rel_err(db[:, 4:6], gv.numpy()) < tol
  # accumulate on top of an existing gradient
  d_fy2, d_fx2 = d_fy.clone(), d_fx.clone()
  ops.extract_patch_bwd(_g(pad(G_np)), x_patch, xs, canvas, chan_map, box_d, fy, fx, d_fy=d_fy2, d_fx=d_fx2)
  torch.cuda.synchronize()
  assert rel_err(d_fy2.cpu().numpy(), 2 * d_fy.cpu().numpy()) < 1e-6 and rel_err(d_fx2.cpu().numpy(), 2 * d_fx.cpu().numpy()) < 1e-6


@pytest.mark.parametrize('arch,H,W,B', [('kitti', 64, 128, 3), ('cvppp', 64, 64, 2), ('cityscapes', 128, 128, 2)])
def test_controller_backward(cuda, arch, H, W, B):
  """BPTT through the controller (read-out, LSTM, glimpse-MLP softmax x5, head, box maths; full_model.py:668-725)
  against torch.autograd through oracle.model.controller_step fed with the same feature map.  The tape kernel's
  forward is also checked against the production controller kernel."""
  from unittest import mock
  import rec_attend_b200 as ra
  from rec_attend_b200 import _lib, ops
  opt = ra.config.full_model_opt(arch, H, W, 2)
  w = ra.synthetic.make_weights(opt, seed=7)
  gh, gw = H // 32, W // 32
  P, Cf, Hd = gh * gw, opt['ctrl_cnn_depth'][-1], opt['ctrl_rnn_hid_dim']
  rng = np.random.default_rng(B + P)
  feat_np = np.abs(rng.standard_normal((B, P, Cf))).astype(np.float32)
  keys = [k for k in w if k.startswith(('ctrl_lstm', 'glimpse_mlp', 'ctrl_mlp'))]
  wt = {k: torch.from_numpy(np.asarray(v, np.float32)).requires_grad_(k in keys) for k, v in w.items()}
  feat = torch.from_numpy(feat_np).requires_grad_(True)
  with mock.patch.object(OM, 'run_cnn', lambda *a, **k: [feat.reshape(B, gh, gw, Cf)]):
    c = OM.controller_step(opt, wt, torch.zeros(B, H, W, 4), 0)
  r = {k: rng.standard_normal(tuple(c[k].shape)).astype(np.float32) for k in ('h', 'ctr', 'size', 'lg_var')}
  rg = rng.standard_normal((B, 3)).astype(np.float32)
  gam = torch.cat([torch.exp(c['lg_gamma']), torch.exp(c['box_lg_gamma']), torch.exp(c['y_lg_gamma'])], 1)
  loss = sum((c[k] * torch.from_numpy(r[k])).sum() for k in r) + (gam * torch.from_numpy(rg)).sum()
  grads = torch.autograd.grad(loss, [feat] + [wt[k] for k in keys], allow_unused=True)
  ref = dict(zip(['feat'] + keys, [g.numpy() for g in grads]))

  flags = 0
  if opt.get('squash_ctrl_params', False):
    flags |= _lib.CTRL_SQUASH
  if opt.get('fixed_var', False):
    flags |= _lib.CTRL_FIXED_VAR
  if opt.get('dynamic_var', False):
    flags |= _lib.CTRL_DYNAMIC_VAR
  if opt.get('fixed_gamma', False):
    flags |= _lib.CTRL_FIXED_GAMMA
  gates = 'ifou'
  dw = {'lstm_wx': _g(np.stack([w['ctrl_lstm_w_x' + g] for g in gates])),
        'lstm_wh': _g(np.stack([w['ctrl_lstm_w_h' + g] for g in gates])),
        'lstm_b': _g(np.stack([w['ctrl_lstm_b_' + g] for g in gates])),
        'gmlp_w0': _g(w['glimpse_mlp_w_0']), 'gmlp_b0': _g(w['glimpse_mlp_b_0']), 'gmlp_w1': _g(w['glimpse_mlp_w_1']),
        'gmlp_b1': _g(w['glimpse_mlp_b_1']), 'cmlp_w': _g(w['ctrl_mlp_w_0']), 'cmlp_b': _g(w['ctrl_mlp_b_0'])}
  order = ('lstm_wx', 'lstm_wh', 'lstm_b', 'gmlp_w0', 'gmlp_b0', 'gmlp_w1', 'gmlp_b1', 'cmlp_w', 'cmlp_b')
  h_fwd, ctrl_fwd, _, box = ops.controller_step(_g(feat_np), *[dw[k] for k in order], H, W, opt['filter_height'],
                                                opt['filter_width'], flags)
  d_box = _g(np.concatenate([r['ctr'], r['size'], r['lg_var']], 1))
  out = ops.controller_bwd(_g(feat_np), box, *[dw[k] for k in order], H, W, flags, d_box, _g(rg), d_h=_g(r['h']))
  torch.cuda.synchronize()
  # the tape's forward = the oracle's = the production kernel's
  assert rel_err(out['h'].cpu().numpy(), c['h'].detach().numpy()) < 1e-4
  assert rel_err(out['h'].cpu().numpy(), h_fwd.cpu().numpy()) < 1e-4
  assert rel_err(out['ctrl_out'].cpu().numpy(), ctrl_fwd.cpu().numpy()) < 1e-4
  tol = 1e-3
  assert rel_err(out['d_feat'].cpu().numpy(), ref['feat']) < tol
  for gi, g in enumerate(gates):
    assert rel_err(out['lstm_wx'][gi].cpu().numpy(), ref['ctrl_lstm_w_x' + g]) < tol, g
    assert rel_err(out['lstm_wh'][gi].cpu().numpy(), ref['ctrl_lstm_w_h' + g]) < tol, g
    assert rel_err(out['lstm_b'][gi].cpu().numpy(), ref['ctrl_lstm_b_' + g]) < tol, g
  for ours, theirs in (('gmlp_w0', 'glimpse_mlp_w_0'), ('gmlp_b0', 'glimpse_mlp_b_0'), ('gmlp_w1', 'glimpse_mlp_w_1'),
                       ('gmlp_b1', 'glimpse_mlp_b_1'), ('cmlp_w', 'ctrl_mlp_w_0'), ('cmlp_b', 'ctrl_mlp_b_0')):
    assert rel_err(out[ours].cpu().numpy(), ref[theirs]) < tol, ours

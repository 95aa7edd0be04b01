"""ORACLE (test infrastructure only): the backward pass of the training step ASSEMBLED BY HAND from the same per-block
formulas the CUDA building blocks implement (csrc/conv_bwd.cu, loss_bwd.cu, attn_bwd.cu, ctrl_bwd.cu), chained over
the T-step decode loop.  It is the executable specification of the GPU assembly that does not exist yet (DESIGN.md
§4.8b): every function below maps to one C-ABI entry point, and the chaining (skip-connection routing, channel
order of the first controller layer, accumulation of the filter gradients over their three consumers, the
scheduled-sampling mixing factors, per-(layer, step) BN parameters vs. shared conv weights) is checked end to end
against oracle/grads.py (autograd through the same forward) in tests/test_backward_manual.py.

Key structural fact: the canvas is behind tf.stop_gradient (full_model.py:846-848) and the matchings are constants,
so NO gradient flows between decode steps - the backward of step t needs only d y_out[t], d attn_box[t], d s_out[t].

numpy, float64-capable (the dtype follows the inputs).  Forward intermediates come from oracle.model's TAPE.
"""
import numpy as np
import torch

from . import model as OM

BN_EPS = OM.BN_EPS


def _np(t):
  return t.detach().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def deconv_to_conv(w):
  """[kh,kw,Cout,Cin] (TF conv2d_transpose layout) <-> the conv-form HWIO filter over the zero-inserted input:
  out[ky,kx,a,b] = w[2-ky,2-kx,b,a].  The same map carries gradients, and it is its own inverse up to the swap."""
  return np.ascontiguousarray(w[::-1, ::-1].transpose(0, 1, 3, 2))


# ----------------------------------------------------------------------------- conv block  (csrc/conv_bwd.cu)
def bn_train_block_bwd(raw, dy, gamma, beta, mean, var, pool, relu):
  """ra_bn_train_block_bwd_f32: dy [B,H/p,W/p,C] -> (d_raw, dgamma, dbeta)."""
  B, H, W, C = raw.shape
  rstd = 1.0 / np.sqrt(var + BN_EPS)
  inv = gamma * rstd
  z = raw * inv + (beta - mean * inv)
  if pool == 2:
    zw = z.reshape(B, H // 2, 2, W // 2, 2, C).transpose(0, 1, 3, 5, 2, 4).reshape(B, H // 2, W // 2, C, 4)
    pos = zw.argmax(axis=4)  # first maximum in (py, px) row-major order
    best = np.take_along_axis(zw, pos[..., None], 4)[..., 0]
    g = np.where(best > 0, dy, 0.0) if relu else dy
    onehot = (np.arange(4) == pos[..., None]) * g[..., None]
    dbn = onehot.reshape(B, H // 2, W // 2, C, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(B, H, W, C)
  else:
    dbn = np.where(z > 0, dy, 0.0) if relu else dy
  xhat = (raw - mean) * rstd
  dbeta = dbn.sum(axis=(0, 1, 2))
  dgamma = (dbn * xhat).sum(axis=(0, 1, 2))
  n = B * H * W
  return inv / n * (n * dbn - dbeta - xhat * dgamma), dgamma, dbeta


def conv3x3_bwd_weight(x, d_raw, up):
  """ra_conv3x3_bwd_weight_f32: conv-form dW [3,3,Cin,Cout] and db; a tap reads Z[Y+ky-up, X+kx-up], Z[2y,2x] = x."""
  B, H, W, Cin = x.shape
  Ho, Wo = H * up, W * up
  Z = np.zeros((B, Ho, Wo, Cin), x.dtype)
  Z[:, ::up, ::up] = x
  Zp = np.pad(Z, ((0, 0), (up, 2), (up, 2), (0, 0)))
  dW = np.empty((3, 3, Cin, d_raw.shape[3]), x.dtype)
  for ky in range(3):
    for kx in range(3):
      dW[ky, kx] = np.einsum('byxi,byxo->io', Zp[:, ky:ky + Ho, kx:kx + Wo], d_raw)
  return dW, d_raw.sum(axis=(0, 1, 2))


def conv3x3_bwd_data(d_raw, w_conv, up):
  """ra_filter_flip_transpose_f32 + ra_conv3x3_f32 (+ ra_subsample2_f32 with off = 1 for up = 2)."""
  wb = np.ascontiguousarray(w_conv[::-1, ::-1].transpose(0, 1, 3, 2))
  dt = torch.from_numpy(np.ascontiguousarray(d_raw))
  full = OM.conv2d_same(dt, torch.from_numpy(wb).to(dt.dtype), torch.zeros(wb.shape[3], dtype=dt.dtype)).numpy()
  return full if up == 1 else full[:, 1::2, 1::2]


def conv_block_bwd(rec, dy, w_conv, gamma, beta, pool, relu, up, want_dx=True):
  """ops.conv3x3_block_train_bwd on one tape record {x, skip, raw, mean, var}."""
  x = _np(rec['x']) if rec['skip'] is None else np.concatenate([_np(rec['x']), _np(rec['skip'])], 3)
  d_raw, dgamma, dbeta = bn_train_block_bwd(_np(rec['raw']), dy, gamma, beta, _np(rec['mean']), _np(rec['var']), pool,
                                            relu)
  dW, db = conv3x3_bwd_weight(x, d_raw, up)
  dx = conv3x3_bwd_data(d_raw, w_conv, up) if want_dx else None
  return dx, dW, db, dgamma, dbeta


# ----------------------------------------------------------------------------- loss block  (csrc/loss_bwd.cu)
def iou_loss_bwd(a, g, match, scale=1.0):
  """ra_iou_loss_bwd_f32: gradient of -(scale/B) sum_b 1/cnt_b sum_nm match * I/U w.r.t. a [B,N,H,W]; g [B,M,H,W]."""
  B, N, H, W = a.shape
  da = np.zeros_like(a)
  for b in range(B):
    cnt = max(1.0, float(match[b].sum()))
    for n, m in zip(*np.nonzero(match[b])):
      inter = (a[b, n] * g[b, m]).sum()
      U = a[b, n].sum() + g[b, m].sum() - inter + H * W * 1e-5
      w = scale * match[b, n, m] / (B * cnt)
      da[b, n] += -w * (U + inter) / U**2 * g[b, m] + w * inter / U**2
  return da


def conf_loss_bwd(s, match, scale=1.0):
  """ra_conf_loss_bwd_f32."""
  B, T = s.shape
  ds = np.zeros_like(s)
  ms = match.sum(axis=2)
  for b in range(B):
    run, arg = np.inf, 0
    for t in range(T):
      if s[b, t] <= run:
        run, arg = s[b, t], t
      ds[b, arg] += scale / (B * T) * (-ms[b, t] / (run + 1e-5))
    run, arg = -np.inf, T - 1
    for t in range(T - 1, -1, -1):
      if s[b, t] >= run:
        run, arg = s[b, t], t
      ds[b, arg] += scale / (B * T) * ((1.0 - ms[b, t]) / (1.0 - run + 1e-5))
  return ds


# ----------------------------------------------------------------------------- attention  (csrc/attn_bwd.cu)
def paste_back_bwd(d_out, out, P, fy, fx, gamma):
  """ra_paste_back_bwd_f32: d_out, out [B,H,W]; P [B,F,F] or None (ones); fy [B,H,F], fx [B,W,F] (oracle layout).
  Returns (d_P or None, d_fy, d_fx, d_gamma [B])."""
  B = d_out.shape[0]
  Fh = fy.shape[2]
  if P is None:
    Pm = np.ones((B, Fh, fx.shape[2]), d_out.dtype)
  else:
    Pm = P
  R = np.einsum('bij,bxj->bix', Pm, fx)
  V = np.einsum('byi,bix->byx', fy, R)
  dZ = d_out * out * (1.0 - out)
  d_gamma = (dZ * V).sum(axis=(1, 2))
  dV = gamma[:, None, None] * dZ
  d_fy = np.einsum('byx,bix->byi', dV, R)
  A = np.einsum('byi,byx->bix', fy, dV)
  d_fx = np.einsum('bix,bij->bxj', A, Pm)
  d_P = np.einsum('bix,bxj->bij', A, fx) if P is not None else None
  return d_P, d_fy, d_fx, d_gamma


def extract_bwd(G, X, fy, fx, gamma, x_patch):
  """ra_gaussian_extract_bwd_f32: G = d x_patch [B,F,F,D]; X [B,H,W,D] -> (d_fy, d_fx, d_gamma)."""
  T = np.einsum('byi,byxc->bixc', fy, X)
  d_fx = gamma[:, None, None] * np.einsum('bijc,bixc->bxj', G, T)
  S = gamma[:, None, None, None] * np.einsum('bijc,bxj->bixc', G, fx)
  d_fy = np.einsum('bixc,byxc->byi', S, X)
  d_gamma = (G * x_patch).sum(axis=(1, 2, 3)) / gamma
  return d_fy, d_fx, d_gamma


def filters_bwd(ctr, size, lg_var, filt, d_filt):
  """ra_gaussian_filters_bwd_f32 for ONE axis: filt, d_filt [B,L,F] -> (d_ctr, d_size, d_lg_var) [B]."""
  B, L, F = filt.shape
  tap = np.arange(F)[None, None, :] - (F - 1) / 2.0
  mu = ctr[:, None, None] + (size[:, None, None] + 1.0) / F * tap
  s2 = np.exp(lg_var)[:, None, None]
  d = np.arange(L)[None, :, None] - mu
  t = d_filt * filt
  dmu = t * d / s2
  return dmu.sum(axis=(1, 2)), (dmu * tap / F).sum(axis=(1, 2)), (t * (d * d / (2.0 * s2) - 0.5)).sum(axis=(1, 2))


def iou_box_coord_bwd(ctr, size, tl_gt, br_gt, wgt):
  """Gradient of sum_m wgt[b,m] * modellib.f_iou_box(box_b, gt_m) (modellib.py:206-238) w.r.t. the box centre and
  size [B,2] (tl = ctr - size/2, br = ctr + size/2).  No CUDA twin yet: a [B,T] elementwise pass."""
  tl, br = (ctr - size / 2.0)[:, None, :], (ctr + size / 2.0)[:, None, :]  # [B,1,2]
  lo, hi = np.maximum(tl, tl_gt), np.minimum(br, br_gt)  # [B,M,2]
  ext = hi - lo
  flag = (ext[..., 0] > 0) & (ext[..., 1] > 0)
  inter = np.where(flag, ext[..., 0] * ext[..., 1], 0.0)
  area_a = (br - tl)[..., 0] * (br - tl)[..., 1]
  area_b = (br_gt - tl_gt)[..., 0] * (br_gt - tl_gt)[..., 1]
  union = area_a + area_b - inter
  d_iou = wgt
  d_inter = d_iou * (1.0 / union + inter / union**2)  # d(I/U)/dI with U = A + B - I
  d_area_a = -d_iou * inter / union**2
  d_tl, d_br = np.zeros(tl.shape[:1] + (2,), ctr.dtype), np.zeros(tl.shape[:1] + (2,), ctr.dtype)
  for ax in range(2):
    other = 1 - ax
    d_ext = np.where(flag, d_inter * ext[..., other], 0.0)  # [B,M]
    # hi = min(br, br_gt), lo = max(tl, tl_gt): the gradient goes to the box only where the box is the binding side
    d_br[:, ax] += (d_ext * (br[..., ax] <= br_gt[..., ax])).sum(axis=1)
    d_tl[:, ax] += (-d_ext * (tl[..., ax] >= tl_gt[..., ax])).sum(axis=1)
    side = (br - tl)[..., other]  # [B,1]
    d_br[:, ax] += (d_area_a * side).sum(axis=1)
    d_tl[:, ax] += (-d_area_a * side).sum(axis=1)
  return d_tl + d_br, (d_br - d_tl) / 2.0


# ----------------------------------------------------------------------------- controller  (csrc/ctrl_bwd.cu)
def _sig(x):
  return 1.0 / (1.0 + np.exp(-x))


def controller_bwd(opt, w, feat, ctrl_out, size, gam3, d_box6, d_gamma3, d_h):
  """ra_controller_tape_f32 + ra_controller_head_bwd_f32 + ra_controller_bwd_f32 + ra_outer_sum_f32.
  feat [B,P,Cf]; size [B,2] (the controller's own); gam3 [B,3] = exp'd gains; d_box6 = (d_ctr, d_size, d_lg_var).
  Returns (d_feat, dict of weight gradients by reference key)."""
  H, W = opt['inp_height'], opt['inp_width']
  B, P, Cf = feat.shape
  Hd, n_iter = opt['ctrl_rnn_hid_dim'], opt['num_ctrl_rnn_iter']
  gates = 'ifou'
  wx = [w['ctrl_lstm_w_x' + g] for g in gates]
  wh = [w['ctrl_lstm_w_h' + g] for g in gates]
  bg = [w['ctrl_lstm_b_' + g] for g in gates]
  w0, b0, w1, b1 = w['glimpse_mlp_w_0'], w['glimpse_mlp_b_0'], w['glimpse_mlp_w_1'], w['glimpse_mlp_b_1']
  cw = w['ctrl_mlp_w_0']
  # tape (forward recompute)
  tape = []
  h = np.zeros((B, Hd), feat.dtype)
  cc = np.zeros((B, Hd), feat.dtype)
  mp = np.full((B, P), 1.0 / P, feat.dtype)
  for _ in range(n_iter):
    rec = {'map': mp, 'h_prev': h, 'c_prev': cc}
    gl = np.einsum('bpc,bp->bc', feat, mp)
    pre = [gl @ wx[g] + h @ wh[g] + bg[g] for g in range(4)]
    gi, gf, go, gu = _sig(pre[0]), _sig(pre[1]), _sig(pre[2]), np.tanh(pre[3])
    cc = gf * cc + gi * gu
    h = go * np.tanh(cc)
    a1 = np.maximum(h @ w0 + b0, 0.0)
    lg = a1 @ w1 + b1
    e = np.exp(lg - lg.max(axis=1, keepdims=True))
    mp = e / e.sum(axis=1, keepdims=True)
    rec.update(glimpse=gl, gates=(gi, gf, go, gu), c=cc, h=h, a1=a1, map_next=mp)
    tape.append(rec)
  # head / box maths
  S = np.array([H, W], feat.dtype)
  d_ctr, d_size, d_lgv = d_box6[:, 0:2], d_box6[:, 2:4].copy(), d_box6[:, 4:6]
  d = np.zeros((B, 9), feat.dtype)
  if opt.get('dynamic_var', False):
    d[:, 4:6] = d_lgv
  elif not opt.get('fixed_var', False):
    d_size = d_size + d_lgv / size
  d_n, d_l = d_ctr * S / 2.0, d_size * size
  if opt.get('squash_ctrl_params', False):
    t = np.tanh(ctrl_out[:, 0:2])
    d_n = d_n * (1.0 - t * t)
    d_l = d_l * -_sig(ctrl_out[:, 2:4])
  d[:, 0:2], d[:, 2:4] = d_n, d_l
  fixed = bool(opt.get('fixed_gamma', False))
  d[:, 6] = 0.0 if fixed else d_gamma3[:, 0] * gam3[:, 0]
  d[:, 7] = d_gamma3[:, 1] * gam3[:, 1]
  d[:, 8] = 0.0 if fixed else d_gamma3[:, 2] * gam3[:, 2]
  # BPTT
  dh = d_h + d @ cw.T
  dc = np.zeros((B, Hd), feat.dtype)
  d_feat = np.zeros_like(feat)
  dG, dA1, dLog = [None] * n_iter, [None] * n_iter, [None] * n_iter
  dmap = None
  for k in range(n_iter - 1, -1, -1):
    rec = tape[k]
    if k < n_iter - 1:
      m = rec['map_next']
      dlog = m * (dmap - (m * dmap).sum(axis=1, keepdims=True))
      da1 = (dlog @ w1.T) * (rec['a1'] > 0)
      dh = dh + da1 @ w0.T
    else:
      dlog, da1 = np.zeros((B, P), feat.dtype), np.zeros((B, Hd), feat.dtype)
    dLog[k], dA1[k] = dlog, da1
    gi, gf, go, gu = rec['gates']
    tc = np.tanh(rec['c'])
    dcc = dc + dh * go * (1.0 - tc * tc)
    p = [dcc * gu * gi * (1.0 - gi), dcc * rec['c_prev'] * gf * (1.0 - gf), dh * tc * go * (1.0 - go),
         dcc * gi * (1.0 - gu * gu)]
    dG[k] = p
    dc = dcc * gf
    dh = sum(p[g] @ wh[g].T for g in range(4))
    dgl = sum(p[g] @ wx[g].T for g in range(4))
    d_feat += rec['map'][:, :, None] * dgl[:, None, :]
    dmap = np.einsum('bpc,bc->bp', feat, dgl)
  grads = {}
  for gi_, g in enumerate(gates):
    grads['ctrl_lstm_w_x' + g] = sum(tape[k]['glimpse'].T @ dG[k][gi_] for k in range(n_iter))
    grads['ctrl_lstm_w_h' + g] = sum(tape[k]['h_prev'].T @ dG[k][gi_] for k in range(n_iter))
    grads['ctrl_lstm_b_' + g] = sum(dG[k][gi_].sum(axis=0) for k in range(n_iter))
  grads['glimpse_mlp_w_0'] = sum(tape[k]['h'].T @ dA1[k] for k in range(n_iter))
  grads['glimpse_mlp_b_0'] = sum(dA1[k].sum(axis=0) for k in range(n_iter))
  grads['glimpse_mlp_w_1'] = sum(tape[k]['a1'].T @ dLog[k] for k in range(n_iter))
  grads['glimpse_mlp_b_1'] = sum(dLog[k].sum(axis=0) for k in range(n_iter))
  grads['ctrl_mlp_w_0'] = tape[-1]['h'].T @ d
  grads['ctrl_mlp_b_0'] = d.sum(axis=0)
  return d_feat, grads


# ----------------------------------------------------------------------------- assembly over the decode loop
def _acc(grads, key, val):
  grads[key] = val if key not in grads else grads[key] + val


def full_model_backward(opt, weights, batch, draws=None, dtype=np.float32, model_module=OM):
  """Gradients of the DATA loss (box + segm + mix * conf; the weight-decay term is the optimiser's, wd * w) of one
  training-mode forward, by reference weight key - the quantity oracle.grads.full_model_grads(...,
  include_weight_decay=False) returns, assembled block by block.  Returns (grads, forward outputs)."""
  w = {k: np.asarray(v, dtype) for k, v in weights.items()}
  tape = {}
  model_module.TAPE = tape
  try:
    with torch.no_grad():
      out = model_module.full_model_forward(opt, w, {k: np.asarray(v, dtype) for k, v in batch.items()},
                                            phase_train=True,
                                            draws=None if draws is None else {k: np.asarray(v, dtype)
                                                                              for k, v in draws.items()})
  finally:
    model_module.TAPE = None
  T, H, W = opt['timespan'], opt['inp_height'], opt['inp_width']
  B = _np(out['y_out']).shape[0]
  y_gt = np.asarray(batch['y_gt'], dtype)
  # ---- loss block: gradients at the model outputs (ra_iou_loss_bwd_f32 x2, ra_conf_loss_bwd_f32)
  match, match_box = _np(out['match']).astype(dtype), _np(out['match_box']).astype(dtype)
  d_y = iou_loss_bwd(_np(out['y_out']).astype(dtype), y_gt, match)
  use_knob = 'iou_soft_box_steps' in out
  coord_box_loss = use_knob and bool(opt.get('use_iou_box', False))
  # with use_knob the box loss uses the per-step IoUs of the decode loop (full_model.py:926-929): the same soft IoU of
  # attn_box[t] against the same clean GT boxes (identical get_gt_box arguments, :561-567), so the same derivative -
  # unless use_iou_box, where those IoUs are modellib.f_iou_box of the controller's (ctr, size): the attention box
  # then gets no gradient and the box loss reaches the controller through the coordinates (iou_box_coord_bwd below)
  box_gt = _np(out['attn_box_gt']).astype(dtype)
  if coord_box_loss:
    d_box_out = np.zeros_like(_np(out['attn_box']).astype(dtype))
    tl_gt, br_gt = _np(out['attn_top_left_gt']).astype(dtype), _np(out['attn_bot_right_gt']).astype(dtype)
    cnt_box = np.maximum(match_box.sum(axis=(1, 2)), 1.0)
  else:
    d_box_out = iou_loss_bwd(_np(out['attn_box']).astype(dtype), box_gt, match_box)
  d_s = conf_loss_bwd(_np(out['s_out']).astype(dtype), match, scale=opt['loss_mix_ratio'])
  grads = {}
  n_c, n_a, n_d = len(opt['ctrl_cnn_filter_size']), len(opt['attn_cnn_filter_size']), len(opt['attn_dcnn_filter_size'])
  add_skip = bool(opt.get('add_skip_conn', True))
  Hd = opt['ctrl_rnn_hid_dim']
  for tt in range(T):  # independent steps (stop-gradient canvas)
    st = {k: (_np(v).astype(dtype) if isinstance(v, torch.Tensor) else v) for k, v in tape[('step', tt)].items()}
    f_y, f_x = st['f_y'], st['f_x']
    g_attn, g_box, g_y = np.exp(st['lg_gamma'][:, 0]), np.exp(st['box_lg_gamma'][:, 0]), np.exp(st['y_lg_gamma'][:, 0])
    # mask write: y_out = sigmoid(g_y * Fy P Fx^T - 5)
    d_P, d_fy, d_fx, dg_y = paste_back_bwd(d_y[:, tt], st['y_out'][:, 0], st['y_patch'][..., 0], f_y, f_x, g_y)
    # attention box: written from the controller's OWN box, before the scheduled-sampling mix (full_model.py:738-741)
    if use_knob:
      to_t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
      f_y0 = _np(model_module.get_gaussian_filter(to_t(st['ctr_ctrl'][:, 0]), to_t(st['size_ctrl'][:, 0]),
                                                  to_t(st['lg_var'][:, 0]), H, opt['filter_height'])).astype(dtype)
      f_x0 = _np(model_module.get_gaussian_filter(to_t(st['ctr_ctrl'][:, 1]), to_t(st['size_ctrl'][:, 1]),
                                                  to_t(st['lg_var'][:, 1]), W, opt['filter_width'])).astype(dtype)
      _, d_fy0, d_fx0, dg_box = paste_back_bwd(d_box_out[:, tt], st['attn_box'][:, 0], None, f_y0, f_x0, g_box)
    else:
      _, d_fy_b, d_fx_b, dg_box = paste_back_bwd(d_box_out[:, tt], st['attn_box'][:, 0], None, f_y, f_x, g_box)
      d_fy, d_fx = d_fy + d_fy_b, d_fx + d_fx_b
    # deconv mask head, last layer first; skips hand their share back to the attention CNN / the glimpse
    d_skip = {}
    dcur = d_P[..., None]
    for i in range(n_d - 1, -1, -1):
      rec = tape[('attn_dcnn', i, tt)]
      w_conv = deconv_to_conv(w['attn_dcnn_w_%d' % i])
      k = 'attn_dcnn_%d_%d_' % (i, tt)
      dx, dW, db, dgm, dbt = conv_block_bwd(rec, dcur, w_conv, w[k + 'gamma'], w[k + 'beta'], 1, True,
                                            opt['attn_dcnn_pool'][i])
      _acc(grads, 'attn_dcnn_w_%d' % i, deconv_to_conv(dW))
      _acc(grads, 'attn_dcnn_b_%d' % i, db)
      grads[k + 'gamma'], grads[k + 'beta'] = dgm, dbt
      c1 = _np(rec['x']).shape[3]
      dcur = dx[..., :c1]
      if rec['skip'] is not None:
        d_skip[i] = dx[..., c1:]
    d_core = dcur  # gradient of h_acnn[-1] from the deconv head
    # score head: s = sigmoid([h, core] W + b)
    s = st['s_out'][:, 0]
    dpre = (d_s[:, tt] * s * (1.0 - s))[:, None]
    inp = np.concatenate([st['h'], st['h_core']], 1)
    _acc(grads, 'score_mlp_w_0', inp.T @ dpre)
    _acc(grads, 'score_mlp_b_0', dpre.sum(axis=0))
    d_inp = dpre @ w['score_mlp_w_0'].T
    d_h = d_inp[:, :Hd]
    d_core = d_core + d_inp[:, Hd:].reshape(d_core.shape)
    # attention CNN; skip list [None, h_acnn[4], ..., h_acnn[0], x_patch] (full_model.py:799-803)
    d_acnn = [None] * n_a
    d_xpatch = 0.0
    if add_skip:
      srcs = [None] + list(range(n_a - 2, -1, -1)) + ['x_patch']
      for i, g in d_skip.items():
        if srcs[i] == 'x_patch':
          d_xpatch = d_xpatch + g
        else:
          d_acnn[srcs[i]] = g
    dcur = d_core
    for i in range(n_a - 1, -1, -1):
      if d_acnn[i] is not None and i < n_a - 1:
        dcur = dcur + d_acnn[i]
      rec = tape[('attn_cnn', i, tt)]
      k = 'attn_cnn_%d_%d_' % (i, tt)
      dx, dW, db, dgm, dbt = conv_block_bwd(rec, dcur, w['attn_cnn_w_%d' % i], w[k + 'gamma'], w[k + 'beta'],
                                            opt['attn_cnn_pool'][i], True, 1)
      _acc(grads, 'attn_cnn_w_%d' % i, dW)
      _acc(grads, 'attn_cnn_b_%d' % i, db)
      grads[k + 'gamma'], grads[k + 'beta'] = dgm, dbt
      dcur = dx
    d_xpatch = d_xpatch + dcur
    # glimpse: x_patch = g_attn * Fy^T X Fx  (no gradient to X)
    d_fy_e, d_fx_e, dg_attn = extract_bwd(d_xpatch, st['acnn_inp'], f_y, f_x, g_attn, st['x_patch'])
    d_fy, d_fx = d_fy + d_fy_e, d_fx + d_fx_e
    # Gaussian filters -> box parameters
    d_box6 = np.zeros((B, 6), dtype)
    for axis, (filt, dfilt) in enumerate(((f_y, d_fy), (f_x, d_fx))):
      dc_, ds_, dv_ = filters_bwd(st['ctr'][:, axis], st['size'][:, axis], st['lg_var'][:, axis], filt, dfilt)
      d_box6[:, 0 + axis], d_box6[:, 2 + axis], d_box6[:, 4 + axis] = dc_, ds_, dv_
    if use_knob:
      # ctr = kb * ctr_gt + (1 - kb) * ctr_ctrl (same for size; lg_var is not mixed, full_model.py:760-776): the mixed
      # box passes (1 - kb) of its centre / size gradient on; the pre-mix filters of the attention box add theirs
      keep = 1.0 - st['kb']  # [B,1]
      d_box6[:, 0:2] *= keep
      d_box6[:, 2:4] *= keep
      for axis, (filt, dfilt) in enumerate(((f_y0, d_fy0), (f_x0, d_fx0))):
        dc_, ds_, dv_ = filters_bwd(st['ctr_ctrl'][:, axis], st['size_ctrl'][:, axis], st['lg_var'][:, axis], filt, dfilt)
        d_box6[:, 0 + axis] += dc_
        d_box6[:, 2 + axis] += ds_
        d_box6[:, 4 + axis] += dv_
      if coord_box_loss:
        # box loss = -(1/B) sum_b 1/cnt_b sum_m match_box[b,t,m] * f_iou_box(box_t, gt_m): weights of this step's row
        wgt = -match_box[:, tt, :] / (B * cnt_box[:, None])
        dctr, dsize = iou_box_coord_bwd(st['ctr_ctrl'], st['size_ctrl'], tl_gt, br_gt, wgt)
        d_box6[:, 0:2] += dctr
        d_box6[:, 2:4] += dsize
    # controller
    feat4 = st['feat']
    feat = feat4.reshape(B, -1, feat4.shape[3])
    d_feat, g_ctrl = controller_bwd(opt, w, feat, st['ctrl_out'], st['size_ctrl'], np.stack([g_attn, g_box, g_y], 1),
                                    d_box6, np.stack([dg_attn, dg_box, dg_y], 1), d_h)
    for k_, v in g_ctrl.items():
      _acc(grads, k_, v)
    # controller CNN, last layer first; the first layer needs no data gradient (image / stop-gradient canvas)
    dcur = d_feat.reshape(feat4.shape)
    for i in range(n_c - 1, -1, -1):
      rec = tape[('ctrl_cnn', i, tt)]
      k = 'ctrl_cnn_%d_%d_' % (i, tt)
      dx, dW, db, dgm, dbt = conv_block_bwd(rec, dcur, w['ctrl_cnn_w_%d' % i], w[k + 'gamma'], w[k + 'beta'],
                                            opt['ctrl_cnn_pool'][i], True, 1, want_dx=(i > 0))
      _acc(grads, 'ctrl_cnn_w_%d' % i, dW)
      _acc(grads, 'ctrl_cnn_b_%d' % i, db)
      grads[k + 'gamma'], grads[k + 'beta'] = dgm, dbt
      dcur = dx
  return grads, out

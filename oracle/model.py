"""CPU oracle of the recurrent-attention model graph — TEST INFRASTRUCTURE ONLY.

A structure-faithful PyTorch-CPU fp32 restatement of the reference model graph (full_model.py / box_model.py /
modellib.py / nnlib.py): same T-step loop, same per-channel ``bmm`` pair in ``extract_patch``, same T-way
pairwise-IoU loop, TF 'SAME' padding, TF ``conv2d_transpose`` cropping, per (layer, timestep) batch-norm copies,
sequential Hungarian (oracle/hungarian_ref.c).

PARITY STATUS: TensorFlow 0.12 + Python 2.7 cannot run here and the reference ships no golden outputs for the graph,
but its .py files parse as Python 3, so they are EXECUTED AS THEY ARE over a numpy stand-in for the TensorFlow-0.12 ops
they use (tests/golden/tf012_shim) and this module is held to the results (tests/golden/make_*_golden.py ->
tests/test_modellib_golden.py, test_nnlib_golden.py, test_full_model_golden.py):
  * full_model_forward == the reference's own full_model.get_model(opt) to 1e-7 in float64 - three architectures,
    training mode, with and without scheduled sampling (the graph's random draws replayed), incl. use_iou_box;
  * the layer functions == the reference's nn.cnn / nn.dcnn / nn.mlp / nn.lstm / batch_norm (train and eval);
  * the math library == the reference's modellib.py function by function.
What remains on trust is TensorFlow's own kernel semantics (SAME padding, conv2d_transpose cropping, softmax ...): the
shim states each in one line of numpy / torch.  box_model_forward is pinned the same way (the reference's own
box_model.get_model, training mode, with and without use_iou_box), and so is fg_model_forward (the reference's
fg_model.get_model for the three shipped FCN architectures at reduced width; fg_model.py imports `image_ops_old`,
which the reference does not ship - the shim aliases it to the reference's image_ops.py, the only repair needed).

Layouts follow the reference: images/features NHWC, mask stacks [B,T,H,W].
Every function cites the reference lines it restates (paths relative to /root/reference).
Eval mode (``phase_train=False``) and training mode with explicitly supplied random draws are covered.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import hungarian as _hung

BN_EPS = 1e-3  # nnlib.py:119

# Optional recording of forward intermediates for oracle/backward_manual.py (the executable spec of the backward
# assembly): when TAPE is a dict, conv layers record their input / skip / pre-BN output / batch statistics under
# (scope, layer, copy) and full_model_forward records one dict per decode step under ('step', tt).
TAPE = None


# ----------------------------------------------------------------------------- nnlib.py
def conv2d_same(x, w, b):
  """nnlib.py:6-12 + :229.  x [B,H,W,Cin], w [3,3,Cin,Cout] (TF HWIO), stride 1, SAME."""
  kh, kw = w.shape[0], w.shape[1]
  assert kh == 3 and kw == 3
  y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), bias=None, stride=1, padding=1)
  return y.permute(0, 2, 3, 1) + b


def max_pool_same(x, ratio):
  """nnlib.py:15-25.  SAME pooling with k = s = ratio; even sizes only."""
  assert x.shape[1] % ratio == 0 and x.shape[2] % ratio == 0
  return F.max_pool2d(x.permute(0, 3, 1, 2), ratio, ratio).permute(0, 2, 3, 1)


def batch_norm_eval(x, p):
  """nnlib.py:65-128 with phase_train=False: the EMA shadows are used (:113-119).
  tf.nn.batch_normalization: inv = rsqrt(var+eps)*gamma; x*inv + (beta - mean*inv)."""
  inv = torch.rsqrt(p['ema_var'] + BN_EPS) * p['gamma']
  return x * inv + (p['beta'] - p['ema_mean'] * inv)


def random_transformation(x, padding, offset, vflip=False, hflip=False, transpose=False, y=None, d=None, c=None):
  """image_ops.random_transformation (image_ops.py:9-113) with phase_train=True and the random draws supplied
  (offset = the tf.random_uniform([2], maxval=2*padding) draw; flips / transpose = the thresholded uniform draws)."""
  def crop_nhwc(t):
    tp = F.pad(t, (0, 0, padding, padding, padding, padding))
    return tp[:, offset[0]:offset[0] + t.shape[1], offset[1]:offset[1] + t.shape[2], :]

  res = {}
  xr = crop_nhwc(x)
  yr = None
  if y is not None:
    yp = F.pad(y, (padding, padding, padding, padding))
    yr = yp[:, :, offset[0]:offset[0] + y.shape[2], offset[1]:offset[1] + y.shape[3]]
  if d is None:
    if vflip:
      xr = torch.flip(xr, [1])
      yr = torch.flip(yr, [2]) if yr is not None else None
    if hflip:
      xr = torch.flip(xr, [2])
      yr = torch.flip(yr, [3]) if yr is not None else None
    if transpose:
      xr = xr.permute(0, 2, 1, 3)
      yr = yr.permute(0, 1, 3, 2) if yr is not None else None
  res['x'] = xr.contiguous()
  if yr is not None:
    res['y'] = yr.contiguous()
  if d is not None:
    res['d'] = crop_nhwc(d).contiguous()
  if c is not None:
    res['c'] = crop_nhwc(c).contiguous()
  return res


def batch_norm_train(x, p, decay=0.9):
  """nnlib.py:65-128 with phase_train=True: batch moments over (B,H,W) (tf.nn.moments, biased variance), the EMA
  shadows move by (1-decay) toward them (tf.train.ExponentialMovingAverage, decay = 1 - 0.1, :101-108), and the
  batch statistics normalise.  Returns (normed, batch_mean, batch_var, new_ema_mean, new_ema_var)."""
  mean = x.mean(dim=(0, 1, 2))
  var = ((x - mean)**2).mean(dim=(0, 1, 2))
  inv = torch.rsqrt(var + BN_EPS) * p['gamma']
  normed = x * inv + (p['beta'] - mean * inv)
  new_mean = p['ema_mean'] - (1.0 - decay) * (p['ema_mean'] - mean)
  new_var = p['ema_var'] - (1.0 - decay) * (p['ema_var'] - var)
  return normed, mean, var, new_mean, new_var


def conv2d_transpose_same(x, w, b, stride):
  """nnlib.py:372-376.  w is [3,3,Cout,Cin] (TF layout for conv2d_transpose), output
  spatial = in*stride, padding SAME.  For stride 2 this is the full transposed conv
  cropped at the END (forward SAME conv pads 0 before / 1 after); for stride 1 a
  symmetric crop of 1."""
  xn = x.permute(0, 3, 1, 2)
  wt = w.permute(3, 2, 0, 1)  # [Cin, Cout, kh, kw], no spatial flip
  if stride == 1:
    y = F.conv_transpose2d(xn, wt, stride=1, padding=1)
  else:
    assert stride == 2
    y = F.conv_transpose2d(xn, wt, stride=2, padding=0)
    y = y[:, :, :2 * x.shape[1], :2 * x.shape[2]]
  return y.permute(0, 2, 3, 1) + b


def _batch_norm(y, weights, scope, layer, copy, ema_out):
  """Eval mode (ema_out is None): EMA statistics.  Training mode: batch statistics; the moved EMA shadows are
  recorded in ema_out under the weight-dict keys (nnlib.py:101-127)."""
  p = _bn(weights, scope, layer, copy)
  if ema_out is None:
    return batch_norm_eval(y, p)
  normed, b_mean, b_var, new_mean, new_var = batch_norm_train(y, p)
  if TAPE is not None:
    TAPE.setdefault((scope, layer, copy), {}).update(raw=y.detach(), mean=b_mean.detach(), var=b_var.detach())
  k = '{}_{}_{}_'.format(scope, layer, copy)
  ema_out[k + 'ema_mean'] = new_mean
  ema_out[k + 'ema_var'] = new_var
  return normed


def run_cnn(x, weights, scope, nlayers, pool, copy, ema_out=None):
  """nnlib.py:214-255.  conv+b -> BN(copy) -> relu -> maxpool; returns every layer."""
  h = []
  for ii in range(nlayers):
    inp = x if ii == 0 else h[-1]
    if TAPE is not None:
      TAPE.setdefault((scope, ii, copy), {}).update(x=inp.detach(), skip=None)
    y = conv2d_same(inp, weights['{}_w_{}'.format(scope, ii)], weights['{}_b_{}'.format(scope, ii)])
    y = _batch_norm(y, weights, scope, ii, copy, ema_out)
    y = torch.relu(y)
    if pool[ii] > 1:
      y = max_pool_same(y, pool[ii])
    h.append(y)
  return h


def run_dcnn(x, weights, scope, nlayers, unpool, copy, skip=None, ema_out=None):
  """nnlib.py:339-402.  concat(prev, skip) -> conv2d_transpose+b -> BN(copy) -> relu
  (relu also on the last layer, full_model.py:487)."""
  h = []
  for ii in range(nlayers):
    inp = x if ii == 0 else h[-1]
    if TAPE is not None:
      has = skip is not None and skip[ii] is not None
      TAPE.setdefault((scope, ii, copy), {}).update(x=inp.detach(), skip=skip[ii].detach() if has else None)
    if skip is not None and skip[ii] is not None:
      inp = torch.cat([inp, skip[ii]], 3)
    y = conv2d_transpose_same(inp, weights['{}_w_{}'.format(scope, ii)], weights['{}_b_{}'.format(scope, ii)],
                              unpool[ii])
    y = _batch_norm(y, weights, scope, ii, copy, ema_out)
    y = torch.relu(y)
    h.append(y)
  return h


def _bn(weights, scope, layer, copy):
  k = '{}_{}_{}_'.format(scope, layer, copy)
  return {n: weights[k + n] for n in ('beta', 'gamma', 'ema_mean', 'ema_var')}


def lstm_step(inp, state, weights, hid, scope='ctrl_lstm'):
  """nnlib.py:637-649 (unroll); state = concat(c, h)."""
  c, h = state[:, :hid], state[:, hid:]
  w = lambda n: weights['{}_{}'.format(scope, n)]
  g_i = torch.sigmoid(inp @ w('w_xi') + h @ w('w_hi') + w('b_i'))
  g_f = torch.sigmoid(inp @ w('w_xf') + h @ w('w_hf') + w('b_f'))
  g_o = torch.sigmoid(inp @ w('w_xo') + h @ w('w_ho') + w('b_o'))
  u = torch.tanh(inp @ w('w_xu') + h @ w('w_hu') + w('b_u'))
  c = g_f * c + g_i * u
  h = g_o * torch.tanh(c)
  return torch.cat([c, h], 1)


def run_mlp(x, weights, scope, acts):
  """nnlib.py:476-493."""
  h = []
  for ii, act in enumerate(acts):
    inp = x if ii == 0 else h[-1]
    y = inp @ weights['{}_w_{}'.format(scope, ii)] + weights['{}_b_{}'.format(scope, ii)]
    if act == 'relu':
      y = torch.relu(y)
    elif act == 'softmax':
      y = torch.softmax(y, dim=1)
    elif act == 'sigmoid':
      y = torch.sigmoid(y)
    h.append(y)
  return h


# -------------------------------------------------------------------------- modellib.py
def f_inter(a, b):
  """modellib.py:104-107."""
  return (a * b).sum(dim=(-2, -1))


def f_union(a, b, eps=1e-5):
  """modellib.py:110-114: eps is added PER PIXEL inside the sum."""
  return (a + b - (a * b) + eps).sum(dim=(-2, -1))


def f_iou_pairwise(a, b):
  """modellib.py:138-153: T-way loop, a [B,N,H,W], b [B,M,H,W] -> [B,N,M]."""
  out = []
  bb = b.unsqueeze(1)  # [B,1,M,H,W]
  for ii in range(a.shape[1]):
    aa = a[:, ii:ii + 1].unsqueeze(2)  # [B,1,1,H,W]
    out.append(f_inter(aa, bb) / f_union(aa, bb))
  return torch.cat(out, 1)


def f_dice_pairwise(a, b):
  """modellib.py:81-97."""
  out = []
  bb = b.unsqueeze(1)
  card_b = (bb + 1e-5).sum(dim=(3, 4))
  for ii in range(a.shape[1]):
    aa = a[:, ii:ii + 1].unsqueeze(2)
    out.append(2 * f_inter(aa, bb) / ((aa + 1e-5).sum(dim=(3, 4)) + card_b))
  return torch.cat(out, 1)


def f_iou_box(tl_a, br_a, tl_b, br_b):
  """modellib.py:206-238 (coordinate IoU, no eps: 0/0 -> NaN for degenerate boxes)."""
  y1a, x1a, y2a, x2a = tl_a[..., 0], tl_a[..., 1], br_a[..., 0], br_a[..., 1]
  y1b, x1b, y2b, x2b = tl_b[..., 0], tl_b[..., 1], br_b[..., 0], br_b[..., 1]
  x1, y1 = torch.maximum(x1a, x1b), torch.maximum(y1a, y1b)
  x2, y2 = torch.minimum(x2a, x2b), torch.minimum(y2a, y2b)
  flag = (x1 < x2).float() * (y1 < y2).float()
  inter = flag * (x2 - x1) * (y2 - y1)
  union = (x2a - x1a) * (y2a - y1a) + (x2b - x1b) * (y2b - y1b) - inter
  return inter / union


def f_greedy_match(score, matched):
  """modellib.py:366-379: one-hot of the row max, ties share 1/k."""
  score = score * (1.0 - matched)
  mx = score.max(dim=1, keepdim=True)[0]
  m = (score == mx).float()
  return m / m.sum(dim=1, keepdim=True)


def tf_round(x):
  """tf.round in TF 0.12 is floor(x + 0.5)."""
  return torch.floor(x + 0.5)


def segm_match_weights(iou, s_gt):
  """modellib.py:395-406: the fp32 matrix handed to the Hungarian op."""
  mask_x = s_gt.unsqueeze(1)
  mask_y = s_gt.unsqueeze(2)
  iou_mask = iou * mask_x * mask_y
  precision = torch.tensor(1e6, dtype=torch.float32)
  iou_mask = tf_round(iou_mask * precision) / precision
  return iou_mask + torch.tensor(1e-5, dtype=torch.float32)


def f_segm_match(iou, s_gt):
  """modellib.py:382-415."""
  w = segm_match_weights(iou, s_gt)
  m = torch.from_numpy(_hung.hungarian(w.detach().numpy())[0])  # ops.NoGradient("Hungarian"), modellib.py:11
  return m * s_gt.unsqueeze(1) * s_gt.unsqueeze(2)


def f_cum_min(s):
  """modellib.py:39-52."""
  return torch.cummin(s, dim=1)[0]


def f_cum_max(s):
  """modellib.py:55-68 (cumulative max from the END)."""
  return torch.flip(torch.cummax(torch.flip(s, [1]), dim=1)[0], [1])


def f_conf_loss(s_out, match):
  """modellib.py:316-339 with use_cum_min=True, and :430-437."""
  eps = 1e-5
  B, T = s_out.shape
  match_sum = match.sum(dim=2)
  s_min, s_max = f_cum_min(s_out), f_cum_max(s_out)
  bce = -match_sum * torch.log(s_min + eps) - (1 - match_sum) * torch.log(1 - s_max + eps)
  return bce.sum() / B / T


def f_coverage_weight(y_gt):
  """modellib.py:278-289."""
  s = y_gt.sum(dim=(2, 3))
  ss = s.sum(dim=1, keepdim=True) + (s == 0).float()
  return s / ss


def f_weighted_coverage(iou, y_gt):
  """modellib.py:292-302."""
  cov = iou.max(dim=1)[0]
  return (cov * f_coverage_weight(y_gt)).sum() / y_gt.shape[0]


def f_unweighted_coverage(iou, count):
  """modellib.py:305-313."""
  cov = iou.max(dim=1)[0]
  return (cov.sum(dim=1) / count).sum() / iou.shape[0]


def f_count_acc(s_out, s_gt):
  """modellib.py:482-494."""
  return ((s_out > 0.5).float().sum(1) == s_gt.sum(1)).float().sum() / s_out.shape[0]


def f_dic(s_out, s_gt, use_abs=False):
  """modellib.py:497-511."""
  d = (s_out > 0.5).float().sum(1) - s_gt.sum(1)
  if use_abs:
    d = d.abs()
  return d.sum() / s_out.shape[0]


def get_gaussian_filter(center, size, lg_var, image_size, filter_size):
  """modellib.py:581-612 -> [B, L, F]."""
  span_filter = torch.arange(filter_size, dtype=torch.float32).view(1, 1, -1)
  center = center.reshape(-1, 1, 1)
  size = size.reshape(-1, 1, 1)
  mu = center + (size + 1) / filter_size * (span_filter - (filter_size - 1) / 2.0)
  lg_var = lg_var.reshape(-1, 1, 1)
  span = torch.arange(image_size, dtype=torch.float32).view(1, image_size, 1)
  two_pi_sqrt = torch.sqrt(torch.tensor(2 * np.pi, dtype=torch.float32))
  filt = (1 / torch.sqrt(torch.exp(lg_var)) / two_pi_sqrt) * torch.exp(-0.5 * (span - mu) * (span - mu) /
                                                                       torch.exp(lg_var))
  return filt


def extract_patch(x, f_y, f_x, nchannels):
  """modellib.py:615-641: per channel slice -> batch_matmul(f_y^T, x_ch) -> batch_matmul(., f_x)."""
  patch = []
  for dd in range(nchannels):
    x_ch = x[:, :, :, dd]
    patch.append(torch.bmm(torch.bmm(f_y.transpose(1, 2), x_ch), f_x).unsqueeze(3))
  return torch.cat(patch, 3)


def get_gt_box(y_gt, padding_ratio=0.0, center_shift_ratio=0.0, min_padding=10.0):
  """modellib.py:663-701 with get_idx_map (:704-729) and get_filled_box_idx (:732-749).
  y_gt [B,T,H,W] -> top_left [B,T,2], bot_right [B,T,2], box [B,T,H,W]; index 0 = y."""
  B, T, H, W = y_gt.shape
  idx_y = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1, 1).expand(B, T, H, W, 1)
  idx_x = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W, 1).expand(B, T, H, W, 1)
  idx = torch.cat([idx_y, idx_x], 4)
  not_zero = (y_gt.sum(dim=(2, 3)) > 0).float().unsqueeze(2)
  idx_min = idx + ((1.0 - y_gt) * float(H * W)).unsqueeze(4)
  idx_max = idx * y_gt.unsqueeze(4)
  top_left = idx_min.amin(dim=(2, 3))
  bot_right = idx_max.amax(dim=(2, 3))
  size = bot_right - top_left
  top_left = top_left + center_shift_ratio * size
  top_left = top_left - torch.clamp(padding_ratio * size, min=min_padding)
  bot_right = bot_right + center_shift_ratio * size
  bot_right = bot_right + torch.clamp(padding_ratio * size, min=min_padding)
  tl = top_left.view(B, T, 1, 1, 2)
  br = bot_right.view(B, T, 1, 1, 2)
  box = (idx >= tl).float().prod(4) * (idx <= br).float().prod(4)
  top_left = top_left * not_zero
  bot_right = not_zero * bot_right + (1 - not_zero) * (2 * min_padding)
  return top_left, bot_right, box


# ------------------------------------------------------------- full_model.py / box_model.py
def _opt(opt, key, default):
  return opt[key] if key in opt else default


def _t(a):
  if isinstance(a, torch.Tensor):  # tensors pass through (oracle/grads.py hands in leaves that require grad)
    return a if a.dtype == torch.float32 else a.to(torch.float32)
  return torch.as_tensor(np.asarray(a), dtype=torch.float32)


def controller_step(opt, weights, ccnn_inp, tt, ema_out=None):
  """full_model.py:663-725 == box_model.py:413-470: ctrl CNN, 5 glimpse iterations of the
  LSTM, controller head and the box-parameter maths.  Returns a dict of per-step tensors."""
  H, W = opt['inp_height'], opt['inp_width']
  Fh, Fw = opt['filter_height'], opt['filter_width']
  hid = opt['ctrl_rnn_hid_dim']
  n_iter = opt['num_ctrl_rnn_iter']
  B = ccnn_inp.shape[0]
  nl = len(opt['ctrl_cnn_filter_size'])
  h_ccnn = run_cnn(ccnn_inp, weights, 'ctrl_cnn', nl, opt['ctrl_cnn_pool'], tt, ema_out=ema_out)
  feat = h_ccnn[-1]
  gdim = feat.shape[1] * feat.shape[2]
  crnn_inp = feat.reshape(B, gdim, feat.shape[3])  # p = y*w' + x
  state = torch.zeros(B, 2 * hid)  # reset every outer step (full_model.py:674)
  gmap = torch.ones(B, gdim, 1) / gdim
  gmaps = []
  n_g = opt['num_glimpse_mlp_layers']
  for tt2 in range(n_iter):
    gmaps.append(gmap[:, :, 0])
    glimpse = (crnn_inp * gmap).sum(dim=1)
    state = lstm_step(glimpse, state, weights, hid)
    h = state[:, hid:]
    h_gmlp = run_mlp(h, weights, 'glimpse_mlp', ['relu'] * (n_g - 1) + ['softmax'])
    if tt2 < n_iter - 1:
      gmap = h_gmlp[-1].unsqueeze(2)
  n_c = opt['num_ctrl_mlp_layers']
  ctrl_out = run_mlp(h, weights, 'ctrl_mlp', ['relu'] * (n_c - 1) + [None])[-1]

  res = {'h_ccnn': h_ccnn, 'h': h, 'ctrl_out': ctrl_out, 'glimpse_map': torch.stack(gmaps, 1)}
  res.update(box_params(opt, ctrl_out))
  return res


def box_params(opt, ctrl_out):
  """full_model.py:691-725 with modellib.py:752-856: the nine controller outputs -> box centre / size / log-variance
  and the three log-gains."""
  H, W = opt['inp_height'], opt['inp_width']
  Fh, Fw = opt['filter_height'], opt['filter_width']
  B = ctrl_out.shape[0]
  ctr_norm = ctrl_out[:, 0:2]
  lg_size = ctrl_out[:, 2:4]
  if opt['squash_ctrl_params']:
    ctr_norm = torch.tanh(ctr_norm)
    lg_size = -F.softplus(lg_size)
  img_size = torch.tensor([H, W], dtype=torch.float32)
  ctr = (ctr_norm + 1.0) * (img_size / 2.0)  # modellib.get_unnormalized_center, modellib.py:752-764
  size = torch.exp(lg_size) * img_size  # modellib.get_unnormalized_size, modellib.py:812-825
  if opt['fixed_var']:
    lg_var = torch.zeros(B, 2)
  else:
    lg_var = torch.log(size) - torch.log(torch.tensor([Fh, Fw], dtype=torch.float32))  # modellib.py:782-793
  if opt['dynamic_var']:
    lg_var = ctrl_out[:, 4:6]
  if opt.get('fixed_gamma', False):
    lg_gamma = torch.zeros(B, 1)
    y_lg_gamma = torch.full((B, 1), 2.0)
  else:
    lg_gamma = ctrl_out[:, 6:7]
    y_lg_gamma = ctrl_out[:, 8:9]
  box_lg_gamma = ctrl_out[:, 7:8]
  return {'ctr_norm': ctr_norm, 'lg_size': lg_size, 'ctr': ctr, 'size': size, 'lg_var': lg_var, 'lg_gamma': lg_gamma,
          'box_lg_gamma': box_lg_gamma, 'y_lg_gamma': y_lg_gamma}


def top_left_pred(ctr, size):
  return ctr - size / 2.0


def bot_right_pred(ctr, size):
  return ctr + size / 2.0


def attn_box_from(opt, ctr, size, lg_var, box_lg_gamma):
  """full_model.py:728-741 / box_model.py:473-487."""
  H, W = opt['inp_height'], opt['inp_width']
  Fh, Fw = opt['filter_height'], opt['filter_width']
  B = ctr.shape[0]
  f_y = get_gaussian_filter(ctr[:, 0], size[:, 0], lg_var[:, 0], H, Fh)
  f_x = get_gaussian_filter(ctr[:, 1], size[:, 1], lg_var[:, 1], W, Fw)
  ones = torch.ones(B, Fh, Fw, 1) * torch.exp(box_lg_gamma).view(-1, 1, 1, 1)
  box = extract_patch(ones, f_y.transpose(1, 2), f_x.transpose(1, 2), 1)
  box = torch.sigmoid(box - 5.0)
  return box.reshape(B, 1, H, W), f_y, f_x


def knob_probability(opt, global_step, offset_key):
  """full_model.py:598-625: min(1, knob_base * knob_decay^(max(0, step - offset) / steps_per_knob_decay) * time scale),
  time scale = 1 + log(1 + 3t) when knob_use_timescale.  Returns [T] probabilities (fp32 like the graph)."""
  T = opt['timespan']
  step = max(0.0, float(global_step) - float(opt[offset_key]))
  p = np.float32(opt['knob_base']) * np.power(np.float32(opt['knob_decay']),
                                               np.float32(step / float(opt['steps_per_knob_decay'])))
  scale = (1.0 + np.log(1.0 + np.arange(T, dtype=np.float32) * 3.0)) if opt.get('knob_use_timescale', True) \
      else np.ones(T, np.float32)
  return np.minimum(np.float32(1.0), p * scale.astype(np.float32)).astype(np.float32)


def full_model_forward(opt, weights, batch, with_loss=True, phase_train=False, draws=None):
  """full_model.get_model in eval mode (phase_train=False, random_transformation is the
  identity, image_ops.py:70-112; the knob terms vanish because phase_train_f = 0), or - phase_train=True - the
  training-mode forward with the identity draw of the augmentation: every BN layer normalises with the batch
  statistics and moves its EMA shadows (nnlib.py:96-119); the result then carries 'ema_updates'.

  Scheduled sampling (opt['use_knob'] with phase_train, full_model.py:589-625,744-785,826-843): TensorFlow's random
  streams cannot be reproduced, so the draws are inputs - draws = {'gt_knob_box' [B,T] in {0,1},
  'gt_knob_segm' [B,T] in {0,1}, 'gt_box_pad' [B,T,1] (padding ratio of the noisy GT box), 'gt_box_ctr_shift' [B,T,2],
  'gt_segm_noise' [B,T,H,W] in [0, opt['gt_segm_noise'])}.

  batch: dict of numpy/torch fp32: x [B,H,W,3], y_gt [B,T,H,W], s_gt [B,T], optional
  d_in [B,H,W,8], y_in [B,H,W,C].  weights: dict keyed like full_model_read.py:32-70.
  Returns a dict with the reference's output keys (full_model.py:853-910, 941-1081)."""
  weights = {k: _t(v) for k, v in weights.items()}
  x = _t(batch['x'])
  B = x.shape[0]
  T, H, W = opt['timespan'], opt['inp_height'], opt['inp_width']
  Fh, Fw = opt['filter_height'], opt['filter_width']
  add_d_out = _opt(opt, 'add_d_out', False)
  attn_add_d = _opt(opt, 'attn_add_d_out', add_d_out)
  attn_add_y = _opt(opt, 'attn_add_y_out', _opt(opt, 'add_y_out', False))
  attn_add_inp = _opt(opt, 'attn_add_inp', True)
  attn_add_canvas = _opt(opt, 'attn_add_canvas', True)
  ctrl_add_d = _opt(opt, 'ctrl_add_d_out', add_d_out)
  ctrl_add_y = _opt(opt, 'ctrl_add_y_out', _opt(opt, 'add_y_out', False))
  ctrl_add_inp = _opt(opt, 'ctrl_add_inp', not ctrl_add_d)
  ctrl_add_canvas = _opt(opt, 'ctrl_add_canvas', not ctrl_add_d)
  add_skip = _opt(opt, 'add_skip_conn', True)
  disable_overwrite = _opt(opt, 'disable_overwrite', True)
  d_in = _t(batch['d_in']) if add_d_out else None
  y_in = _t(batch['y_in']) if add_d_out else None

  ema_out = {} if phase_train else None
  # The knob terms are part of the GRAPH whenever opt['use_knob'] is set (full_model.py:744-759): the per-step IoUs of
  # the decode loop then feed the box loss in evaluation too (:926-929; with use_iou_box they are the coordinate IoUs
  # of modellib.f_iou_box, not the pixel IoUs); only the MIXING is switched off by phase_train_f = 0 (:767-776).
  knob_graph = bool(_opt(opt, 'use_knob', False)) and ('y_gt' in batch)
  use_knob = bool(phase_train and knob_graph)
  if knob_graph:
    y_gt_k = _t(batch['y_gt'])
    minpad = opt['padding'] + 4
    # full_model.py:561-580: the clean GT boxes (for the greedy match) and the noisy ones (mixed into the controller)
    tl_gt_k, br_gt_k, box_gt_k = get_gt_box(y_gt_k, padding_ratio=opt['attn_box_padding_ratio'],
                                            center_shift_ratio=0.0, min_padding=minpad)
    iou_box_steps = []
  if use_knob:
    if draws is None:
      raise ValueError('use_knob needs the random draws as inputs')
    tl_n, br_n, _ = get_gt_box(y_gt_k, padding_ratio=_t(draws['gt_box_pad']),
                               center_shift_ratio=_t(draws['gt_box_ctr_shift']), min_padding=minpad)
    ctr_gt_noise, size_gt_noise = (tl_n + br_n) / 2.0, br_n - tl_n
    knob_box, knob_segm = _t(draws['gt_knob_box']), _t(draws['gt_knob_segm'])
    segm_noise = _t(draws['gt_segm_noise'])
  canvas = torch.zeros(B, H, W, 1)
  n_acnn = len(opt['attn_cnn_filter_size'])
  n_adcnn = len(opt['attn_dcnn_filter_size'])
  out = {k: [] for k in ('y_out', 's_out', 'attn_box', 'x_patch', 'y_out_patch', 'attn_ctr', 'attn_size',
                         'attn_top_left', 'attn_bot_right', 'glimpse_map', 'ctrl_out', 'attn_lg_var', 'h_ctrl')}

  for tt in range(T):
    ccnn_list, acnn_list = [], []  # full_model.py:640-661 (order: x, canvas, d_in, y_in)
    if ctrl_add_inp: ccnn_list.append(x)
    if attn_add_inp: acnn_list.append(x)
    if ctrl_add_canvas: ccnn_list.append(canvas)
    if attn_add_canvas: acnn_list.append(canvas)
    if ctrl_add_d: ccnn_list.append(d_in)
    if attn_add_d: acnn_list.append(d_in)
    if ctrl_add_y: ccnn_list.append(y_in)
    if attn_add_y: acnn_list.append(y_in)
    acnn_inp = torch.cat(acnn_list, 3)
    ccnn_inp = torch.cat(ccnn_list, 3)

    c = controller_step(opt, weights, ccnn_inp, tt, ema_out=ema_out)
    ctr, size = c['ctr'], c['size']
    box, f_y, f_x = attn_box_from(opt, ctr, size, c['lg_var'], c['box_lg_gamma'])
    if knob_graph:
      # full_model.py:744-785: greedy-match the predicted box to a GT box (grd_match_cum stays zero),
      # mix the matched noisy GT box into centre / size where the Bernoulli draw says so, rebuild the filters
      if _opt(opt, 'use_iou_box', False):
        iou_t = f_iou_box(top_left_pred(ctr, size).unsqueeze(1), bot_right_pred(ctr, size).unsqueeze(1), tl_gt_k, br_gt_k)
      else:
        iou_t = f_inter(box, box_gt_k) / f_union(box, box_gt_k)
      iou_box_steps.append(iou_t.unsqueeze(1))
    if use_knob:
      grd = f_greedy_match(iou_t, torch.zeros(B, T))
      kb = knob_box[:, tt:tt + 1]
      ctr = kb * (grd.unsqueeze(2) * ctr_gt_noise).sum(1) + (1.0 - kb) * ctr
      size = kb * (grd.unsqueeze(2) * size_gt_noise).sum(1) + (1.0 - kb) * size
      f_y = get_gaussian_filter(ctr[:, 0], size[:, 0], c['lg_var'][:, 0], H, Fh)
      f_x = get_gaussian_filter(ctr[:, 1], size[:, 1], c['lg_var'][:, 1], W, Fw)
    top_left, bot_right = ctr - size / 2.0, ctr + size / 2.0

    # full_model.py:788-807
    x_patch = torch.exp(c['lg_gamma']).view(-1, 1, 1, 1) * extract_patch(acnn_inp, f_y, f_x, acnn_inp.shape[3])
    h_acnn = run_cnn(x_patch, weights, 'attn_cnn', n_acnn, opt['attn_cnn_pool'], tt, ema_out=ema_out)
    h_core = h_acnn[-1].reshape(B, -1)  # (h, w, c) flatten order, full_model.py:794
    if add_skip:
      # full_model.py:497-499,799-803: every layer >= 1 gets a skip
      skip = [None] + (h_acnn[::-1][1:] + [x_patch])[:n_adcnn - 1]
    else:
      skip = None
    h_adcnn = run_dcnn(h_acnn[-1], weights, 'attn_dcnn', n_adcnn, opt['attn_dcnn_pool'], tt, skip=skip,
                       ema_out=ema_out)

    # full_model.py:810-822
    y = extract_patch(h_adcnn[-1], f_y.transpose(1, 2), f_x.transpose(1, 2), 1)
    y = torch.sigmoid(torch.exp(c['y_lg_gamma']).view(-1, 1, 1, 1) * y - 5.0)
    y = y.reshape(B, 1, H, W)
    if disable_overwrite:
      y = (1 - canvas).reshape(B, 1, H, W) * y
    s = run_mlp(torch.cat([c['h'], h_core], 1), weights, 'score_mlp', ['sigmoid'])[-1]
    y_canvas = y.reshape(B, H, W, 1)
    if use_knob:
      # full_model.py:826-843: the canvas may be written from the greedily matched GT mask (x (1 - noise)) instead
      y_gtm = (grd.view(B, T, 1, 1) * y_gt_k).sum(1)
      y_gtm = (y_gtm - y_gtm * segm_noise[:, tt]).reshape(B, H, W, 1)
      ks = knob_segm[:, tt].view(B, 1, 1, 1)
      y_canvas = ks * y_gtm + (1.0 - ks) * y_canvas
    if TAPE is not None:
      det = lambda v: v.detach() if isinstance(v, torch.Tensor) else v
      TAPE[('step', tt)] = {k: det(v) for k, v in dict(
          ccnn_inp=ccnn_inp, acnn_inp=acnn_inp, feat=c['h_ccnn'][-1], ctrl_out=c['ctrl_out'], h=c['h'],
          ctr_ctrl=c['ctr'], size_ctrl=c['size'], lg_var=c['lg_var'], lg_gamma=c['lg_gamma'],
          box_lg_gamma=c['box_lg_gamma'], y_lg_gamma=c['y_lg_gamma'], ctr=ctr, size=size, f_y=f_y, f_x=f_x,
          kb=(knob_box[:, tt:tt + 1] if use_knob else None), x_patch=x_patch, h_core=h_core,
          y_patch=h_adcnn[-1], y_out=y, attn_box=box, s_out=s).items()}
    canvas = torch.maximum(y_canvas, canvas)  # full_model.py:843-845
    if _opt(opt, 'stop_canvas_grad', True):
      canvas = canvas.detach()  # tf.stop_gradient, full_model.py:846-848

    out['y_out'].append(y)
    out['s_out'].append(s)
    out['attn_box'].append(box)
    out['x_patch'].append(x_patch.unsqueeze(1))
    out['y_out_patch'].append(h_adcnn[-1].unsqueeze(1))
    out['attn_ctr'].append(ctr.unsqueeze(1))
    out['attn_size'].append(size.unsqueeze(1))
    out['attn_top_left'].append(top_left.unsqueeze(1))
    out['attn_bot_right'].append(bot_right.unsqueeze(1))
    out['glimpse_map'].append(c['glimpse_map'].unsqueeze(1))
    out['ctrl_out'].append(c['ctrl_out'].unsqueeze(1))
    out['attn_lg_var'].append(c['lg_var'].unsqueeze(1))
    out['h_ctrl'].append(c['h'].unsqueeze(1))

  model = {k: torch.cat(v, 1) for k, v in out.items()}
  gm = model.pop('glimpse_map')  # full_model.py:897-907: [B, T, n_iter, h', w']
  sub = int(np.prod(opt['ctrl_cnn_pool']))
  model['ctrl_rnn_glimpse_map'] = gm.reshape(B, T, gm.shape[2], H // sub, W // sub)
  model['canvas'] = canvas
  if knob_graph:
    model['iou_soft_box_steps'] = torch.cat(iou_box_steps, 1)  # [B,T(step),T(gt)], full_model.py:926-929
  if ema_out is not None:
    model['ema_updates'] = ema_out
  if not with_loss:
    return model

  y_gt, s_gt = _t(batch['y_gt']), _t(batch['s_gt'])
  model.update(full_model_loss(opt, weights, model, y_gt, s_gt))
  return model


def weight_decay_loss(opt, weights):
  """nnlib.py:59-61: wd * l2_loss(w) on conv / mlp / lstm weight matrices only."""
  wd = opt['weight_decay']
  tot = torch.zeros(())
  for k, v in weights.items():
    is_w = ('_w_' in k) and not k.endswith(('_beta', '_gamma', '_ema_mean', '_ema_var'))
    if is_w:
      tot = tot + wd * (v * v).sum() / 2
  return tot


def full_model_loss(opt, weights, model, y_gt, s_gt):
  """full_model.py:916-1081 with fixed_order=False, box_loss_fn=segm_loss_fn='iou'."""
  T = opt['timespan']
  B = y_gt.shape[0]
  r = {}
  # GT attention boxes, full_model.py:561-567
  tl_gt, br_gt, box_gt = get_gt_box(
      y_gt, padding_ratio=opt['attn_box_padding_ratio'], center_shift_ratio=0.0, min_padding=opt['padding'] + 4)
  r['attn_top_left_gt'], r['attn_bot_right_gt'], r['attn_box_gt'] = tl_gt, br_gt, box_gt
  r['attn_ctr_gt'], r['attn_size_gt'] = (tl_gt + br_gt) / 2.0, br_gt - tl_gt

  if 'iou_soft_box_steps' in model:
    iou_box = model['iou_soft_box_steps']  # use_knob: the per-step IoUs of the decode loop, full_model.py:926-929
  else:
    iou_box = f_iou_pairwise(model['attn_box'], box_gt)  # :931
  match_box = f_segm_match(iou_box, s_gt)  # :939
  r['iou_soft_box_pairwise'] = iou_box
  r['match_box'] = match_box
  cnt_box = torch.clamp(match_box.sum(dim=(1, 2)), min=1.0)
  iou_soft_box = ((iou_box * match_box).sum(dim=(1, 2)) / cnt_box).sum() / B
  r['box_loss'] = -iou_soft_box
  if _opt(opt, 'box_loss_fn', 'iou') == 'wt_cov':  # full_model.py:967
    r['box_loss'] = -f_weighted_coverage(iou_box, box_gt)

  iou_soft_pw = f_iou_pairwise(model['y_out'], y_gt)  # :983
  match = f_segm_match(iou_soft_pw, s_gt)
  r['iou_soft_pairwise'] = iou_soft_pw
  r['match'] = match
  cnt = torch.clamp(match.sum(dim=(1, 2)), min=1.0)
  r['wt_cov_soft'] = f_weighted_coverage(iou_soft_pw, y_gt)
  r['unwt_cov_soft'] = f_unweighted_coverage(iou_soft_pw, cnt)
  r['iou_soft'] = ((iou_soft_pw * match).sum(dim=(1, 2)) / cnt).sum() / B
  r['segm_loss'] = -r['iou_soft']
  if _opt(opt, 'segm_loss_fn', 'iou') == 'wt_cov':  # full_model.py:1013-1014
    r['segm_loss'] = -r['wt_cov_soft']
  r['conf_loss'] = f_conf_loss(model['s_out'], match)
  r['loss'] = r['box_loss'] + r['segm_loss'] + opt['loss_mix_ratio'] * r['conf_loss'] + \
      weight_decay_loss(opt, weights)

  y_hard = (model['y_out'] > 0.5).float()  # :1063-1081
  iou_hard = f_iou_pairwise(y_hard, y_gt)
  r['iou_hard_pairwise'] = iou_hard
  r['wt_cov_hard'] = f_weighted_coverage(iou_hard, y_gt)
  r['unwt_cov_hard'] = f_unweighted_coverage(iou_hard, cnt)
  r['iou_hard'] = ((iou_hard * match).sum(dim=(1, 2)) / cnt).sum() / B
  dice = f_dice_pairwise(y_hard, y_gt)
  r['dice'] = ((dice * match).sum(dim=(1, 2)) / cnt).sum() / B
  r['count_acc'] = f_count_acc(model['s_out'], s_gt)
  r['dic'] = f_dic(model['s_out'], s_gt, False)
  r['dic_abs'] = f_dic(model['s_out'], s_gt, True)
  return r


def box_model_forward(opt, weights, batch, canvas_noise=None, phase_train=False):
  """box_model.get_model (box_model.py:403-629), eval mode or - phase_train=True - with batch-statistics BN.  The
  canvas is driven by the
  greedily matched GT masks in eval too (:484-504); the per-step U[0,0.3) noise (:501-502)
  is an explicit input ``canvas_noise`` [B,T,H,W] (zeros when None)."""
  weights = {k: _t(v) for k, v in weights.items()}
  x = _t(batch['x'])
  y_gt, s_gt = _t(batch['y_gt']), _t(batch['s_gt'])
  B = x.shape[0]
  T, H, W = opt['timespan'], opt['inp_height'], opt['inp_width']
  add_d_out = _opt(opt, 'add_d_out', False)
  nsc = _opt(opt, 'num_semantic_classes', 1)
  assert nsc == 1
  opt = dict(opt)
  opt.setdefault('fixed_var', True)  # box_model.py:58-61
  opt.setdefault('dynamic_var', False)
  opt['fixed_gamma'] = True  # box_model has no gamma outputs besides the box gamma
  d_in = _t(batch['d_in']) if add_d_out else None
  y_in = _t(batch['y_in']) if add_d_out else None
  noise = torch.zeros(B, T, H, W) if canvas_noise is None else _t(canvas_noise)

  tl_gt, br_gt, box_gt = get_gt_box(y_gt, padding_ratio=opt['attn_box_padding_ratio'], center_shift_ratio=0.0)
  canvas = torch.zeros(B, H, W, 1)
  ema_out = {} if phase_train else None  # training mode: batch-statistics BN (nnlib.py:96-119)
  grd_match_cum = torch.zeros(B, T)  # never updated (box_model.py:398,496)
  out = {k: [] for k in ('s_out', 'attn_box', 'attn_ctr', 'attn_size', 'attn_top_left', 'attn_bot_right',
                         'iou_soft_box', 'glimpse_map', 'ctrl_out')}
  for tt in range(T):
    lst = [x, canvas]
    if add_d_out:
      lst += [d_in, y_in]
    ccnn_inp = torch.cat(lst, 3)
    c = controller_step(opt, weights, ccnn_inp, tt, ema_out=ema_out)
    box, _, _ = attn_box_from(opt, c['ctr'], c['size'], c['lg_var'], c['box_lg_gamma'])
    if _opt(opt, 'use_iou_box', False):
      tl, br = c['ctr'] - c['size'] / 2.0, c['ctr'] + c['size'] / 2.0
      iou_t = f_iou_box(tl.unsqueeze(1), br.unsqueeze(1), tl_gt, br_gt)
    else:
      iou_t = f_inter(box, box_gt) / f_union(box, box_gt)  # [B,T]
    grd = f_greedy_match(iou_t, grd_match_cum)
    y_sel = (grd.view(B, T, 1, 1) * y_gt).sum(dim=1).unsqueeze(3)
    y_sel = y_sel - y_sel * noise[:, tt].unsqueeze(3)
    canvas = torch.maximum(y_sel, canvas)
    s = torch.sigmoid(run_mlp(c['h'], weights, 'score_mlp', [None])[-1])
    out['s_out'].append(s)
    out['attn_box'].append(box)
    out['attn_ctr'].append(c['ctr'].unsqueeze(1))
    out['attn_size'].append(c['size'].unsqueeze(1))
    out['attn_top_left'].append((c['ctr'] - c['size'] / 2.0).unsqueeze(1))
    out['attn_bot_right'].append((c['ctr'] + c['size'] / 2.0).unsqueeze(1))
    out['iou_soft_box'].append(iou_t.unsqueeze(1))
    out['glimpse_map'].append(c['glimpse_map'].unsqueeze(1))
    out['ctrl_out'].append(c['ctrl_out'].unsqueeze(1))
  model = {k: torch.cat(v, 1) for k, v in out.items()}
  gm = model.pop('glimpse_map')  # full_model.py:897-907: [B, T, n_iter, h', w']
  sub = int(np.prod(opt['ctrl_cnn_pool']))
  model['ctrl_rnn_glimpse_map'] = gm.reshape(B, T, gm.shape[2], H // sub, W // sub)
  model['canvas'] = canvas
  model['attn_top_left_gt'], model['attn_bot_right_gt'], model['attn_box_gt'] = tl_gt, br_gt, box_gt

  iou_box = model.pop('iou_soft_box')  # [B,T,T]
  model['iou_soft_box_pairwise'] = iou_box
  match_box = f_segm_match(iou_box, s_gt)
  model['match_box'] = match_box
  cnt = torch.clamp(match_box.sum(dim=(1, 2)), min=1.0)
  model['box_loss'] = -(((iou_box * match_box).sum(dim=(1, 2)) / cnt).sum() / B)
  model['conf_loss'] = f_conf_loss(model['s_out'], match_box)
  model['loss'] = model['box_loss'] + model['conf_loss'] + weight_decay_loss(opt, weights)
  if ema_out is not None:
    model['ema_updates'] = ema_out
  return model


# ----------------------------------------------------------------------------- fg_model.py
def f_iou_all(a, b):
  """modellib.py:171-181: one IoU over every element, eps on the union only."""
  inter = (a * b).sum()
  return inter / (a.sum() + b.sum() - inter + 1e-5)


def f_ce(y_out, y_gt):
  """modellib.py:418-421."""
  return -y_gt * torch.log(y_out + 1e-5)


def f_bce(y_out, y_gt):
  """modellib.py:424-427."""
  return -y_gt * torch.log(y_out + 1e-5) - (1 - y_gt) * torch.log(1 - y_out + 1e-5)


def fg_skip_lists(opt, x, h_cnn):
  """fg_model.py:122-146: the CNN-side mask picks activations from [x] + h_cnn[:-1] in order, the DCNN-side mask
  hands them out from the last one backwards; DCNN layer 0 has no skip."""
  if not _opt(opt, 'add_skip_conn', False):
    return None
  n = len(opt['cnn_depth'])
  cnn_mask = opt['cnn_skip_mask'] if 'cnn_skip_mask' in opt else opt.get('cnn_skip', [True] * n)
  dcnn_mask = opt['dcnn_skip_mask'] if 'dcnn_skip_mask' in opt else list(cnn_mask)[::-1]
  layers = [h for sk, h in zip(cnn_mask, [x] + h_cnn[:-1]) if sk]
  skip = [None]
  counter = len(layers) - 1
  for sk in dcnn_mask:
    if sk:
      skip.append(layers[counter])
      counter -= 1
    else:
      skip.append(None)
  return skip


def fg_model_forward(opt, weights, batch, phase_train=False):
  """fg_model.get_model (fg_model.py:11-245) in eval mode (phase_train=False: random_transformation is the identity
  crop, batch norm uses the EMA shadows) or, with phase_train=True, with batch-statistics BN and the identity crop: CNN -> DCNN with skip concatenation (last layer: no BN, no activation,
  :121,148) -> sigmoid / softmax foreground head and softmax orientation head (:174-192) -> IoU / BCE / CE losses
  (:194-236).  batch: x [B,H,W,3], y_gt [B,H,W,nsc][, d_gt [B,H,W,8]]."""
  weights = {k: _t(v) for k, v in weights.items()}
  x, y_gt = _t(batch['x']), _t(batch['y_gt'])
  nsc = _opt(opt, 'num_semantic_classes', 1)
  ori = bool(_opt(opt, 'add_orientation', False))
  nori = opt['num_orientation_classes'] if ori else 0
  n_c, n_d = len(opt['cnn_depth']), len(opt['dcnn_depth'])
  if opt['dcnn_depth'][-1] != nsc + nori:
    raise ValueError('Expecting last channel to be {}'.format(nsc + nori))  # fg_model.py:162-172
  ema_out = {} if phase_train else None  # training mode: batch-statistics BN
  h_cnn = run_cnn(x, weights, 'cnn', n_c, opt['cnn_pool'], 0, ema_out=ema_out)
  skip = fg_skip_lists(opt, x, h_cnn)
  h = h_cnn[-1]
  for ii in range(n_d):  # nnlib.py:339-402 with act = [relu]*(n-1) + [None], use_bn = [True]*(n-1) + [False]
    inp = h
    if skip is not None and skip[ii] is not None:
      inp = torch.cat([inp, skip[ii]], 3)
    h = conv2d_transpose_same(inp, weights['dcnn_w_%d' % ii], weights['dcnn_b_%d' % ii], opt['dcnn_pool'][ii])
    if ii < n_d - 1:
      h = torch.relu(_batch_norm(h, weights, 'dcnn', ii, 0, ema_out))
  model = {'logits': h}
  y_out = h[..., :nsc]
  if ori:
    d_out = torch.softmax(h[..., nsc:], dim=3)
    model['d_out'] = d_out
  y_out = torch.sigmoid(y_out) if nsc == 1 else torch.softmax(y_out, dim=3)
  model['y_out'] = y_out
  B, H, W = x.shape[0], x.shape[1], x.shape[2]
  num_pixel = float(B * H * W)
  y_gt_mask = y_gt[..., 1:nsc].max(dim=3, keepdim=True)[0] if nsc > 1 else y_gt
  if nsc == 1:
    y_hard = (y_out > 0.5).float()
    model['iou_soft'] = f_iou_all(y_out, y_gt)
    model['iou_hard'] = f_iou_all(y_hard, y_gt)
    segloss = f_bce(y_out, y_gt).sum() / num_pixel
  else:
    y_hard = (y_out == y_out.max(dim=3, keepdim=True)[0]).float()
    model['iou_soft'] = f_iou_all(y_out[..., 1:], y_gt[..., 1:])
    model['iou_hard'] = f_iou_all(y_hard[..., 1:], y_gt[..., 1:])
    segloss = f_ce(y_out, y_gt).sum() / num_pixel
  model['y_out_hard'] = y_hard
  fn = _opt(opt, 'segm_loss_fn', 'iou')
  loss = -model['iou_soft'] if fn == 'iou' else segloss
  model['foreground_loss'] = loss
  if ori:
    d_gt = _t(batch['d_gt'])
    num_pixel_ori = y_gt_mask.sum()
    model['orientation_ce'] = (f_ce(d_out, d_gt) * y_gt_mask).sum() / num_pixel_ori
    loss = loss + model['orientation_ce']
    correct = (d_out.argmax(dim=3) == d_gt.argmax(dim=3)).float()
    model['orientation_acc'] = (correct * y_gt_mask[..., 0]).sum() / y_gt_mask.sum()
  model['loss'] = loss
  if ema_out is not None:
    model['ema_updates'] = ema_out  # the moved EMA shadows of a training-mode forward (nnlib.py:104-108)
  return model

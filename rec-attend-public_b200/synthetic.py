"""Deterministic synthetic inputs and random-init weights for the hot path.

There are no datasets or checkpoints in this environment, so benches and parity tests use
seeded synthetic batches shaped like the reference's batch contract
(data_api/ins_seg_dataset.py:169-172,267-271: GT masks sorted by area descending,
``s_gt`` = first min(num_obj, T) ones) and weights keyed like the reference's flat
``weights.h5`` (full_model_read.py:32-70, box_model_read.py:31-52), in TF layouts.
numpy only: the same bits feed the CUDA path and the CPU oracle.
"""
import numpy as np

from .config import input_depths


def make_batch(opt, batch_size, seed=1234):
  """x [B,H,W,3] U[0,1); y_gt [B,T,H,W] disjoint filled ellipses sorted by area; s_gt [B,T];
  d_in [B,H,W,8], y_in [B,H,W,C] when the architecture takes them."""
  rng = np.random.default_rng(seed)
  H, W, T = opt['inp_height'], opt['inp_width'], opt['timespan']
  B = batch_size
  x = rng.random((B, H, W, 3), dtype=np.float32)
  y_gt = np.zeros((B, T, H, W), np.float32)
  s_gt = np.zeros((B, T), np.float32)
  yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
  for b in range(B):
    k = int(rng.integers(max(2, T // 4), T + 1))
    if b == 0:
      k = T
    if b == 1:
      k = 2
    occupied = np.zeros((H, W), bool)
    masks = []
    for _ in range(k):
      cy, cx = rng.uniform(0, H), rng.uniform(0, W)
      ay, ax = rng.uniform(H / 16.0, H / 4.0), rng.uniform(W / 16.0, W / 4.0)
      m = (((yy - cy) / ay)**2 + ((xx - cx) / ax)**2 <= 1.0) & ~occupied
      if m.sum() == 0:
        continue
      occupied |= m
      masks.append(m)
    masks.sort(key=lambda m: -int(m.sum()))
    for t, m in enumerate(masks):
      y_gt[b, t] = m
      s_gt[b, t] = 1.0
  batch = {'x': x, 'y_gt': y_gt, 's_gt': s_gt}
  if opt.get('add_d_out', False):
    nsc = opt.get('num_semantic_classes', 1)
    fg = y_gt.max(axis=1)[..., None]  # [B,H,W,1]
    z = rng.standard_normal((B, H, W, 8)).astype(np.float32)
    e = np.exp(z - z.max(axis=3, keepdims=True))
    batch['d_in'] = (e / e.sum(axis=3, keepdims=True) * fg).astype(np.float32)
    y_in = np.clip(fg * rng.uniform(0.5, 1.0, (B, H, W, nsc)).astype(np.float32), 0, 1)
    if nsc > 1:
      y_in[..., 0:1] = 1.0 - fg
    batch['y_in'] = y_in.astype(np.float32)
  return batch


def _normal(rng, shape, fan_in):
  return (rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)


def _bn(rng, w, scope, layer, t, ch):
  k = '{}_{}_{}_'.format(scope, layer, t)
  w[k + 'gamma'] = rng.uniform(0.5, 1.5, ch).astype(np.float32)
  w[k + 'beta'] = (rng.standard_normal(ch) * 0.5).astype(np.float32)
  w[k + 'ema_mean'] = (rng.standard_normal(ch) * 0.1).astype(np.float32)
  w[k + 'ema_var'] = rng.uniform(0.5, 1.5, ch).astype(np.float32)


def make_weights(opt, seed=4321, model='full'):
  """Random-init weights with the reference's key schema.  w ~ N(0, 1/fan_in) (the
  reference's own sigma=0.01 init, nnlib.py:54, gives T identical boxes — degenerate ties);
  LSTM b_f = 1 (nnlib.py:564-569); per-(layer,timestep) BN parameters and EMA shadows are
  randomised so that decode steps differ.  The controller-head bias is set so that boxes
  cover a fraction of the image and the sigmoid(gamma*v - 5) outputs are not flat."""
  rng = np.random.default_rng(seed)
  T = opt['timespan']
  H, W = opt['inp_height'], opt['inp_width']
  hid = opt['ctrl_rnn_hid_dim']
  d_ctrl, d_attn, _ = input_depths(opt)
  w = {}

  ch = [d_ctrl] + list(opt['ctrl_cnn_depth'])
  for i in range(len(ch) - 1):
    w['ctrl_cnn_w_%d' % i] = _normal(rng, (3, 3, ch[i], ch[i + 1]), 9 * ch[i])
    w['ctrl_cnn_b_%d' % i] = (rng.standard_normal(ch[i + 1]) * 0.1).astype(np.float32)
    for t in range(T):
      _bn(rng, w, 'ctrl_cnn', i, t, ch[i + 1])
  feat = ch[-1]
  sub = int(np.prod(opt['ctrl_cnn_pool']))
  gdim = (H // sub) * (W // sub)

  for g in 'ifuo':
    w['ctrl_lstm_w_x' + g] = _normal(rng, (feat, hid), feat)
    w['ctrl_lstm_w_h' + g] = _normal(rng, (hid, hid), hid)
    w['ctrl_lstm_b_' + g] = np.full(hid, 1.0 if g == 'f' else 0.0, np.float32)

  n_g = opt['num_glimpse_mlp_layers']
  gd = [hid] * n_g + [gdim]
  for i in range(n_g):
    scale = 4.0 if i == n_g - 1 else 1.0  # peaky glimpse maps
    w['glimpse_mlp_w_%d' % i] = _normal(rng, (gd[i], gd[i + 1]), gd[i]) * np.float32(scale)
    w['glimpse_mlp_b_%d' % i] = (rng.standard_normal(gd[i + 1]) * 0.1).astype(np.float32)

  n_c = opt['num_ctrl_mlp_layers']
  cd = [hid] + [opt['ctrl_mlp_dim']] * (n_c - 1) + [9]
  for i in range(n_c):
    w['ctrl_mlp_w_%d' % i] = _normal(rng, (cd[i], cd[i + 1]), cd[i])
    w['ctrl_mlp_b_%d' % i] = np.zeros(cd[i + 1], np.float32)
  last = n_c - 1
  w['ctrl_mlp_w_%d' % last][:, 0:2] *= 1.5  # spread the box centres
  w['ctrl_mlp_w_%d' % last][:, 2:4] *= 2.0
  #            ctr_y ctr_x lg_sy lg_sx lg_vy lg_vx  lg_g_attn lg_g_box lg_g_y
  w['ctrl_mlp_b_%d' % last][:] = [0.0, 0.0, -1.3, -1.3, 1.0, 1.0, 0.0, 3.0, 3.5]

  if model == 'box':
    w['score_mlp_w_0'] = _normal(rng, (hid, 1), hid)
    w['score_mlp_b_0'] = np.zeros(1, np.float32)
    return w

  ach = [d_attn] + list(opt['attn_cnn_depth'])
  for i in range(len(ach) - 1):
    w['attn_cnn_w_%d' % i] = _normal(rng, (3, 3, ach[i], ach[i + 1]), 9 * ach[i])
    w['attn_cnn_b_%d' % i] = (rng.standard_normal(ach[i + 1]) * 0.1).astype(np.float32)
    for t in range(T):
      _bn(rng, w, 'attn_cnn', i, t, ach[i + 1])

  asub = int(np.prod(opt['attn_cnn_pool']))
  core_dim = (opt['filter_height'] // asub) * (opt['filter_width'] // asub) * ach[-1]
  w['score_mlp_w_0'] = _normal(rng, (hid + core_dim, 1), hid + core_dim)
  w['score_mlp_b_0'] = np.zeros(1, np.float32)

  dch = [ach[-1]] + list(opt['attn_dcnn_depth'])
  skip_ch = dcnn_skip_channels(opt)
  in_ch = dch[0]
  for i in range(len(dch) - 1):
    in_ch_i = in_ch + skip_ch[i]
    # conv2d_transpose filter layout [kh, kw, Cout, Cin+skip] (nnlib.py:320-325)
    w['attn_dcnn_w_%d' % i] = _normal(rng, (3, 3, dch[i + 1], in_ch_i), 9 * in_ch_i / float(opt['attn_dcnn_pool'][i]**2))
    w['attn_dcnn_b_%d' % i] = (rng.standard_normal(dch[i + 1]) * 0.1).astype(np.float32)
    for t in range(T):
      _bn(rng, w, 'attn_dcnn', i, t, dch[i + 1])
      if i == len(dch) - 2:  # keep the 1-channel mask patch away from all-zero after the ReLU
        w['attn_dcnn_%d_%d_beta' % (i, t)] += np.float32(0.7)
    in_ch = dch[i + 1]
  return w


def dcnn_skip_channels(opt):
  """Skip-connection widths of the deconv mask head (full_model.py:494-502, with the
  ``attn_cnn_skip`` string bug of full_model_train.py:599,640: when ``add_skip_conn`` is on,
  EVERY layer >= 1 gets a skip — SURVEY §9.5)."""
  n = len(opt['attn_dcnn_filter_size'])
  if not opt.get('add_skip_conn', True):
    return [0] * n
  _, d_attn, _ = input_depths(opt)
  ach = [d_attn] + list(opt['attn_cnn_depth'])
  rev = ach[::-1][1:] + [d_attn]
  return ([0] + rev)[:n]


def knob_probability(opt, global_step, offset_key):
  """full_model.py:598-625: min(1, knob_base * knob_decay^(max(0, step - offset) / steps_per_knob_decay) * (1 + log(1 + 3t)))
  per decode step t (time scale only with knob_use_timescale)."""
  T = opt['timespan']
  step = max(0.0, float(global_step) - float(opt[offset_key]))
  p = np.float32(opt['knob_base']) * np.power(np.float32(opt['knob_decay']),
                                               np.float32(step / float(opt['steps_per_knob_decay'])))
  scale = (1.0 + np.log(1.0 + np.arange(T, dtype=np.float32) * 3.0)) if opt.get('knob_use_timescale', True) \
      else np.ones(T, np.float32)
  return np.minimum(np.float32(1.0), p * scale.astype(np.float32)).astype(np.float32)


def make_knob_draws(opt, batch_size, global_step=0, seed=99, device=None):
  """The random draws of one training step of the scheduled-sampling knob as explicit arrays (TensorFlow's streams
  cannot be reproduced): Bernoulli box / mask switches (full_model.py:608-610,622-624), the padding
  ratio and centre shift of the noisy GT boxes (:573-579) and the per-step canvas noise (:836-837).
  device=None: numpy arrays from a seeded numpy generator (parity tests: the oracle and the CUDA path see the same
  bits).  device='cuda...': torch tensors drawn ON the device, like the reference's tf.random_uniform nodes inside
  the graph - a training loop must not ship B*T*H*W floats of noise over PCIe every step."""
  B, T, H, W = batch_size, opt['timespan'], opt['inp_height'], opt['inp_width']
  pb = knob_probability(opt, global_step, 'knob_box_offset').reshape(1, T)
  ps = knob_probability(opt, global_step, 'knob_segm_offset').reshape(1, T)
  r = opt['attn_box_padding_ratio']
  if device is not None:
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))

    def uni(shape, lo, hi):
      return torch.rand(shape, generator=g, device=device, dtype=torch.float32) * (hi - lo) + lo

    pb_d = torch.from_numpy(pb.astype(np.float32)).to(device)
    ps_d = torch.from_numpy(ps.astype(np.float32)).to(device)
    return {
        'gt_knob_box': (uni((B, T), 0.0, 1.0) <= pb_d).to(torch.float32),
        'gt_knob_segm': (uni((B, T), 0.0, 1.0) <= ps_d).to(torch.float32),
        'gt_box_pad': uni((B, T, 1), r - opt['gt_box_pad_noise'], r + opt['gt_box_pad_noise']),
        'gt_box_ctr_shift': uni((B, T, 2), -opt['gt_box_ctr_noise'], opt['gt_box_ctr_noise']),
        'gt_segm_noise': uni((B, T, H, W), 0.0, opt['gt_segm_noise']),
    }
  rng = np.random.default_rng(seed)
  return {
      'gt_knob_box': (rng.random((B, T)) <= pb).astype(np.float32),
      'gt_knob_segm': (rng.random((B, T)) <= ps).astype(np.float32),
      'gt_box_pad': rng.uniform(r - opt['gt_box_pad_noise'], r + opt['gt_box_pad_noise'], (B, T, 1)).astype(np.float32),
      'gt_box_ctr_shift': rng.uniform(-opt['gt_box_ctr_noise'], opt['gt_box_ctr_noise'], (B, T, 2)).astype(np.float32),
      'gt_segm_noise': rng.uniform(0.0, opt['gt_segm_noise'], (B, T, H, W)).astype(np.float32),
  }


def make_fg_weights(opt, seed=4321):
  """Random-init weights of the foreground / orientation FCN (fg_model.py) under the keys its graph registers
  (nnlib scopes 'cnn' and 'dcnn', one BN copy: ``cnn_w_i``, ``cnn_b_i``, ``cnn_{i}_0_{beta,gamma,ema_mean,ema_var}``,
  ``dcnn_w_i`` [3,3,Cout,Cin+skip] ...; the last DCNN layer has no BN, fg_model.py:148)."""
  from .config import fg_skip_wiring
  rng = np.random.default_rng(seed)
  w = {}
  ch = [opt['inp_depth']] + list(opt['cnn_depth'])
  for i in range(len(ch) - 1):
    w['cnn_w_%d' % i] = _normal(rng, (3, 3, ch[i], ch[i + 1]), 9 * ch[i]) * np.float32(1.4)
    w['cnn_b_%d' % i] = (rng.standard_normal(ch[i + 1]) * 0.1).astype(np.float32)
    _bn(rng, w, 'cnn', i, 0, ch[i + 1])
  _, skip_ch = fg_skip_wiring(opt)
  dch = [ch[-1]] + list(opt['dcnn_depth'])
  n_d = len(dch) - 1
  for i in range(n_d):
    cin = dch[i] + skip_ch[i]
    w['dcnn_w_%d' % i] = _normal(rng, (3, 3, dch[i + 1], cin), 9 * cin / float(opt['dcnn_pool'][i]**2)) * np.float32(1.4)
    w['dcnn_b_%d' % i] = (rng.standard_normal(dch[i + 1]) * 0.1).astype(np.float32)
    if i < n_d - 1:
      _bn(rng, w, 'dcnn', i, 0, dch[i + 1])
    else:
      w['dcnn_w_%d' % i] *= np.float32(0.3)  # moderate logits: the heads are not saturated
  return w


def make_fg_batch(opt, batch_size, seed=1234):
  """x [B,H,W,3]; y_gt [B,H,W,nsc] (one channel: foreground; several: one-hot classes, channel 0 = background);
  d_gt [B,H,W,8] one-hot orientation inside the foreground when the model has the orientation head."""
  rng = np.random.default_rng(seed)
  H, W, B = opt['inp_height'], opt['inp_width'], batch_size
  nsc = opt.get('num_semantic_classes', 1)
  x = rng.random((B, H, W, 3), dtype=np.float32)
  yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
  cls = np.zeros((B, H, W), np.int64)
  for b in range(B):
    for _ in range(4):
      cy, cx = rng.uniform(0, H), rng.uniform(0, W)
      ay, ax = rng.uniform(H / 8.0, H / 3.0), rng.uniform(W / 8.0, W / 3.0)
      m = ((yy - cy) / ay)**2 + ((xx - cx) / ax)**2 <= 1.0
      cls[b][m] = 1 if nsc == 1 else int(rng.integers(1, nsc))
  if nsc == 1:
    y_gt = (cls > 0).astype(np.float32)[..., None]
  else:
    y_gt = np.eye(nsc, dtype=np.float32)[cls]
  batch = {'x': x, 'y_gt': y_gt}
  if opt.get('add_orientation', False):
    no = opt['num_orientation_classes']
    ang = np.arctan2(yy - H / 2.0, xx - W / 2.0)
    k = ((ang + np.pi) / (2 * np.pi) * no).astype(np.int64) % no
    batch['d_gt'] = (np.eye(no, dtype=np.float32)[k][None] * (cls > 0)[..., None]).astype(np.float32)
  return batch

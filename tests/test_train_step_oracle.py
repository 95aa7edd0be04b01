"""The training-step oracle (oracle/grads.py): gradients of the training-mode loss through the oracle's forward graph
(torch.autograd standing in for TensorFlow's autodiff of the same graph), clip + Adam, EMA shadow update.

The gradients are pinned to the forward oracle by central differences in float64; the stop-gradients of the reference
(canvas, full_model.py:846-848; Hungarian, modellib.py:11) are tested explicitly.  CPU only."""
import numpy as np
import pytest
import torch

import rec_attend_b200 as ra
from conftest import oracle_fp64
from oracle import grads as OG
from oracle import model as OM


def _setup(arch, H, W, T, B, knob, **over):
  opt = ra.config.full_model_opt(arch, H, W, T, use_knob=knob, **over)
  batch = ra.synthetic.make_batch(opt, B, seed=21)
  weights = ra.synthetic.make_weights(opt, seed=4321)
  draws = ra.synthetic.make_knob_draws(opt, B, global_step=9000, seed=3) if knob else None
  if knob:
    draws['gt_knob_box'][:, 0] = [1, 0][:B] + [0] * (B - 2)
    draws['gt_knob_segm'][:, 0] = [0, 1][:B] + [0] * (B - 2)
  return opt, batch, weights, draws


PROBES = ['ctrl_cnn_w_0', 'ctrl_cnn_b_3', 'ctrl_cnn_2_1_gamma', 'ctrl_cnn_5_0_beta', 'ctrl_lstm_w_xi', 'ctrl_lstm_w_hf',
          'ctrl_lstm_b_o', 'glimpse_mlp_w_0', 'glimpse_mlp_w_1', 'ctrl_mlp_w_0', 'ctrl_mlp_b_0', 'attn_cnn_w_1',
          'attn_cnn_3_0_gamma', 'attn_dcnn_w_2', 'attn_dcnn_b_6', 'attn_dcnn_6_1_beta', 'score_mlp_w_0']


@pytest.mark.parametrize('arch,H,W,knob', [('cvppp', 64, 64, False), ('kitti', 64, 64, True)])
def test_gradients_match_central_differences_in_float64(arch, H, W, knob):
  """With the canvas gradient flowing (stop_canvas_grad=False) the autograd result is the true gradient of the
  training-mode loss and must agree with central differences.  The loss has high curvature in the controller
  weights (sigmoid box edges: d loss / d ctrl_mlp_b is O(100)), so the step is 1e-9 in float64."""
  T, B = 2, 2
  opt, batch, weights, draws = _setup(arch, H, W, T, B, knob, stop_canvas_grad=False)
  O64 = oracle_fp64()
  w64 = {k: np.asarray(v, np.float64) for k, v in weights.items()}
  b64 = {k: np.asarray(v, np.float64) for k, v in batch.items()}
  d64 = None if draws is None else {k: np.asarray(v, np.float64) for k, v in draws.items()}
  torch.set_default_dtype(torch.float64)
  try:
    grads, out = OG.full_model_grads(opt, w64, b64, draws=d64, model_module=O64, dtype=torch.float64)

    def loss_at(key, idx, delta):
      w = dict(w64)
      a = w64[key].copy()
      a[idx] += delta
      w[key] = a
      with torch.no_grad():
        return float(O64.full_model_forward(opt, w, b64, phase_train=True, draws=d64)['loss'])

    rng = np.random.default_rng(0)
    checked = 0
    for key in PROBES:
      g = grads[key]
      assert g is not None and g.shape == w64[key].shape, key
      # probe the entry with the largest gradient and one random entry
      for idx in (np.unravel_index(np.abs(g).argmax(), g.shape), tuple(rng.integers(0, s) for s in g.shape)):
        eps = 1e-9 * max(1.0, abs(float(w64[key][idx])))
        fd = (loss_at(key, idx, eps) - loss_at(key, idx, -eps)) / (2 * eps)
        assert fd == pytest.approx(float(g[idx]), rel=1e-3, abs=1e-4), (key, idx, fd, float(g[idx]))  # abs: round-off of the loss / eps
        checked += 1
    assert checked == 2 * len(PROBES)
  finally:
    torch.set_default_dtype(torch.float32)
  # the fp32 gradients the optimiser sees point the same way as the float64 ones.  They cannot agree tightly: the
  # loss is stiff in the controller weights (the central differences above needed a 1e-9 step: the gradient itself
  # changes by a factor of 4 over a 1e-4 weight perturbation), so fp32 round-off in the forward pass moves them by
  # ~10 %.  A conv bias in front of a batch-statistics BN has an exactly zero gradient (pure round-off): skipped.
  g32, _ = OG.full_model_grads(opt, weights, batch, draws=draws)
  for key in PROBES:
    a, b = g32[key].astype(np.float64).ravel(), grads[key].ravel()
    if '_cnn_b_' in key or 'dcnn_b_' in key:
      assert float(np.abs(b).max()) < 1e-9, key
      continue
    cos = float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
    assert cos > 0.95, (key, cos)
  assert set(g32) == set(OG.trainable_keys(weights)) and not any(k.endswith(('_ema_mean', '_ema_var')) for k in g32)


def test_stop_gradients_of_the_reference():
  T, B = 2, 2
  opt, batch, weights, _ = _setup('cvppp', 64, 64, T, B, False)
  g_stop, _ = OG.full_model_grads(opt, weights, batch)
  g_flow, _ = OG.full_model_grads(dict(opt, stop_canvas_grad=False), weights, batch)
  # with the canvas gradient stopped, the mask head of step 0 gets no gradient from step 1's input canvas
  assert not np.allclose(g_stop['attn_dcnn_6_0_beta'], g_flow['attn_dcnn_6_0_beta'])
  # the last step's BN copy never feeds a later canvas: identical either way
  assert np.allclose(g_stop['attn_dcnn_6_1_beta'], g_flow['attn_dcnn_6_1_beta'], rtol=1e-5, atol=1e-9)
  # weight decay: differentiating total_loss adds exactly wd * w on weight matrices, nothing elsewhere (nnlib.py:59-61)
  g_data, _ = OG.full_model_grads(opt, weights, batch, include_weight_decay=False)
  wd = np.float32(opt['weight_decay'])
  assert np.allclose(g_stop['ctrl_lstm_w_hi'] - g_data['ctrl_lstm_w_hi'], wd * weights['ctrl_lstm_w_hi'], atol=1e-7)
  assert np.allclose(g_stop['ctrl_cnn_b_0'], g_data['ctrl_cnn_b_0'], atol=1e-9)
  assert np.allclose(g_stop['ctrl_cnn_0_0_gamma'], g_data['ctrl_cnn_0_0_gamma'], atol=1e-9)
  # frozen variables (checkpoint.apply_pretrained) are not differentiated at all
  g_frozen, _ = OG.full_model_grads(opt, weights, batch, frozen=['ctrl_lstm_w_hi', 'ctrl_cnn_w_0'])
  assert 'ctrl_lstm_w_hi' not in g_frozen and 'ctrl_cnn_b_0' in g_frozen


def test_train_step_updates():
  T, B = 2, 2
  opt, batch, weights, _ = _setup('cvppp', 64, 64, T, B, False)
  keys = OG.trainable_keys(weights)
  m = {k: np.zeros_like(weights[k]) for k in keys}
  v = {k: np.zeros_like(weights[k]) for k in keys}
  w1, m1, v1, out = OG.train_step(opt, weights, batch, m, v, 0)
  lr = opt['base_learn_rate']
  # first Adam step: |delta| = lr * |g| / (|g| + eps') <= lr, and = lr wherever the gradient is not tiny
  d = np.abs(w1['ctrl_lstm_w_xi'] - weights['ctrl_lstm_w_xi'])
  assert float(d.max()) <= lr * (1 + 1e-4) and float(np.median(d)) > 0.5 * lr
  # EMA shadows moved by the training-mode forward, 0.1 of the way to the batch statistics (nnlib.py:101-108)
  k = 'ctrl_cnn_2_1_ema_var'
  assert not np.allclose(w1[k], weights[k])
  assert np.allclose(w1[k], out['ema_updates'][k].numpy())
  # loss goes down over a few steps on a fixed batch
  losses = [float(out['loss'])]
  w, mm, vv = w1, m1, v1
  for step in range(1, 6):
    w, mm, vv, o = OG.train_step(opt, w, batch, mm, vv, step)
    losses.append(float(o['loss']))
  assert losses[-1] < losses[0], losses
  # data parallel: two ranks with the same gradient = one rank (mean over ranks, clip after averaging)
  g, _ = OG.full_model_grads(opt, weights, batch, include_weight_decay=False)
  w2, _, _, _ = OG.train_step(opt, weights, batch, m, v, 0, world_grads=[g])
  assert all(np.allclose(w2[k], w1[k], atol=1e-8) for k in keys)
  # frozen variables keep their values and have no slots
  fr = ['ctrl_cnn_w_0', 'ctrl_cnn_b_0']
  keys_f = OG.trainable_keys(weights, fr)
  w3, m3, _, _ = OG.train_step(opt, weights, batch, {k: m[k] for k in keys_f}, {k: v[k] for k in keys_f}, 0, frozen=fr)
  assert np.array_equal(w3['ctrl_cnn_w_0'], weights['ctrl_cnn_w_0']) and 'ctrl_cnn_w_0' not in m3


@pytest.mark.parametrize('name', ['cvppp', 'kitti'])
def test_gradient_oracle_equals_derivatives_of_the_reference_graph(name):
  """tests/golden/reference_fd_golden.npz holds central differences of the REFERENCE'S OWN loss (full_model.get_model
  executed unmodified over the TF-0.12 stand-in, float64, canvas gradient not stopped).  torch.autograd through the
  oracle, in the same setting, must give the same derivatives: the gradient oracle is then pinned to the reference's
  code (up to where it stops gradients, full_model.py:846-848, restated by inspection)."""
  import json
  import os
  G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_fd_golden.npz'))
  meta = json.loads(str(G[name + '/meta']))
  opt = ra.config.full_model_opt(meta['arch'], meta['H'], meta['W'], meta['T'], **meta['overrides'])
  assert opt['stop_canvas_grad'] is False
  batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_batch(opt, meta['B'], seed=meta['batch_seed']).items()}
  w64 = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_weights(opt, seed=meta['weight_seed']).items()}
  O64 = oracle_fp64()
  torch.set_default_dtype(torch.float64)
  try:
    grads, _ = OG.full_model_grads(opt, w64, batch, include_weight_decay=True, model_module=O64, dtype=torch.float64)
  finally:
    torch.set_default_dtype(torch.float32)
  checked = 0
  for row in meta['rows']:
    g = float(grads[row['key']][tuple(row['idx'])])
    assert row['fd'] == pytest.approx(g, rel=2e-3, abs=2e-4), (row['key'], row['idx'], row['fd'], g)
    checked += 1
  assert checked == len(meta['rows']) >= 30
  big = [abs(r['fd']) for r in meta['rows']]
  assert max(big) > 1.0  # the probes are not all trivially small

"""oracle.model.full_model_forward against golden vectors produced by EXECUTING THE REFERENCE'S OWN
full_model.get_model(opt) (tests/golden/make_full_model_golden.py: full_model.py, nnlib.py, modellib.py, image_ops.py
imported unmodified and run eagerly over the numpy stand-in of tests/golden/tf012_shim).  Training mode (batch-
statistics BN), three architectures, with and without scheduled sampling (the graph's own random draws are replayed
into the oracle), incl. the use_iou_box form.  This pins the whole T-step decode + matching loss block of the oracle
- hence the CUDA path that is held to the oracle - to the reference's code; TensorFlow's kernel semantics (the
shim's one-line ops) are what remains on trust.

Both sides run in DOUBLE precision (the shim with TF012_SHIM_DTYPE=float64, the oracle as its float64 twin): the
training-mode decode loop amplifies fp32 round-off ~10x per step, which would blur a float32 comparison; in float64
the two implementations must agree to ~1e-8 everywhere, matchings exactly."""
import json
import os

import numpy as np
import pytest
import torch

import rec_attend_b200 as ra
from conftest import oracle_fp64

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'full_model_golden.npz'))
CASES = sorted({k.split('/')[0] for k in G.files})


def rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  assert a.shape == b.shape, (a.shape, b.shape)
  return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-9))


@pytest.mark.parametrize('name', CASES)
def test_oracle_equals_the_reference_graph(name):
  meta = json.loads(str(G[name + '/meta']))
  opt = ra.config.full_model_opt(meta['arch'], meta['H'], meta['W'], meta['T'], **meta['overrides'])
  B, T = meta['B'], meta['T']
  batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_batch(opt, B, seed=meta['batch_seed']).items()}
  w = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_weights(opt, seed=meta['weight_seed']).items()}
  assert sum(float(np.abs(v).sum()) for v in w.values()) == pytest.approx(float(G[name + '/weights_checksum']), rel=1e-12), \
      'synthetic.make_weights changed: regenerate the fixture'
  phase = meta.get('phase_train', True)
  if not phase:  # eval mode: the EMA shadows of the reference's fresh graph are zero
    w = {k: (np.zeros_like(v) if k.endswith(('_ema_mean', '_ema_var')) else v) for k, v in w.items()}
  draws = None
  if opt['use_knob'] and phase:
    step = meta['global_step']
    p_box = ra.synthetic.knob_probability(opt, step, 'knob_box_offset')
    p_segm = ra.synthetic.knob_probability(opt, step, 'knob_segm_offset')
    draws = {'gt_box_pad': G[name + '/draw_box_pad'].astype(np.float64),
             'gt_box_ctr_shift': G[name + '/draw_ctr_shift'].astype(np.float64),
             'gt_knob_box': (G[name + '/draw_knob_box_u'][..., 0] <= p_box[None, :]).astype(np.float64),
             'gt_knob_segm': (G[name + '/draw_knob_segm_u'][..., 0] <= p_segm[None, :]).astype(np.float64),
             'gt_segm_noise': G[name + '/draw_segm_noise'].astype(np.float64)}
    # both branches of both switches occur in the replayed draws
    assert 0 < draws['gt_knob_box'].mean() < 1 or 0 < draws['gt_knob_segm'].mean() < 1
  O64 = oracle_fp64()
  torch.set_default_dtype(torch.float64)
  try:
    with torch.no_grad():
      o = O64.full_model_forward(opt, w, batch, phase_train=phase, draws=draws)
  finally:
    torch.set_default_dtype(torch.float32)
  ref = lambda k: G['%s/%s' % (name, k)]
  for k in ('match', 'match_box'):
    assert np.array_equal(o[k].numpy(), ref(k)), k
  for k in ('iou_hard', 'wt_cov_hard', 'unwt_cov_hard', 'dice', 'count_acc', 'dic', 'dic_abs', 'loss', 'box_loss',
            'segm_loss', 'conf_loss', 'iou_soft', 'wt_cov_soft', 'unwt_cov_soft'):
    assert float(o[k]) == pytest.approx(float(ref(k)), rel=1e-7, abs=1e-9), k
  for k in ('y_out', 's_out', 'attn_box', 'x_patch', 'y_out_patch', 'attn_ctr', 'attn_size', 'attn_top_left',
            'attn_bot_right', 'ctrl_rnn_glimpse_map'):
    if '%s/%s' % (name, k) not in G.files:
      continue
    tol = 1e-6 if ref(k).dtype == np.float32 else 1e-7  # the big image tensors are stored in float32
    assert rel(o[k].numpy(), ref(k)) < tol, (k, rel(o[k].numpy(), ref(k)))
  assert float(ref('learn_rate')) == pytest.approx(
      opt['base_learn_rate'] * opt['learn_rate_decay']**(meta['global_step'] // opt['steps_per_learn_rate_decay']), rel=1e-6)


GB = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'box_model_golden.npz'))


@pytest.mark.parametrize('name', sorted({k.split('/')[0] for k in GB.files}))
def test_box_oracle_equals_the_reference_graph(name):
  """oracle.model.box_model_forward against the reference's own box_model.get_model(opt) executed over the shim
  (tests/golden/make_box_model_golden.py), float64, training mode, the graph's canvas noise replayed."""
  meta = json.loads(str(GB[name + '/meta']))
  opt = ra.config.box_model_opt(meta['H'], meta['W'], meta['T'], **meta['overrides'])
  batch = {k: np.asarray(v, np.float64) for k, v in ra.synthetic.make_batch(opt, meta['B'], seed=meta['batch_seed']).items()}
  w = {k: np.asarray(v, np.float64)
       for k, v in ra.synthetic.make_weights(opt, seed=meta['weight_seed'], model='box').items()}
  assert sum(float(np.abs(v).sum()) for v in w.values()) == pytest.approx(float(GB[name + '/weights_checksum']), rel=1e-12)
  O64 = oracle_fp64()
  torch.set_default_dtype(torch.float64)
  try:
    with torch.no_grad():
      o = O64.box_model_forward(opt, w, batch, canvas_noise=GB[name + '/draw_canvas_noise'].astype(np.float64),
                                phase_train=True)
  finally:
    torch.set_default_dtype(torch.float32)
  assert np.array_equal(o['match_box'].numpy(), GB[name + '/match_box'])
  for k in ('loss', 'box_loss', 'conf_loss'):
    assert float(o[k]) == pytest.approx(float(GB['%s/%s' % (name, k)]), rel=1e-7, abs=1e-9), k
  for k in ('attn_box', 's_out', 'attn_ctr', 'attn_size', 'attn_top_left', 'attn_bot_right', 'attn_top_left_gt',
            'attn_bot_right_gt', 'ctrl_rnn_glimpse_map'):
    ref = GB['%s/%s' % (name, k)]
    tol = 1e-6 if ref.dtype == np.float32 else 1e-7
    assert rel(o[k].numpy().reshape(ref.shape), ref) < tol, (k, rel(o[k].numpy().reshape(ref.shape), ref))

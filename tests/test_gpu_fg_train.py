"""Training mode of the foreground / orientation FCN (fg_model.py:96-172,249-266) on the GPU against
the oracle: batch-statistics forward, every gradient tensor against autograd through oracle.model.fg_model_forward,
the Adam update against oracle.optim, the moved EMA shadows and the eval forward after training."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import grads as OG
from oracle import model as OM
from oracle import optim as OO

pytestmark = pytest.mark.gpu


def _small_opt(ra, ori, nsc, loss):
  opt = ra.config.fg_model_opt('kitti', 32, 64)
  opt.update({'cnn_depth': [8, 8, 16, 16], 'cnn_pool': [1, 2, 1, 2], 'dcnn_depth': [16, 8, 8, nsc + (8 if ori else 0)],
              'dcnn_pool': [2, 1, 2, 1], 'num_semantic_classes': nsc, 'add_orientation': ori, 'segm_loss_fn': loss})
  for k in ('cnn_skip_mask', 'dcnn_skip_mask', 'cnn_skip'):
    opt.pop(k, None)
  return opt


@pytest.mark.parametrize('ori,nsc,loss', [(True, 1, 'bce'), (False, 1, 'iou'), (True, 3, 'iou'), (False, 3, 'bce')])
def test_fg_training_forward_gradients_and_step(cuda, ori, nsc, loss):
  import rec_attend_b200 as ra
  from rec_attend_b200.fg_model import FgModel
  opt = _small_opt(ra, ori, nsc, loss)
  B = 3
  weights = ra.synthetic.make_fg_weights(opt, seed=11)
  batch = ra.synthetic.make_fg_batch(opt, B, seed=12)
  grads, ref = OG.fg_model_grads(opt, weights, batch)
  g64, _ = OG.fg_model_grads(opt, weights, batch, dtype=torch.float64)
  model = FgModel(opt).load_weights(weights)
  out, tape = model._forward_train(batch)
  torch.cuda.synchronize()
  for k in ('y_out', 'logits'):
    assert rel_err(out[k].cpu().numpy(), ref[k].numpy()) < 1e-3, k
  if ori:
    assert rel_err(out['d_out'].cpu().numpy(), ref['d_out'].numpy()) < 1e-3
  assert abs(float(out['loss']) - float(ref['loss'])) <= 1e-4 * max(1.0, abs(float(ref['loss'])))
  grad = model._backward(tape)
  torch.cuda.synchronize()
  adam = model.optimizer
  gpu = adam.flat.unflatten(grad.cpu().numpy())
  assert set(gpu) == set(grads)
  n_d = len(opt['dcnn_depth'])
  bn_bias = set('cnn_b_%d' % i for i in range(len(opt['cnn_depth']))) | set('dcnn_b_%d' % i for i in range(n_d - 1))
  bad = []
  for k in sorted(gpu):
    r64 = g64[k]
    scale = max(float(np.abs(r64).max()), 1e-12)
    e_gpu = float(np.abs(gpu[k] - r64).max())
    e_ref = float(np.abs(grads[k] - r64).max())
    if k in bn_bias:
      # a bias in front of a batch-statistics BN: the mean removes it, the gradient is zero up to round-off in ANY
      # implementation (1e-17 in float64) - only its smallness relative to the filter gradient can be asserted
      assert float(np.abs(gpu[k]).max()) <= 1e-3 * max(float(np.abs(gpu[k.replace('_b_', '_w_')]).max()), 1e-6), k
      continue
    if e_gpu > max(1e-3 * scale, 20.0 * e_ref):
      bad.append((k, e_gpu / scale, e_ref / scale))
  assert not bad, bad
  # one full step: parameters bit-equal to the oracle's Adam fed with the GPU's gradients (no clipping)
  before = {k: np.asarray(weights[k], np.float32) for k in adam.flat.keys}
  model2 = FgModel(opt).load_weights(weights)
  res = model2.train_step(batch)
  torch.cuda.synchronize()
  assert res['global_step'] == 1 and abs(float(res['loss']) - float(ref['loss'])) <= 1e-4 * max(1.0, abs(float(ref['loss'])))
  new_w = model2.export_weights()
  lr = OO.learn_rate(opt['base_learn_rate'], opt['learn_rate_decay'], opt['steps_per_learn_rate_decay'], 0)
  from rec_attend_b200.optim import has_weight_decay
  keys = adam.flat.keys
  wd = {k: (np.float32(opt['weight_decay']) if has_weight_decay(k) else 0.0) for k in keys}
  zeros = {k: np.zeros_like(before[k]) for k in keys}
  want, _, _ = OO.adam_step(before, {k: gpu[k] for k in keys}, zeros, dict(zeros), wd, lr, 1, clip=0.0)
  for k in keys:
    assert np.array_equal(new_w[k], want[k]), k
  # EMA shadows moved like the oracle's
  assert len(ref['ema_updates']) == 2 * (4 + 3)
  for k, v in ref['ema_updates'].items():
    assert rel_err(new_w[k], v.numpy()) < 2e-3, k
  # eval forward after training uses the new weights and shadows
  ev = model2.forward(batch)
  ref_eval = OM.fg_model_forward(opt, new_w, batch)
  assert rel_err(ev['y_out'].cpu().numpy(), ref_eval['y_out'].numpy()) < 1e-3


def test_fg_training_loss_falls(cuda):
  import rec_attend_b200 as ra
  from rec_attend_b200.fg_model import FgModel
  opt = _small_opt(ra, True, 1, 'bce')
  opt['base_learn_rate'] = 1e-3
  model = FgModel(opt).load_weights(ra.synthetic.make_fg_weights(opt, seed=3))
  batch = ra.synthetic.make_fg_batch(opt, 4, seed=4)
  losses = [float(model.train_step(batch)['loss']) for _ in range(12)]
  assert losses[-1] < losses[0], losses

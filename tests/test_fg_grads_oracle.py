"""The gradient oracle of the foreground / orientation FCN (oracle/grads.py:fg_model_grads) against central differences
of the float64 oracle forward (which is itself pinned to the reference's fg_model.py executed over the TF-0.12 stand-in,
tests/test_model_oracle.py) - CPU only."""
import importlib

import numpy as np
import pytest
import torch

from oracle import grads as OG


def _fp64_oracle():
  spec = importlib.util.find_spec('oracle.model')
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  mod._t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a), dtype=torch.float64)
  return mod


@pytest.mark.parametrize('ori,nsc,loss', [(True, 1, 'bce'), (True, 3, 'iou'), (False, 1, 'iou')])
def test_fg_gradient_oracle_matches_central_differences(ori, nsc, loss):
  import rec_attend_b200.config as config
  import rec_attend_b200.synthetic as synthetic
  opt = config.fg_model_opt('kitti', 16, 32)
  opt.update({'cnn_depth': [4, 8], 'cnn_pool': [2, 2], 'dcnn_depth': [8, 4, nsc + (8 if ori else 0)],
              'dcnn_pool': [2, 2, 1], 'num_semantic_classes': nsc, 'add_orientation': ori, 'segm_loss_fn': loss})
  for k in ('cnn_skip_mask', 'dcnn_skip_mask', 'cnn_skip'):
    opt.pop(k, None)
  weights = {k: np.asarray(v, np.float64) for k, v in synthetic.make_fg_weights(opt, seed=5).items()}
  batch = {k: np.asarray(v, np.float64) for k, v in synthetic.make_fg_batch(opt, 2, seed=6).items()}
  O64 = _fp64_oracle()
  torch.set_default_dtype(torch.float64)
  try:
    grads, out = OG.fg_model_grads(opt, weights, batch, model_module=O64, dtype=torch.float64)

    def loss_at(w):
      return float(O64.fg_model_forward(opt, w, batch, phase_train=True)['loss'])

    rng = np.random.default_rng(0)
    keys = [k for k in sorted(grads) if '_w_' in k or k.endswith(('_gamma', '_beta'))]
    for k in rng.choice(keys, size=8, replace=False):
      g = grads[k]
      idx = tuple(int(rng.integers(0, n)) for n in g.shape)
      h = 1e-6
      wp = {kk: v.copy() for kk, v in weights.items()}
      wm = {kk: v.copy() for kk, v in weights.items()}
      wp[k][idx] += h
      wm[k][idx] -= h
      fd = (loss_at(wp) - loss_at(wm)) / (2 * h)
      assert abs(fd - g[idx]) <= 1e-5 * max(1.0, abs(fd)) + 1e-7, (k, idx, fd, g[idx])
  finally:
    torch.set_default_dtype(torch.float32)

"""ORACLE (test infrastructure only) of the reference's TRAINING STEP: the gradients TensorFlow's autodiff would hand
to the optimiser block and the resulting update (full_model.py:1039-1057; Trainer.run_step, runner.py:62-89).

  grads = d total_loss / d variable  for every trainable variable        optimizer.compute_gradients, :1049
  g     = clip_by_value(grads, -1, 1)                                    :1050-1055
  var  <- Adam(eps 1e-7, lr = staircase decay)                           :1048,1056

The forward graph is oracle.model.full_model_forward(phase_train=True) — batch-statistics BN (differentiated through
the moments like tf.nn.moments), scheduled sampling with the draws as inputs, tf.stop_gradient on the canvas
(full_model.py:846-848), no gradient through the Hungarian op (modellib.py:11) — and torch.autograd stands in for
TensorFlow's autodiff of the same graph.  The forward graph is pinned to the reference's own source (see oracle/model.py); TensorFlow's
autodiff itself cannot run here, so the gradients are pinned (tests/test_train_step_oracle.py) by central differences
in float64 - of the forward oracle, and of the REFERENCE GRAPH'S OWN LOSS (full_model.get_model executed over the
TF-0.12 stand-in, tests/golden/make_reference_fd_golden.py).  What neither can see is where the reference stops
gradients (two lines, full_model.py:846-848), restated by inspection.

The CUDA backward pass does not exist yet; this module is the checker it will be built against, and
`train_step` + oracle.optim.adam_step is the complete CPU restatement of one training step today.
"""
import numpy as np
import torch

from . import model as OM
from . import optim as OO

NON_TRAINABLE = ('_ema_mean', '_ema_var')


def trainable_keys(weights, frozen=()):
  frozen = set(frozen)
  return sorted(k for k in weights if not k.endswith(NON_TRAINABLE) and k not in frozen)


def has_weight_decay(key):
  """nnlib.py:59-61: wd * l2_loss on conv / mlp / lstm weight matrices only."""
  return '_w_' in key and not key.endswith(('_beta', '_gamma'))


def full_model_grads(opt, weights, batch, draws=None, include_weight_decay=True, frozen=(), model_module=OM,
                     dtype=torch.float32):
  """Gradients of the training-mode loss by weight key.

  include_weight_decay=True differentiates total_loss as the reference builds it (data loss + sum wd*||w||^2/2);
  False leaves the decay term out - the form `ra_adam_step_f32` expects, which adds wd*w itself.
  A variable the loss does not depend on gets None ('grad is None', full_model.py:1051-1055) - e.g. nothing here,
  but the second glimpse-MLP layer would if the controller ran a single glimpse iteration.
  Returns (grads, forward outputs incl. 'loss' and 'ema_updates')."""
  keys = trainable_keys(weights, frozen)
  leaves = {}
  for k, v in weights.items():
    t = torch.as_tensor(np.asarray(v), dtype=dtype).clone()
    if k in keys:
      t.requires_grad_(True)
    leaves[k] = t
  out = model_module.full_model_forward(opt, leaves, batch, phase_train=True, draws=draws)
  loss = out['loss']
  if not include_weight_decay:
    loss = loss - model_module.weight_decay_loss(opt, leaves)
  g = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
  grads = {k: (None if gi is None else gi.detach().numpy()) for k, gi in zip(keys, g)}
  det = lambda v: v.detach() if isinstance(v, torch.Tensor) else v
  return grads, {k: ({kk: det(vv) for kk, vv in v.items()} if isinstance(v, dict) else det(v)) for k, v in out.items()}


def box_model_grads(opt, weights, batch, canvas_noise=None, include_weight_decay=True, frozen=(), model_module=OM,
                    dtype=torch.float32):
  """Gradients of the box model's training-mode loss (box_model.py:560-634, optimiser block :635-652) by weight key;
  same conventions as full_model_grads."""
  keys = trainable_keys(weights, frozen)
  leaves = {}
  for k, v in weights.items():
    t = torch.as_tensor(np.asarray(v), dtype=dtype).clone()
    if k in keys:
      t.requires_grad_(True)
    leaves[k] = t
  out = model_module.box_model_forward(opt, leaves, batch, canvas_noise=canvas_noise, phase_train=True)
  loss = out['loss']
  if not include_weight_decay:
    loss = loss - model_module.weight_decay_loss(opt, leaves)
  g = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
  grads = {k: (None if gi is None else gi.detach().numpy()) for k, gi in zip(keys, g)}
  det = lambda v: v.detach() if isinstance(v, torch.Tensor) else v
  return grads, {k: ({kk: det(vv) for kk, vv in v.items()} if isinstance(v, dict) else det(v)) for k, v in out.items()}


def fg_model_grads(opt, weights, batch, model_module=OM, dtype=torch.float32):
  """Gradients of the foreground / orientation FCN's training-mode loss (fg_model.py:174-250: batch-statistics BN,
  loss = foreground loss [+ orientation cross-entropy]) by weight key, WITHOUT the weight-decay term (the optimiser
  adds wd * w, like for the instance model).  The reference minimises total_loss with Adam and no gradient clipping
  (fg_model.py:258-265)."""
  keys = trainable_keys(weights)
  leaves = {}
  for k, v in weights.items():
    t = torch.as_tensor(np.asarray(v), dtype=dtype).clone()
    if k in keys:
      t.requires_grad_(True)
    leaves[k] = t
  b = {k: torch.as_tensor(np.asarray(v), dtype=dtype) for k, v in batch.items()}
  out = model_module.fg_model_forward(opt, leaves, b, phase_train=True)
  g = torch.autograd.grad(out['loss'], [leaves[k] for k in keys], allow_unused=True)
  grads = {k: (None if gi is None else gi.detach().numpy()) for k, gi in zip(keys, g)}
  det = lambda v: v.detach() if isinstance(v, torch.Tensor) else v
  return grads, {k: ({kk: det(vv) for kk, vv in v.items()} if isinstance(v, dict) else det(v)) for k, v in out.items()}


def train_step(opt, weights, batch, adam_m, adam_v, global_step, draws=None, frozen=(), world_grads=None):
  """One ``sess.run([loss, train_step])`` (runner.py:98-105): forward + backward + clip + Adam + EMA shadow update.

  weights / adam_m / adam_v: dicts by weight key (slots only for trainable keys; zeros at step 0).
  global_step: the step counter BEFORE this step (the reference's global_step variable).
  world_grads: optional list of gradient dicts from the other data-parallel ranks; the update then uses the mean over
  all ranks (clip AFTER averaging).
  Returns (new_weights, new_m, new_v, forward outputs)."""
  grads, out = full_model_grads(opt, weights, batch, draws=draws, include_weight_decay=False, frozen=frozen)
  keys = trainable_keys(weights, frozen)
  if world_grads:
    n = 1 + len(world_grads)
    for k in keys:
      parts = [g[k] for g in [grads] + list(world_grads) if g.get(k) is not None]
      grads[k] = None if not parts else (np.sum(parts, axis=0, dtype=np.float32) / np.float32(n)).astype(np.float32)
  var = {k: np.asarray(weights[k], np.float32) for k in keys}
  wd = {k: (np.float32(opt['weight_decay']) if has_weight_decay(k) else 0.0) for k in keys}
  lr = OO.learn_rate(opt['base_learn_rate'], opt['learn_rate_decay'], opt['steps_per_learn_rate_decay'], global_step)
  m = {k: np.asarray(adam_m[k], np.float32) for k in keys}
  v = {k: np.asarray(adam_v[k], np.float32) for k in keys}
  var, m, v = OO.adam_step(var, grads, m, v, wd, lr, global_step + 1, clip=float(opt.get('clip_gradient', 1.0)))
  new_w = {k: np.asarray(a, np.float32) for k, a in weights.items()}
  new_w.update(var)
  for k, a in out.get('ema_updates', {}).items():  # the EMA shadows moved by this forward (nnlib.py:101-108)
    new_w[k] = a.numpy()
  return new_w, m, v, out

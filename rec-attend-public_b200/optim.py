"""Optimiser block of the reference (full_model.py:1039-1057, box_model.py:635-652) and the gradient all-reduce
of the data-parallel training step, on ONE flat fp32 bucket.

Every trainable tensor of the reference's weight dict (conv / mlp / lstm weights and biases, per-(layer, step) BN
beta / gamma; the EMA shadows are not trainable, nnlib.py:104-127) is laid out in one contiguous CUDA buffer, in
sorted key order.  A training step then needs exactly one collective and one kernel:

  grads (flat, this rank's shard of the batch)  --NCCL all-reduce(SUM)-->  x 1/world  -->  ra_adam_step_f32
  (weight-decay gradient, clip to [-1, 1], Adam with eps 1e-7, TF-0.12 form)

Clip-after-average keeps the reference's global-batch semantics.  The gradients come from the backward pass of
`train.py` (FullModel / BoxModel.train_step) or `fg_model.py`; this module is the tail of the step.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops

ADAM_EPS = 1e-7  # full_model.py:1046
ADAM_BETA1, ADAM_BETA2 = 0.9, 0.999  # tf.train.AdamOptimizer defaults


def trainable_keys(weights, frozen=()):
  """Sorted names of the trainable tensors: everything but the BN EMA shadows and the tensors the freeze_* flags of
  the pretrained hand-off exclude (`checkpoint.apply_pretrained`; tf.Variable(trainable=False), nnlib.py:185,455,527)."""
  frozen = set(frozen)
  return sorted(k for k in weights if not k.endswith(('_ema_mean', '_ema_var')) and k not in frozen)


def has_weight_decay(key):
  """nnlib.py:59-61: wd * l2_loss is attached to conv / mlp / lstm WEIGHT matrices only."""
  return '_w_' in key and not key.endswith(('_beta', '_gamma'))


class FlatParams(object):
  """Layout of the trainable tensors in one flat buffer: key -> (offset, shape)."""

  def __init__(self, weights, frozen=()):
    self.keys = trainable_keys(weights, frozen)
    self.layout = {}
    off = 0
    for k in self.keys:
      shape = tuple(np.asarray(weights[k]).shape)
      n = int(np.prod(shape)) if shape else 1
      self.layout[k] = (off, shape)
      off += n
    self.numel = off

  def flatten(self, tensors, dtype=np.float32):
    """dict of arrays -> flat numpy vector (missing keys = zeros: 'grad is None', full_model.py:1051-1055)."""
    out = np.zeros(self.numel, dtype)
    for k in self.keys:
      if k in tensors and tensors[k] is not None:
        off, shape = self.layout[k]
        out[off:off + int(np.prod(shape)) if shape else off + 1] = np.asarray(tensors[k], dtype).reshape(-1)
    return out

  def unflatten(self, flat):
    flat = np.asarray(flat)
    return {k: flat[off:off + (int(np.prod(shape)) if shape else 1)].reshape(shape)
            for k, (off, shape) in self.layout.items()}

  def weight_decay_vector(self, wd):
    out = np.zeros(self.numel, np.float32)
    for k in self.keys:
      if has_weight_decay(k):
        off, shape = self.layout[k]
        out[off:off + int(np.prod(shape))] = wd
    return out


def learn_rate(opt, global_step):
  """tf.train.exponential_decay(base, global_step, steps_per_decay, decay, staircase=True), full_model.py:1039-1044
  (fp32 like the graph)."""
  p = math.floor(float(global_step) / float(opt['steps_per_learn_rate_decay']))
  return float(np.float32(opt['base_learn_rate']) * np.power(np.float32(opt['learn_rate_decay']), np.float32(p)))


def all_reduce_sum_(flat):
  """SUM all-reduce of the flat gradient bucket over the data-parallel ranks (NCCL on GPUs, gloo in the CPU tests).
  One collective per step on 2-4 MB: latency-bound over NVLink/NVSwitch, launched on the compute stream.
  Returns the world size (1 when torch.distributed is not initialised)."""
  if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return dist.get_world_size()
  return 1


class AdamOptimizer(object):
  """The reference's train_step on the flat bucket.  State (params, m, v, wd vector) lives on the GPU."""

  def __init__(self, opt, weights, device=None, frozen=()):
    if not torch.cuda.is_available():
      raise _lib.RecAttendError('rec_attend_b200.optim needs a CUDA device (there is no CPU fallback)')
    _lib.lib()
    self.opt = dict(opt)
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    self.flat = FlatParams(weights, frozen)
    self.params = torch.from_numpy(self.flat.flatten(weights)).to(self.device)
    self.m = torch.zeros_like(self.params)
    self.v = torch.zeros_like(self.params)
    self.wd = torch.from_numpy(self.flat.weight_decay_vector(np.float32(opt['weight_decay']))).to(self.device)
    self.clip = float(opt.get('clip_gradient', 1.0))
    self.global_step = 0  # the reference keeps it as a float variable (full_model.py:587); an int here

  def step(self, grad_flat, grad_scale=None):
    """One train_step: all-reduce(SUM) of `grad_flat` in place, then clip + Adam in one launch.
    `grad_flat` = this rank's gradient of the DATA loss of its batch shard, each rank's loss being the mean over its
    own examples (so the average over ranks is the global-batch mean for equal shards)."""
    ops._chk(grad_flat)
    if grad_flat.numel() != self.params.numel():
      raise _lib.RecAttendError('gradient bucket has {} elements, expected {}'.format(grad_flat.numel(),
                                                                                    self.params.numel()))
    world = all_reduce_sum_(grad_flat)
    lr = learn_rate(self.opt, self.global_step)
    self.global_step += 1
    # grad_scale: this rank's weight in the global-batch mean AFTER the SUM all-reduce; 1/world for equal shards.
    # Unequal shards: every rank passes its own shard_size / global_batch BEFORE calling (scale the bucket) - see
    # dist_util.shard, which refuses uneven splits for training.
    scale = (1.0 / world) if grad_scale is None else float(grad_scale)
    _lib.call('ra_adam_step_f32', ops._p(self.params), ops._p(grad_flat), ops._p(self.m), ops._p(self.v),
              ops._p(self.wd), self.params.numel(), scale, lr, ADAM_BETA1, ADAM_BETA2, ADAM_EPS, self.clip,
              self.global_step, ops._stream())
    return lr

  def state(self):
    """(m, v, global_step) as host arrays, for `checkpoint.pack_state`."""
    return self.m.cpu().numpy(), self.v.cpu().numpy(), self.global_step

  def load_state(self, m, v, global_step):
    """Resume from a checkpoint written with `state()` (same weight schema and frozen set)."""
    for name, src in (('m', m), ('v', v)):
      src = np.asarray(src, np.float32).reshape(-1)
      if src.size != self.params.numel():
        raise _lib.RecAttendError('Adam slot {} has {} elements, expected {}'.format(name, src.size,
                                                                                   self.params.numel()))
      getattr(self, name).copy_(torch.from_numpy(src))
    self.global_step = int(global_step)

  def export_weights(self, weights):
    """The updated trainable tensors merged back into a weight dict (reference key schema)."""
    out = dict(weights)
    out.update(self.flat.unflatten(self.params.cpu().numpy()))
    return out

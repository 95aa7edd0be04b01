"""ORACLE (test infrastructure only).  numpy fp32 restatement of the reference's optimiser block
(full_model.py:1039-1057): per-variable gradient of total_loss = data gradient + wd * w on weight matrices
(nnlib.py:59-61), tf.clip_by_value(g, -1, 1), TensorFlow-0.12 ApplyAdam:
  alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m += (g - m)(1 - beta1);  v += (g^2 - v)(1 - beta2);
  var -= m * alpha / (sqrt(v) + eps)
PARITY UNPINNED against TensorFlow itself (not runnable here); the update rule is the published one."""
import math

import numpy as np

F = np.float32


def learn_rate(base, decay, steps_per_decay, global_step):
  return F(base) * np.power(F(decay), F(math.floor(global_step / float(steps_per_decay))))


def adam_step(var, grad, m, v, wd, lr, t, clip=1.0, beta1=0.9, beta2=0.999, eps=1e-7, grad_scale=1.0):
  """One apply_gradients over a dict of per-variable arrays (in place on copies; returns var, m, v)."""
  b1p, b2p = F(1), F(1)
  for _ in range(t):
    b1p = F(b1p * F(beta1))
    b2p = F(b2p * F(beta2))
  alpha = F(F(lr) * np.sqrt(F(1) - b2p) / (F(1) - b1p))
  var, m, v = {k: a.astype(F).copy() for k, a in var.items()}, dict(m), dict(v)
  for k in var:
    g = grad.get(k)
    g = np.zeros_like(var[k]) if g is None else (g.astype(F) * F(grad_scale)).astype(F)
    if wd.get(k, 0.0):
      g = (g + F(wd[k]) * var[k]).astype(F)
    if clip > 0:
      g = np.clip(g, F(-clip), F(clip))
    m[k] = (m[k] + (g - m[k]) * (F(1) - F(beta1))).astype(F)
    v[k] = (v[k] + (g * g - v[k]) * (F(1) - F(beta2))).astype(F)
    var[k] = (var[k] - (m[k] * alpha) / (np.sqrt(v[k]) + F(eps))).astype(F)
  return var, m, v

"""numpy stand-in for the TensorFlow-0.12 ops used by /root/reference/modellib.py (see README.md).  Eager, float32."""
import os

import numpy as np

# The working precision.  float32 like the reference by default; TF012_SHIM_DTYPE=float64 (set before import) runs the
# same graph in double precision, which separates STRUCTURAL agreement from fp32 round-off in ill-conditioned graphs.
F32 = np.dtype(os.environ.get('TF012_SHIM_DTYPE', 'float32')).type


def _axes(a):
  if a is None:
    return None
  return tuple(int(i) for i in np.ravel(np.asarray(a)))


def _f(x):
  return np.asarray(x, F32)


# ---- conversions / constructors
def to_float(x):
  return np.asarray(x).astype(F32)


def constant(value, dtype=None, shape=None):
  a = np.asarray(value, dtype=np.dtype(dtype) if dtype else None)
  if a.dtype == np.float64:
    a = a.astype(F32)
  return np.broadcast_to(a, shape).copy() if shape is not None else a


def zeros(shape, dtype='float32'):
  return np.zeros(_axes(shape), np.dtype(dtype))


def ones(shape, dtype='float32'):
  return np.ones(_axes(shape), np.dtype(dtype))


def range(start, limit=None, delta=1):  # noqa: A001  (tf.range)
  s, l = (0, start) if limit is None else (start, limit)
  return np.arange(int(s), int(l), int(delta), dtype=np.int32)


def shape(x):
  return np.asarray(np.shape(x), np.int32)


def size(x):
  return np.int32(np.size(x))


def pack(values):  # tf.pack: stack along a new first axis
  return np.stack([np.asarray(v) for v in values])


# ---- shape manipulation (TF-0.12 argument orders)
def reshape(x, shp):
  return np.reshape(x, _axes(shp))


def expand_dims(x, dim):
  return np.expand_dims(x, int(dim))


def concat(concat_dim, values):  # TF 0.12: dimension FIRST
  return np.concatenate([np.asarray(v) for v in values], axis=int(concat_dim))


def split(split_dim, num_split, value):  # TF 0.12: (dim, num, value)
  return np.split(value, int(num_split), axis=int(split_dim))


def slice(x, begin, size):  # noqa: A001  (tf.slice; -1 = to the end)
  idx = tuple(np.s_[int(b):] if int(s) == -1 else np.s_[int(b):int(b) + int(s)] for b, s in zip(begin, size))
  return np.asarray(x)[idx]


def tile(x, multiples):
  return np.tile(x, _axes(multiples))


def transpose(x, perm=None):
  return np.transpose(x, perm)


# ---- elementwise
def mul(a, b, name=None):
  return _f(a) * _f(b)


def div(a, b):
  return a / b


def maximum(a, b):
  return np.maximum(a, b)


def minimum(a, b):
  return np.minimum(a, b)


def equal(a, b):
  return np.equal(a, b)


def abs(x):  # noqa: A001
  return np.abs(x)


def exp(x):
  return np.exp(_f(x))


def log(x):
  return np.log(_f(x))


def sqrt(x):
  return np.sqrt(_f(x))


def round(x):  # noqa: A001  TF 0.12's Round functor is floor(x + 0.5) (half up), not round-half-even
  return np.floor(_f(x) + F32(0.5))


# ---- reductions: (x, reduction_indices=None, keep_dims=False)
def _reduce(fn, x, reduction_indices, keep_dims):
  return fn(np.asarray(x), axis=_axes(reduction_indices), keepdims=bool(keep_dims))


def reduce_sum(x, reduction_indices=None, keep_dims=False):
  return _reduce(np.sum, x, reduction_indices, keep_dims)


def reduce_prod(x, reduction_indices=None, keep_dims=False):
  return _reduce(np.prod, x, reduction_indices, keep_dims)


def reduce_max(x, reduction_indices=None, keep_dims=False):
  return _reduce(np.max, x, reduction_indices, keep_dims)


def reduce_min(x, reduction_indices=None, keep_dims=False):
  return _reduce(np.min, x, reduction_indices, keep_dims)


# ---- linear algebra
def batch_matmul(x, y, adj_x=False, adj_y=False):
  x = np.swapaxes(x, -1, -2) if adj_x else x
  y = np.swapaxes(y, -1, -2) if adj_y else y
  return np.matmul(_f(x), _f(y))


# ---- the custom op
class _HungarianModule(object):

  @staticmethod
  def hungarian(weights):
    from oracle import hungarian as H  # the C restatement of hungarian.cc, pinned by the reference's own KATs
    m, cx, cy = H.hungarian(np.asarray(weights, F32))
    return m, cx, cy


def load_op_library(path):
  return _HungarianModule()


# =====================================================================================================================
# Additions for /root/reference/nnlib.py (layer factories).  Eager evaluation: "building" and "running" the graph are
# the same thing here, with one exception - tf.cond must BUILD both branches (TensorFlow creates the variables of both,
# e.g. the EMA shadows) but RUN only the selected one; _EXEC tells stateful ops whether they are being run.
# =====================================================================================================================
import contextlib as _ctx

import torch as _torch
import torch.nn.functional as _F

_EXEC = [True]
_COLLECTIONS = {}


class Tensor(np.ndarray):
  """ndarray with the few Tensor methods the reference calls (set_shape)."""

  def set_shape(self, shape):
    assert tuple(self.shape) == tuple(int(s) for s in shape), (self.shape, shape)


def _t(a):
  return np.asarray(a, F32).view(Tensor)


VARIABLE_OVERRIDES = {}  # name -> initial value (e.g. 'global_step': the graph creates it as 0.0)


def Variable(initial_value, name=None, trainable=True):
  if name in VARIABLE_OVERRIDES:
    initial_value = VARIABLE_OVERRIDES[name]
  return np.array(initial_value, dtype=F32, copy=True).view(Tensor)


def truncated_normal_initializer(stddev=1.0, seed=0):
  rng = np.random.default_rng(seed)

  def init(shape):
    v = rng.standard_normal(tuple(int(s) for s in shape))
    bad = np.abs(v) > 2
    while bad.any():
      v[bad] = rng.standard_normal(int(bad.sum()))
      bad = np.abs(v) > 2
    return (v * stddev).astype(np.float32).astype(F32)  # float32-representable: fixtures store weights losslessly

  return init


def constant_initializer(value=0.0):
  return lambda shape: np.full(tuple(int(s) for s in shape), value, F32)


@_ctx.contextmanager
def variable_scope(name, *a, **k):
  yield


@_ctx.contextmanager
def control_dependencies(ops_):
  yield


def identity(x, name=None):
  return x


def cond(pred, fn1, fn2):
  """Build both branches, run the selected one (see the header of this section)."""
  sel, other = (fn1, fn2) if bool(pred) else (fn2, fn1)
  prev = _EXEC[0]
  _EXEC[0] = False
  try:
    other()
  finally:
    _EXEC[0] = prev
  return sel()


def add_to_collection(name, value):
  _COLLECTIONS.setdefault(name, []).append(value)


def get_collection(name):
  return list(_COLLECTIONS.get(name, []))


def matmul(a, b):
  return _t(np.matmul(_f(a), _f(b)))


def sigmoid(x):
  with np.errstate(over='ignore'):  # exp overflow -> inf -> sigmoid 0: the saturated value, as in TensorFlow
    return _t(1.0 / (1.0 + np.exp(-_f(x))))


def tanh(x):
  return _t(np.tanh(_f(x)))


def _same_pad(size, k, s):
  """TensorFlow 'SAME': output ceil(size / s); total padding split with the extra element at the END."""
  out = -(-size // s)
  total = max((out - 1) * s + k - size, 0)
  return total // 2, total - total // 2


class _NN(object):

  @staticmethod
  def relu(x):
    return _t(np.maximum(_f(x), 0))

  @staticmethod
  def softmax(x):  # over the last dimension
    x = _f(x)
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return _t(e / e.sum(axis=-1, keepdims=True))

  @staticmethod
  def l2_loss(x):
    return F32(np.sum(_f(x)**2) / 2)

  @staticmethod
  def conv2d(x, w, strides, padding):
    assert padding == 'SAME' and strides[0] == strides[3] == 1
    xt = _torch.from_numpy(np.ascontiguousarray(_f(x))).permute(0, 3, 1, 2)
    wt = _torch.from_numpy(np.ascontiguousarray(_f(w))).permute(3, 2, 0, 1)  # HWIO -> OIHW
    pt, pb = _same_pad(xt.shape[2], wt.shape[2], strides[1])
    pl, pr = _same_pad(xt.shape[3], wt.shape[3], strides[2])
    y = _F.conv2d(_F.pad(xt, (pl, pr, pt, pb)), wt, stride=(strides[1], strides[2]))
    return _t(y.permute(0, 2, 3, 1).numpy())

  @staticmethod
  def conv2d_transpose(x, w, output_shape, strides, padding='SAME'):
    """Gradient of conv2d(SAME) w.r.t. its input: the full transposed convolution cropped by the forward padding.
    w is [kh, kw, out_ch, in_ch]."""
    assert padding == 'SAME'
    xt = _torch.from_numpy(np.ascontiguousarray(_f(x))).permute(0, 3, 1, 2)
    wt = _torch.from_numpy(np.ascontiguousarray(_f(w))).permute(3, 2, 0, 1)  # [in_ch, out_ch, kh, kw]
    oh, ow = int(np.asarray(output_shape)[1]), int(np.asarray(output_shape)[2])
    full = _F.conv_transpose2d(xt, wt, stride=(strides[1], strides[2]))
    pt, _ = _same_pad(oh, wt.shape[2], strides[1])
    pl, _ = _same_pad(ow, wt.shape[3], strides[2])
    full = _F.pad(full, (0, max(0, pl + ow - full.shape[3]), 0, max(0, pt + oh - full.shape[2])))
    return _t(full[:, :, pt:pt + oh, pl:pl + ow].permute(0, 2, 3, 1).numpy())

  @staticmethod
  def max_pool(x, ksize, strides, padding):
    assert padding == 'SAME'
    xt = _torch.from_numpy(np.ascontiguousarray(_f(x))).permute(0, 3, 1, 2)
    pt, pb = _same_pad(xt.shape[2], ksize[1], strides[1])
    pl, pr = _same_pad(xt.shape[3], ksize[2], strides[2])
    y = _F.max_pool2d(_F.pad(xt, (pl, pr, pt, pb), value=float('-inf')), (ksize[1], ksize[2]), (strides[1], strides[2]))
    return _t(y.permute(0, 2, 3, 1).numpy())

  @staticmethod
  def moments(x, axes, name=None):
    x = _f(x)
    mean = x.mean(axis=tuple(axes), dtype=np.float64)
    var = ((x - mean.astype(F32))**2).mean(axis=tuple(axes), dtype=np.float64)  # biased, like tf.nn.moments
    return _t(mean), _t(var)

  @staticmethod
  def batch_normalization(x, mean, variance, offset, scale, variance_epsilon):
    inv = (1.0 / np.sqrt(_f(variance) + F32(variance_epsilon))).astype(F32) * _f(scale)  # rsqrt(var + eps) * gamma
    return _t(_f(x) * inv + (_f(offset) - _f(mean) * inv))

  @staticmethod
  def dropout(x, keep_prob):
    raise NotImplementedError('dropout is dead in every shipped config (dropout_keep=None)')


nn = _NN()


class _EMA(object):
  """tf.train.ExponentialMovingAverage over TENSORS: apply() creates a zero-initialised shadow per tensor and moves it,
  shadow -= (1 - decay) * (shadow - value); average() returns the shadow."""

  def __init__(self, decay):
    self.decay = decay
    self.shadow = {}

  def apply(self, var_list):
    for v in var_list:
      key = id(v)
      if key not in self.shadow:
        self.shadow[key] = (np.zeros_like(_f(v)).view(Tensor), v)  # keep v alive: ids must stay unique
      if _EXEC[0]:
        s = self.shadow[key][0]
        s -= (F32(1.0) - F32(self.decay)) * (s - _f(v))
    return None

  def average(self, v):
    return self.shadow[id(v)][0] if id(v) in self.shadow else None


class _Train(object):
  ExponentialMovingAverage = _EMA


train = _Train()


# =====================================================================================================================
# Additions for /root/reference/full_model.py + image_ops.py: get_model(opt) is executed EAGERLY - placeholders return
# the arrays queued in FEED (in creation order: x, y_gt, s_gt[, d_in, y_in], phase_train), random draws come from a
# seeded generator and are LOGGED so that the caller can hand the very same numbers to the oracle, the optimiser is a
# stub (the backward pass is not part of this comparison).
# =====================================================================================================================
FEED = []
RANDOM_LOG = []
_RNG = [np.random.default_rng(0)]


def reset(feed, seed=0):
  del FEED[:]
  FEED.extend(feed)
  del RANDOM_LOG[:]
  _RNG[0] = np.random.default_rng(seed)
  _COLLECTIONS.clear()


def placeholder(dtype, shape=None, name=None):
  name_, value = FEED.pop(0)
  assert name is None or name == name_, (name, name_)
  if dtype == 'bool':
    return bool(value)
  return _t(value)


RANDOM_OVERRIDES = []  # values handed out, in call order, before the generator is consulted


def random_uniform(shape, minval=0, maxval=None, dtype='float32', seed=None, name=None):
  shp = tuple(int(s) for s in np.ravel(np.asarray(shape)))
  if RANDOM_OVERRIDES:
    v = np.broadcast_to(np.asarray(RANDOM_OVERRIDES.pop(0)), shp).astype(np.int32 if np.dtype(dtype).kind == 'i' else F32)
    RANDOM_LOG.append({'shape': shp, 'min': minval, 'max': maxval, 'value': v.copy()})
    return v.view(Tensor) if v.dtype.kind == 'f' else v
  if np.dtype(dtype).kind == 'i':
    # the only integer draw is the crop offset of image_ops.random_transformation: always the centre (identity crop)
    v = np.full(shp, int(maxval) // 2, np.int32)
  else:
    hi = 1.0 if maxval is None else maxval
    lo_a, hi_a = np.asarray(minval, F32), np.asarray(hi, F32)
    # drawn in float32 so that the logged values are exactly representable in a float32 fixture
    v = (lo_a + (hi_a - lo_a) * _RNG[0].random(shp, dtype=np.float32).astype(F32)).astype(np.float32).astype(F32)
  RANDOM_LOG.append({'shape': shp, 'min': minval, 'max': maxval, 'value': v.copy()})
  return v.view(Tensor) if v.dtype.kind == 'f' else v


def pad(x, paddings):
  return _t(np.pad(_f(x), [tuple(int(v) for v in p) for p in paddings]))


def reverse(x, dims):  # TF 0.12: a boolean per dimension
  x = np.asarray(x)
  for ax, flag in enumerate(np.ravel(np.asarray(dims))):
    if bool(flag):
      x = np.flip(x, ax)
  return _t(x)


def cast(x, dtype):
  return np.asarray(x).astype(np.dtype(dtype))


def clip_by_value(x, lo, hi):
  return _t(np.clip(_f(x), lo, hi))


def stop_gradient(x):
  return x


def add_n(values, name=None):
  out = values[0]
  for v in values[1:]:
    out = out + v
  return out


_NN.softplus = staticmethod(lambda x: _t(np.log1p(np.exp(-np.abs(_f(x)))) + np.maximum(_f(x), 0)))


def _exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
  p = np.asarray(global_step, F32) / F32(decay_steps)
  if staircase:
    p = np.floor(p)
  return F32(learning_rate) * np.power(F32(decay_rate), p).astype(F32)


class _AdamStub(object):

  def __init__(self, learning_rate, epsilon=1e-8):
    self.learning_rate = learning_rate

  def compute_gradients(self, loss):
    return []

  def apply_gradients(self, gvs, global_step=None):
    return None

  def minimize(self, loss, global_step=None):  # fg_model.py
    return None


class _MomentumStub(_AdamStub):

  def __init__(self, learning_rate, momentum=0.9):
    self.learning_rate = learning_rate


_Train.exponential_decay = staticmethod(_exponential_decay)
_Train.AdamOptimizer = _AdamStub
_Train.MomentumOptimizer = _MomentumStub


def argmax(x, dimension):
  return np.argmax(np.asarray(x), axis=int(dimension))


def squeeze(x, squeeze_dims=None):
  return np.squeeze(np.asarray(x), axis=None if squeeze_dims is None else tuple(int(d) for d in squeeze_dims)).view(Tensor)


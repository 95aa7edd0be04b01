"""numpy stand-in for the TensorFlow-0.12 ops used by /root/reference/modellib.py (see README.md).  Eager, float32."""
import numpy as np

F32 = np.float32


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
def mul(a, b):
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

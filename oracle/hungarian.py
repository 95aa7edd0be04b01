"""ctypes front-end of the C oracle for the Hungarian op (test infrastructure).

``hungarian(W)`` follows /root/reference/hungarian.cc literally (see
oracle/hungarian_ref.c); ``hungarian_bitset(W)`` is the CPU model of the
duplicate-free search the CUDA kernel uses (oracle/hungarian_bitset.c).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ST_OUTER_CAP = 1
ST_BFS_CAP = 2
ST_FLOW_CAP = 4
ST_EQUALIZE_CAP = 8
ST_TOO_LARGE = 16


def build(force=False):
  """Compile oracle/_build/liboracle.so with gcc (idempotent)."""
  so = os.path.join(_HERE, '_build', 'liboracle.so')
  srcs = [os.path.join(_HERE, f) for f in ('hungarian_ref.c', 'hungarian_bitset.c')]
  stale = force or not os.path.exists(so) or any(
      os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
  if stale:
    subprocess.check_call(['make', '-s', '-C', _HERE, '_build/liboracle.so'])
  return so


def _lib():
  global _LIB
  if _LIB is None:
    lib = ctypes.CDLL(build())
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int)
    lp = ctypes.POINTER(ctypes.c_long)
    lib.ra_oracle_hungarian_f32.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp, ip, lp,
                                            ctypes.c_int]
    lib.ra_oracle_hungarian_f32.restype = ctypes.c_int
    lib.ra_oracle_hungarian_bitset_f32.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp, fp, ip]
    lib.ra_oracle_hungarian_bitset_f32.restype = ctypes.c_int
    _LIB = lib
  return _LIB


def _prep(W):
  W = np.ascontiguousarray(W, dtype=np.float32)
  if W.ndim == 2:
    Wb = W[None]
  elif W.ndim == 3:
    Wb = W
  else:
    raise ValueError('Must have dimension 3 or 2.')  # hungarian.cc:61-64
  B, nx, ny = Wb.shape
  M = np.zeros((B, nx, ny), np.float32)
  cx = np.zeros((B, nx, 1), np.float32)
  cy = np.zeros((B, 1, ny), np.float32)
  st = np.zeros((B,), np.int32)
  return W, Wb, B, nx, ny, M, cx, cy, st


def _ptr(a, t=ctypes.c_float):
  return a.ctypes.data_as(ctypes.POINTER(t))


def hungarian(W, stop_at_fatal=False, return_info=False):
  """Literal restatement. Returns (matching, cover_x, cover_y) shaped like the op's outputs
  (hungarian.cc:52-76): same shape as W, [..., nx, 1], [..., 1, ny]."""
  W, Wb, B, nx, ny, M, cx, cy, st = _prep(W)
  pops = np.zeros((B,), np.int64)
  _lib().ra_oracle_hungarian_f32(_ptr(Wb), B, nx, ny, _ptr(M), _ptr(cx), _ptr(cy), _ptr(st, ctypes.c_int),
                                 _ptr(pops, ctypes.c_long), int(stop_at_fatal))
  if W.ndim == 2:
    M, cx, cy = M[0], cx[0], cy[0]
  if return_info:
    return M, cx, cy, {'status': st, 'max_bfs_pops': pops}
  return M, cx, cy


def hungarian_bitset(W, return_info=False):
  W, Wb, B, nx, ny, M, cx, cy, st = _prep(W)
  _lib().ra_oracle_hungarian_bitset_f32(_ptr(Wb), B, nx, ny, _ptr(M), _ptr(cx), _ptr(cy), _ptr(st, ctypes.c_int))
  if W.ndim == 2:
    M, cx, cy = M[0], cx[0], cy[0]
  if return_info:
    return M, cx, cy, {'status': st}
  return M, cx, cy


_REF = None
REF_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libhungarian_ref.so')


def reference_available():
  """True when oracle/_ref/libhungarian_ref.so (the reference's own hungarian.cc, see oracle/Makefile) is present."""
  return os.path.exists(REF_PATH)


def hungarian_reference(W):
  """THE REFERENCE'S OWN HungarianOp (hungarian.cc compiled unmodified against oracle/ref_shim).  Returns
  (matching, cover_x, cover_y, fatal) with the op's output shapes; fatal = text of the LOG(FATAL) the reference would
  have aborted with (BFS / max-flow caps, hungarian.cc:124-127,186-188), else None."""
  global _REF
  if _REF is None:
    lib = ctypes.CDLL(REF_PATH)
    lib.ref_hungarian_f32.restype = ctypes.c_int
    _REF = lib
  W = np.ascontiguousarray(W, dtype=np.float32)
  if W.ndim == 3:
    B, nx, ny = W.shape
  elif W.ndim == 2:
    B, (nx, ny) = 0, W.shape
  else:
    raise ValueError('Must have dimension 3 or 2.')
  nb = max(B, 1)
  M = np.zeros((nb, nx, ny), np.float32)
  cx = np.zeros((nb, nx, 1), np.float32)
  cy = np.zeros((nb, 1, ny), np.float32)
  msg = ctypes.create_string_buffer(512)
  vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
  rc = _REF.ref_hungarian_f32(vp(W), B, nx, ny, vp(M), vp(cx), vp(cy), msg, 512)
  if W.ndim == 2:
    M, cx, cy = M[0], cx[0], cy[0]
  return M, cx, cy, (msg.value.decode('utf-8', 'replace') if rc != 0 else None)

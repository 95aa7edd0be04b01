"""ORACLE (test infrastructure only — never on the product path).

numpy restatement of the reference's ``utils/postprocess.py`` functions that
``full_model_eval.py:112-125`` chains after the decode path, function by function.  PINNED: checked against
outputs of the reference module itself (imported in the build container by
``tests/golden/make_postprocess_golden.py``; vectors in ``tests/golden/postprocess_golden.npz``).
"""
import numpy as np


def apply_confidence(y_out, s_out):
  """utils/postprocess.py:17-31: weight by the confidence score; s_out_hard = s_out > 0.5 (float64)."""
  s_mask = np.reshape(s_out, [-1, s_out.shape[1], 1, 1])
  return y_out * s_mask, (s_out > 0.5).astype('float')


def apply_one_label(y_out):
  """utils/postprocess.py:34-55: keep, per pixel, only the arg-max instance (first maximum). float64 out."""
  res = []
  for _y in y_out:
    idx = np.argmax(_y, axis=0)
    _y2 = np.zeros(_y.shape)
    for jj in range(_y.shape[0]):
      _y2[jj] = (idx == jj).astype('float32') * _y[jj]
    res.append(_y2)
  return res


def apply_threshold(y_out, thresh):
  """utils/postprocess.py:6-14."""
  return [(_y > thresh).astype('float32') for _y in y_out]


def mask_foreground(y_out, fg):
  """utils/postprocess.py:146-155."""
  return [_y * _fg for _y, _fg in zip(y_out, fg)]


def remove_tiny(y_out, conf, threshold=200):
  """utils/postprocess.py:106-143 (conf is updated in place, like the reference)."""
  if threshold == 0:
    return y_out, conf
  res = []
  for ii, _y in enumerate(y_out):
    size = _y.sum(axis=1, keepdims=True).sum(axis=2, keepdims=True)
    not_tiny = (size > threshold).astype('float32')
    conf[ii] = conf[ii] * np.reshape(not_tiny, [-1])
    res.append(_y * not_tiny)
  return res, conf


def eval_chain(y_out, s_out, thresh, fg=None, remove_tiny_threshold=0):
  """full_model_eval.py:112-125 without the cv2 steps (upsample, morph): returns the dense thresholded masks
  [B,T,H,W] float32, the confidences [B,T] and the per-instance sizes before tiny removal [B,T]."""
  y, s_hard = apply_confidence(y_out, s_out)
  y = apply_one_label(list(y))
  y = apply_threshold(y, thresh)
  if fg is not None:
    y = mask_foreground(y, list(fg))
  area = np.stack([_y.sum(axis=1).sum(axis=1) for _y in y]).astype(np.float32)
  if fg is not None or remove_tiny_threshold:
    y, s_hard = remove_tiny(y, s_hard, threshold=remove_tiny_threshold)
  return np.stack(y).astype(np.float32), np.asarray(s_hard, np.float32), area


def label_map(y_dense):
  """Dense one-label masks [B,T,H,W] -> int32 label map [B,H,W] (0 background, t+1 instance t)."""
  B, T, H, W = y_dense.shape
  on = y_dense != 0
  assert (on.sum(axis=1) <= 1).all()
  return (on * (np.arange(T).reshape(1, T, 1, 1) + 1)).sum(axis=1).astype(np.int32)

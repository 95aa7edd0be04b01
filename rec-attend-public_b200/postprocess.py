"""Host mirror of the reference's ``utils/postprocess.py`` as chained by ``full_model_eval.py:112-125``
(SURVEY.md §8f rank 2): confidence weighting, one label per pixel, threshold, foreground mask, tiny-region
removal — one C-ABI call (``ra_postprocess_f32``, csrc/postprocess.cu) on device-resident ``y_out`` / ``s_out``.

The reference functions each return dense ``[T,H,W]`` float arrays per example; the natural output here is the
int32 label map (0 = background, t+1 = instance t) — the thing SURVEY §8(d) grades "bit-exact labels" on — and the
dense form is available on request (``want_dense``).  The cv2 steps of the reference chain (``upsample`` with a
bilateral filter, ``morph``) are not part of this path.
"""
import torch

from . import _lib, ops


def postprocess(y_out, s_out, thresh=0.3, fg=None, remove_tiny=0, want_dense=False):
  """apply_confidence -> apply_one_label -> apply_threshold [-> mask_foreground -> remove_tiny].

  y_out [B,T,H,W], s_out [B,T] (CUDA fp32); fg [B,H,W] or None (full_model_eval.py:121-124 runs mask_foreground
  and remove_tiny only when a foreground map exists; pass ``remove_tiny`` > 0 without fg to remove tiny regions
  anyway).  thresh is the Python float of ``threshold_list`` (default [0.3], full_model_eval.py:193-194).
  Returns dict: label int32 [B,H,W], conf [B,T] (= s_out_hard * is_not_tiny), area [B,T]
  (instance sizes before removal), y_out_thresh [B,T,H,W] when want_dense."""
  ops._chk(y_out, s_out, fg)
  B, T, H, W = y_out.shape
  if tuple(s_out.shape) != (B, T):
    raise _lib.RecAttendError('s_out must be [B,T]')
  if fg is not None and tuple(fg.shape) != (B, H, W):
    raise _lib.RecAttendError('fg must be [B,H,W]')
  dev = y_out.device
  n_ws = _lib.lib().ra_postprocess_workspace(B, T)
  ws = torch.empty((max(1, n_ws // 8),), device=dev, dtype=torch.float64)
  label = torch.empty((B, H, W), device=dev, dtype=torch.int32)
  conf = torch.empty((B, T), device=dev, dtype=torch.float32)
  area = torch.empty((B, T), device=dev, dtype=torch.float32)
  dense = torch.empty((B, T, H, W), device=dev, dtype=torch.float32) if want_dense else None
  _lib.call('ra_postprocess_f32', ops._p(y_out), ops._p(s_out), ops._p(fg), B, T, H, W, float(thresh),
            float(remove_tiny), ops._p(ws), ops._p(label), ops._p(dense), ops._p(conf), ops._p(area), ops._stream())
  out = {'label': label, 'conf': conf, 'area': area}
  if want_dense:
    out['y_out_thresh'] = dense
  return out

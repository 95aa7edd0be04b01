"""Generates tests/golden/postprocess_golden.npz by running the REFERENCE's own utils/postprocess.py
(imported from /root/reference in the build container; its only Python-2-ism is `xrange`) on small seeded inputs.
The chain is full_model_eval.py:112-125 without the cv2 steps.  Run:  python tests/golden/make_postprocess_golden.py
"""
import builtins
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
  builtins.xrange = range
  spec = importlib.util.spec_from_file_location('ref_postprocess', '/root/reference/utils/postprocess.py')
  m = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(m)
  return m


def make_case(rng, B, T, H, W, with_fg, binary_fg=True):
  # blobby soft masks with overlaps, exact ties (duplicated channels) and confidences around 0.5
  yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
  y = np.zeros((B, T, H, W), np.float32)
  for b in range(B):
    for t in range(T):
      cy, cx = rng.uniform(0, H), rng.uniform(0, W)
      r = rng.uniform(2, max(3, min(H, W) / 2.5))
      y[b, t] = 1.0 / (1.0 + np.exp(((yy - cy)**2 + (xx - cx)**2 - r * r) / (2.0 * r)))
  y += rng.uniform(0, 0.02, y.shape).astype(np.float32)
  if T >= 3:
    y[:, 2] = y[:, 1]  # exact tie between two instances: np.argmax must pick the first
  s = rng.uniform(0.2, 1.0, (B, T)).astype(np.float32)
  s[0, 0] = 0.5  # s_out > 0.5 is strict
  if T >= 3:
    s[:, 2] = s[:, 1]
  fg = None
  if with_fg:
    fg = (rng.uniform(0, 1, (B, H, W)) > 0.25).astype(np.float32)
    if not binary_fg:
      fg = fg * (np.round(rng.uniform(0, 255, (B, H, W))) / 255.0).astype(np.float32)
  return y.astype(np.float32), s, fg


def run_reference(pp, y_out, s_out, thresh, fg, tiny):
  y, s_hard = pp.apply_confidence(y_out, s_out)
  y = pp.apply_one_label(list(y))
  y = pp.apply_threshold(y, thresh)
  if fg is not None:
    y = pp.mask_foreground(y, list(fg))
  area = np.stack([_y.sum(axis=1).sum(axis=1) for _y in y]).astype(np.float32)
  if fg is not None or tiny:
    y, s_hard = pp.remove_tiny(y, s_hard, threshold=tiny)
  return np.stack(y).astype(np.float32), np.asarray(s_hard, np.float32), area


def main():
  pp = load_reference()
  rng = np.random.default_rng(20260101)
  out = {}
  cases = [  # B, T, H, W, with_fg, binary_fg, thresh, remove_tiny
      (2, 5, 24, 32, False, True, 0.3, 0),
      (2, 5, 24, 32, True, True, 0.3, 40),
      (1, 8, 16, 20, False, True, 0.5, 25),
      (3, 4, 12, 16, True, False, 0.1, 10),
      (1, 1, 8, 8, False, True, 0.3, 0),
  ]
  for i, (B, T, H, W, with_fg, binary_fg, thresh, tiny) in enumerate(cases):
    y, s, fg = make_case(rng, B, T, H, W, with_fg, binary_fg)
    dense, conf, area = run_reference(pp, y, s, thresh, fg, tiny)
    p = 'c%d_' % i
    out[p + 'y_out'], out[p + 's_out'] = y, s
    if fg is not None:
      out[p + 'fg'] = fg
    out[p + 'thresh'], out[p + 'tiny'] = np.float64(thresh), np.int64(tiny)
    out[p + 'dense'], out[p + 'conf'], out[p + 'area'] = dense, conf, area
  out['n_cases'] = np.int64(len(cases))
  np.savez_compressed(os.path.join(HERE, 'postprocess_golden.npz'), **out)
  print('wrote', os.path.join(HERE, 'postprocess_golden.npz'))


if __name__ == '__main__':
  main()

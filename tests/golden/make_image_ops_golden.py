"""Generates tests/golden/image_ops_golden.npz by EXECUTING THE REFERENCE'S OWN image_ops.random_transformation
(/root/reference/image_ops.py, unmodified) over the numpy stand-in of tests/golden/tf012_shim with chosen random draws
(crop offset, flip / transpose coin tosses), in training and in eval mode.  Pins oracle.model.random_transformation
(and through it ra_random_transformation_f32).   Run:  python tests/golden/make_image_ops_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
# name, H, W, padding, offset, (hflip, vflip, transpose), with orientation input
CASES = [('crop', 12, 16, 4, (1, 7), (False, False, False), False), ('flip_h', 12, 16, 3, (3, 3), (True, False, False), False),
         ('flip_v_t', 12, 12, 2, (0, 4), (False, True, True), False), ('all', 10, 10, 5, (9, 2), (True, True, True), False),
         ('orientation', 12, 16, 4, (6, 0), (False, False, False), True), ('eval', 12, 16, 4, (1, 7), (True, True, False), False)]


def main():
  sys.path.insert(0, ROOT)
  sys.path.insert(0, REF)
  sys.path.insert(0, os.path.join(HERE, 'tf012_shim'))
  import tensorflow as tf
  import image_ops as IO  # the reference source file itself
  assert os.path.dirname(os.path.abspath(IO.__file__)) == REF, IO.__file__
  rng = np.random.default_rng(99)
  out = {}
  B, T = 2, 3
  for name, H, W, pad, off, (hf, vf, tr), with_d in CASES:
    x = rng.random((B, H, W, 3)).astype(np.float32)
    y = (rng.random((B, T, H, W)) > 0.5).astype(np.float32)
    d = rng.random((B, H, W, 8)).astype(np.float32) if with_d else None
    c = rng.random((B, H, W, 2)).astype(np.float32) if with_d else None
    tf.reset([], seed=0)
    # draws in call order: offset [2] int; then (only without d) rand_h, rand_v, rand_t in [1 - flag, 1]: < 0.5 = do it
    tf.RANDOM_OVERRIDES[:] = [np.array(off)] + ([] if with_d else [0.25 if hf else 0.75, 0.25 if vf else 0.75,
                                                                  0.25 if tr else 0.75])
    r = IO.random_transformation(x, pad, name != 'eval', rnd_vflip=not with_d, rnd_hflip=not with_d,
                                 rnd_transpose=not with_d, rnd_colour=False, y=y, d=d, c=c)
    assert not tf.RANDOM_OVERRIDES
    out[name + '/in_x'], out[name + '/in_y'] = x, y
    if with_d:
      out[name + '/in_d'], out[name + '/in_c'] = d, c
    out[name + '/params'] = np.array([H, W, pad, off[0], off[1], int(hf), int(vf), int(tr), int(with_d), int(name != 'eval')])
    for k, v in r.items():
      out['%s/out_%s' % (name, k)] = np.asarray(v, np.float32)
  path = os.path.join(HERE, 'image_ops_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()

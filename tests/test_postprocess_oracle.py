"""CPU: the numpy restatement of utils/postprocess.py (oracle/postprocess.py) against golden vectors produced by
the reference module itself (tests/golden/make_postprocess_golden.py)."""
import os

import numpy as np

from oracle import postprocess as OP

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'postprocess_golden.npz')


def golden_cases():
  g = np.load(GOLD)
  for i in range(int(g['n_cases'])):
    p = 'c%d_' % i
    yield {
        'y_out': g[p + 'y_out'], 's_out': g[p + 's_out'], 'fg': g[p + 'fg'] if p + 'fg' in g.files else None,
        'thresh': float(g[p + 'thresh']), 'tiny': int(g[p + 'tiny']), 'dense': g[p + 'dense'], 'conf': g[p + 'conf'],
        'area': g[p + 'area']
    }


def test_oracle_matches_reference_golden():
  n = 0
  for c in golden_cases():
    dense, conf, area = OP.eval_chain(c['y_out'], c['s_out'], c['thresh'], c['fg'], c['tiny'])
    assert (dense == c['dense']).all()
    assert (conf == c['conf']).all()
    assert (area == c['area']).all()
    n += 1
  assert n == 5


def test_one_label_first_max_and_strict_threshold():
  y = np.zeros((1, 3, 1, 4), np.float32)
  y[0, :, 0, 0] = [0.4, 0.4, 0.2]  # tie -> instance 0
  y[0, :, 0, 1] = [0.1, 0.6, 0.6]  # tie -> instance 1
  y[0, :, 0, 2] = [0.3, 0.2, 0.1]  # max == thresh (as float32 -> double) is not above it
  y[0, :, 0, 3] = [0.0, 0.0, 0.0]
  s = np.ones((1, 3), np.float32)
  dense, conf, area = OP.eval_chain(y, s, float(np.float32(0.3)))
  assert (OP.label_map(dense)[0, 0] == [1, 2, 0, 0]).all()
  assert (area[0] == [1, 1, 0]).all() and (conf == 1).all()


def test_remove_tiny_updates_confidence():
  rng = np.random.default_rng(0)
  y = rng.random((2, 4, 8, 8)).astype(np.float32)
  s = np.array([[0.9, 0.4, 0.8, 0.7], [0.6, 0.6, 0.2, 0.9]], np.float32)
  dense, conf, area = OP.eval_chain(y, s, 0.0, None, 12)
  keep = area > 12
  assert (conf == (s > 0.5) * keep).all()
  assert (dense.sum(axis=(2, 3)) == area * keep).all()

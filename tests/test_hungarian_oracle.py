"""CPU tests of the Hungarian oracle against the reference's own vectors
(/root/reference/hungarian_tf_tests.py -> tests/golden/hungarian_kat.json)."""
import json
import os

import numpy as np
import pytest

from oracle import hungarian as H

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'hungarian_kat.json')


def _cases():
  return json.load(open(GOLD))['cases']


@pytest.mark.parametrize('case', _cases(), ids=lambda c: c['name'])
def test_known_answers_and_termination(case):
  W = np.frombuffer(bytes.fromhex(case['W_f32_hex']), np.float32).reshape(case['shape'])
  M, cx, cy, info = H.hungarian(W, stop_at_fatal=True, return_info=True)
  assert (info['status'] == 0).all(), 'reference iteration caps must not trigger on its own test inputs'
  if case['kind'] == 'known_answer':
    assert (M == np.array(case['M'], np.float32)).all()
    if 'cover_x' in case:  # hungarian_tf_tests.py:9-67 assert both covers exactly
      assert (cx.reshape(-1) == np.array(case['cover_x'], np.float32).reshape(-1)).all()
      assert (cy.reshape(-1) == np.array(case['cover_y'], np.float32).reshape(-1)).all()
  else:
    # termination-only regression inputs: a perfect matching must come back
    assert M.sum() == min(W.shape[-2:]) and (M.sum(-1) <= 1).all() and (M.sum(-2) <= 1).all()
  # the duplicate-free search used by the CUDA kernel gives the same bits
  M2, cx2, cy2 = H.hungarian_bitset(W)
  assert (M == M2).all() and (cx == cx2).all() and (cy == cy2).all()


def test_output_shapes_follow_the_op():
  W = np.random.default_rng(0).random((4, 5, 7)).astype(np.float32)
  M, cx, cy = H.hungarian(W)
  assert M.shape == (4, 5, 7) and cx.shape == (4, 5, 1) and cy.shape == (4, 1, 7)  # hungarian.cc:52-76
  M, cx, cy = H.hungarian(W[0])
  assert M.shape == (5, 7) and cx.shape == (5, 1) and cy.shape == (1, 7)
  with pytest.raises(ValueError):
    H.hungarian(W[0, 0])


def _random_weights(rng, kind, B, nx, ny):
  if kind == 0:
    W = rng.random((B, nx, ny))
  elif kind == 1:
    W = rng.integers(0, 6, (B, nx, ny)).astype(float)
  elif kind == 2:  # training-like: K real objects, the rest on the 1e-5 floor (modellib.py:403-406)
    W = np.zeros((B, nx, ny))
    for b in range(B):
      k = rng.integers(0, min(nx, ny) + 1)
      W[b, :k, :k] = rng.random((k, k))**3
    W = np.floor(W * 1e6 + 0.5) / 1e6 + 1e-5
  elif kind == 3:  # (nearly) identical rows: hungarian_tf_tests.py:93-163,239-276
    W = np.repeat(rng.random((B, 1, ny)), nx, 1) + (rng.random((B, nx, ny)) < 0.1) * 1e-6
    W = np.floor(W * 1e6 + 0.5) / 1e6 + 1e-5
  else:
    W = rng.choice([1e-5, 0.5, 0.25, 0.75], (B, nx, ny))
  return W.astype(np.float32)


def test_bitset_search_equals_literal_search():
  """oracle/hungarian_bitset.c (the CUDA kernel's algorithm) == oracle/hungarian_ref.c (literal),
  including rectangular, tied and degenerate inputs and the cases where the reference's
  1000-pop BFS cap would have aborted it."""
  rng = np.random.default_rng(7)
  n_abort = 0
  for it in range(120):
    kind = it % 5
    nx, ny = int(rng.integers(1, 25)), int(rng.integers(1, 25))
    if it % 3 == 0:
      ny = nx
    if it % 40 == 0:
      nx = ny = 32
    if it == 119:
      nx = ny = 64
    B = 32 if nx < 40 else 2
    W = _random_weights(rng, kind, B, nx, ny)
    M, cx, cy, i1 = H.hungarian(W, return_info=True)
    M2, cx2, cy2, i2 = H.hungarian_bitset(W, return_info=True)
    assert (M == M2).all() and (cx == cx2).all() and (cy == cy2).all(), (kind, nx, ny)
    assert ((i1['status'] & 1) == (i2['status'] & 1)).all()
    n_abort += int(((i1['status'] & H.ST_BFS_CAP) > 0).sum())
  assert n_abort >= 0


def test_matching_is_optimal_on_generic_inputs():
  from scipy.optimize import linear_sum_assignment
  rng = np.random.default_rng(3)
  for _ in range(20):
    n = int(rng.integers(2, 13))
    W = rng.random((n, n)).astype(np.float32)
    M, _, _ = H.hungarian(W)
    r, c = linear_sum_assignment(-W.astype(np.float64))
    assert abs(float((M * W).sum()) - float(W[r, c].sum())) < 1e-4


def test_float_thresholds_equal_the_double_literal():
  """hungarian.cc compares fp32 values against the DOUBLE literal 1e-6 (:18,318,428).  The CUDA
  kernel uses `<= 1e-6f`; both select the same set of floats."""
  f = np.float32(1e-6)
  assert float(f) < 1e-6 < float(np.nextafter(f, np.float32(1)))
  for v in (f, np.nextafter(f, np.float32(0)), np.nextafter(f, np.float32(1))):
    assert (float(v) <= 1e-6) == (v <= f)
    assert (float(v) < 1e-6) == (v <= f)


def test_bitset_rejects_oversize():
  W = np.zeros((1, 65, 3), np.float32)
  _, _, _, info = H.hungarian_bitset(W, return_info=True)
  assert (info['status'] == H.ST_TOO_LARGE).all()


needs_ref = pytest.mark.skipif(not H.reference_available(),
                               reason='oracle/_ref/libhungarian_ref.so is built where /root/reference exists (make -C oracle)')


@needs_ref
@pytest.mark.parametrize('case', _cases(), ids=lambda c: c['name'])
def test_compiled_reference_reproduces_its_own_test_vectors(case):
  """oracle/_ref = the reference's own hungarian.cc compiled unmodified against stand-in TF / Eigen headers
  (oracle/ref_shim).  First make sure that build behaves like the op: it must pass the reference's own tests."""
  W = np.frombuffer(bytes.fromhex(case['W_f32_hex']), np.float32).reshape(case['shape'])
  M, cx, cy, fatal = H.hungarian_reference(W)
  assert fatal is None
  if case['kind'] == 'known_answer':
    assert (M == np.array(case['M'], np.float32)).all()
    if 'cover_x' in case:
      assert (cx.reshape(-1) == np.array(case['cover_x'], np.float32).reshape(-1)).all()
      assert (cy.reshape(-1) == np.array(case['cover_y'], np.float32).reshape(-1)).all()
  else:
    assert M.sum() == min(W.shape[-2:])


@needs_ref
def test_restatement_equals_the_compiled_reference_bit_for_bit():
  """The literal C restatement (and the duplicate-free search the CUDA kernel uses) against the REAL reference code
  on 20k random / degenerate / rectangular problems: matching and both covers identical to the bit; where the
  reference would abort on its BFS cap (LOG(FATAL), hungarian.cc:124-127) the restatement reports the same."""
  rng = np.random.default_rng(2026)
  n = fatal_seen = outer_cap_seen = 0
  for trial in range(2500):
    nx, ny = int(rng.integers(1, 25)), int(rng.integers(1, 25))
    if trial % 50 == 0:
      nx = ny = int(rng.integers(26, 34))  # large enough for the reference's 1000-pop BFS cap to matter
    W = _random_weights(rng, trial % 5, 8, nx, ny)
    for b in range(W.shape[0]):
      Mr, cxr, cyr, fatal = H.hungarian_reference(W[b])
      Mo, cxo, cyo, info = H.hungarian(W[b], stop_at_fatal=True, return_info=True)
      n += 1
      if fatal is not None:
        fatal_seen += 1
        assert int(info['status'][0]) & ~1, ('the restatement must flag what the reference aborts on', fatal)
        continue
      # bit 1 = the outer loop ran its 1000 rounds: the reference logs an ERROR and returns the unfinished matching
      # (hungarian.cc:362-375) - not fatal, and the unfinished matching must be the same one
      assert int(info['status'][0]) & ~1 == 0
      outer_cap_seen += int(info['status'][0]) & 1
      assert np.array_equal(Mr, Mo) and np.array_equal(cxr, cxo) and np.array_equal(cyr, cyo), (trial, b, nx, ny)
      Mb, cxb, cyb = H.hungarian_bitset(W[b])
      assert np.array_equal(Mr, Mb) and np.array_equal(cxr, cxb) and np.array_equal(cyr, cyb), (trial, b, nx, ny)
  assert n == 20000
  # batched (rank-3) entry of the op
  W = _random_weights(rng, 2, 6, 9, 12)
  Mr, cxr, cyr, fatal = H.hungarian_reference(W)
  Mo, cxo, cyo = H.hungarian(W)
  assert fatal is None and np.array_equal(Mr, Mo) and np.array_equal(cxr, cxo) and np.array_equal(cyr, cyo)

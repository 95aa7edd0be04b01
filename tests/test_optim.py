"""Optimiser block (full_model.py:1039-1057) + gradient all-reduce.
CPU: flat-bucket layout, learning-rate schedule, the oracle's Adam algebra, a 2-rank gloo all-reduce of the bucket.
GPU: ra_adam_step_f32 against the oracle over several steps."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import rel_err
from oracle import optim as OO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _weights():
  import rec_attend_b200 as ra
  opt = ra.config.full_model_opt('cvppp', 32, 32, 2)
  return opt, ra.synthetic.make_weights(opt)


def test_flat_layout_and_weight_decay_mask():
  from rec_attend_b200 import optim
  opt, w = _weights()
  fp = optim.FlatParams(w)
  assert not any(k.endswith(('_ema_mean', '_ema_var')) for k in fp.keys)
  assert 'ctrl_cnn_0_1_gamma' in fp.layout and 'ctrl_lstm_w_xi' in fp.layout and 'ctrl_cnn_b_0' in fp.layout
  flat = fp.flatten(w)
  back = fp.unflatten(flat)
  assert all((back[k] == np.asarray(w[k], np.float32)).all() for k in fp.keys)
  assert flat.size == sum(np.asarray(w[k]).size for k in fp.keys)
  wd = fp.weight_decay_vector(5e-5)
  on = fp.unflatten(wd)
  assert (on['ctrl_cnn_w_0'] == np.float32(5e-5)).all() and (on['ctrl_lstm_w_hf'] == np.float32(5e-5)).all()
  assert (on['ctrl_cnn_b_0'] == 0).all() and (on['ctrl_cnn_0_0_beta'] == 0).all() and (on['ctrl_lstm_b_f'] == 0).all()
  g = fp.flatten({'ctrl_cnn_b_0': np.ones_like(w['ctrl_cnn_b_0'])})  # other grads are None -> zeros
  assert g.sum() == w['ctrl_cnn_b_0'].size


def test_learn_rate_staircase():
  from rec_attend_b200 import optim
  opt = {'base_learn_rate': 0.001, 'learn_rate_decay': 0.85, 'steps_per_learn_rate_decay': 5000}
  assert optim.learn_rate(opt, 0) == pytest.approx(0.001)
  assert optim.learn_rate(opt, 4999) == pytest.approx(0.001)
  assert optim.learn_rate(opt, 5000) == pytest.approx(0.00085)
  assert optim.learn_rate(opt, 12345) == pytest.approx(0.001 * 0.85**2)
  assert float(OO.learn_rate(0.001, 0.85, 5000, 12345)) == pytest.approx(optim.learn_rate(opt, 12345), rel=1e-6)


def test_oracle_adam_first_step_is_sign_descent():
  # t = 1, m = v = 0: update = lr * g / (|g| + eps*...) ~ lr * sign(g); clipping caps |g| at 1 first
  var = {'w': np.array([1.0, -2.0, 0.5], np.float32)}
  grad = {'w': np.array([10.0, -0.01, 0.0], np.float32)}
  z = {'w': np.zeros(3, np.float32)}
  out, m, v = OO.adam_step(var, grad, z, z, {}, 0.001, 1)
  assert out['w'][0] == pytest.approx(1.0 - 0.001, abs=1e-6) and out['w'][1] == pytest.approx(-2.0 + 0.001, abs=1e-6)
  assert out['w'][2] == 0.5 and m['w'][0] == pytest.approx(0.1, rel=1e-5) and v['w'][0] == pytest.approx(0.001, rel=1e-4)


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, q):
  os.environ.update({'RANK': str(rank), 'LOCAL_RANK': str(rank), 'WORLD_SIZE': str(world),
                     'MASTER_ADDR': '127.0.0.1', 'MASTER_PORT': str(port)})
  sys.path.insert(0, ROOT)
  from rec_attend_b200 import dist_util, optim
  dist_util.init('gloo')
  g = torch.full((1000,), float(rank + 1))
  g[rank] = 100.0
  world_seen = optim.all_reduce_sum_(g)
  q.put((rank, world_seen, float(g[0]), float(g[1]), float(g[5])))
  dist_util.finalize()


def test_two_rank_gradient_bucket_all_reduce():
  world, port = 2, _free_port()
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in range(world))
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for r in res:
    assert r[1] == 2 and r[2] == 102.0 and r[3] == 101.0 and r[4] == 3.0  # SUM over ranks; mean = x 1/world later


@pytest.mark.gpu
def test_adam_step_kernel_vs_oracle(cuda):
  from rec_attend_b200 import optim
  opt, w = _weights()
  opt = dict(opt, steps_per_learn_rate_decay=2)  # decay inside the test
  o = optim.AdamOptimizer(opt, w)
  fp = o.flat
  rng = np.random.default_rng(0)
  var = {k: np.asarray(w[k], np.float32) for k in fp.keys}
  m = {k: np.zeros_like(var[k]) for k in fp.keys}
  v = {k: np.zeros_like(var[k]) for k in fp.keys}
  wd = {k: (np.float32(opt['weight_decay']) if optim.has_weight_decay(k) else 0.0) for k in fp.keys}
  for step in range(5):
    # heavy-tailed gradients so that the clip is exercised; one variable without a gradient
    grad = {k: (rng.standard_normal(var[k].shape) * rng.choice([1e-3, 0.3, 5.0])).astype(np.float32) for k in fp.keys}
    grad['ctrl_cnn_b_1'] = None
    lr = o.step(torch.from_numpy(fp.flatten(grad)).cuda())
    assert lr == pytest.approx(float(OO.learn_rate(opt['base_learn_rate'], opt['learn_rate_decay'], 2, step)), rel=1e-6)
    var, m, v = OO.adam_step(var, grad, m, v, wd, lr, step + 1)
    got = fp.unflatten(o.params.cpu().numpy())
    worst = max(rel_err(got[k], var[k]) for k in fp.keys)
    assert worst < 1e-6, (step, worst)
  assert rel_err(fp.unflatten(o.m.cpu().numpy())['ctrl_lstm_w_xi'], m['ctrl_lstm_w_xi']) < 1e-6
  new_w = o.export_weights(w)
  assert set(new_w) == set(w) and (new_w['ctrl_cnn_0_0_ema_var'] == w['ctrl_cnn_0_0_ema_var']).all()


@pytest.mark.gpu
def test_adam_step_errors(cuda):
  from rec_attend_b200 import _lib, optim
  opt, w = _weights()
  o = optim.AdamOptimizer(opt, w)
  with pytest.raises(_lib.RecAttendError):
    o.step(torch.zeros(3, device='cuda'))


@pytest.mark.gpu
def test_adam_frozen_set_and_resume(cuda):
  """Pretrained hand-off (checkpoint.apply_pretrained): frozen tensors stay out of the bucket and keep their values;
  a run resumed from (weights, m, v, step) continues bit-identically to the uninterrupted one."""
  from rec_attend_b200 import checkpoint as ck, optim
  opt, w = _weights()
  _, frozen = ck.apply_pretrained(dict(opt, freeze_ctrl_cnn=True, freeze_ctrl_rnn=True), w)
  a = optim.AdamOptimizer(opt, w, frozen=frozen)
  assert a.params.numel() == sum(np.asarray(w[k]).size for k in optim.trainable_keys(w, frozen))
  rng = np.random.default_rng(1)
  grads = [torch.from_numpy(rng.standard_normal(a.params.numel()).astype(np.float32)).cuda() for _ in range(4)]
  for g in grads[:2]:
    a.step(g.clone())
  w_mid = a.export_weights(w)
  for k in frozen:
    assert np.array_equal(w_mid[k], np.asarray(w[k], np.float32)), k
  assert not np.array_equal(w_mid['ctrl_cnn_0_0_gamma'], w['ctrl_cnn_0_0_gamma'])  # BN stays trainable
  m, v, step = a.state()
  st = ck.pack_state(w_mid, m, v, step)
  w2, m2, v2, step2 = ck.unpack_state(st)
  b = optim.AdamOptimizer(opt, w2, frozen=frozen)
  b.load_state(m2, v2, step2)
  for g in grads[2:]:
    a.step(g.clone())
    b.step(g.clone())
  assert torch.equal(a.params, b.params) and torch.equal(a.v, b.v) and a.global_step == b.global_step == 4
  with pytest.raises(Exception):
    b.load_state(m2[:-1], v2, 0)
